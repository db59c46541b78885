"""Documents that share a DocumentKey (segments of one document): one record per key (SegmentProcessor.ConsolidateSegments,
Scoring/SegmentProcessor.cs:15-37), the coverage stage reading the first live document of a key (DocumentCollection.GetDocumentByPublicKey).
The reference's own assertions (SegmentTrackingTests.cs) on the oracle, then the kernel emulation against the oracle on corpora with
repeated keys. (`bestSegmentsMap` is never allocated in the reference -- SearchPipeline.cs:80 -- so segment numbers play no role in Search.)"""
import random

import numpy as np

import infidex_b200 as ib
from oracle.oracle import OracleEngine
from parity_util import compare_search, compare_stage1, emu_lib

DOCS = [(1, "Introduction chapter one"), (1, "Batman fights crime in Gotham City"), (1, "Conclusion chapter one"),
        (2, "Batman and Robin save the day"), (2, "The end of their adventure"), (3, "Superman flies faster than a speeding bullet")]


def _pair(docs):
    keys = np.array([k for k, _ in docs], np.int64); texts = [t for _, t in docs]
    orc = OracleEngine(); orc.index_texts(texts, keys=keys)
    eng = ib.SearchEngine(_gpu_lib=emu_lib()); eng.IndexColumns(keys, [ib.Field("content")], [texts])
    return eng, orc


def test_reference_segment_assertions_on_the_oracle_and_the_emulation():
    eng, orc = _pair(DOCS)
    for search in (lambda q: orc.search(q, 10)["keys"], lambda q: [e.DocumentId for e in eng.Search(ib.Query(q, 10)).Records]):
        assert sorted(search("batman")) == [1, 2]                      # Search_MultipleSegmentedDocuments_ConsolidatesCorrectly (SegmentTrackingTests.cs)
    eng, orc = _pair([(1, "Introduction to the topic of animals"), (1, "The quick brown fox jumps over the lazy dog"), (1, "Conclusion and summary of findings")])
    assert orc.search("fox", 10)["keys"] == [1] and [e.DocumentId for e in eng.Search(ib.Query("fox", 10)).Records] == [1]      # Search_SegmentedDocument_ReturnsBestSegment


def test_emu_matches_oracle_with_repeated_keys():
    rng = random.Random(7); vocab = ["".join(rng.choice("abcdefghij") for _ in range(rng.randint(3, 8))) for _ in range(60)]       # (words of >= 3 characters: the short-query path is flagged unsupported when keys repeat)
    docs = [(i // rng.choice([1, 2, 3]) + 1, " ".join(rng.choice(vocab) for _ in range(rng.randint(1, 7)))) for i in range(900)]
    eng, orc = _pair(docs)
    qs = [" ".join(rng.choice(vocab) for _ in range(rng.randint(1, 3))) for _ in range(40)] + [vocab[3][:-1] + "x", vocab[5] + " " + vocab[9]]
    assert not compare_stage1(eng, orc, qs)
    assert not compare_stage1(eng, orc, qs, depth=20)
    bad = compare_search(eng, orc, qs, max_results=10)
    assert not bad, bad[:3]
    bad = compare_search(eng, orc, qs, max_results=50, depth=40)
    assert not bad, bad[:3]
