"""GPU parity at scale (collected before the other GPU files on purpose: these are the tests that reach heap saturation, MaxScore
skips, several containers, skip tables / bitmaps / forward index, the AND tiers and > 1024 LD1 matches on hardware).
Every comparison goes through the C-ABI (ifx_search_batch / ifx_stage1_batch) against the oracle: DocumentId order, float32 Score
bits, Tiebreaker bytes, TotalCandidates, facet tables."""
import os
import threading

import numpy as np
import pytest

import infidex_b200 as ib
from infidex_b200 import dist as ifxd
from infidex_b200 import synth
from parity_util import build_pair, compare_search, compare_search_batch, compare_stage1

pytestmark = pytest.mark.gpu
C4_FILTER = "year >= 2000 AND rating > 7.0"


@pytest.mark.parametrize("multi", [False, True])
def test_stage1_synthetic(multi):
    vocab = synth.make_vocab(100_000)
    n = 300_000 if not multi else 100_000
    docs = synth.gen_docs(n, vocab, with_description=multi)
    qs = synth.gen_queries(600, docs, vocab)
    schema, cols = synth.schema_and_columns(docs, multi)
    eng, orc = build_pair(docs["keys"], schema, cols)
    bad = compare_stage1(eng, orc, qs)
    assert not bad, bad[:5]
    bad = compare_stage1(eng, orc, qs[:100], depth=50)     # small K: the pruning heap saturates early (Q1b / Q2 paths)
    assert not bad, bad[:5]


@pytest.mark.parametrize("multi", [False, True])
def test_search_synthetic(multi):
    vocab = synth.make_vocab(100_000)
    n = 300_000 if not multi else 100_000
    docs = synth.gen_docs(n, vocab, with_description=multi)
    qs = synth.gen_queries(600, docs, vocab)
    schema, cols = synth.schema_and_columns(docs, multi)
    eng, orc = build_pair(docs["keys"], schema, cols)
    bad = compare_search(eng, orc, qs)
    assert not bad, bad[:5]
    if multi:   # BASELINE.json configs[3]: Filter.Parse("year >= 2000 AND rating > 7.0") + EnableFacets
        flt = ib.Filter.Parse(C4_FILTER)
        bad = compare_search(eng, orc, qs[:300], flt=flt, facets=True)
        assert not bad, bad[:5]
        bad = compare_search(eng, orc, qs[:100], flt=ib.Filter.Parse("genre = 'drama' OR year < 1960"), facets=True, max_results=50)
        assert not bad, bad[:5]


def test_bench_corpus_configs1_timed_queries():
    """BASELINE.json configs[1] itself: the 1M-document corpus of `bench.py --workload c2` and the 1 000 queries of its first timed
    step (same generator, same seeds), full pipeline."""
    vocab = synth.make_vocab(400_000)
    docs = synth.gen_docs(1_000_000, vocab)
    schema, cols = synth.schema_and_columns(docs, False)
    eng, orc = build_pair(docs["keys"], schema, cols)
    qs = synth.gen_queries(1000, docs, vocab, seed=ifxd.rank_batch_seed(synth.SEED, 3, 0))     # step index 3 = first timed step at --warmup 3
    bad = compare_search_batch(eng, orc, qs)
    assert not bad, (len(bad), bad[:5])
    bad = compare_search(eng, orc, qs[:60])            # + TotalCandidates
    assert not bad, bad[:5]


def test_bench_corpus_multifield_slice_with_filter_and_facets():
    """A slice of the configs[2]/[3] corpus (same generator and schema as `bench.py --workload c3`; IFX_TEST_C3_DOCS documents,
    default 2M = 31 containers): plain queries, then Filter.Parse("year >= 2000 AND rating > 7.0") + EnableFacets."""
    n = int(os.environ.get("IFX_TEST_C3_DOCS", "2000000"))
    vocab = synth.make_vocab(400_000)
    docs = synth.gen_docs(n, vocab, with_description=True)
    schema, cols = synth.schema_and_columns(docs, True)
    eng, orc = build_pair(docs["keys"], schema, cols)
    qs = synth.gen_queries(400, docs, vocab, seed=ifxd.rank_batch_seed(synth.SEED, 3, 0))
    bad = compare_search_batch(eng, orc, qs)
    assert not bad, (len(bad), bad[:5])
    flt = ib.Filter.Parse(C4_FILTER)
    bad = compare_search_batch(eng, orc, qs, flt=flt)
    assert not bad, (len(bad), bad[:5])
    bad = compare_search(eng, orc, qs[:40], flt=flt, facets=True)
    assert not bad, bad[:5]


def test_many_unknown_words_per_query():
    """A batch whose queries carry many unknown words of >= 4 characters (up to MAX_FUZZY = 16 LD1 expansions each): the expansion
    work list must hold them all (it used to be sized for 4 per query)."""
    vocab = synth.make_vocab(30_000)
    docs = synth.gen_docs(60_000, vocab)
    schema, cols = synth.schema_and_columns(docs, False)
    eng, orc = build_pair(docs["keys"], schema, cols)
    rng = np.random.default_rng(5); words = vocab["words"]
    qs = []
    for i in range(100):
        ws = []
        for _ in range(int(rng.integers(6, 13))):
            w = words[int(rng.integers(0, len(words)))]
            if len(w) >= 4:
                p = int(rng.integers(0, len(w))); w = w[:p] + "q" + w[p + 1:]
            ws.append(w)
        qs.append(" ".join(ws))
    bad = compare_stage1(eng, orc, qs)
    assert not bad, bad[:5]
    bad = compare_search(eng, orc, qs[:30])
    assert not bad, bad[:5]


def test_concurrent_callers_on_gpu(movie_titles):
    """ThreadSafetyTests.cs in spirit, on hardware: several host threads call Search on one engine; every caller gets exactly its own answer."""
    eng = ib.SearchEngine.CreateDefault()
    eng.IndexColumns(np.arange(len(movie_titles)), [ib.Field("content")], [movie_titles])
    qs = ["star wars", "the matrix", "godfather", "lord rings", "toy story", "batman", "alien", "love", "night", "dark knight"]
    want = {q: [(e.DocumentId, e.Score, e.Tiebreaker) for e in eng.Search(ib.Query(q, 10)).Records] for q in qs}
    errors = []

    def worker(k):
        try:
            for i in range(40):
                q = qs[(i * 7 + k) % len(qs)]
                got = [(e.DocumentId, e.Score, e.Tiebreaker) for e in eng.Search(ib.Query(q, 10)).Records]
                if got != want[q]:
                    errors.append((k, q))
        except Exception as e:      # noqa: BLE001
            errors.append((k, repr(e)))
    ts = [threading.Thread(target=worker, args=(k,)) for k in range(8)]
    [t.start() for t in ts]; [t.join() for t in ts]
    assert not errors, errors[:3]
