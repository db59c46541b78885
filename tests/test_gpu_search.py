"""GPU parity tests, full search path (Stage 1 -> WordMatcher -> coverage/fusion -> truncation -> filter -> facets)
through ifx_search_batch vs the oracle: identical DocumentId order, Score bits, tie bytes, totals and facets."""
import json
import os

import numpy as np
import pytest

import infidex_b200 as ib
from conftest import REFERENCE_10
from infidex_b200 import synth
from oracle.oracle import OracleEngine
from parity_util import build_pair, compare_search
from test_gpu_stage1 import MOVIE_QUERIES
from test_oracle_known_answers import check_movie_case

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_readme_and_reference_corpus():
    eng = ib.SearchEngine.CreateDefault()
    eng.IndexDocuments([ib.Document(i + 1, t) for i, t in enumerate(REFERENCE_10[:3])])
    assert [r.DocumentId for r in eng.Search(ib.Query("quik fox", 10)).Records] == [1]          # BASELINE.json configs[0]
    eng = ib.SearchEngine.CreateDefault()
    eng.IndexDocuments([ib.Document(i + 1, t) for i, t in enumerate(REFERENCE_10)])
    expect = {"qick fux": [5, 1], "battamam": [6], "new york": [8], "speeding": [7]}            # ReferenceMatchingTests.cs:51-98
    for q, keys in expect.items():
        assert [r.DocumentId for r in eng.Search(ib.Query(q, 10)).Records] == keys
    assert eng.Search(ib.Query("batman", 10)).Records[0].DocumentId == 6


def test_movies_known_answers_and_oracle(movie_titles, oracle_movies):
    eng = ib.SearchEngine.CreateDefault()
    eng.IndexColumns(np.arange(len(movie_titles)), [ib.Field("content")], [movie_titles])
    for case in json.load(open(os.path.join(HERE, "golden", "movie_known_answers.json")))["cases"]:
        r = eng.Search(ib.Query(case["query"], case["max"]))
        check_movie_case(case, [x.DocumentId for x in r.Records], [x.Score for x in r.Records], movie_titles)
    bad = compare_search(eng, oracle_movies, MOVIE_QUERIES + ["sap", "two fo", "two f", "shawsh"])
    assert not bad, bad[:5]
    bad = compare_search(eng, oracle_movies, ["star", "sap", "the"], max_results=500)
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
        flt = ib.Filter.Parse("year >= 2000 AND rating > 7.0")
        bad = compare_search(eng, orc, qs[:300], flt=flt, facets=True)
        assert not bad, bad[:5]
        bad = compare_search(eng, orc, qs[:100], flt=ib.Filter.Parse("genre = 'drama' OR year < 1960"), facets=True, max_results=50)
        assert not bad, bad[:5]
