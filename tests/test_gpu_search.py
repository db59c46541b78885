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


@pytest.fixture(scope="module")
def movie_engine(movie_titles):
    eng = ib.SearchEngine.CreateDefault()
    eng.IndexColumns(np.arange(len(movie_titles)), [ib.Field("content")], [movie_titles])
    return eng


MOVIE_CASES = json.load(open(os.path.join(HERE, "golden", "movie_known_answers.json")))["cases"]


@pytest.mark.parametrize("case", MOVIE_CASES, ids=[c["query"].strip() or "blank" for c in MOVIE_CASES])
def test_movie_known_answer(case, movie_engine, movie_titles):
    """One test per case of MovieSearchParityTests.cs (tests/golden/movie_known_answers.json)."""
    r = movie_engine.Search(ib.Query(case["query"], case["max"]))
    assert r.Status & ~8 == 0
    check_movie_case(case, [x.DocumentId for x in r.Records], [x.Score for x in r.Records], movie_titles)


SHORT_QUERIES = ["a", "x", "th", "io", "as am", "a b", "é", "of", "I", "to be", "x y z", "q", "zz", "it", "9"]      # no word of >= 3 characters


def test_short_queries_vs_oracle(movie_engine, oracle_movies):
    """ShortQueryResolver champion lists (max 10), the single-character scan (max 100), SearchShortQuery incl. its fuzzy fallback."""
    for mx in (10, 100):
        bad = compare_search(movie_engine, oracle_movies, SHORT_QUERIES, max_results=mx)
        assert not bad, (mx, bad[:3])


def test_movies_vs_oracle(movie_engine, oracle_movies):
    eng = movie_engine
    bad = compare_search(eng, oracle_movies, MOVIE_QUERIES + ["sap", "two fo", "two f", "shawsh"] + SHORT_QUERIES)
    assert not bad, bad[:5]
    bad = compare_search(eng, oracle_movies, ["star", "sap", "the"], max_results=500)
    assert not bad, bad[:5]
