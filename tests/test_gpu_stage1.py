"""GPU parity tests, Stage 1 (BM25 backbone): the CUDA path through the C-ABI vs the oracle, bit-exact
(DocumentId order and float32 score bits). Run on the B200 box with `pytest -m gpu`."""
import numpy as np
import pytest

import infidex_b200 as ib
from infidex_b200 import synth
from parity_util import build_pair, compare_stage1

pytestmark = pytest.mark.gpu

MOVIE_QUERIES = ["redemption sh", "Shawshank", "Shaaawshank", "the amtrix", "star", "the hear", "fellowship of the ring", "te matri", "san a",
                 "batman", "the", "love", "new york", "harry potter and the", "x-men", "zzzzqqq", "matrix reloaded", "lord of the rings",
                 "the lord of the rings the return of the king", "a", "  ", "spider-man", "o'brien", "Música", "amelie"]


def test_stage1_movies(movie_titles, oracle_movies):
    eng = ib.SearchEngine.CreateDefault()
    eng.IndexColumns(np.arange(len(movie_titles)), [ib.Field("content")], [movie_titles])
    bad = compare_stage1(eng, oracle_movies, MOVIE_QUERIES)
    assert not bad, bad[:5]


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
