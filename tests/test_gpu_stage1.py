"""GPU parity tests, Stage 1 (BM25 backbone): the CUDA path through the C-ABI vs the oracle, bit-exact
(DocumentId order and float32 score bits). Run on the B200 box with `pytest -m gpu`."""
import numpy as np
import pytest

import infidex_b200 as ib
from infidex_b200 import synth
from oracle.oracle import Field as OField
from oracle.oracle import OracleEngine

pytestmark = pytest.mark.gpu


def _compare(eng, orc, queries, depth=500):
    keys, scores, n, status = eng.Stage1Batch(queries, depth)
    bad = []
    for i, q in enumerate(queries):
        r = orc.stage1(q, depth)
        if r["status"] != 0 or status[i] != 0:
            if (r["status"] != 0) != (status[i] != 0):
                bad.append((q, "status", r["status"], int(status[i])))
            continue
        ok = n[i] == len(r["keys"]) and np.array_equal(keys[i, : n[i]], r["keys"]) and \
            np.array_equal(scores[i, : n[i]].view(np.uint32), r["scores"].view(np.uint32))
        if not ok:
            bad.append((q, int(n[i]), len(r["keys"]), r["path"], r["candidates"]))
    assert not bad, bad[:10]


def test_stage1_movies(movie_titles, oracle_movies):
    eng = ib.SearchEngine.CreateDefault()
    eng.IndexColumns(np.arange(len(movie_titles)), [ib.Field("content")], [movie_titles])
    qs = ["redemption sh", "Shawshank", "Shaaawshank", "the amtrix", "star", "the hear", "fellowship of the ring", "te matri", "san a",
          "batman", "the", "love", "new york", "harry potter and the", "x-men", "zzzzqqq", "matrix reloaded", "lord of the rings",
          "the lord of the rings the return of the king", "a", "  ", "spider-man", "o'brien", "Música", "amelie"]
    _compare(eng, oracle_movies, qs)


@pytest.mark.parametrize("multi", [False, True])
def test_stage1_synthetic(multi):
    vocab = synth.make_vocab(100_000)
    n = 300_000 if not multi else 100_000
    docs = synth.gen_docs(n, vocab, with_description=multi)
    qs = synth.gen_queries(600, docs, vocab)
    schema, cols = synth.schema_and_columns(docs, multi)
    eng = ib.SearchEngine.CreateDefault(); eng.IndexColumns(docs["keys"], schema, cols)
    orc = OracleEngine([OField(f.Name, f.Weight, f.Indexable, f.Filterable, f.Facetable) for f in schema]); orc.index_columns(docs["keys"], cols)
    _compare(eng, orc, qs)
    _compare(eng, orc, qs[:100], depth=50)     # small K: pruning heap saturates early (Q1b / Q2 paths)
