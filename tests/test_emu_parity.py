"""CPU-side logic tests: the kernel sources compiled as a single-threaded host emulation (tests/emu) against the oracle.
This checks the *logic* of the device code without a GPU; the real parity tests are the `-m gpu` ones."""
import numpy as np
import pytest

import infidex_b200 as ib
from conftest import REFERENCE_10
from infidex_b200 import synth
from oracle.oracle import OracleEngine
from parity_util import build_pair, compare_search, compare_stage1, emu_lib


@pytest.fixture(scope="module")
def emu():
    return emu_lib()


def test_emu_reference_corpus(emu):
    eng = ib.SearchEngine(_gpu_lib=emu)
    eng.IndexDocuments([ib.Document(i + 1, t) for i, t in enumerate(REFERENCE_10)])
    orc = OracleEngine(); orc.index_texts(REFERENCE_10, keys=np.arange(1, 11))
    qs = ["batman", "qick fux", "battamam", "new york", "speeding", "quik fox", "the", "fox", "gotham cty", "a", "wonder woman protects", ""]
    assert not compare_search(eng, orc, qs)
    assert not compare_stage1(eng, orc, qs)


def test_emu_movies(emu, movie_titles, oracle_movies):
    eng = ib.SearchEngine(_gpu_lib=emu)
    eng.IndexColumns(np.arange(len(movie_titles)), [ib.Field("content")], [movie_titles])
    qs = ["redemption sh", "Shaaawshank", "the amtrix", "star", "fellowship of the ring", "te matri", "san a", "the", "new york", "x-men", "Música"]
    assert not compare_stage1(eng, oracle_movies, qs)
    assert not compare_search(eng, oracle_movies, qs)


@pytest.mark.parametrize("multi", [False, True])
def test_emu_synthetic(emu, multi):
    vocab = synth.make_vocab(30_000)
    docs = synth.gen_docs(30_000 if not multi else 15_000, vocab, with_description=multi)
    qs = synth.gen_queries(150, docs, vocab)
    schema, cols = synth.schema_and_columns(docs, multi)
    eng, orc = build_pair(docs["keys"], schema, cols, gpu_lib=emu)
    assert not compare_stage1(eng, orc, qs)
    assert not compare_search(eng, orc, qs)
    if multi:
        flt = ib.Filter.Parse("year >= 2000 AND rating > 7.0")
        assert not compare_search(eng, orc, qs[:80], flt=flt, facets=True)
