"""The host index builder (csrc/ifx_host_build.cpp -> ifx_index_image) against the oracle's index, structure by structure:
term dictionary in TermCollection order, df, posting lists with their tf bytes, document lengths, avgdl. No GPU involved."""
import ctypes as C

import numpy as np
import pytest

import infidex_b200 as ib
from infidex_b200 import synth
from oracle.oracle import Field as OField
from oracle.oracle import OracleEngine
from parity_util import emu_lib


class _Strings(C.Structure):
    _fields_ = [("chars", C.POINTER(C.c_uint16)), ("off", C.POINTER(C.c_uint32)), ("n", C.c_int32)]


class _Docset(C.Structure):
    _fields_ = [("keys", _Strings), ("row_ptr", C.POINTER(C.c_int64)), ("doc_id", C.POINTER(C.c_int32))]


class _Image(C.Structure):      # include/infidex_gpu.h: ifx_index_image, field for field (up to the parts compared here)
    _fields_ = [("n_docs", C.c_int32), ("n_live", C.c_int32), ("avgdl", C.c_float),
                ("doc_key", C.POINTER(C.c_int64)), ("deleted", C.POINTER(C.c_uint8)), ("doc_len", C.POINTER(C.c_float)),
                ("text_chars", C.POINTER(C.c_uint16)), ("text_off", C.POINTER(C.c_int64)), ("first_token", _Strings),
                ("token_count", C.POINTER(C.c_uint16)), ("terms", _Strings), ("df", C.POINTER(C.c_int32)),
                ("row_ptr", C.POINTER(C.c_int64)), ("post_doc", C.POINTER(C.c_int32)), ("post_tf", C.POINTER(C.c_uint8)),
                ("words", _Strings), ("word_idf", C.POINTER(C.c_float)), ("prefix", _Docset)]


def _strings(s, i):
    a, b = s.off[i], s.off[i + 1]
    return np.ctypeslib.as_array(s.chars, shape=(s.off[s.n],))[a:b].tobytes().decode("utf-16-le") if b > a else ""


@pytest.mark.parametrize("multi", [False, True])
def test_image_matches_oracle_index(multi):
    vocab = synth.make_vocab(20_000)
    docs = synth.gen_docs(4_000, vocab, with_description=multi)
    schema, cols = synth.schema_and_columns(docs, multi)
    eng = ib.SearchEngine(_gpu_lib=emu_lib()); eng.IndexColumns(docs["keys"], schema, cols)
    orc = OracleEngine([OField(f.Name, f.Weight, f.Indexable, f.Filterable, f.Facetable) for f in schema]); orc.index_columns(docs["keys"], cols)
    img = C.cast(eng.image_ptr(), C.POINTER(_Image)).contents
    n = img.n_docs
    assert n == docs["n"] == img.n_live
    assert np.float32(img.avgdl).view(np.uint32) == np.float32(orc.avgdl()).view(np.uint32)
    assert np.array_equal(np.ctypeslib.as_array(img.doc_len, shape=(n,)).view(np.uint32), orc.doc_lens().view(np.uint32))
    T = img.terms.n
    assert T == orc.num_terms()
    row = np.ctypeslib.as_array(img.row_ptr, shape=(T + 1,)); P = int(row[T])
    pd = np.ctypeslib.as_array(img.post_doc, shape=(max(P, 1),)); pt = np.ctypeslib.as_array(img.post_tf, shape=(max(P, 1),))
    df = np.ctypeslib.as_array(img.df, shape=(T,))
    rng = np.random.Generator(np.random.PCG64(11))
    for t in list(range(min(T, 300))) + rng.integers(0, T, 700).tolist():          # the first ordinals and a random sample
        text, odf, odocs, ow = orc.term(int(t))
        assert _strings(img.terms, t) == text, t
        assert int(df[t]) == odf, (t, text)
        a, b = int(row[t]), int(row[t + 1])
        assert np.array_equal(pd[a:b], odocs) and np.array_equal(pt[a:b], ow), (t, text)
