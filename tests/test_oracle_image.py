"""The oracle's benchmark-scale shortcut (Index::load_image, used by bench.py only): searches over an oracle whose index STATE was
taken from the flattened image of the product's host builder must equal searches over the oracle's own add_document/build."""
import numpy as np
import pytest

import infidex_b200 as ib
from infidex_b200 import synth
from oracle.oracle import Field as OField
from oracle.oracle import OracleEngine


@pytest.mark.parametrize("multi", [False, True])
def test_oracle_from_image_equals_own_build(multi):
    vocab = synth.make_vocab(20_000)
    docs = synth.gen_docs(12_000 if multi else 30_000, vocab, with_description=multi)
    schema, cols = synth.schema_and_columns(docs, multi)
    of = [OField(f.Name, f.Weight, f.Indexable, f.Filterable, f.Facetable) for f in schema]
    own = OracleEngine(of); own.index_columns(docs["keys"], cols)
    eng = ib.SearchEngine.__new__(ib.SearchEngine); eng._host = ib.engine._load_host(); eng._builder = None; eng._index = None; eng._gpu = None   # host builder only, no device
    eng.IndexColumns(docs["keys"], schema, cols, upload=False)
    img = OracleEngine(of); img.load_image(eng.image_ptr())
    qs = synth.gen_queries(150, docs, vocab) + ["zzzzqqq", "  "]
    flt = ib.Filter.Parse("year >= 2000 AND rating > 7.0").bytecode() if multi else None
    for q in qs:
        a = own.search(q, 10); b = img.search(q, 10)
        assert a["keys"] == b["keys"] and np.array_equal(a["scores"].view(np.uint32), b["scores"].view(np.uint32)) and a["ties"] == b["ties"] and a["total"] == b["total"], q
        x = own.stage1(q, 500); y = img.stage1(q, 500)
        assert np.array_equal(x["keys"], y["keys"]) and np.array_equal(x["scores"].view(np.uint32), y["scores"].view(np.uint32)), q
    if multi:
        for q in qs[:60]:
            a = own.search(q, 10, filter_bytes=flt, facets=True); b = img.search(q, 10, filter_bytes=flt, facets=True)
            assert a["keys"] == b["keys"] and a["facets"] == b["facets"] and a["total"] == b["total"], q
    assert img.search("ab", 10)["status"] != 0          # short queries need token positions: flagged, not answered, from an image
    eng._host.ifx_builder_destroy(__import__("ctypes").c_void_p(eng._builder)); eng._builder = None
