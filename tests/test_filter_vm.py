"""Infiscript VM known answers restated from /root/reference/src/Infidex.Tests/BytecodeVMTests.cs (lines cited), checked on
(1) the oracle's VM and (2) the product's device VM (kernel emulation build here; the CUDA build in the gpu tests)."""
import numpy as np
import pytest

import infidex_b200 as ib
from infidex_b200 import Filter as F
from oracle.oracle import Field as OField
from oracle.oracle import OracleEngine
from parity_util import compare_search, emu_lib

# (filter, document fields, expected) -- BytecodeVMTests.cs
CASES = [
    (F.Value("genre", "Fantasy"), {"genre": "Fantasy"}, True),                     # :65-75
    (F.Value("genre", "Fantasy"), {"genre": "Horror"}, False),                     # :78-88
    (F.Value("genre", "fantasy"), {"genre": "FANTASY"}, True),                     # :91-101
    (F.Range("year", 2000, 2020), {"year": 2010}, True),                           # :108-118
    (F.Range("year", 2000, None), {"year": 2015}, True),                           # :121-131
    (F.Range("year", None, 2020), {"year": 2015}, True),                           # :134-144
    (F.Range("year", 2000, 2010), {"year": 2020}, False),                          # :147-157
    (F.String("title", "CONTAINS", "ring"), {"title": "The Fellowship of the Ring"}, True),      # :164-174
    (F.String("title", "STARTS_WITH", "the"), {"title": "The Matrix"}, True),      # :177-187
    (F.String("title", "ENDS_WITH", "matrix"), {"title": "The Matrix"}, True),     # :190-200
    (F.String("title", "LIKE", "The%Ring"), {"title": "The Fellowship of the Ring"}, True),       # :203-213
    (F.In("genre", ["Fantasy", "SciFi"]), {"genre": "SciFi"}, True),               # :250-260
    (F.In("genre", ["Fantasy", "SciFi"]), {"genre": "Horror"}, False),             # :263-273
    (F.And(F.Value("genre", "Fantasy"), F.Range("year", 2000, None)), {"genre": "Fantasy", "year": 2010}, True),    # :310-325
    (F.And(F.Value("genre", "Fantasy"), F.Range("year", 2000, None)), {"genre": "Fantasy", "year": 1990}, False),   # :328-343
    (F.Or(F.Value("genre", "Fantasy"), F.Value("genre", "Horror")), {"genre": "Horror", "year": 1}, True),          # :346-360
    (F.Or(F.Value("genre", "Fantasy"), F.Value("genre", "Horror")), {"genre": "Romance", "year": 1}, False),        # :363-377
    (F.Not(F.Value("genre", "Fantasy")), {"genre": "Horror"}, True),               # :380-390
    (F.Value("genre", ""), {"genre": ""}, True),                                   # :783-794
]
COMPLEX = F.Or(F.And(F.Value("genre", "Fantasy"), F.Range("year", 2000, None)), F.And(F.Value("genre", "Horror"), F.Range("year", 1980, None)))
CASES += [(COMPLEX, {"genre": "Fantasy", "year": 2010}, True), (COMPLEX, {"genre": "Horror", "year": 1990}, True),
          (COMPLEX, {"genre": "Romance", "year": 2000}, False)]                    # :396-432


def _engine_for(fields):
    names = sorted(fields)
    schema = [OField(n, 1, True, True, False) for n in names]
    e = OracleEngine(schema)
    cols = [np.array([fields[n]], np.int64) if isinstance(fields[n], int) else [fields[n]] for n in names]
    e.index_columns(np.array([1]), cols)
    return e


@pytest.mark.parametrize("idx", range(len(CASES)))
def test_oracle_vm_known_answers(idx):
    flt, fields, expected = CASES[idx]
    assert _engine_for(fields).filter_eval(flt.bytecode(), 0) == (1 if expected else 0)


def test_oracle_vm_missing_field_is_false():          # BytecodeVMTests.cs:757-767
    assert _engine_for({"other": "x"}).filter_eval(F.Value("genre", "Fantasy").bytecode(), 0) == 0


def test_serialized_header_and_parse():               # BytecodeVMTests.cs:451-460, 559-613
    code = F.Parse("genre = 'Fantasy' AND year >= 2000").bytecode()
    assert code.startswith(b"INFISCRIPT-V1") and code[13:15] == b"\x01\x00"
    assert F.Parse("year >= 2000 AND rating > 7.0").bytecode() == F.And(F.Range("year", min="2000"), F.Range("rating", min="7.0", include_min=False)).bytecode()


def vm_corpus_check(gpu_lib):
    """Device VM vs oracle VM through the whole search: every doc contains 'item', filters vary."""
    rng = np.random.default_rng(7); n = 400
    genres = ["Fantasy", "horror", "SciFi", "Romance", ""]
    title = ["item %s number%d" % (["alpha", "beta", "gamma"][i % 3], i) for i in range(n)]
    genre = [genres[i] for i in rng.integers(0, len(genres), n)]
    year = rng.integers(1950, 2025, n).astype(np.int64); rating = np.round(rng.integers(10, 101, n) / 10.0, 1)
    schema = [ib.Field("title", None, ib.Weight.High), ib.Field("genre", None, indexable=False, filterable=True, facetable=True),
              ib.Field("year", None, indexable=False, filterable=True, facetable=True), ib.Field("rating", None, indexable=False, filterable=True)]
    cols = [title, genre, year, rating]
    eng = ib.SearchEngine(_gpu_lib=gpu_lib) if gpu_lib else ib.SearchEngine.CreateDefault(); eng.IndexColumns(np.arange(n), schema, cols)
    orc = OracleEngine([OField(f.Name, f.Weight, f.Indexable, f.Filterable, f.Facetable) for f in schema]); orc.index_columns(np.arange(n), cols)
    filters = ["year >= 2000 AND rating > 7.0", "genre = 'fantasy'", "genre != 'Horror' AND year < 1990", "genre IN ('SciFi', 'romance') OR rating <= 2.5",
               "year BETWEEN 1980 AND 1999", "genre IS NULL", "genre IS NOT NULL AND NOT (year > 1960)", "genre CONTAINS 'an'", "genre STARTS WITH 'sci'",
               "genre ENDS WITH 'OR'", "genre LIKE '%o_ance'", "rating > 'abc'", "missing = 'x'"]
    for f in filters:
        flt = F.Parse(f)
        bad = compare_search(eng, orc, ["item", "item alpha", "number7"], max_results=500, flt=flt, facets=True)
        assert not bad, (f, bad[:3])


def test_product_vm_matches_oracle_on_a_corpus():
    vm_corpus_check(emu_lib())


@pytest.mark.gpu
def test_product_vm_matches_oracle_on_a_corpus_gpu():
    vm_corpus_check(None)
