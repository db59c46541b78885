"""The 18-book library of the reference's FacetingTests.cs (multi-field documents: title High, author Med + facetable, year Low and
NOT indexable + facetable, genre Low + facetable, description Med): the reference's own assertions on the oracle, then the product
logic (kernel emulation) against the oracle with filters and facets. Fixture: tests/golden/books.json (make_book_fixture.py)."""
import json
import os

import numpy as np
import pytest

import infidex_b200 as ib
from oracle.oracle import Field as OField
from oracle.oracle import OracleEngine
from parity_util import compare_search, compare_stage1, emu_lib

BOOKS = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "books.json"), encoding="utf-8"))
F = ib.Filter


def _schema():
    return [ib.Field("title", None, ib.Weight.High), ib.Field("author", None, ib.Weight.Med, facetable=True),
            ib.Field("year", None, ib.Weight.Low, indexable=False, facetable=True), ib.Field("genre", None, ib.Weight.Low, facetable=True),
            ib.Field("description", None, ib.Weight.Med)]


def _columns():
    return [[b[1] for b in BOOKS], [b[2] for b in BOOKS], [b[3] for b in BOOKS], [b[4] for b in BOOKS], [b[5] for b in BOOKS]]


@pytest.fixture(scope="module")
def oracle_books():
    o = OracleEngine([OField(f.Name, f.Weight, f.Indexable, f.Filterable, f.Facetable) for f in _schema()])
    o.index_columns(np.array([b[0] for b in BOOKS], np.int64), _columns())
    return o


def _by_id(i):
    return next(b for b in BOOKS if b[0] == i)


def test_reference_assertions(oracle_books):
    r = oracle_books.search("magic", 20, facets=True)                                   # Facets_BookSearch_ShowsAuthorYearGenreFacets (:107-132)
    assert r["keys"] and r["facets"]
    r = oracle_books.search("harry potter", 20, facets=True)                            # Facets_BookSearch_AuthorFaceting (:134-158)
    assert len(r["keys"]) >= 3
    r = oracle_books.search("magic fantasy adventure", 30, facets=True, filter_bytes=F.Range("year", "2000", None).bytecode())     # :160-236
    assert r["keys"] and all(int(_by_id(k)[3]) >= 2000 for k in r["keys"])
    flt = F.And(F.Value("genre", "Fantasy"), F.Range("year", "2000", None))             # CompositeFilter_FantasyAfter2000 (:258-313)
    r = oracle_books.search("magic adventure", 30, facets=True, filter_bytes=flt.bytecode())
    assert r["keys"] and all(_by_id(k)[4] == "Fantasy" and int(_by_id(k)[3]) >= 2000 for k in r["keys"])
    flt = F.Or(F.Value("author", "J.K. Rowling"), F.Value("author", "Stephen King"))    # CompositeFilter_RowlingOrKing (:315-)
    r = oracle_books.search("magic dark", 30, facets=True, filter_bytes=flt.bytecode())
    assert r["keys"] and all(_by_id(k)[2] in ("J.K. Rowling", "Stephen King") for k in r["keys"])
    assert oracle_books.search("stone philosopher", 10)["keys"][0] == 1                 # Facets_BookSearch_RecentPublications (:238-256)


def test_emu_matches_oracle_on_the_book_library(oracle_books):
    eng = ib.SearchEngine(_gpu_lib=emu_lib())
    eng.IndexColumns(np.array([b[0] for b in BOOKS], np.int64), _schema(), _columns())
    qs = ["magic", "harry potter", "magic fantasy adventure", "magic adventure", "magic dark", "stone philosopher", "wizard school",
          "rowling", "stephen king horror", "hogwarts", "fantasy", "dune", "chamber secrets", "magik", "potter stone"]
    assert not compare_stage1(eng, oracle_books, qs)
    assert not compare_search(eng, oracle_books, qs, max_results=20, facets=True)
    for flt in (F.Range("year", "2000", None), F.And(F.Value("genre", "Fantasy"), F.Range("year", "2000", None)),
                F.Or(F.Value("author", "J.K. Rowling"), F.Value("author", "Stephen King")), F.Not(F.Value("genre", "Fantasy")),
                F.In("genre", ["Horror", "Mystery"]), F.String("author", "CONTAINS", "king")):
        assert not compare_search(eng, oracle_books, qs, max_results=30, flt=flt, facets=True), flt


def test_emu_blank_query_with_facets_browses_the_corpus(oracle_books):
    """SearchEngine.HandleEmptyQueryWithFacets (SearchEngine.cs:321-346): a blank query with EnableFacets returns the first live documents
    (id order, score 65535) that pass the filter, cut to max, and the facets over them; without EnableFacets the result stays empty."""
    eng = ib.SearchEngine(_gpu_lib=emu_lib())
    eng.IndexColumns(np.array([b[0] for b in BOOKS], np.int64), _schema(), _columns())
    for mx in (5, 10, 50):
        assert not compare_search(eng, oracle_books, ["", "   "], max_results=mx, facets=True)
        for flt in (F.Range("year", "2000", None), F.Value("genre", "Fantasy"), F.Not(F.Value("genre", "Fantasy"))):
            assert not compare_search(eng, oracle_books, ["", " "], max_results=mx, flt=flt, facets=True), (mx, flt)
    r = eng.Search(ib.Query("", 10)); assert r.Records == [] and not r.Facets
    q = ib.Query("  ", 7); q.EnableFacets = True
    r = eng.Search(q); assert len(r.Records) == 7 and all(e.Score == 65535.0 for e in r.Records) and r.Facets and r.TotalCandidates == 0
