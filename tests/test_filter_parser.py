"""Filter.Parse of the Python mirror against the reference's FilterParserTests.cs: structure (operator precedence, aliases,
case-insensitive keywords), the error cases, and -- through the oracle's VM -- the meaning of the parsed filters."""
import numpy as np
import pytest

import infidex_b200 as ib
from oracle.oracle import Field as OField
from oracle.oracle import OracleEngine

P = ib.Filter.Parse


def test_precedence_and_structure():                      # FilterParserTests.cs:207-260
    f = P("a = '1' OR b = '2' AND c = '3'"); assert f.kind == "or" and f.right.kind == "and"
    f = P("(a = '1' OR b = '2') AND c = '3'"); assert f.kind == "and" and f.left.kind == "or"
    f = P("((a = '1' OR b = '2') AND c = '3') OR d = '4'"); assert f.kind == "or" and f.left.kind == "and" and f.left.left.kind == "or"
    f = P("NOT a = '1'"); assert f.kind == "not" and f.left.kind == "value"
    f = P("NOT (status = 'inactive' OR deleted = 'true')"); assert f.kind == "not" and f.left.kind == "or"
    f = P("status != 'inactive'"); assert f.kind == "not" and f.left.kind == "value" and f.left.value == "inactive"     # FilterParser.cs:445


@pytest.mark.parametrize("a,b", [
    ("genre = 'Fantasy' AND year >= '2000'", "genre = 'Fantasy' && year >= '2000'"), ("genre = 'Fantasy' AND year >= '2000'", "genre = 'Fantasy' & year >= '2000'"),
    ("genre = 'Fantasy' AND year >= '2000'", "  genre   =   'Fantasy'   AND   year   >=   '2000'  "), ("genre = 'Fantasy' AND year >= '2000'", "genre='Fantasy'AND year>='2000'"),
    ("author = 'Rowling' OR author = 'King'", "author = 'Rowling' || author = 'King'"), ("author = 'Rowling' OR author = 'King'", "author = 'Rowling' | author = 'King'"),
    ("NOT status = 'inactive'", "! status = 'inactive'"), ("NOT a = '1'", "not a = '1'"), ("a = '1' AND b = '2'", "a = '1' and b = '2'"), ("a = '1' AND b = '2'", "a = '1' And b = '2'"),
    ("a = '1' OR b = '2'", "a = '1' or b = '2'"), ("title CONTAINS 'test'", "title contains 'test'"), ("title LIKE '%test%'", "title like '%test%'"),
    ("name STARTS WITH 'John'", "name starts with 'John'"), ("email ENDS WITH '.com'", "email Ends With '.com'"), ("name MATCHES '^John'", "name matches '^John'"),
    ("genre IN ('Fantasy')", "genre in ('Fantasy')"), ("description IS NULL", "description is null"), ("author IS NOT NULL", "author Is Not Null"),
    ("price BETWEEN '10' AND '100'", "price between '10' and '100'"), ("name = 'Jane Smith'", 'name = "Jane Smith"'),
    ("(genre = 'Fantasy' && year >= '2000') || (genre = 'Horror' & year >= '1970')", "(genre = 'Fantasy' AND year >= '2000') OR (genre = 'Horror' AND year >= '1970')"),
])
def test_aliases_and_case_insensitive_keywords(a, b):     # :60-205, 262-420: same filter, hence the same Infiscript bytecode
    assert P(a).bytecode() == P(b).bytecode()


@pytest.mark.parametrize("expr", ["", "name = 'unterminated", "(genre = 'Fantasy'", "genre =", "genre 'Fantasy'", "genre === 'Fantasy'"])
def test_errors(expr):                                    # :612-655 (FilterParseException)
    with pytest.raises(ib.FilterParseError):
        P(expr)


def test_meaning_through_the_oracle_vm():
    """The parsed filters evaluated by the oracle's FilterVM on a few documents (fields are read as strings, numeric compare
    when both sides parse as numbers: FilterVM.cs:48-357)."""
    rows = [dict(genre="Fantasy", year="2005", author="Rowling", price="50", title="Harry Potter", email="a@x.com", description="d"),
            dict(genre="Horror", year="1975", author="King", price="150", title="The Shining", email="k@y.org", description="d"),
            dict(genre="SciFi", year="1999", author="Herbert", price="7", title="Dune test", email="h@z.com", description="d")]
    names = sorted(rows[0]); o = OracleEngine([OField(n, 1, n == "title", True, False) for n in names])
    o.index_columns(np.arange(len(rows)), [[r[n] for r in rows] for n in names])

    def sel(expr): return [i for i in range(len(rows)) if o.filter_eval(P(expr).bytecode(), i)]
    assert sel("genre = 'Fantasy' AND year >= '2000'") == [0]
    assert sel("(genre = 'Fantasy' && year >= '2000') || (genre = 'Horror' & year >= '1970')") == [0, 1]
    assert sel("author = 'Rowling' | author = 'King'") == [0, 1]
    assert sel("! genre = 'Fantasy'") == [1, 2] and sel("genre != 'Fantasy'") == [1, 2]
    assert sel("price > 100") == [1] and sel("price < '500'") == [0, 1, 2] and sel("price BETWEEN '10' AND '100'") == [0]
    assert sel("genre IN ('Fantasy', 'SciFi', 'Horror')") == [0, 1, 2] and sel("genre IN ('Fantasy')") == [0]
    assert sel("title CONTAINS 'test'") == [2] and sel("title STARTS WITH 'The'") == [1] and sel("email ENDS WITH '.com'") == [0, 2]
    assert sel("title LIKE '%Potter%'") == [0] and sel("description IS NOT NULL") == [0, 1, 2] and sel("description IS NULL") == []
    assert sel("a = '1' OR genre = 'Horror' AND year >= '2000'") == [] and sel("(a = '1' OR genre = 'Horror') AND year >= '1970'") == [1]


# ---- TernaryFilterTests.cs: cond ? a : b (lowest precedence, right-associative), literal branches ---------------------------------------
def test_ternary_structure():
    f = P("score >= 90 ? status = 'premium' : status = 'basic'"); assert f.kind == "ternary" and f.cond.kind == "range" and f.left.kind == "value"   # :26-44
    f = P("a = '1' ? b = '2' : c = '3' ? d = '4' : e = '5'"); assert f.kind == "ternary" and f.right.kind == "ternary" and f.left.kind == "value"  # :228-246
    f = P("a = '1' OR b = '2' ? c = '3' : d = '4'"); assert f.kind == "ternary" and f.cond.kind == "or"                                             # :197-226
    f = P("(a = '1' ? b = '2' : c = '3') AND d = '4'"); assert f.kind == "and" and f.left.kind == "ternary"                                          # :136-157
    f = P("age >= 18 ? 'adult' : 'minor'"); assert (f.left.kind, f.left.value, f.right.value) == ("literal", "adult", "minor")                        # :340-359
    f = P("premium = 'yes' ? 100 : 50"); assert (f.left.value, f.right.value) == (100.0, 50.0)                                                        # :377-394
    f = P("available = 'yes' ? price >= 100 : 'unavailable'"); assert f.left.kind == "range" and f.right.kind == "literal"                            # :396-
    assert b"VIP" in P("premium = 'yes' ? 'VIP' : 'Standard'").bytecode() and b"Standard" in P("premium = 'yes' ? 'VIP' : 'Standard'").bytecode()     # :361-375


@pytest.mark.parametrize("expr", ["score >= 90 ? 'high'", "? 'yes' : 'no'", "score >= 90 ? : 'low'", "score >= 90 ? 'high' :"])
def test_ternary_errors(expr):                            # :247-282
    with pytest.raises(ib.FilterParseError):
        P(expr)


def test_ternary_meaning_oracle_and_product():
    """Execute_SimpleTernary_True (:46-60) and friends on the oracle's VM, then the product's VM (kernel emulation) on a small corpus."""
    from parity_util import compare_search, emu_lib
    rows = [dict(title="alpha premium member", score="95", status="premium"), dict(title="alpha basic member", score="40", status="basic"),
            dict(title="alpha odd member", score="95", status="basic"), dict(title="alpha other member", score="10", status="premium")]
    names = ["title", "score", "status"]
    o = OracleEngine([OField("title", 1, True, False, False), OField("score", 1, False, True, False), OField("status", 1, False, True, True)])
    keys = np.arange(1, len(rows) + 1); cols = [[r[n] for r in rows] for n in names]; o.index_columns(keys, cols)
    sel = lambda e: [i + 1 for i in range(len(rows)) if o.filter_eval(P(e).bytecode(), i)]
    assert sel("score >= 90 ? status = 'premium' : status = 'basic'") == [1, 2]
    assert sel("score >= 90 ? status = 'premium' : score < 20 ? status = 'premium' : status = 'basic'") == [1, 2, 4]
    assert sel("score >= 90 ? 'high' : 'low'") == []            # a literal on top of the stack is not `true` (FilterVM.cs:26-46)
    eng = ib.SearchEngine(_gpu_lib=emu_lib())
    eng.IndexColumns(keys, [ib.Field("title"), ib.Field("score", None, ib.Weight.Med, indexable=False, filterable=True),
                            ib.Field("status", None, ib.Weight.Med, indexable=False, filterable=True, facetable=True)], cols)
    for e in ("score >= 90 ? status = 'premium' : status = 'basic'", "score >= 90 ? status = 'premium' : score < 20 ? status = 'premium' : status = 'basic'",
              "score >= 90 ? 'high' : 'low'"):
        assert not compare_search(eng, o, ["alpha member", "premium"], flt=P(e), facets=True), e
    # numeric constants in the pool are not executed on the device (their double.ToString() form would be needed): flagged, not guessed
    q = ib.Query("alpha member", 10); q.Filter = P("status = 'basic' ? 1 : score > 50")
    assert eng.Search(q).Status & 2
