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
