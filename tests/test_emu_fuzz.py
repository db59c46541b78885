"""Randomised parity: the kernel sources (host emulation) against the oracle on many small random corpora -- alphabets with diacritics,
upper case, Cyrillic and delimiter characters; 1 to a few thousand documents; single- and multi-field; typos, short words, unknown
words; varying result limits and coverage depths. Bit-exact comparison as everywhere (tests/parity_util.py)."""
import random

import numpy as np
import pytest

import infidex_b200 as ib
from oracle.oracle import Field as OField
from oracle.oracle import OracleEngine
from parity_util import compare_search, compare_stage1, emu_lib

ALPHABETS = ["abcdefghij", "abcdeéèüößñ", "abcABCdeF", "aábcčdďeěéiíjklmnoópqrřsštťuúůvwxyýzž", "абвгдежзиклмн", "abc-de/fg.hi_jk"]


def _word(rng, alph, lo=1, hi=9):
    return "".join(rng.choice(alph) for _ in range(rng.randint(lo, hi)))


def _case(seed, emu):
    rng = random.Random(seed)
    alph = rng.choice(ALPHABETS); vocab = [_word(rng, alph) for _ in range(rng.choice([5, 20, 80, 400]))]
    nd = rng.choice([1, 3, 17, 120, 700, 3000]); multi = rng.random() < 0.3

    def title():
        return rng.choice([" ", "  ", "-", ", "]).join(rng.choice(vocab) for _ in range(rng.randint(1, rng.choice([3, 6, 25]))))
    titles = [title() for _ in range(nd)]
    keys = np.array(rng.sample(range(1, 10 * nd + 10), nd), np.int64)
    if multi:
        schema = [ib.Field("title", None, ib.Weight.High), ib.Field("description", None, ib.Weight.Low)]
        cols = [titles, [title() + " " + title() for _ in range(nd)]]
    else:
        schema = [ib.Field("content", None, ib.Weight.Med)]; cols = [titles]
    eng = ib.SearchEngine(_gpu_lib=emu); eng.IndexColumns(keys, schema, cols)
    orc = OracleEngine([OField(f.Name, f.Weight, f.Indexable, f.Filterable, f.Facetable) for f in schema]); orc.index_columns(keys, cols)
    qs = []
    for _ in range(12):
        r = rng.random()
        if r < 0.5:
            ws = rng.choice(titles).split(); q = " ".join(ws[: rng.randint(1, min(3, len(ws)))]) if ws else ""
        elif r < 0.8:
            q = " ".join(rng.choice(vocab) for _ in range(rng.randint(1, 3)))
        else:
            q = _word(rng, alph, 1, 12)
        if q and rng.random() < 0.4:       # one edit
            p = rng.randrange(len(q)); op = rng.randint(0, 2)
            q = q[:p] + (rng.choice(alph) + q[p + 1:] if op == 0 else (q[p + 1:] if op == 1 else rng.choice(alph) + q[p:]))
        qs.append(q)
    mr = rng.choice([1, 3, 10, 50]); depth = rng.choice([5, 50, 500])
    return compare_stage1(eng, orc, qs, depth=depth) + compare_search(eng, orc, qs, max_results=mr, depth=depth)


@pytest.mark.parametrize("block", range(4))
def test_emu_random_corpora(block):
    emu = emu_lib()
    for seed in range(block * 40, block * 40 + 40):
        bad = _case(seed, emu)
        assert not bad, (seed, bad[:1])


def _q(text, flt):
    q = ib.Query(text, 10); q.Filter = flt
    return q


def _filter_case(seed, emu):
    """Multi-field documents with numeric / string columns, random Infiscript filters, facets on."""
    rng = random.Random(10_000 + seed)
    alph = rng.choice(ALPHABETS[:4]); vocab = [_word(rng, alph, 2, 8) for _ in range(rng.choice([10, 60]))]
    nd = rng.choice([5, 60, 400])
    titles = [" ".join(rng.choice(vocab) for _ in range(rng.randint(1, 5))) for _ in range(nd)]
    years = np.array([rng.randint(1950, 2024) for _ in range(nd)], np.int64)
    ratings = np.array([round(rng.randint(10, 100) / 10.0, 1) for _ in range(nd)], np.float64)
    genres = [rng.choice(["Fantasy", "Horror", "SciFi", "drama", "", "Sci-Fi Noir"]) for _ in range(nd)]
    schema = [ib.Field("title", None, ib.Weight.High), ib.Field("year", None, ib.Weight.Med, indexable=False, filterable=True, facetable=True),
              ib.Field("rating", None, ib.Weight.Med, indexable=False, filterable=True), ib.Field("genre", None, ib.Weight.Low, filterable=True, facetable=True)]
    cols = [titles, years, ratings, genres]; keys = np.arange(100, 100 + nd, dtype=np.int64)
    eng = ib.SearchEngine(_gpu_lib=emu); eng.IndexColumns(keys, schema, cols)
    orc = OracleEngine([OField(f.Name, f.Weight, f.Indexable, f.Filterable, f.Facetable) for f in schema]); orc.index_columns(keys, cols)
    qs = [" ".join(rng.choice(titles).split()[:2]) for _ in range(8)]
    exprs = ["year >= %d" % rng.randint(1950, 2024), "year BETWEEN 1980 AND %d" % rng.randint(1981, 2024), "rating > %.1f AND year < %d" % (rng.randint(10, 90) / 10, rng.randint(1960, 2024)),
             "genre = 'fantasy' OR genre = 'HORROR'", "NOT genre = ''", "genre IN ('SciFi', 'drama') && rating <= 7.5", "genre CONTAINS 'sci'", "genre STARTS WITH 'sci' | genre ENDS WITH 'noir'",
             "genre LIKE '%%i%%' AND ! (year < 1990)", "rating != 5.0", "genre IS NOT NULL AND (year > 2000 OR rating >= 9)"]
    with pytest.raises(ValueError):        # `title` is indexable only: no device column -- the mirror must refuse, not answer "no match"
        eng.Search(_q(qs[0], ib.Filter.Parse("title CONTAINS 'a'")))
    bad = []
    for e in rng.sample(exprs, 4):
        bad += compare_search(eng, orc, qs, max_results=rng.choice([3, 10, 40]), flt=ib.Filter.Parse(e), facets=True)
    return bad


@pytest.mark.parametrize("block", range(2))
def test_emu_random_filters_and_facets(block):
    emu = emu_lib()
    for seed in range(block * 30, block * 30 + 30):
        bad = _filter_case(seed, emu)
        assert not bad, (seed, bad[:1])
