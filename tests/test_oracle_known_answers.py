"""Pins the ORACLE (CPU restatement) against the reference's own known-answer tests.

Every expectation below is restated from a test in /root/reference/src/Infidex.Tests (file:line cited);
none comes from the oracle itself. The reference is C#/.NET and cannot run in this image, so these
known answers (exact result lists on the 10-doc corpus, top-1 / ordering relations on movies.csv) are
the strongest pin available; complete ranked lists and Score constants remain "parity unpinned" (SURVEY 8c).
"""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


# ReferenceMatchingTests.cs:40-98
@pytest.mark.parametrize("query,expected,exact", [
    ("batman", [6], False), ("qick fux", [5, 1], True), ("battamam", [6], True), ("new york", [8], True), ("speeding", [7], True)])
def test_reference_matching(oracle_ref10, query, expected, exact):
    r = oracle_ref10.search(query, 10)
    assert r["status"] == 0
    if exact:
        assert r["keys"] == expected
    else:
        assert r["keys"][: len(expected)] == expected


# README 3-doc example (BASELINE.json configs[0]; SURVEY 8c C1 hand trace: exactly one record, DocumentId 1)
def test_readme_quik_fox():
    e = O.OracleEngine()
    e.index_texts(["The quick brown fox jumps over the lazy dog", "A journey of a thousand miles begins with a single step",
                   "To be or not to be, that is the question"], keys=np.arange(1, 4))
    r = e.search("quik fox", 10)
    assert r["keys"] == [1]


# FuzzyRegressionTests.cs:20-52
def test_fuzzy_regression_matrix_above_mat():
    e = O.OracleEngine()
    e.index_texts(["The Mat", "The Matrix", "The Matriarx", "The Match", "The Meatrix"], keys=np.arange(1, 6))
    r = e.search("the matrx", 10)
    s = dict(zip(r["keys"], r["scores"]))
    assert 2 in s and (1 not in s or s[2] > s[1])


# LevenshteinDistanceTests.cs:11-79
@pytest.mark.parametrize("a,b,d", [("hello", "hello", 0), ("hello", "hallo", 1), ("bat", "brat", 1), ("batman", "batma", 1), ("", "", 0),
                                   ("hello", "", 5), ("", "hello", 5), ("kitten", "sitting", 3), ("saturday", "sunday", 3), ("abc", "xyz", 3),
                                   ("a" * 70 + "test", "a" * 70 + "best", 1)])
def test_levenshtein(a, b, d):
    assert O.levenshtein(a, b) == d


def test_levenshtein_within():
    assert O.levenshtein("batman", "batmam", 1) <= 1
    assert O.levenshtein("batman", "ratmin", 1) > 1


# CoverageEngineTests.cs:18-120, all seven tests (engine without corpus statistics: IDF falls back to log2(len+1))
def test_coverage_engine_known_answers():
    e = O.OracleEngine()
    c = e.coverage("hello world", "this is hello world text")
    assert c["coverage"] > 200 and c["word_hits"] == 2
    assert e.coverage("xyz abc", "hello world test")["coverage"] < 100
    c = e.coverage("hello world test", "hello world")
    assert c["coverage"] > 100 and c["word_hits"] == 2
    c = e.coverage("batmam", "batman is a superhero")
    assert c["coverage"] > 150 and c["word_hits"] > 0
    assert e.coverage("new york", "I live in newyork city")["coverage"] > 100          # :77-90 JoinedWords_DetectsCompound
    assert e.coverage("bat", "batman is a superhero")["coverage"] > 50                 # :92-105 PrefixMatch_FindsPartialWord
    c = e.coverage("", "hello world"); assert c["coverage"] == 0 and c["word_hits"] == 0   # :107-120 EmptyQuery_ReturnsZero


# MovieSearchParityTests.cs (33 tests; the n-gram-path ones are restated in movie_known_answers.json)
def _cases():
    return json.load(open(os.path.join(HERE, "golden", "movie_known_answers.json")))["cases"]


@pytest.mark.parametrize("case", _cases(), ids=lambda c: c["query"])
def test_movie_known_answers(oracle_movies, movie_titles, case):
    r = oracle_movies.search(case["query"], case["max"])
    assert r["status"] == 0
    check_movie_case(case, r["keys"], r["scores"], movie_titles)


def check_movie_case(c, keys, scores, titles_all):
    titles = [titles_all[k] for k in keys]
    if "top1" in c:
        assert titles and titles[0] == c["top1"]
    if "top1_contains" in c:
        assert titles and c["top1_contains"].lower() in titles[0].lower()
    if c.get("strict_gt_second"):
        assert len(titles) >= 2 and scores[0] > scores[1]
    if "min_results" in c:
        assert len(titles) >= c["min_results"]
    if "a_before_b_score" in c:
        a, b = c["a_before_b_score"]
        assert a in titles and b in titles and scores[titles.index(a)] > scores[titles.index(b)]
    if "a_before_b_rank" in c:
        a, b = c["a_before_b_rank"]
        assert a in titles and b in titles and titles.index(a) < titles.index(b)
    if "title_within_top" in c:
        a, k = c["title_within_top"]
        assert a in titles and titles.index(a) < k


def test_short_query_known_answers(oracle_movies, movie_titles):
    """Queries without a word of >= 3 characters (ShortQueryProcessor / ShortQueryResolver, oracle/shortquery.hpp): the reference's
    own assertions, MovieSearchParityTests.cs:557-622."""
    def titles_of(q, k=10):
        r = oracle_movies.search(q, k); assert r["status"] == 0
        return [movie_titles[i] for i in r["keys"]]
    t = titles_of("a"); assert t                                           # Search_SingleLetter_ReturnsResults
    for title in t[:5]:
        low = title.lower(); assert low.startswith("a") or " a" in low
    assert titles_of("x")[0] == "X"                                        # SingleLetter_X_PrefersExactTitle
    assert titles_of("th")                                                 # Search_TwoLetters_ReturnsResults
    assert titles_of("io")[0] == "IO"                                      # Io_PrefersExactTitleOverPrefixes
    assert oracle_movies.search("as am", 20)["status"] == 0                # several short words: SearchShortQuery, no coverage stage


def test_short_query_small_corpora():
    """MovieSearchParityTests.cs:1085-1140 (ShortQuery_SingleLetter_ReturnsAllMatches, ShortQuery_TwoLetters_NoExactMatch_ReturnsPartial)."""
    import numpy as np
    from oracle.oracle import OracleEngine
    o = OracleEngine(); o.index_texts(["alpha", "beta", "gamma", "delta"], keys=np.arange(1, 5))
    r = o.search("a", 10); assert r["status"] == 0 and len(r["keys"]) >= 3
    o = OracleEngine(); o.index_texts(["table", "chair", "desk", "lamp"], keys=np.arange(1, 5))
    r = o.search("ab", 10); assert r["status"] == 0 and len(r["keys"]) > 0 and r["keys"][0] == 1     # only "table" has both letters


# ---- SearchEngineTests.cs:12-129 and QueryTests.cs:128-215: small corpora through the whole pipeline --------------------------------
def _engine(texts, first_key=1):
    import numpy as np
    from oracle.oracle import OracleEngine
    o = OracleEngine(); o.index_texts(texts, keys=np.arange(first_key, first_key + len(texts)))
    return o


def test_search_engine_tests():
    o = _engine(["The quick brown fox jumps over the lazy dog", "A journey of a thousand miles begins with a single step",
                 "To be or not to be that is the question", "The fox was quick and clever"])
    r = o.search("fox", 10); assert r["keys"] and 1 in r["keys"] and 4 in r["keys"]                     # IndexAndSearch_FindsDocuments
    o = _engine(["hello world", "goodbye world", "hello there"])
    r = o.search("hello world", 10); assert r["keys"][0] == 1 and r["scores"][0] > 200                   # Search_ExactMatch_ReturnsHighScore
    o = _engine(["batman and robin", "superman flies high", "spiderman swings"])
    assert o.search("batmam", 10)["keys"][0] == 1                                                        # Search_FuzzyMatch_FindsSimilar
    assert o.search("", 10)["keys"] == []                                                                # Search_EmptyQuery_ReturnsNoResults
    o = _engine(["hello world", "goodbye world"])
    r = o.search("xyzabc", 10); assert not r["keys"] or r["scores"][0] < 50                              # Search_NoMatches_ReturnsEmptyResults
    o = _engine(["the quick brown fox", "the lazy brown dog", "a quick decision", "quick brown"])
    assert o.search("quick brown", 10)["keys"][0] in (4, 1)                                              # Search_MultiWordQuery_RanksRelevance


def test_query_tests_result_limits():
    o = _engine(["The quick brown fox", "The lazy dog", "Quick thinking"])
    assert o.search("quick", 10)["keys"]                                                                 # SearchEngine_QuerySearch_ReturnsResult
    o = _engine(["batman saves the day"] * 20, first_key=0)
    assert len(o.search("batman", 5)["keys"]) == 5                                                       # ..._LimitsResults_IdenticalDocuments
    o = _engine(["batman saves the day story %d" % i for i in range(20)], first_key=0)
    assert len(o.search("batman", 8)["keys"]) == 8                                                       # ..._LimitsResults_VariedDocuments


def test_word_matcher_tests():
    """WordMatcherTests.cs:9-73 (default WordMatcherSetup sizes: exact 2..8, LD1 3..8, affix >= 3); doc ids are insertion indices."""
    o = _engine(["hello world test", "goodbye world"]); assert o.wm_lookup("world") == [0, 1]           # Lookup_ExactMatch_FindsDocument
    o = _engine(["batman is here"]); assert 0 in o.wm_lookup("batmam")                                   # Lookup_LD1Support_FindsFuzzyMatches
    o = _engine(["batman superman spiderman"]); assert 0 in o.wm_lookup("bat", affix=True)               # LookupAffix_FindsPrefixMatches


def test_tokenizer_tests_through_the_index():
    """TokenizerTests.cs:20-33 (words of >= 3 characters are tokens) with configuration 400 (3-grams after two pad characters,
    Tokenizer.cs:89-139): observable through the term dictionary of an indexed document."""
    o = _engine(["hello world"])
    for term in ("hello", "world", "hel", "ell", "llo", "lo ", "o w", " wo", "wor", "orld"[:3], "rld", "￿￿h", "￿he"):
        assert o.lookup_term(term) >= 0, term
    assert o.lookup_term("￿￿￿") < 0 and o.lookup_term("he") < 0 and o.lookup_term("hello world") < 0


def test_fst_index_tests():
    """FstIndexTests.cs:20-124 -- exact counts of the Myers / trie LD1 match (the "search variant" the LD1 kernel reproduces)."""
    from oracle.oracle import fst_query
    terms = {"apple": 10, "apples": 20, "apply": 30, "bpple": 40}
    n, r = fst_query(terms, "applz"); assert n == 2 and sorted(r) == [10, 30]                       # MatchWithinEditDistance1_FindsMatches
    n, r = fst_query(terms, "apple"); assert n == 4 and sorted(r) == [10, 20, 30, 40]
    n, r = fst_query({"apple": 1, "apply": 2, "bpple": 3}, "apple", cap=1); assert n == 3           # ..._BufferOverflow: total count
    n, r = fst_query({"apple": 1, "apply": 2, "bpple": 3}, "app", cap=1, prefix=True); assert n == 1 and r[0] in (1, 2)   # GetByPrefix_FillsBufferAndStops
    n, r = fst_query({"apple": 1, "apply": 2, "bpple": 3}, "app", cap=5, prefix=True); assert n == 2 and sorted(r) == [1, 2]


def test_bug_reproduction_prefix_preference():
    """BugReproductionTests.cs:12-67 (CoverageEngine + FusionScorer with a fixed word-idf cache, bm25 0.5): "the matrix rev" must
    score "The Matrix Revisited" above "The Matrix Reloaded"."""
    from oracle.oracle import OracleEngine
    o = OracleEngine(); o.set_word_idf({"the": 1.574, "matrix": 9.544, "rev": 9.515})
    reloaded = o.coverage("the matrix rev", "The Matrix Reloaded", 0.0, 0.5); revisited = o.coverage("the matrix rev", "The Matrix Revisited", 0.0, 0.5)
    assert revisited["score"] > reloaded["score"]


# ---- the MovieSearchParityTests whose assertions are invariants over the whole list (not expressible in movie_known_answers.json) -------
def _movie_search(oracle_movies, movie_titles, q, k):
    r = oracle_movies.search(q, k); assert r["status"] == 0
    return [movie_titles[i] for i in r["keys"]], list(r["scores"])


def test_movie_sap_and_de_prefix_at_title_start(oracle_movies, movie_titles):
    t, _ = _movie_search(oracle_movies, movie_titles, "sap", 200); assert t            # Sap_PrefersPrefixAtTitleStart (:381-427)
    seen_other = False
    for title in t:
        low = title.lower(); starts = low.startswith("sap") and (len(low) == 3 or not low[3].isalpha())
        if not starts: seen_other = True
        else: assert not seen_other, title
    t, _ = _movie_search(oracle_movies, movie_titles, "de", 200); assert t              # De_PrefersPrefixAtTitleStart (:510-555)
    seen_other = False
    for title in t:
        if not title.lower().startswith("de"): seen_other = True
        else: assert not seen_other, title


def test_movie_eatrix_f(oracle_movies, movie_titles):                                   # EatrixF_PrefersBeatrixFarrand (:469-508)
    for q in ("eatrix f", "eatrix fe", "eatrix fea", "eatrix fer"):
        t, _ = _movie_search(oracle_movies, movie_titles, q, 10); assert t, q
        if len(q.split()[-1]) >= 3:
            assert "beatrix" in t[0].lower() and "farrand" in t[0].lower(), (q, t[0])


def test_movie_two_f_and_two_fo(oracle_movies, movie_titles):
    import re
    t, _ = _movie_search(oracle_movies, movie_titles, "two f", 10)                      # Search_TwoF_PrefersStrictPrefixMatch (:661-694)
    assert len(t) >= 2 and t[0].lower().startswith("two ") and re.search(r"\bTwo\s+[Ff]", t[0], re.I)
    t, s = _movie_search(oracle_movies, movie_titles, "two fo", 20)                     # Search_TwoFo_AllExactPrefixesBeforePartialMatches (:696-776)
    assert len(t) >= 5
    pre = [x.lower().startswith("two fo") for x in t]
    if False in pre and pre.index(False) > 0:
        i = pre.index(False); assert pre[i - 1] and s[i - 1] > s[i]
    low = [x.lower() for x in t]
    if "tea for two" in low:
        assert any(pre[: low.index("tea for two")])


def test_short_query_two_letters_small_corpora():
    o = _engine(["cat", "dog", "ape"])                                                  # ShortQuery_TwoLetters_ReturnsPartialMatch (:999-1042)
    r = o.search("va", 10); assert r["status"] == 0 and r["keys"] and r["keys"][0] in (1, 3) and all(r["scores"][0] >= x for x in r["scores"])
    o = _engine(["apple", "banana", "cherry", "grape", "orange"])                       # ShortQuery_TwoLetters_MultiplePartialMatches (:1044-1085)
    r = o.search("ra", 10); assert r["status"] == 0 and r["keys"] and {3, 4, 5} & set(r["keys"])
