// ORACLE (test infrastructure, NOT product code): CPU restatement of lofcz/Infidex Stage 2
// (WordMatcher lookup, lexical coverage, fusion signals, fusion score) and the pipeline orchestration.
//
// Follows (under /root/reference/src/Infidex):
//   Scoring/WordMatcherLookup.cs:11-69, WordMatcher/WordMatcher.cs:201-354
//   Coverage/CoverageEngine.cs:61-126,222-427, CoverageTokenizer.cs:7-107
//   Coverage/WholeWordMatcher.cs, JoinedWordMatcher.cs, PrefixSuffixMatcher.cs, FuzzyWordMatcher.cs
//   Coverage/CoverageScorer.cs:5-198, Coverage/FusionSignalComputer.cs:53-568
//   Metrics/LevenshteinDistance.cs:181-341, Metrics/StringMetrics.cs:12-36
//   Scoring/FusionScorer.cs:19-396
//   Scoring/SearchPipeline.cs:49-206,298-576, Scoring/ResultProcessor.cs:146-178
#pragma once
#include "stage1.hpp"
#include "shortquery.hpp"

namespace ifxo {

// ---- LevenshteinDistance.Calculate (early-exit row DP; ignoreCase folds with ToUpperInvariant)
inline int lev(sv pattern, sv text, int max_errors, bool ic) {
    if (pattern.empty()) return (int)text.size();
    if (text.empty()) return (int)pattern.size();
    if (pattern.size() > text.size()) std::swap(pattern, text);
    int m = (int)pattern.size(), n = (int)text.size();
    std::vector<int> costs(m + 1);
    for (int i = 0; i <= m; i++) costs[i] = i;
    for (int j = 0; j < n; j++) {
        char16_t tv = ic ? up(text[j]) : text[j];
        int diag = costs[0]; costs[0] = j + 1; int minc = costs[0];
        for (int i = 0; i < m; i++) {
            int left = costs[i + 1], upv = costs[i];
            char16_t pv = ic ? up(pattern[i]) : pattern[i];
            int cost;
            if (tv == pv) cost = diag;
            else { cost = upv + 1; if (left + 1 < cost) cost = left + 1; if (diag + 1 < cost) cost = diag + 1; }
            diag = left; costs[i + 1] = cost;
            if (cost < minc) minc = cost;
        }
        if (minc > max_errors) return max_errors + 1;
    }
    return costs[m];
}
// LevenshteinDistance.CalculateDamerau (restricted: one transposition at the first mismatch)
inline int damerau(sv s, sv t, int maxd, bool ic) {
    int ld = std::abs((int)s.size() - (int)t.size());
    if (ld > maxd) return maxd + 1;
    int dist = lev(s, t, maxd + 1, ic);
    if (dist <= maxd) return dist;
    if (dist <= maxd + 1) {
        int len = (int)s.size();
        for (int i = 0; i < len - 1; i++) {
            if (i >= (int)t.size()) break;
            char16_t s1 = ic ? lo(s[i]) : s[i], t1 = ic ? lo(t[i]) : t[i];
            if (s1 != t1) {
                if (i + 1 >= (int)t.size()) break;
                char16_t s2 = ic ? lo(s[i + 1]) : s[i + 1], t2 = ic ? lo(t[i + 1]) : t[i + 1];
                if (s1 == t2 && s2 == t1) {
                    int rem = maxd - 1; if (rem < 0) return maxd + 1;
                    sv sr = (i + 2 < len) ? s.substr(i + 2) : sv(); sv tr = (i + 2 < (int)t.size()) ? t.substr(i + 2) : sv();
                    int rd = lev(sr, tr, rem, ic);
                    if (rd <= rem) return 1 + rd;
                }
                break;
            }
        }
    }
    return dist;
}
// StringMetrics.Lcs (containment, else common prefix + tolerance) on lower-cased inputs (SegmentProcessor.CalculateLcs)
inline int lcs_metric(sv q_in, sv r_in, int tol) {
    str q = to_lower(q_in), r = to_lower(r_in);
    if (q.empty() || r.empty()) return 0;
    if (q == r) return (int)q.size();
    if (r.find(q) != str::npos) return (int)q.size();
    int pl = 0, len = (int)std::min(q.size(), r.size());
    for (int i = 0; i < len; i++) { if (q[i] != r[i]) break; pl++; }
    return pl == 0 ? 0 : std::min(pl + tol, len);
}

struct Slice { int off = 0, len = 0, pos = 0; };
inline std::vector<Slice> cov_tokenize(sv text, int min_size, size_t max_tokens) {   // CoverageTokenizer.TokenizeToSpan
    std::vector<Slice> out; size_t i = 0;
    while (i < text.size()) {
        while (i < text.size() && is_delim(text[i])) i++;
        if (i >= text.size()) break;
        size_t b = i; while (i < text.size() && !is_delim(text[i])) i++;
        if ((int)(i - b) >= min_size && out.size() < max_tokens) out.push_back({(int)b, (int)(i - b), (int)b});
    }
    return out;
}
inline std::vector<Slice> cov_dedupe(const std::vector<Slice>& in, sv text) {
    std::vector<Slice> out;
    for (auto& c : in) { bool dup = false; for (auto& e : out) if (e.len == c.len && eq_ic(text.substr(e.off, e.len), text.substr(c.off, c.len))) { dup = true; break; } if (!dup) out.push_back(c); }
    return out;
}

struct CoverageSetup {   // Coverage/CoverageSetup.cs defaults
    int min_word_size = 2, lev_max_word = 20, num_typos = 2, min_len_one = 3, min_len_two = 7;
    int min_hits_abs = 1, min_hits_rel = 0, qlimit_tol = 5; double lcs_tol_rel = 0.2;
    bool whole_query = true, whole = true, fuzzy = true, joined = true, prefix_suffix = true, truncate = true;
    uint8_t truncation_score = 254;
};

struct QueryCtx {   // CoverageQueryContext
    str query; std::vector<Slice> tok; std::vector<float> term_idf; std::vector<float> word_idf; bool has_word_idf = false;
};

struct FusionSignals { int unfiltered_q = 0; bool lex_prefix_last = false, all_prec_exact = false, perfect_doc = false, stem_evidence = false, anchor_stem = false; uint8_t trailing_density = 0, single_sim = 0; int single_char_boost = 0; };

struct Features {
    uint8_t coverage_score = 0; int terms_count = 0, any_match = 0, fully = 0, strict = 0, prefix_matched = 0, first_match = -1;
    float sum_ci = 0; int word_hits = 0, doc_tokens = 0, longest_prefix_run = 0, suffix_prefix_run = 0, phrase_span = 0, preceding_strict = 0;
    bool last_has_prefix = false; float last_ci = 0, weighted_cov = 0; bool last_typeahead = false;
    float idf_cov = 0, total_idf = 0, missing_idf = 0; std::vector<float> term_idf, term_ci; bool has_arrays = false;
    FusionSignals fs;
};

struct Coverage {
    const Index& ix; CoverageSetup setup;
    explicit Coverage(const Index& i) : ix(i) {}

    float compute_term_idf(sv term) const {   // CoverageEngine.ComputeTermIdf
        if (ix.live_count == 0) return std::log2((float)(term.size() + 1));
        float sum = 0.f; int cnt = 0;
        if ((int)term.size() >= NGRAM) for (size_t i = 0; i + NGRAM <= term.size(); i++) {
            auto it = ix.term_ids.find(str(term.substr(i, NGRAM)));
            if (it != ix.term_ids.end() && ix.terms[it->second].df > 0) { sum += compute_idf(ix.live_count, ix.terms[it->second].df); cnt++; }
        }
        return cnt > 0 ? sum / cnt : std::log2((float)(term.size() + 1));
    }
    QueryCtx prepare(sv query) const {   // PrepareQuery
        QueryCtx c; c.query = str(query);
        if (query.empty()) return c;
        auto raw = cov_tokenize(query, setup.min_word_size, query.size() / 2 + 1);
        if (raw.empty()) return c;
        c.tok = cov_dedupe(raw, query);
        for (auto& t : c.tok) c.term_idf.push_back(compute_term_idf(query.substr(t.off, t.len)));
        c.has_word_idf = true;
        for (auto& t : c.tok) { auto it = ix.word_idf.find(to_lower(query.substr(t.off, t.len))); c.word_idf.push_back(it != ix.word_idf.end() ? it->second : 0.f); }
        return c;
    }

    struct State {
        sv q, d; const std::vector<Slice>* qt; std::vector<Slice> dt;
        std::vector<char> qa, da; std::vector<float> matched; std::vector<int> maxc; std::vector<char> has_whole, has_joined, has_prefix; std::vector<int> first_pos;
        int word_hits = 0; double num_whole = 0, num_joined = 0, num_fuzzy = 0, num_ps = 0; uint8_t penalty = 0;
        sv qtext(int i) const { return q.substr((*qt)[i].off, (*qt)[i].len); }
        sv dtext(int j) const { return d.substr(dt[j].off, dt[j].len); }
        void pos_min(int i, int pos) { if (first_pos[i] == -1 || pos < first_pos[i]) first_pos[i] = pos; }
    };

    static void match_whole(State& s) {
        int qc = (int)s.qt->size(), dc = (int)s.dt.size(); int pinc = qc > 1 ? 1 : 0;
        for (int i = 0; i < qc; i++) {
            int mi = -1;
            for (int j = 0; j < dc; j++) if (s.da[j] && s.dt[j].len == (*s.qt)[i].len && eq_ic(s.qtext(i), s.dtext(j))) { mi = j; break; }
            if (mi == -1) continue;
            int ql = (*s.qt)[i].len;
            s.word_hits++; s.num_whole += ql; s.matched[i] += (float)ql; s.has_whole[i] = 1; s.has_prefix[i] = 1;
            s.pos_min(i, s.dt[mi].pos);
            if (dc > i) { if (s.dt[i].len != ql || !eq_ic(s.qtext(i), s.dtext(i))) s.penalty++; } else s.penalty++;
            if (i < qc - 1) s.num_whole += pinc;
            s.qa[i] = 0; s.da[mi] = 0;
        }
    }
    static void match_joined(State& s) {
        int qc = (int)s.qt->size(), dc = (int)s.dt.size();
        for (int i = 0; i < qc - 1; i++) {
            if (!s.qa[i] || !s.qa[i + 1]) continue;
            int nx = -1; for (int k = i + 1; k < qc; k++) if (s.qa[k]) { nx = k; break; }
            if (nx == -1) break;
            int jl = (*s.qt)[i].len + (*s.qt)[nx].len; int mi = -1;
            for (int j = 0; j < dc; j++) if (s.da[j] && s.dt[j].len == jl && starts_ic(s.dtext(j), s.qtext(i)) && ends_ic(s.dtext(j), s.qtext(nx))) { mi = j; break; }
            if (mi == -1) continue;
            s.num_joined += jl; s.word_hits += 2;
            s.matched[i] += (float)(*s.qt)[i].len; s.has_joined[i] = 1; s.has_prefix[i] = 1; int pos = s.dt[mi].pos; s.pos_min(i, pos);
            s.matched[nx] += (float)(*s.qt)[nx].len; s.has_joined[nx] = 1; s.pos_min(nx, pos);
            s.qa[i] = 0; s.qa[nx] = 0; s.da[mi] = 0;
        }
        for (int i = 0; i < dc - 1; i++) {
            if (!s.da[i]) continue;
            int nx = -1; for (int k = i + 1; k < dc; k++) if (s.da[k]) { nx = k; break; }
            if (nx == -1) break;
            int jl = s.dt[i].len + s.dt[nx].len; int mi = -1;
            for (int j = 0; j < qc; j++) if (s.qa[j] && (*s.qt)[j].len == jl && starts_ic(s.qtext(j), s.dtext(i)) && ends_ic(s.qtext(j), s.dtext(nx))) { mi = j; break; }
            if (mi == -1) continue;
            s.num_joined += jl; s.word_hits += 1;
            s.matched[mi] += (float)jl; s.has_joined[mi] = 1; s.has_prefix[mi] = 1; s.pos_min(mi, s.dt[i].pos);
            s.qa[mi] = 0; s.da[i] = 0; s.da[nx] = 0;
        }
    }
    static void sort_len_desc(std::vector<int>& idx, const std::vector<Slice>& toks) {   // stable insertion sort
        for (size_t i = 1; i < idx.size(); i++) { int c = idx[i], cl = toks[c].len; int j = (int)i - 1; while (j >= 0 && toks[idx[j]].len < cl) { idx[j + 1] = idx[j]; j--; } idx[j + 1] = c; }
    }
    static void match_prefix_suffix(State& s) {
        int qc = (int)s.qt->size(), dc = (int)s.dt.size();
        std::vector<int> qi, di;
        for (int i = 0; i < qc; i++) if (s.qa[i]) qi.push_back(i);
        for (int j = 0; j < dc; j++) if (s.da[j]) di.push_back(j);
        sort_len_desc(qi, *s.qt); sort_len_desc(di, s.dt);
        auto hit = [&](int i, int j, double sc, bool is_prefix) {
            s.num_ps += sc; s.word_hits++; s.matched[i] += (float)sc; if (is_prefix) s.has_prefix[i] = 1;
            s.pos_min(i, s.dt[j].pos); s.qa[i] = 0; s.da[j] = 0;
        };
        for (int i : qi) {   // MatchExact
            if (!s.qa[i]) continue;
            int ql = (*s.qt)[i].len; sv qt = s.qtext(i);
            for (int j : di) {
                if (!s.da[j]) continue;
                int dl = s.dt[j].len; if (ql == dl) continue;
                sv dt = s.dtext(j); bool m = false, pre = false; double sc = 0;
                if (ql < dl) {
                    if (starts_ic(dt, qt)) { sc = ql; m = true; pre = true; }
                    else if (ends_ic(dt, qt)) { sc = std::max(1, ql / 2); m = true; }
                    else if (ql >= 4 && contains_ic(dt, qt)) { sc = ql * 0.6; m = true; }
                } else if (ends_ic(qt, dt)) { sc = dl; m = true; }
                if (m) { hit(i, j, sc, pre); break; }
            }
        }
        for (int i : qi) {   // MatchFuzzyPrefix
            if (!s.qa[i]) continue;
            int ql = (*s.qt)[i].len; sv qt = s.qtext(i);
            if (!(ql >= 4 || (i == qc - 1 && ql >= 2))) continue;
            for (int j : di) {
                if (!s.da[j]) continue;
                int dl = s.dt[j].len; if (ql >= dl) continue;
                sv dt = s.dtext(j); bool m = false; double sc = 0;
                int dist = damerau(qt, dt.substr(0, ql), 1, true);
                if (dist <= 1) { sc = ql - dist; if (sc < 0.1) sc = 0.1; m = true; }
                else if (dl > ql) {
                    dist = damerau(qt, dt.substr(0, ql + 1), 1, true);
                    if (dist <= 1) { sc = ql - dist; if (sc < 0.1) sc = 0.1; m = true; }
                    else if (ql > 1) { dist = damerau(qt, dt.substr(0, ql - 1), 1, true); if (dist <= 1) { sc = ql - 1 - dist; if (sc < 0.1) sc = 0.1; m = true; } }
                }
                if (m) { hit(i, j, sc, false); break; }
            }
        }
    }
    void match_fuzzy(State& s) const {
        int qc = (int)s.qt->size(), dc = (int)s.dt.size();
        int maxq = 0; for (int i = 0; i < qc; i++) if (s.qa[i] && (*s.qt)[i].len > maxq) maxq = (*s.qt)[i].len;
        if (maxq == 0) return;
        int maxe = maxq >= setup.min_len_two ? 2 : (maxq >= setup.min_len_one ? 1 : 0);
        if (maxq == 2 && maxe == 0 && setup.num_typos >= 1) maxe = 1;
        if (maxe > setup.num_typos) maxe = setup.num_typos;
        if (maxe == 0) return;
        for (int e = 1; e <= maxe; e++) {
            bool any = false; for (int i = 0; i < qc; i++) if (s.qa[i]) any = true;
            if (!any) break;
            for (int i = 0; i < qc; i++) {
                if (!s.qa[i]) continue;
                int ql = (*s.qt)[i].len; if (ql < setup.min_word_size) continue;
                int tme = ql >= setup.min_len_two ? 2 : (ql >= setup.min_len_one ? 1 : 0); bool special = false;
                if (ql == 2 && tme == 0 && setup.num_typos >= 1) { tme = 1; special = true; }
                if (tme > setup.num_typos) tme = setup.num_typos;
                if (e > tme) continue;
                if (special && e != 1) continue;
                int minl = std::max(setup.min_word_size, ql - e), maxl = std::min(setup.lev_max_word, ql + e); if (maxl > 63) maxl = 63;
                sv qt = s.qtext(i);
                for (int j = 0; j < dc; j++) {
                    if (!s.da[j]) continue;
                    int dl = s.dt[j].len; if (dl > maxl || dl < minl) continue;
                    sv dt = s.dtext(j);
                    if (special && (dt.empty() || lo(dt[0]) != lo(qt[0]))) continue;
                    int dist = damerau(qt, dt, e, true);
                    if (dist <= e) { s.word_hits++; s.num_fuzzy += (ql - dist); s.matched[i] += (float)(ql - dist); s.pos_min(i, s.dt[j].pos); s.qa[i] = 0; s.da[j] = 0; break; }
                }
            }
        }
    }

    // ---- FusionSignalComputer
    static float single_term_sim(sv query, sv doc, const std::vector<Slice>& dt) {
        int ql = (int)query.size(); if (ql < 3) return 0.f;
        str qlow = to_lower(query); float best = 0.f;
        for (auto& t : dt) {
            if (t.len < 2) continue;
            str tl = to_lower(doc.substr(t.off, t.len));
            size_t idx = qlow.find(tl);
            if (idx != str::npos) { float lf = (float)tl.size() / ql; float pf = 1.f - (float)idx / ql; float sc = lf * pf; if (sc > best) best = sc; continue; }
            int maxk = std::min(ql, (int)tl.size()), bestk = 0;
            for (int len = maxk; len >= 2; len--) if (sv(qlow).substr(ql - len) == sv(tl).substr(0, len)) { bestk = len; break; }
            float ps = bestk > 0 ? (float)bestk / ql : 0.f, fz = 0.f;
            if (tl.size() <= 32) { int dist = damerau(qlow, tl, 2, false); if (dist <= 2) fz = (float)(ql - dist) / ql; }
            float comb = std::max(ps, fz); if (comb > best) best = comb;
        }
        if (ql >= 6) {
            int seg = std::min(6, ql / 2); sv pf = sv(qlow).substr(0, seg), sf = sv(qlow).substr(ql - seg, seg); int pi = -1, si = -1;
            for (int i = 0; i < (int)dt.size(); i++) {
                if (dt[i].len < 3) continue;
                str tl = to_lower(doc.substr(dt[i].off, dt[i].len)); sv t(tl);
                auto sw = [](sv a, sv b) { return a.size() >= b.size() && a.substr(0, b.size()) == b; };
                auto ew = [](sv a, sv b) { return a.size() >= b.size() && a.substr(a.size() - b.size()) == b; };
                if (pi == -1 && (sw(t, pf) || sw(pf, t))) pi = i;
                if (si == -1 && (ew(t, sf) || ew(sf, t))) si = i;
                if (pi != -1 && si != -1) break;
            }
            if (pi != -1 && si != -1 && pi != si) { float ts = std::min(1.f, (float)(pf.size() + sf.size()) / (float)ql); if (ts > best) best = ts; }
        }
        return best;
    }
    static int single_char_last(sv q, sv d, const std::vector<Slice>& qt, const std::vector<Slice>& dt) {
        int qc = (int)qt.size(), dc = (int)dt.size();
        if (qt[qc - 1].len != 1) return 0;
        char16_t target = lo(q[qt[qc - 1].off]); if (!is_letter(target)) return 0;
        int di = 0, first = -1;
        for (int i = 0; i < qc - 1; i++) {
            sv qtm = q.substr(qt[i].off, qt[i].len); bool found = false;
            while (di < dc) { if (index_of_ic(d.substr(dt[di].off, dt[di].len), qtm) >= 0) { found = true; if (first == -1) first = di; break; } di++; }
            if (!found) return 0;
        }
        if (di + 1 < dc) {
            const Slice& nx = dt[di + 1];
            if (nx.len > 0 && lo(d[nx.off]) == target) {
                int end = dt[di].off + dt[di].len; bool broken = false;
                for (int p = end; p < nx.off; p++) if (!is_space(d[p])) { broken = true; break; }
                if (!broken) { int boost = 8 + std::max(0, 16 - first); if (nx.len == 1) boost += 4; return boost; }
            }
        }
        return 0;
    }
    FusionSignals signals(sv q, sv d, const std::vector<Slice>& qt, const std::vector<Slice>& dt, int min_stem, int doc_id) const {
        FusionSignals f; int qc = (int)qt.size(), dc = (int)dt.size(); f.unfiltered_q = qc;
        if (qc == 0 || dc == 0) return f;
        auto Q = [&](int i) { return q.substr(qt[i].off, qt[i].len); }; auto D = [&](int j) { return d.substr(dt[j].off, dt[j].len); };
        // CheckPrefixLastMatch
        if (qc == 1) { for (int i = 0; i < dc; i++) if (starts_ic(D(i), Q(0))) { f.lex_prefix_last = true; f.all_prec_exact = eq_ic(D(i), Q(0)); break; } }
        else {
            bool all = true;
            for (int i = 0; i < qc - 1; i++) { if (Q(i).empty()) continue; bool fe = false; for (int j = 0; j < dc; j++) if (eq_ic(D(j), Q(i))) { fe = true; break; } if (!fe) { all = false; break; } }
            if (all) { sv last = Q(qc - 1); if (last.empty()) { f.lex_prefix_last = true; f.all_prec_exact = true; } else for (int i = 0; i < dc; i++) if (starts_ic(D(i), last)) { f.lex_prefix_last = true; f.all_prec_exact = true; break; } }
        }
        // ComputePerfectDoc
        { bool ok = true; for (int j = 0; j < dc && ok; j++) { bool ex = false; for (int i = 0; i < qc; i++) if (starts_ic(D(j), Q(i)) || starts_ic(Q(i), D(j))) { ex = true; break; } if (!ex) ok = false; } f.perfect_doc = ok; }
        // CheckStemEvidence
        if (qc >= 2) {
            int unmatched = 0, evidence = 0;
            for (int qi = 0; qi < qc; qi++) {
                sv qq = Q(qi); if ((int)qq.size() < min_stem) continue;
                bool wm = false; for (int j = 0; j < dc; j++) { sv dd = D(j); if (dd.empty()) continue; if (eq_ic(dd, qq) || starts_ic(dd, qq)) { wm = true; break; } }
                if (wm) continue;
                unmatched++;
                for (int j = 0; j < dc; j++) {
                    sv dd = D(j); if ((int)dd.size() < min_stem) continue;
                    if (starts_ic(qq, dd)) { evidence++; break; }
                    int mc = (int)std::min(qq.size(), dd.size());
                    if (mc >= min_stem) { int pl = 0; for (int i = 0; i < mc; i++) { if (lo(qq[i]) == lo(dd[i])) pl++; else break; } if (pl >= min_stem) { evidence++; break; } }
                }
            }
            f.stem_evidence = unmatched > 0 && evidence == unmatched;
        }
        // HasAnchorStem (uses DocumentMetadata first token)
        if (qt[0].len >= 3) {
            sv stem = Q(0).substr(0, 3);
            bool has_tokens = doc_id >= 0 && doc_id < (int)ix.token_count.size() && ix.token_count[doc_id] > 0;
            if (has_tokens && ix.first_token[doc_id].size() >= stem.size()) {
                if (starts_ic(ix.first_token[doc_id], stem)) f.anchor_stem = true;
                else for (int i = 1; i < dc; i++) if (D(i).size() >= stem.size() && starts_ic(D(i), stem)) { f.anchor_stem = true; break; }
            } else if (!has_tokens) { for (int i = 0; i < dc; i++) if (D(i).size() >= stem.size() && starts_ic(D(i), stem)) { f.anchor_stem = true; break; } }
        }
        // TrailingMatchDensity
        if (qc >= 2 && qt[qc - 1].len >= 1 && qt[qc - 1].len <= 2) {
            sv last = Q(qc - 1); int cnt = 0;
            for (int i = 0; i < dc; i++) if (starts_ic(D(i), last) || (D(i).size() > last.size() && contains_ic(D(i), last))) cnt++;
            if (cnt > 0) { float dens = (float)cnt / dc; float v = dens * 255.f; v = v < 0.f ? 0.f : (v > 255.f ? 255.f : v); f.trailing_density = (uint8_t)v; }
        }
        if (qc == 1) { float sim = single_term_sim(Q(0), d, dt); float v = sim * 255.f; v = v < 0.f ? 0.f : (v > 255.f ? 255.f : v); f.single_sim = (uint8_t)v; }
        if (qc >= 2) f.single_char_boost = single_char_last(q, d, qt, dt);
        return f;
    }

    // CoverageEngine.CalculateCoverageInternal + CoverageScorer.CalculateFinalScore
    Features features(const QueryCtx& c, sv doc, double lcs_sum, int doc_id) const {
        Features f; int qc = (int)c.tok.size();
        if (qc == 0) { f.first_match = -1; return f; }
        sv q(c.query);
        State s; s.q = q; s.d = doc; s.qt = &c.tok;
        auto raw = cov_tokenize(doc, setup.min_word_size, doc.size() / 2 + 1);
        f.doc_tokens = (int)raw.size();
        s.dt = cov_dedupe(raw, doc); int dc = (int)s.dt.size();
        s.qa.assign(qc, 1); s.da.assign(dc, 1); s.matched.assign(qc, 0.f); s.maxc.resize(qc); for (int i = 0; i < qc; i++) s.maxc[i] = c.tok[i].len;
        s.has_whole.assign(qc, 0); s.has_joined.assign(qc, 0); s.has_prefix.assign(qc, 0); s.first_pos.assign(qc, -1);
        if (setup.whole) match_whole(s);
        if (setup.joined) match_joined(s);
        if (setup.prefix_suffix) match_prefix_suffix(s);
        if (setup.fuzzy) { bool all = true; for (int i = 0; i < qc; i++) if (s.maxc[i] > 0 && s.matched[i] < (float)s.maxc[i]) { all = false; break; } if (!all) match_fuzzy(s); }
        f.word_hits = s.word_hits;
        // CalculateFinalScore
        int qlen = (int)c.query.size();
        if (!setup.whole_query) lcs_sum = 0.0;
        double num = s.num_joined + s.num_whole + s.num_fuzzy + s.num_ps - (double)s.penalty;
        if (num == 0.0 && lcs_sum > 2.0) num = lcs_sum - 2.0;
        f.coverage_score = (uint8_t)(int)std::min(num / qlen * 255.0, 255.0);
        float sum_ci = 0, wsum = 0, tw = 0, idfw = 0, tidf = 0, midf = 0, last_ci = 0, last_idf = 0; int first = -1, minp = INT32_MAX, maxp = -1;
        if (c.has_word_idf) { f.term_ci.assign(qc, 0.f); f.term_idf = c.word_idf; f.has_arrays = true; }
        for (int i = 0; i < qc; i++) {
            if (s.maxc[i] <= 0) continue;
            float ci = std::min(1.0f, s.matched[i] / (float)s.maxc[i]);
            sum_ci += ci; if (f.has_arrays) f.term_ci[i] = ci;
            if (ci > 0) f.any_match++;
            float twt = (float)s.maxc[i]; tw += twt; wsum += ci * twt;
            float idf = c.term_idf[i]; tidf += idf; idfw += ci * idf;
            if (ci < 1.0f) midf += (1.0f - ci) * idf;
            if (i == qc - 1) { last_ci = ci; last_idf = idf; }
            bool full = s.matched[i] >= ((float)s.maxc[i] - 0.01f);
            if (full) f.fully++;
            if ((s.has_whole[i] || s.has_joined[i]) && full) f.strict++;
            if (s.has_prefix[i]) f.prefix_matched++;
            if (s.first_pos[i] >= 0) { if (first == -1 || s.first_pos[i] < first) first = s.first_pos[i]; minp = std::min(minp, s.first_pos[i]); maxp = std::max(maxp, s.first_pos[i]); }
        }
        f.weighted_cov = tw > 0.f ? wsum / tw : 0.f;
        f.idf_cov = tidf > 0.f ? idfw / tidf : 0.f;
        if (qc > 0 && tidf > 0.f) { float share = last_idf / tidf; float th = 1.f / (float)(qc + 1); f.last_typeahead = share <= th; }
        if (qc == 1 && qlen > 0 && lcs_sum > 0.0) { float cl = (float)std::min(1.0, lcs_sum / qlen); if (cl > sum_ci) sum_ci = cl; }
        int run = 0;
        for (int i = 0; i < qc; i++) { bool ph = s.has_prefix[i] && s.maxc[i] > 0 && s.matched[i] > 0; if (ph) { run++; if (run > f.longest_prefix_run) f.longest_prefix_run = run; } else run = 0; }
        int srun = 0; for (int i = qc - 1; i >= 0; i--) { bool ph = s.has_prefix[i] && s.maxc[i] > 0 && s.matched[i] > 0; if (ph) srun++; else break; }
        f.suffix_prefix_run = srun;
        if (minp != INT32_MAX && maxp >= minp && f.any_match >= 2) f.phrase_span = maxp - minp + 1;
        f.last_has_prefix = s.has_prefix[qc - 1] && s.matched[qc - 1] > 0;
        if (qc >= 2) for (int i = 0; i < qc - 1; i++) if ((s.has_whole[i] || s.has_joined[i]) && s.matched[i] >= ((float)s.maxc[i] - 0.01f)) f.preceding_strict++;
        f.terms_count = qc; f.first_match = first; f.sum_ci = sum_ci; f.last_ci = last_ci; f.total_idf = tidf; f.missing_idf = midf;
        // fusion signals: re-tokenise with minWordSize 0, no dedupe; minStemLength = setup.MinWordSize (CoverageEngine.cs:371)
        auto fq = cov_tokenize(q, 0, q.size() / 2 + 1); auto fd = cov_tokenize(doc, 0, doc.size() / 2 + 1);
        f.fs = signals(q, doc, fq, fd, setup.min_word_size, doc_id);
        return f;
    }
};

// FusionScorer.Calculate
inline std::pair<float, uint8_t> fusion_score(sv query, sv doc, const Features& f, float bm25) {
    const FusionSignals& fs = f.fs;
    int n = fs.unfiltered_q > 0 ? fs.unfiltered_q : f.terms_count;
    bool single = n <= 1;
    bool complete = f.terms_count > 0 && f.any_match == f.terms_count;
    bool clean = f.terms_count > 0 && f.prefix_matched == f.terms_count;
    bool exact = f.terms_count > 0 && f.strict == f.terms_count;
    bool at_start = f.first_match == 0;
    bool lpl = fs.lex_prefix_last;
    int preceding = std::max(0, f.terms_count - 1);
    bool cpl = f.terms_count >= 1 && f.preceding_strict == preceding && f.last_has_prefix;
    bool strong = lpl && cpl; bool perfect = fs.perfect_doc;
    int prec = 0, tier = 0;
    if (!single && f.terms_count > 0) { int m = f.any_match, t = f.terms_count; tier = m >= t ? 3 : (m == t - 1 ? 2 : (m * 2 >= t ? 1 : 0)); }
    if (!single && tier > 0) prec |= (tier & 3) << 16;
    if (!single && clean && at_start && lpl && complete) prec |= 1 << 15;
    if (!single && f.doc_tokens > 0 && f.word_hits == f.doc_tokens) prec |= 1 << 14;
    float avg_idf = 0.f;
    if (!single && f.terms_count >= 2) {
        bool dominant = false;
        if (f.has_arrays && (int)f.term_idf.size() == f.terms_count && (int)f.term_ci.size() == f.terms_count) {
            avg_idf = f.total_idf > 0.f && f.terms_count > 0 ? f.total_idf / (float)f.terms_count : 0.f;
            for (int c = 0; c < f.terms_count; c++) {
                float power = f.term_idf[c] * f.term_ci[c];
                if (f.term_ci[c] <= 0.1f || f.term_idf[c] <= 0.f || f.term_idf[c] < avg_idf) continue;
                float other = 0.f; for (int i = 0; i < f.terms_count; i++) if (i != c) other += f.term_idf[i] * f.term_ci[i];
                if (power >= other) { dominant = true; break; }
            }
        }
        bool anchor = fs.anchor_stem && f.has_arrays && f.term_idf.size() >= 1 && f.term_idf[0] >= avg_idf;
        if (dominant || anchor) prec |= 1 << 13;
        int unmatched = f.terms_count - f.any_match;
        if (dominant && unmatched == 1) prec |= 8;
    }
    if (single) {
        if (complete) prec |= 1 << 17;
        if (clean && f.terms_count > 0) prec |= 1 << 16;
        int t = 0; if (complete) { if (at_start) { if (exact) t = 4; else if (clean) t = 3; } else { if (exact) t = 2; else if (clean) t = 1; } }
        prec |= t << 3;
    } else {
        bool anchor_run = fs.anchor_stem && f.longest_prefix_run >= 2;
        int mt = strong ? 3 : (lpl ? 2 : ((perfect || anchor_run) ? 1 : 0));
        if (fs.unfiltered_q > f.terms_count) mt += fs.single_char_boost;
        prec |= mt;
    }
    float ratio = f.terms_count > 0 ? (float)f.any_match / (float)f.terms_count : 0.f;
    bool partial = ratio > 0.f && ratio < 1.f;
    if (partial && n >= 2) {
        if (fs.stem_evidence) prec |= 8;
        else {
            int unmatched = f.terms_count - f.any_match;
            bool last_matched = f.last_has_prefix || (f.terms_count > 0 && f.any_match == f.terms_count);
            bool can = (last_matched || !f.last_typeahead) && f.total_idf > 0.f;
            if (unmatched == 1 && can) { float mr = f.missing_idf / f.total_idf; float gap = 1.f - ratio; if (mr < gap) prec |= 8; }
        }
    }
    // ComputeSemanticScore
    float avg_ci = f.terms_count > 0 ? f.sum_ci / (float)f.terms_count : 0.f, sem;
    if (single) { float ls = (float)fs.single_sim / 255.f; sem = (avg_ci + ls) / 2.f; }
    else if (f.doc_tokens == 0) sem = avg_ci;
    else {
        int unmatched = f.terms_count - f.any_match;
        bool last_matched = f.last_has_prefix || (f.terms_count > 0 && f.any_match == f.terms_count);
        bool can = (last_matched || !f.last_typeahead) && f.total_idf > 0.f;
        bool use_idf = partial && unmatched == 1 && can && f.idf_cov > ratio;
        float base = use_idf ? f.idf_cov : avg_ci;
        float density = (float)f.word_hits / (float)f.doc_tokens;
        sem = base * density;
        if (f.terms_count >= 3) { int sc = (fs.anchor_stem ? 1 : 0) + (f.suffix_prefix_run >= 2 ? 1 : 0); if (sc > 0) { float bonus = 0.15f * (float)sc; sem = std::min(1.f, sem + bonus); } }
        if (f.terms_count >= 2) { float md = (float)fs.trailing_density / 255.f; if (md > 0.f) { float head = 1.f - sem; sem += head * md; } }
    }
    float gap = 1.f - ratio;
    if (partial && bm25 >= gap) sem = ratio * sem + gap * bm25;
    sem = sem < 0.f ? 0.f : (sem > 0.999f ? 0.999f : sem);
    uint8_t tie = 0;
    if (n >= 2 && !doc.empty()) { float focus = std::min(1.f, (float)query.size() / (float)doc.size()); tie = (uint8_t)(focus * 255.f); }
    return {(float)prec + sem, tie};
}

struct SearchOut { std::vector<ScoreEntry> records; int stage = 0; bool unsupported = false; bool short_path = false; };

struct Pipeline {
    const Index& ix; Stage1 s1; Coverage cov; ShortQuery sq;
    explicit Pipeline(const Index& i) : ix(i), s1(i), cov(i), sq(i) {}

    static void union_into(std::vector<int>& acc, const std::vector<int>& b) { std::vector<int> r; r.reserve(acc.size() + b.size()); std::set_union(acc.begin(), acc.end(), b.begin(), b.end(), std::back_inserter(r)); acc.swap(r); }

    std::vector<int> wm_lookup(sv word) const {   // WordMatcher.Lookup
        std::vector<int> r; str nrm = normalize(to_lower(word)); int len = (int)nrm.size();
        auto acc = [&](const StrMap<std::vector<int>>& m, const str& k) { auto it = m.find(k); if (it != m.end()) union_into(r, it->second); };
        acc(ix.wm_exact, nrm);
        if (len >= 3 && len <= 8) { acc(ix.wm_ld1, nrm); for (int i = 0; i < len; i++) { str d(nrm); d.erase(i, 1); acc(ix.wm_ld1, d); acc(ix.wm_exact, d); } }
        return r;
    }
    std::vector<int> wm_affix(sv word) const {    // WordMatcher.LookupAffix
        std::vector<int> r; str nrm = normalize(to_lower(word));
        if (ix.wm_fwd.empty()) return r;
        int pn = nrm.empty() ? -1 : ix.wm_fwd.walk(nrm); str rev(nrm.rbegin(), nrm.rend()); int sn = nrm.empty() ? -1 : ix.wm_rev.walk(rev);
        int pc = ix.wm_fwd.count_outputs(pn), sc = ix.wm_rev.count_outputs(sn), budget = 4096;
        if (pc == 0 && sc == 0) return r;
        std::vector<int> docs;
        if (pc > 0 && budget > 0) { int w = ix.wm_fwd.collect(pn, std::min(pc, budget), docs); budget -= w; }
        if (sc > 0 && budget > 0) { ix.wm_rev.collect(sn, std::min(sc, budget), docs); }
        std::sort(docs.begin(), docs.end()); docs.erase(std::unique(docs.begin(), docs.end()), docs.end());
        return docs;
    }
    std::vector<int> wm_execute(sv query) const { // WordMatcherLookup.Execute
        std::vector<int> r;
        for (sv w : split_words(query)) {
            if (is_blank(w) || w.size() < 2) continue;
            union_into(r, wm_lookup(w));
            if (cov.setup.prefix_suffix) union_into(r, wm_affix(w));
        }
        return r;
    }

    // SearchPipeline.Execute
    SearchOut execute(sv search_in, bool enable_coverage, int depth, int max_results, std::vector<ScoreEntry>* stage1_out = nullptr, Stage1Stats* st = nullptr) const {
        SearchOut out;
        if (is_blank(search_in)) return out;
        str search = normalize(search_in);
        QueryAnalysis qa = analyze_query(search);
        std::vector<ScoreEntry> stage1;
        if (!qa.can_use_ngrams && ix.from_image) { out.unsupported = true; return out; }   // an image carries no token positions (champion lists)
        if (!qa.can_use_ngrams) { out.short_path = true; stage1 = sq.stage1(search, max_results); }   // no word of >= 3 chars: SURVEY 8(f)-1, oracle/shortquery.hpp
        else { str tfidf_q = qa.mixed ? qa.long_words : search; if (is_blank(tfidf_q)) tfidf_q = search; stage1 = s1.search(tfidf_q, depth, st); }
        if (stage1_out) *stage1_out = stage1;
        bool short_q = !search.empty() && search.size() <= 3; for (char16_t c : search) if (is_delim(c)) short_q = false;
        if (short_q && (long long)stage1.size() >= max_results && max_results < INT32_MAX) { stage1.resize(max_results); out.records = stage1; out.stage = 1; return out; }
        int sq_count = 0; bool sq_known = false;
        if (short_q) { auto it = ix.prefix_docs.find(search); sq_count = it == ix.prefix_docs.end() ? 0 : (int)it->second.size(); sq_known = true; }
        bool skip_cap = short_q && sq_known && sq_count > 500;
        bool allow_short_cov = short_q && sq_known && sq_count > 0 && sq_count <= 500;      // SearchPipeline.cs:133-137
        if (!enable_coverage || skip_cap || (!qa.can_use_ngrams && !allow_short_cov)) { out.records = stage1; out.stage = 1; return out; }
        std::vector<ScoreEntry> covr = coverage_stage(search, depth, max_results, stage1);
        if (covr.empty() && !stage1.empty()) { out.records = stage1; out.stage = 1; return out; }
        out.records = covr; out.stage = 2; return out;
    }

    std::vector<ScoreEntry> coverage_stage(const str& search, int depth, int max_results, std::vector<ScoreEntry> top) const {
        const CoverageSetup& cs = cov.setup;
        if ((int)top.size() > depth) top.resize(depth);
        std::vector<int> wm = wm_execute(search);
        // BuildDocumentKeyIndex: keys of top in rank order, then live WM docs ascending
        std::unordered_map<long long, int> key_idx; int next = 0;
        for (auto& e : top) if (!key_idx.count(e.key)) key_idx[e.key] = next++;
        for (int id : wm) if (!ix.docs[id].deleted && !key_idx.count(ix.docs[id].key)) key_idx[ix.docs[id].key] = next++;
        std::vector<uint8_t> lcs_row(2, 0), hits_row(2, 0);   // Span2D height-2 quirk (Q3): only docIndex < 2 are stored
        TopKHeap final_scores(depth); int max_hits = 0;
        QueryCtx ctx = cov.prepare(search);
        std::vector<int> top_ids; for (auto& e : top) { int id = ix.doc_by_key(e.key); if (id >= 0) top_ids.push_back(id); }
        std::sort(top_ids.begin(), top_ids.end());
        std::vector<int> overlap, uniq; for (int id : wm) (std::binary_search(top_ids.begin(), top_ids.end(), id) ? overlap : uniq).push_back(id);
        int wm_limit = std::max(0, depth - (int)overlap.size());
        auto process = [&](int id, float base) {
            const Doc& d = ix.docs[id]; if (d.deleted) return;
            auto it = key_idx.find(d.key); if (it == key_idx.end()) return; int di = it->second;
            str text = normalize(d.indexed_text);
            int lcs = 0;
            if (di < 2) {
                lcs = lcs_row[di];
                if (lcs == 0) { int tol = 0; if ((int)ctx.query.size() >= cs.qlimit_tol) tol = (int)((double)ctx.query.size() * cs.lcs_tol_rel); lcs = lcs_metric(ctx.query, text, tol); lcs_row[di] = (uint8_t)std::min(lcs, 255); }
            }
            Features f = cov.features(ctx, text, (double)lcs, id);
            auto [score, tie] = fusion_score(ctx.query, text, f, base);
            if (di < 2 && hits_row[di] == 0) hits_row[di] = (uint8_t)std::min(f.word_hits, 255);
            max_hits = std::max(max_hits, f.word_hits);
            final_scores.add({score, d.key, tie});
        };
        for (int id : overlap) process(id, 0.f);
        int pu = 0; for (int id : uniq) { if (pu >= wm_limit) break; process(id, 0.f); pu++; }
        for (auto& c : top) {
            int id = ix.doc_by_key(c.key); if (id < 0) continue;
            float mx = !top.empty() ? top[0].score : 1.f; float nb = mx > 0 ? c.score / mx : 0.f;
            process(id, nb);
        }
        if (max_hits == 0 && wm.empty()) return {};
        std::vector<ScoreEntry> fin = consolidate(final_scores.get_top_k());
        int trunc = -1;
        if (cs.truncate && !fin.empty()) {   // ResultProcessor.CalculateTruncationIndex
            int min_hits = std::max(cs.min_hits_abs, max_hits - cs.min_hits_rel);
            for (int i = (int)fin.size() - 1; i >= 0; i--) {
                auto it = key_idx.find(fin[i].key); if (it == key_idx.end()) continue; int di = it->second;
                uint8_t wh = di < 2 ? hits_row[di] : 0, lb = di < 2 ? lcs_row[di] : 0;
                if (wh >= min_hits || lb > 0 || fin[i].score >= (float)cs.truncation_score) { trunc = i; break; }
            }
        }
        int count = (trunc == -1 || !cs.truncate) ? max_results : std::min(std::max(0, trunc) + 1, max_results);
        if ((int)fin.size() > count) fin.resize(count);
        return fin;
    }
};

}  // namespace ifxo
