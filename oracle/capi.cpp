// ORACLE (test infrastructure, NOT product code): C entry points over the CPU restatement, loaded with ctypes by
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs ONLY.
// PARITY PINNING: the reference is C#/.NET and cannot run in this image (no dotnet/mono); the restatement is
// pinned by the reference's own known-answer tests (tests/test_oracle_known_answers.py), not by reference outputs.
//
// Follows /root/reference/src/Infidex/SearchEngine.cs:256-319 (Search) and :124-192 (IndexDocumentsInternal).
#include "filter.hpp"
#include "../include/infidex_gpu.h"     // ifx_index_image: the flattened index format of the boundary (ifxo_load_image, bench only)
#include <thread>
#include <atomic>
#include <string>

using namespace ifxo;

struct Engine {
    Index ix; Pipeline* pipe = nullptr;
    ~Engine() { delete pipe; }
};

struct SearchResult { std::vector<ScoreEntry> recs; int total = 0; int status = 0; std::vector<FacetEntry> facets; };

static SearchResult do_search(const Engine& e, sv raw, int max_results, int depth, bool enable_cov,
                              const CompiledFilter* filt, bool facets, std::vector<ScoreEntry>* stage1 = nullptr, Stage1Stats* st = nullptr) {
    SearchResult r;
    if (!e.ix.built) return r;
    str q = to_lower(normalize(trim(raw)));
    if (is_blank(q)) {
        if (!facets) return r;                      // SearchEngine.cs:295-296 Result.MakeEmptyResult()
        // SearchEngine.HandleEmptyQueryWithFacets (SearchEngine.cs:321-346): every live document with score ushort.MaxValue in id order, the
        // filter, Take(max), facets over what was taken; TotalCandidates is not set on this path
        std::vector<ScoreEntry> all; FilterVM vm;
        for (size_t id = 0; id < e.ix.docs.size(); id++) { const Doc& d = e.ix.docs[id]; if (d.deleted) continue; if (filt && !vm.execute(*filt, e.ix, (int)id)) continue; all.push_back({65535.f, d.key, 0}); if ((int)all.size() >= max_results) break; }
        if (vm.unsupported) r.status = 1;
        r.facets = build_facets(e.ix, all); r.total = 0; r.recs = std::move(all);
        return r;
    }
    SearchOut o = e.pipe->execute(q, enable_cov, depth, max_results, stage1, st);
    if (o.unsupported) { r.status = 1; return r; }
    std::vector<ScoreEntry> res = std::move(o.records);
    if (filt) {
        FilterVM vm; std::vector<ScoreEntry> kept;
        for (auto& s : res) { int id = e.ix.doc_by_key(s.key); if (id < 0) continue; if (vm.execute(*filt, e.ix, id)) kept.push_back(s); }
        if (vm.unsupported) r.status = 1;
        res.swap(kept);
    }
    if (facets) r.facets = build_facets(e.ix, res);
    r.total = (int)res.size();
    if ((int)res.size() > max_results) res.resize(max_results);
    r.recs = std::move(res);
    return r;
}

static std::string utf16_to_utf8(sv s) {
    std::string o;
    for (size_t i = 0; i < s.size(); i++) {
        uint32_t c = s[i];
        if (c >= 0xD800 && c < 0xDC00 && i + 1 < s.size()) { c = 0x10000 + ((c - 0xD800) << 10) + (s[i + 1] - 0xDC00); i++; }
        if (c < 0x80) o.push_back((char)c);
        else if (c < 0x800) { o.push_back((char)(0xC0 | (c >> 6))); o.push_back((char)(0x80 | (c & 0x3F))); }
        else if (c < 0x10000) { o.push_back((char)(0xE0 | (c >> 12))); o.push_back((char)(0x80 | ((c >> 6) & 0x3F))); o.push_back((char)(0x80 | (c & 0x3F))); }
        else { o.push_back((char)(0xF0 | (c >> 18))); o.push_back((char)(0x80 | ((c >> 12) & 0x3F))); o.push_back((char)(0x80 | ((c >> 6) & 0x3F))); o.push_back((char)(0x80 | (c & 0x3F))); }
    }
    return o;
}

extern "C" {

void* ifxo_create() { return new Engine(); }
void ifxo_destroy(void* h) { delete (Engine*)h; }

// flags: bit0 indexable, bit1 filterable, bit2 facetable; weight: 0 High, 1 Med, 2 Low
int ifxo_set_schema(void* h, int nfields, const uint16_t* names, const int* name_off, const int* weight, const int* flags) {
    Engine* e = (Engine*)h; e->ix.schema.clear();
    for (int i = 0; i < nfields; i++) { FieldDef f; f.name = str((const char16_t*)names + name_off[i], name_off[i + 1] - name_off[i]); f.weight = weight[i]; f.indexable = flags[i] & 1; f.filterable = flags[i] & 2; f.facetable = flags[i] & 4; e->ix.schema.push_back(f); }
    return 0;
}
// kinds[f]: 1 string (cols[f]=u16 blob, offs[f]=int64[n+1]), 2 int64 (cols[f]=int64[n]), 3 double (cols[f]=double[n])
int ifxo_add_docs(void* h, int n, const long long* keys, const int* kinds, const void* const* cols, const long long* const* offs) {
    Engine* e = (Engine*)h; int F = (int)e->ix.schema.size();
    for (int d = 0; d < n; d++) {
        std::vector<Value> vals(F);
        for (int f = 0; f < F; f++) {
            Value& v = vals[f];
            if (kinds[f] == 1) { v.kind = 1; const char16_t* b = (const char16_t*)cols[f]; v.s = str(b + offs[f][d], (size_t)(offs[f][d + 1] - offs[f][d])); }
            else if (kinds[f] == 2) { v.kind = 2; v.i = ((const long long*)cols[f])[d]; }
            else if (kinds[f] == 3) { v.kind = 3; v.d = ((const double*)cols[f])[d]; }
        }
        e->ix.add_document(keys[d], vals);
    }
    return 0;
}
// bench.py only: index state from a flattened image instead of add_docs + build (see Index::load_image)
int ifxo_load_image(void* h, const ifx_index_image* img) { Engine* e = (Engine*)h; e->ix.load_image(*img); delete e->pipe; e->pipe = new Pipeline(e->ix); return 0; }
int ifxo_build(void* h) { Engine* e = (Engine*)h; e->ix.build(); delete e->pipe; e->pipe = new Pipeline(e->ix); return 0; }

// status: 0 ok, 1 unsupported query/filter feature (short-query path, regex)
int ifxo_search(void* h, const uint16_t* q, int qlen, int max_results, int depth, int enable_cov,
                const uint8_t* filter, int filter_len, int enable_facets,
                long long* out_keys, float* out_scores, uint8_t* out_ties, int cap, int* out_n, int* out_total,
                char* facet_buf, int facet_cap, int* facet_len) {
    Engine* e = (Engine*)h; CompiledFilter cf; const CompiledFilter* pf = nullptr;
    if (filter && filter_len > 0) { cf = deserialize_filter(filter, (size_t)filter_len); if (!cf.ok) return 2; pf = &cf; }
    SearchResult r = do_search(*e, sv((const char16_t*)q, (size_t)qlen), max_results, depth, enable_cov != 0, pf, enable_facets != 0);
    int n = std::min((int)r.recs.size(), cap);
    for (int i = 0; i < n; i++) { out_keys[i] = r.recs[i].key; out_scores[i] = r.recs[i].score; out_ties[i] = r.recs[i].tie; }
    *out_n = n; if (out_total) *out_total = r.total;
    if (facet_len) {
        std::string fb; for (auto& f : r.facets) { fb += utf16_to_utf8(f.field); fb.push_back('\t'); fb += utf16_to_utf8(f.value); fb.push_back('\t'); fb += std::to_string(f.count); fb.push_back('\n'); }
        int m = std::min((int)fb.size(), facet_cap); if (facet_buf && m > 0) std::memcpy(facet_buf, fb.data(), m); *facet_len = m;
    }
    return r.status;
}

// Stage-1 only (TopKHeap after consolidation), for intermediate parity checks. stats: [path, candidates, streamed_postings, n_terms, n_fuzzy]
int ifxo_stage1(void* h, const uint16_t* q, int qlen, int depth, long long* out_keys, float* out_scores, int cap, int* out_n, long long* stats) {
    Engine* e = (Engine*)h; std::vector<ScoreEntry> s1; Stage1Stats st;
    SearchResult r = do_search(*e, sv((const char16_t*)q, (size_t)qlen), 10, depth, false, nullptr, false, &s1, &st);
    int n = std::min((int)s1.size(), cap);
    for (int i = 0; i < n; i++) { out_keys[i] = s1[i].key; out_scores[i] = s1[i].score; }
    *out_n = n;
    if (stats) { stats[0] = st.path; stats[1] = st.candidates; stats[2] = st.streamed_postings; stats[3] = st.n_terms; stats[4] = st.n_fuzzy; }
    return r.status;
}

// Per-chunk trace of the MaxScore threshold chain of one query (see Stage1Stats::trace); returns the number of chunks, fills <= cap rows of 6 doubles.
int ifxo_stage1_trace(void* h, const uint16_t* q, int qlen, int depth, double* out, int cap, long long* stats) {
    Engine* e = (Engine*)h; std::vector<ScoreEntry> s1; Stage1Stats st; std::vector<double> tr; st.trace = &tr;
    do_search(*e, sv((const char16_t*)q, (size_t)qlen), 10, depth, false, nullptr, false, &s1, &st);
    int n = (int)(tr.size() / 6); std::memcpy(out, tr.data(), sizeof(double) * 6 * (size_t)std::min(n, cap));
    if (stats) { stats[0] = st.path; stats[1] = st.candidates; stats[2] = st.streamed_postings; stats[3] = st.n_terms; stats[4] = st.n_fuzzy; }
    return n;
}

// Batch search over `threads` host threads (queries are independent; the index is immutable) -- CPU baseline timing.
// out arrays are [nq * cap]; out_n[nq]; status[nq]
int ifxo_search_batch(void* h, const uint16_t* qblob, const long long* qoff, int nq, int max_results, int depth, int enable_cov,
                      const uint8_t* filter, int filter_len, int threads,
                      long long* out_keys, float* out_scores, uint8_t* out_ties, int cap, int* out_n, int* status) {
    Engine* e = (Engine*)h; CompiledFilter cf; const CompiledFilter* pf = nullptr;
    if (filter && filter_len > 0) { cf = deserialize_filter(filter, (size_t)filter_len); if (!cf.ok) return 2; pf = &cf; }
    std::atomic<int> next{0};
    auto work = [&]() {
        for (;;) { int i = next.fetch_add(1); if (i >= nq) break;
            SearchResult r = do_search(*e, sv((const char16_t*)qblob + qoff[i], (size_t)(qoff[i + 1] - qoff[i])), max_results, depth, enable_cov != 0, pf, false);
            int n = std::min((int)r.recs.size(), cap);
            for (int k = 0; k < n; k++) { out_keys[(size_t)i * cap + k] = r.recs[k].key; out_scores[(size_t)i * cap + k] = r.recs[k].score; out_ties[(size_t)i * cap + k] = r.recs[k].tie; }
            out_n[i] = n; if (status) status[i] = r.status; }
    };
    if (threads <= 1) work(); else { std::vector<std::thread> ts; for (int t = 0; t < threads; t++) ts.emplace_back(work); for (auto& t : ts) t.join(); }
    return 0;
}

// ---- introspection (index parity between the oracle's builder and the product's host builder)
int ifxo_num_docs(void* h) { return (int)((Engine*)h)->ix.docs.size(); }
int ifxo_num_terms(void* h) { return (int)((Engine*)h)->ix.terms.size(); }
float ifxo_avgdl(void* h) { return ((Engine*)h)->ix.avgdl; }
void ifxo_doc_lens(void* h, float* out) { auto& v = ((Engine*)h)->ix.doc_len; std::memcpy(out, v.data(), v.size() * sizeof(float)); }
int ifxo_term_text(void* h, int t, uint16_t* buf, int cap) { auto& s = ((Engine*)h)->ix.terms[t].text; int n = std::min((int)s.size(), cap); std::memcpy(buf, s.data(), n * 2); return (int)s.size(); }
int ifxo_term_df(void* h, int t) { return ((Engine*)h)->ix.terms[t].df; }
int ifxo_term_npost(void* h, int t) { auto& x = ((Engine*)h)->ix.terms[t]; return x.df > 0 ? (int)x.docs.size() : 0; }
void ifxo_term_postings(void* h, int t, int* docs, uint8_t* w) { auto& x = ((Engine*)h)->ix.terms[t]; if (x.df <= 0) return; std::memcpy(docs, x.docs.data(), x.docs.size() * 4); std::memcpy(w, x.w.data(), x.w.size()); }
int ifxo_lookup_term(void* h, const uint16_t* s, int n) { auto& m = ((Engine*)h)->ix.term_ids; auto it = m.find(str((const char16_t*)s, (size_t)n)); return it == m.end() ? -1 : it->second; }
float ifxo_word_idf(void* h, const uint16_t* s, int n) { auto& m = ((Engine*)h)->ix.word_idf; auto it = m.find(str((const char16_t*)s, (size_t)n)); return it == m.end() ? -1.f : it->second; }
// kind: 0 prefix docset, 1 wm exact, 2 wm ld1; returns count (fills <= cap)
int ifxo_dict_docs(void* h, int kind, const uint16_t* s, int n, int* out, int cap) {
    Engine* e = (Engine*)h; const StrMap<std::vector<int>>& m = kind == 0 ? e->ix.prefix_docs : (kind == 1 ? e->ix.wm_exact : e->ix.wm_ld1);
    auto it = m.find(str((const char16_t*)s, (size_t)n)); if (it == m.end()) return -1;
    int c = std::min((int)it->second.size(), cap); std::memcpy(out, it->second.data(), (size_t)c * 4); return (int)it->second.size();
}
int ifxo_affix_last(void* h, const uint16_t* s, int n) { auto& m = ((Engine*)h)->ix.wm_affix_last; auto it = m.find(str((const char16_t*)s, (size_t)n)); return it == m.end() ? -1 : it->second; }

// ---- component-level entry points (known-answer tests of the reference's unit tests)
// FstIndex over an ad-hoc term list (FstIndexTests.cs): mode 0 MatchWithinEditDistance1, mode 1 GetByPrefix. Returns the reference's
// return value (total match count / number written), fills up to `cap` outputs.
int ifxo_fst_query(const uint16_t* blob, const int* off, const int* outs, int n, int mode, const uint16_t* q, int qn, int* out, int cap) {
    std::vector<std::pair<str, int>> items; for (int i = 0; i < n; i++) items.emplace_back(str((const char16_t*)blob + off[i], (size_t)(off[i + 1] - off[i])), outs[i]);
    std::sort(items.begin(), items.end());
    Trie t; t.build(items); std::vector<int> r; sv qs((const char16_t*)q, (size_t)qn); int ret;
    if (mode == 0) ret = t.match_ld1(qs, cap, r); else { int node = t.walk(qs); ret = t.collect(node, cap, r); }
    for (size_t i = 0; i < r.size() && (int)i < cap; i++) out[i] = r[i];
    return ret;
}
// WordMatcher.Lookup (kind 0) / LookupAffix (kind 1) for one word: ascending internal doc ids
int ifxo_wm_lookup(void* h, int kind, const uint16_t* s, int n, int* out, int cap) {
    Engine* e = (Engine*)h; if (!e->pipe) return -1;
    sv w((const char16_t*)s, (size_t)n); std::vector<int> r = kind == 0 ? e->pipe->wm_lookup(w) : e->pipe->wm_affix(w);
    int m = std::min((int)r.size(), cap); std::memcpy(out, r.data(), (size_t)m * sizeof(int)); return (int)r.size();
}
int ifxo_levenshtein(const uint16_t* a, int na, const uint16_t* b, int nb, int max_errors, int ic) { return lev(sv((const char16_t*)a, na), sv((const char16_t*)b, nb), max_errors, ic != 0); }
int ifxo_damerau(const uint16_t* a, int na, const uint16_t* b, int nb, int maxd, int ic) { return damerau(sv((const char16_t*)a, na), sv((const char16_t*)b, nb), maxd, ic != 0); }
int ifxo_normalize(const uint16_t* a, int na, uint16_t* out, int cap) { str r = normalize(sv((const char16_t*)a, na)); int n = std::min((int)r.size(), cap); std::memcpy(out, r.data(), n * 2); return (int)r.size(); }
// coverage of one (query, doc) pair with the engine's corpus statistics: out = [coverage_score, word_hits, fusion score bits, tie]
// CoverageEngine.SetWordIdfCache (BugReproductionTests.cs:24-31): replace the word-level idf cache of an (empty) engine
int ifxo_set_word_idf(void* h, const uint16_t* blob, const int* off, const float* idf, int n) {
    Engine* e = (Engine*)h; e->ix.word_idf.clear();
    for (int i = 0; i < n; i++) e->ix.word_idf[str((const char16_t*)blob + off[i], (size_t)(off[i + 1] - off[i]))] = idf[i];
    return 0;
}
int ifxo_coverage(void* h, const uint16_t* q, int nq, const uint16_t* d, int nd, double lcs, float bm25, int* out) {
    Engine* e = (Engine*)h; Coverage cov(e->ix); QueryCtx ctx = cov.prepare(sv((const char16_t*)q, nq));
    Features f = cov.features(ctx, sv((const char16_t*)d, nd), lcs, -1);
    auto fs = fusion_score(ctx.query, sv((const char16_t*)d, nd), f, bm25);
    out[0] = f.coverage_score; out[1] = f.word_hits; std::memcpy(&out[2], &fs.first, 4); out[3] = fs.second;
    return 0;
}
int ifxo_filter_eval(void* h, const uint8_t* filter, int len, int doc) {
    Engine* e = (Engine*)h; CompiledFilter cf = deserialize_filter(filter, (size_t)len); if (!cf.ok) return -2;
    FilterVM vm; bool r = vm.execute(cf, e->ix, doc); if (vm.unsupported) return -1; return r ? 1 : 0;
}

}  // extern "C"
