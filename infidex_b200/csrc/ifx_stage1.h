// infidex_b200 -- Stage 1 on device: query term resolution, LD1 expansion, tiered candidate selection,
// BM25+ with MaxScore over 4096-candidate chunks, exact emulation of the reference's pruning heap, top-K.
//
// What each routine replaces in the reference (src/Infidex/...):
//   prepare_query   Scoring/QueryAnalyzer.cs:10-54, Tokenization/Tokenizer.cs:144-200, Indexing/VectorModel.cs:376-563
//   expand_fuzzy    Indexing/Fst/FstIndex.cs:202-352 (MatchWithinEditDistance1), Indexing/VectorModel.cs:643-743
//   stage1_query    Scoring/TieredCandidateSelector.cs:53-532, Indexing/Bm25Scorer.cs:56-445,654-670
// Written against ifx::Ctx (one CTA on the GPU).
#pragma once
#include "ifx_base.h"

namespace ifx {

IFX_FN bool is_delim(const DevIndex& ix, uint16_t c) { return ix.cflags[c] & 4; }
IFX_FN bool is_space(const DevIndex& ix, uint16_t c) { return ix.cflags[c] & 2; }

IFX_FN int cmp_ordinal(const uint16_t* a, int na, const uint16_t* b, int nb) {   // string.CompareOrdinal
    int n = na < nb ? na : nb;
    for (int i = 0; i < n; i++) if (a[i] != b[i]) return a[i] < b[i] ? -1 : 1;
    return na == nb ? 0 : (na < nb ? -1 : 1);
}

IFX_FN float max_term_score(float idf, float avgdl) {   // VectorModel.cs:523-531
    const float maxTf = 255.f, k1 = 1.2f, b = 0.75f, delta = 1.0f;
    float minDlNorm = 1.f - b + b * (1.f / avgdl);
    float core = (maxTf * (k1 + 1.f)) / (maxTf + k1 * minDlNorm);
    return idf * (core + delta);
}

// ---------------------------------------------------------------------------------------------------------------
// prepare_query: one thread per query.
IFX_FN void prepare_query(const DevIndex& ix, const uint16_t* text, int len, int depth, int max_results, int enable_cov,
                          int filter_id, int enable_facets, QueryPlan& p, FuzzyItem* items, int items_cap, BatchCounters* bc, int qi) {
    p.status = 0; p.n_terms = 0; p.n_fuzzy = 0; p.qlen = 0; p.tlen = 0; p.depth = depth; p.max_results = max_results;
    p.enable_coverage = enable_cov; p.filter_id = filter_id; p.enable_facets = enable_facets; p.short_skip_coverage = 0; p.is_short3 = 0;
    if (len > MAX_QLEN || depth > MAX_K || depth < 1) { p.status = 4; return; }
    bool blank = true;
    for (int i = 0; i < len; i++) { p.qtext[i] = text[i]; if (!is_space(ix, text[i])) blank = false; }
    p.qlen = len;
    if (blank) { p.status = 8; return; }
    // QueryAnalyzer.Analyze
    int n_words = 0, n_long = 0, n_short = 0, tl = 0;
    for (int i = 0; i < len;) {
        while (i < len && is_delim(ix, text[i])) i++;
        if (i >= len) break;
        int b = i; while (i < len && !is_delim(ix, text[i])) i++;
        n_words++;
        if (i - b > MAX_TOKLEN) p.status |= 4;      // the coverage kernel's Levenshtein row (ifx_cov.h lev) cannot hold this word: flagged, never silently unmatched
        if (i - b >= 3) { if (n_long > 0) p.ttext[tl++] = u' '; for (int k = b; k < i; k++) p.ttext[tl++] = text[k]; n_long++; } else n_short++;
    }
    bool can_ngrams = n_words == 0 ? len >= 3 : n_long > 0;
    if (!can_ngrams) { p.status |= 1; return; }
    bool mixed = n_short > 0 && n_long > 0;
    if (!mixed) { tl = len; for (int i = 0; i < len; i++) p.ttext[i] = text[i]; }
    p.tlen = tl;
    // SearchPipeline.cs:110-142 short (<= 3 chars, no delimiter) query rules
    bool short3 = len <= 3; for (int i = 0; i < len; i++) if (is_delim(ix, text[i])) short3 = false;
    p.is_short3 = short3;
    if (short3) { int k = dict_lookup(ix.prefix.keys, text, len); if (k >= 0 && ix.prefix.row_ptr[k + 1] - ix.prefix.row_ptr[k] > 500) p.short_skip_coverage = 1; }
    // tokens: words (len >= 3) then padded 3-grams (Tokenizer.EnumerateShinglesForSearch), first 128 kept
    uint16_t padded[MAX_QLEN + 2]; padded[0] = PAD; padded[1] = PAD; for (int i = 0; i < tl; i++) padded[2 + i] = p.ttext[i];
    struct Raw { int32_t id; uint16_t off, len; };
    Raw raw[MAX_RAW_TOKENS]; int nr = 0;
    for (int i = 0; i < tl && nr < MAX_RAW_TOKENS;) {
        while (i < tl && is_delim(ix, p.ttext[i])) i++;
        if (i >= tl) break;
        int b = i; while (i < tl && !is_delim(ix, p.ttext[i])) i++;
        if (i - b >= 3) { raw[nr].off = (uint16_t)(b + 2); raw[nr].len = (uint16_t)(i - b); nr++; }
    }
    for (int k = 0; k + 3 <= tl + 2 && nr < MAX_RAW_TOKENS; k++) {
        if (padded[k] == PAD && padded[k + 1] == PAD && padded[k + 2] == PAD) continue;
        raw[nr].off = (uint16_t)k; raw[nr].len = 3; nr++;
    }
    for (int i = 0; i < nr; i++) raw[i].id = dict_lookup(ix.terms, padded + raw[i].off, raw[i].len);
    // RawToken.CompareTo: (TermId, ordinal text); equal elements are interchangeable so any sort gives the same sequence
    for (int i = 1; i < nr; i++) {
        Raw t = raw[i]; int j = i - 1;
        while (j >= 0) {
            int c = raw[j].id != t.id ? (raw[j].id < t.id ? -1 : 1) : (t.id >= 0 ? 0 : cmp_ordinal(padded + raw[j].off, raw[j].len, padded + t.off, t.len));
            if (c <= 0) break;
            raw[j + 1] = raw[j]; j--;
        }
        raw[j + 1] = t;
    }
    float avgdl = ix.avgdl > 0.f ? ix.avgdl : 1.f;
    int nt = 0;
    for (int i = 0; i < nr; i++) {
        if (i > 0 && raw[i].id == raw[i - 1].id && (raw[i].id >= 0 || cmp_ordinal(padded + raw[i].off, raw[i].len, padded + raw[i - 1].off, raw[i - 1].len) == 0)) continue;
        if (raw[i].id >= 0) {
            int df = ix.df[raw[i].id];
            if (df <= 0 || df > ix.stop_term_limit) continue;
            QTerm& t = p.terms[nt++];
            t.term_id = raw[i].id; t.df = df; t.list_off = ix.row_ptr[raw[i].id]; t.list_len = (int32_t)(ix.row_ptr[raw[i].id + 1] - ix.row_ptr[raw[i].id]);
            t.idf = compute_idf(ix, df); t.max_score = max_term_score(t.idf, avgdl);
        } else if (raw[i].len >= 4) {
            if (p.n_fuzzy >= MAX_FUZZY || raw[i].len > 64) { p.status |= 4; continue; }
            int slot = atomic_add(&bc->n_fuzzy_items, 1);
            if (slot >= items_cap) { p.status |= 4; continue; }
            QTerm& t = p.terms[nt]; t.term_id = -1; t.df = 0; t.list_off = 0; t.list_len = 0; t.idf = 0.f; t.max_score = 0.f;
            FuzzyReq& f = p.fuzzy[p.n_fuzzy++]; f.off = (uint16_t)(raw[i].off - 2); f.len = raw[i].len; f.term_slot = nt;
            items[slot].query = qi; items[slot].slot = p.n_fuzzy - 1;
            nt++;
        }
    }
    p.n_terms = nt;
}

// ---------------------------------------------------------------------------------------------------------------
// block primitives
struct ScanTmp { int w[33]; };

IFX_FN int block_excl_scan(const Ctx& c, int v, ScanTmp& tmp, int& total) {
#ifdef IFX_EMU
    (void)c; (void)tmp; total = v; return 0;
#else
    int incl = v;
    for (int d = 1; d < 32; d <<= 1) { int o = __shfl_up_sync(0xffffffffu, incl, d); if (c.lane() >= d) incl += o; }
    if (c.lane() == 31) tmp.w[c.warp()] = incl;
    c.sync();
    int base = 0, tot = 0, nw = c.nwarps();
    for (int i = 0; i < nw; i++) { int x = tmp.w[i]; if (i < c.warp()) base += x; tot += x; }
    total = tot;
    c.sync();
    return base + incl - v;
#endif
}
// single-barrier variant: callers alternate between two ScanTmp buffers (the barrier of the next call protects reuse)
IFX_FN int block_excl_scan_1b(const Ctx& c, int v, ScanTmp& tmp, int& total) {
#ifdef IFX_EMU
    (void)c; (void)tmp; total = v; return 0;
#else
    int incl = v;
    for (int d = 1; d < 32; d <<= 1) { int o = __shfl_up_sync(0xffffffffu, incl, d); if (c.lane() >= d) incl += o; }
    if (c.lane() == 31) tmp.w[c.warp()] = incl;
    c.sync();
    int base = 0, tot = 0, nw = c.nwarps();
    for (int i = 0; i < nw; i++) { int x = tmp.w[i]; if (i < c.warp()) base += x; tot += x; }
    total = tot;
    return base + incl - v;
#endif
}
IFX_FN int block_sum(const Ctx& c, int v, ScanTmp& tmp) { int t; block_excl_scan(c, v, tmp, t); return t; }

IFX_FN int64_t lower_bound_i32(const int32_t* a, int64_t lo, int64_t hi, int32_t target) {   // first index in [lo,hi) with a[i] >= target
    while (lo < hi) { int64_t mid = lo + ((hi - lo) >> 1); if (a[mid] < target) lo = mid + 1; else hi = mid; }
    return lo;
}

// lower bound executed by one full warp: 32-way splits instead of binary halving (log32 n dependent loads).
// Every lane of the calling warp must participate; the result is uniform across the warp.
IFX_FN int64_t warp_lower_bound(const Ctx& c, const int32_t* a, int64_t lo, int64_t hi, int32_t target) {
#ifdef IFX_EMU
    (void)c; return lower_bound_i32(a, lo, hi, target);
#else
    while (hi - lo > 32) {   // invariant: a[x] < target for x < lo, a[x] >= target for x >= hi
        int64_t step = (hi - lo + 31) / 32; int64_t p = lo + (int64_t)(c.lane() + 1) * step - 1; if (p > hi - 1) p = hi - 1;
        unsigned m = __ballot_sync(0xffffffffu, a[p] < target); int k = __popc(m);          // probes are monotone: lanes [0,k) see "less"
        int64_t pk = lo + (int64_t)(k + 1) * step - 1; if (pk > hi - 1) pk = hi - 1;          // first probe that is >= target (k < 32)
        int64_t pk1 = lo + (int64_t)k * step - 1; if (pk1 > hi - 1) pk1 = hi - 1;            // last probe that is < target (k > 0)
        if (k == 32) { lo = hi; break; }
        if (k > 0) lo = pk1 + 1;
        hi = pk;
    }
    int64_t i = lo + c.lane(); bool less = i < hi && a[i] < target;
    unsigned m = __ballot_sync(0xffffffffu, less);
    return lo + __popc(m);
#endif
}

// per-CTA global workspace
struct S1Workspace {
    unsigned* bits;        // candidate bitset over the shard's docs (all zero between uses)
    unsigned* bits2;       // membership bitset of the running AND-tier intersection (all zero between uses)
    int32_t* cand;         // sorted candidate ids
    int32_t* buf_a; int32_t* buf_b;   // AND-tier ping-pong arrays
    unsigned long long* surv_g;       // [CHUNK] flush survivors of one chunk when they exceed the shared staging buffer
    int64_t cand_cap, buf_cap;
};

struct TermS {             // term as seen by the scorer
    const int32_t* docs; const uint8_t* tf; const int32_t* skip; const unsigned* bm; const int32_t* bmr; int32_t len; int32_t df; float idf, max_score, suffix_after; int32_t term_id; int64_t cursor, s0, s1;
};

constexpr int S1_TILE = 6;

constexpr int SURV_CAP = 512;
constexpr int QH_SIZE = 256;
constexpr int SMALL_CHUNK = 512;                               // chunks up to this size are scored by a single warp, without block barriers
constexpr int SMALL_TERMS = S1_TILE * CHUNK / SMALL_CHUNK;      // ... when all their terms fit the tile buffer re-cut as [term][SMALL_CHUNK]
IFX_FN unsigned qh_hash(int32_t term_id) { return ((unsigned)term_id * 2654435761u) >> 24; }     // 8 bits = QH_SIZE
struct S1Shared {
    TermS terms[MAX_TERMS];
    int order[MAX_TERMS];
    int n_terms;
    alignas(16) float score[CHUNK];
    alignas(16) uint8_t tfm[S1_TILE][CHUNK];   // per-tile tf of (term, candidate slot); 0 = no match
    int32_t cand_s[CHUNK]; alignas(16) float nv_s[CHUNK];   // per-slot length norm of the vector form (the scalar form is needed for < 8 matches per term and chunk: recomputed)
    uint16_t cpref[2048];                // rank directory of `cbits`
    unsigned ballots[2][CHUNK / Ctx::WS + 8]; int bprefix[CHUNK / Ctx::WS + 8];
    int heap_size; float thr;
    // .NET PriorityQueue nodes packed as (doc << 32 | float bits of the priority), stored with a +3 shift so the four children of
    // node i (4i+1..4i+4) form one aligned 32-byte group; slots beyond the current size hold +huge sentinels
    alignas(16) float heap_pr[MAX_K + 8]; int32_t heap_doc[MAX_K + 8];   // split so that one 16-byte load fetches the four child priorities
    int32_t qh_key[QH_SIZE]; uint8_t qh_slot[QH_SIZE];   // term id -> slot in `terms` (open addressing; terms with idf > 0 only), for the forward-index lookups
    unsigned long long surv[SURV_CAP];     // (doc, score) of the last chunk's flush survivors, drained into the heap while the next chunk is staged
    union {                               // never live at the same time: selection/compaction vs. chunk scoring
        uint8_t dirty[MAX_CONTAINERS];    // containers of the global bitset touched by the current set operation (all zero between uses)
        unsigned cbits[2048];             // container-local bitmap of the current chunk's candidates (stream mode)
    };
    ScanTmp scan; ScanTmp scan2[2]; alignas(16) unsigned scan3[32][4];   // scan3: packed per-warp totals of the batched rank scan
    int bcast[8]; long long bcast64[4];
    unsigned long long streamed_mask[2];   // terms whose list the selector streamed in full (roofline accounting)
    unsigned long long peq[128];           // Myers pattern masks of the word being expanded (ASCII fast path)
};

// Stream a sorted id list and hand it to `put(word, mask)` as per-32-bit-word masks (word = id >> 5).
// Long lists: every thread takes runs of 8 consecutive ids (two 16-byte loads) and merges the ids that fall into the same word
// before calling `put` -- dense lists average several ids per word, so this cuts the global atomics behind `put` several-fold.
template <class Put>
IFX_FN void stream_list_words(const Ctx& c, const int32_t* list, int64_t n, Put put) {
    const int NT = c.nthreads(); int64_t done = 0;
    if (n >= 4096) {
        int64_t pre = (int64_t)((0 - (reinterpret_cast<uintptr_t>(list) >> 2)) & 3);     // ids in front of the first 16-byte boundary
        for (int64_t i = c.tid(); i < pre; i += NT) { int d = list[i]; put(d >> 5, 1u << (d & 31)); }
        struct alignas(16) Id4 { int32_t v[4]; };
        const Id4* p4 = reinterpret_cast<const Id4*>(list + pre); const int64_t n8 = (n - pre) >> 3;
        for (int64_t g0 = c.tid(); g0 < n8; g0 += 2LL * NT) {      // two runs (four loads) in flight per thread
            Id4 q[4]; const int64_t g1 = g0 + NT; const bool two = g1 < n8;
            q[0] = p4[2 * g0]; q[1] = p4[2 * g0 + 1]; if (two) { q[2] = p4[2 * g1]; q[3] = p4[2 * g1 + 1]; }
            for (int h = 0; h < (two ? 2 : 1); h++) {
                int word = q[2 * h].v[0] >> 5; unsigned mask = 0;
                for (int k = 0; k < 8; k++) { int d = q[2 * h + (k >> 2)].v[k & 3]; if ((d >> 5) != word) { put(word, mask); word = d >> 5; mask = 0; } mask |= 1u << (d & 31); }
                put(word, mask);
            }
        }
        done = pre + (n8 << 3);
    }
    const int64_t NT4 = 4LL * NT;
    for (int64_t i0 = done + c.tid(); i0 < n; i0 += NT4) {      // four independent loads in flight per thread
        int dd[4];
        for (int u = 0; u < 4; u++) { int64_t i = i0 + (int64_t)u * NT; dd[u] = i < n ? list[i] : -1; }
        for (int u = 0; u < 4; u++) if (dd[u] >= 0) put(dd[u] >> 5, 1u << (dd[u] & 31));
    }
}

// OR a sorted id list into the CTA's bitset; returns the number of newly set docs (block-wide).
IFX_FN int or_list_into_bits(const Ctx& c, const int32_t* list, int64_t n, S1Workspace& ws, S1Shared& sh) {
    int fresh = 0;
    stream_list_words(c, list, n, [&](int word, unsigned mask) { unsigned old = atomic_or(&ws.bits[word], mask); fresh += popc(mask & ~old); sh.dirty[word >> 11] = 1; });
    c.sync();
    return block_sum(c, fresh, sh.scan);
}

// Warp-aggregated unordered append (all lanes of the warp must call it together).
IFX_FN void warp_append(const Ctx& c, bool pred, int32_t value, int32_t* arr, int* counter) {
    unsigned m = c.ballot(pred); if (m == 0) return;
    int leader = ffs32(m) - 1; int base = 0;
    if (c.lane() == leader) base = atomic_add(counter, popc(m));
    base = c.shfl(base, leader);
    if (pred) arr[base + popc(m & c.lanemask_lt())] = value;
}

// AND of the first `cnt` terms of sh.order (TieredCandidateSelector.IntersectTerms). The running intersection lives both as an
// unordered id array (ping-pong ws.buf_a / ws.buf_b) and as the membership bitset ws.bits2. Each further list either streams
// past the bitset (coalesced, when it is not much longer than the running set) or is probed per surviving id (binary search).
// Returns the size (-1: buffer overflow); `res` points at the surviving ids; ws.bits2 is left all-zero.
IFX_FN int64_t intersect_terms(const Ctx& c, const DevIndex& ix, S1Workspace& ws, S1Shared& sh, int cnt, const int32_t*& res) {
    const int NT = c.nthreads();
    int by_len[MAX_TERMS];
    for (int i = 0; i < cnt; i++) by_len[i] = sh.order[i];
    for (int i = 1; i < cnt; i++) { int x = by_len[i]; int j = i - 1; while (j >= 0 && sh.terms[by_len[j]].len > sh.terms[x].len) { by_len[j + 1] = by_len[j]; j--; } by_len[j + 1] = x; }
    const TermS& t0 = sh.terms[by_len[0]];
    int64_t n = t0.len; if (n > ws.buf_cap) return -1;
    int32_t* cur = ws.buf_a; int32_t* nxt = ws.buf_b;
    const int bw = (int)(((int64_t)ix.n_docs + 31) >> 5);
    if (cnt > 1 && 4 * n >= bw && ws.buf_cap >= bw) {
        // Large running sets (>= 1/128 of the shard): keep the intersection as a bitset only. Lists with a membership bitmap are
        // ANDed word by word; the others (fuzzy unions, mid-size lists) are streamed against the bitset into a scratch bitmap
        // (buf_b) that then replaces it. No per-id probes or appends; ids come out ascending at the end.
        unsigned* acc = ws.bits2; unsigned* tmp = reinterpret_cast<unsigned*>(nxt);
        if (t0.bm) { for (int w = c.tid(); w < bw; w += NT) acc[w] = t0.bm[w]; }
        else stream_list_words(c, t0.docs, t0.len, [&](int word, unsigned mask) { atomic_or(&acc[word], mask); });
        c.sync();
        for (int li = 1; li < cnt; li++) {
            const TermS& t = sh.terms[by_len[li]];
            if (t.bm) { for (int w = c.tid(); w < bw; w += NT) acc[w] &= t.bm[w]; }
            else {
                for (int w = c.tid(); w < bw; w += NT) tmp[w] = 0u;
                c.sync();
                stream_list_words(c, t.docs, t.len, [&](int word, unsigned mask) { unsigned hit = acc[word] & mask; if (hit) atomic_or(&tmp[word], hit); });
                c.sync();
                for (int w = c.tid(); w < bw; w += NT) acc[w] = tmp[w];
            }
            c.sync();
        }
        int64_t total = 0;
        for (int w0 = 0; w0 < bw; w0 += 4 * NT) {
            unsigned v[4]; int mine = 0; const int wb = w0 + c.tid() * 4;
            for (int u = 0; u < 4; u++) { int w = wb + u; unsigned x = 0; if (w < bw) { x = acc[w]; acc[w] = 0u; } v[u] = x; mine += popc(x); }
            int tot; int off = block_excl_scan(c, mine, sh.scan, tot);
            int64_t o = total + off;
            for (int u = 0; u < 4; u++) { unsigned x = v[u]; while (x) { int b = ffs32(x) - 1; x &= x - 1; cur[o++] = ((wb + u) << 5) | b; } }
            total += tot;
        }
        res = cur; c.sync();
        return total;
    }
    for (int64_t i = c.tid(); i < n; i += NT) { int d = t0.docs[i]; cur[i] = d; if (cnt > 1) atomic_or(&ws.bits2[d >> 5], 1u << (d & 31)); }
    c.sync();
    for (int li = 1; li < cnt && n > 0; li++) {
        const TermS& t = sh.terms[by_len[li]];
        if (c.tid() == 0) sh.bcast[6] = 0;
        c.sync();
        if (!t.bm && (int64_t)t.len <= 32 * n) {             // stream the list past the membership bitset (dense terms are probed through their bitmap instead)
            const int64_t rounds = ((int64_t)t.len + NT - 1) / NT;
            for (int64_t r = 0; r < rounds; r += 4) {         // four independent loads in flight per thread
                int64_t i0 = r * NT + c.tid(); int32_t d[4]; bool in[4];
                for (int u = 0; u < 4; u++) { int64_t i = i0 + (int64_t)u * NT; d[u] = (r + u < rounds && i < t.len) ? t.docs[i] : -1; }
                for (int u = 0; u < 4; u++) in[u] = d[u] >= 0 && ((ws.bits2[d[u] >> 5] >> (d[u] & 31)) & 1u);
                for (int u = 0; u < 4; u++) warp_append(c, in[u], d[u], nxt, &sh.bcast[6]);
            }
            c.sync();
            for (int64_t i = c.tid(); i < n; i += NT) { int d = cur[i]; atomic_and(&ws.bits2[d >> 5], ~(1u << (d & 31))); }   // drop the old set ...
            c.sync();
            const int64_t nn = sh.bcast[6];
            if (li + 1 < cnt) for (int64_t i = c.tid(); i < nn; i += NT) { int d = nxt[i]; atomic_or(&ws.bits2[d >> 5], 1u << (d & 31)); }   // ... keep the survivors
            n = nn;
        } else {                                              // probe the (much longer) list once per surviving id
            const int64_t rounds = (n + NT - 1) / NT;
            for (int64_t r = 0; r < rounds; r += 4) {        // four independent probes in flight per thread
                int dd[4]; bool found[4]; unsigned wv[4];
                for (int u = 0; u < 4; u++) { int64_t i = (r + u) * NT + c.tid(); dd[u] = (r + u < rounds && i < n) ? cur[i] : -1; }
                if (t.bm) { for (int u = 0; u < 4; u++) wv[u] = dd[u] >= 0 ? t.bm[dd[u] >> 5] : 0u; for (int u = 0; u < 4; u++) found[u] = dd[u] >= 0 && ((wv[u] >> (dd[u] & 31)) & 1u); }
                else for (int u = 0; u < 4; u++) { found[u] = false; if (dd[u] >= 0) { int d = dd[u]; int64_t lo = 0, hi = t.len; if (t.skip) { lo = t.skip[d >> 16]; hi = t.skip[(d >> 16) + 1]; } int64_t p = lower_bound_i32(t.docs, lo, hi, d); found[u] = p < hi && t.docs[p] == d; } }
                for (int u = 0; u < 4; u++) { if (dd[u] >= 0 && !found[u]) atomic_and(&ws.bits2[dd[u] >> 5], ~(1u << (dd[u] & 31))); if (r + u < rounds) warp_append(c, found[u], dd[u], nxt, &sh.bcast[6]); }
            }
            c.sync();
            n = sh.bcast[6];
            if (li + 1 == cnt) { for (int64_t i = c.tid(); i < n; i += NT) { int d = nxt[i]; atomic_and(&ws.bits2[d >> 5], ~(1u << (d & 31))); } }
        }
        c.sync();
        int32_t* tmp = cur; cur = nxt; nxt = tmp;
    }
    if (cnt > 1 && n == 0) { /* bitset already empty: every id was cleared when it dropped out */ }
    res = cur;
    c.sync();
    return n;
}

// Expand the dirty containers of the bitset into ws.cand (ascending) and clear them. Returns the count.
IFX_FN int64_t compact_bits(const Ctx& c, const DevIndex& ix, S1Workspace& ws, S1Shared& sh, int32_t* out, int64_t out_cap, bool& overflow) {
    int ncont = (ix.n_docs + 65535) >> 16; int64_t total = 0; int64_t nwords = ((int64_t)ix.n_docs + 31) >> 5;
    overflow = false;
    for (int k = 0; k < ncont; k++) {
        if (!sh.dirty[k]) continue;                    // uniform across the CTA (shared flag, synced by callers)
        int64_t w0 = (int64_t)k * 2048, w1 = w0 + 2048; if (w1 > nwords) w1 = nwords;
        int per = (int)((w1 - w0 + c.nthreads() - 1) / c.nthreads());
        int64_t my0 = w0 + (int64_t)c.tid() * per, my1 = my0 + per; if (my1 > w1) my1 = w1;
        int cnt = 0; for (int64_t w = my0; w < my1; w++) cnt += popc(ws.bits[w]);
        int tot; int off = block_excl_scan(c, cnt, sh.scan, tot);
        if (total + tot > out_cap) { overflow = true; }
        else { int64_t o = total + off;
            for (int64_t w = my0; w < my1; w++) { unsigned v = ws.bits[w]; while (v) { int b = ffs32(v) - 1; out[o++] = (int32_t)((w << 5) | b); v &= v - 1; } } }
        for (int64_t w = my0; w < my1; w++) ws.bits[w] = 0;
        total += tot;
        c.sync();
        if (c.tid() == 0) sh.dirty[k] = 0;
    }
    c.sync();
    return total;
}

// .NET ArraySortHelper<T>.IntrospectiveSort with comparison (b.Idf.CompareTo(a.Idf)) over term indices -- unstable,
// reproduced exactly because idf ties decide which lists the selector unions (TieredCandidateSelector.cs:128,253).
struct IdfSorter {
    const TermS* t;
    IFX_FN int cmp(int a, int b) const { float x = t[b].idf, y = t[a].idf; return x < y ? -1 : (x > y ? 1 : 0); }
    IFX_FN void swap_if_greater(int* k, int i, int j) const { if (cmp(k[i], k[j]) > 0) { int x = k[i]; k[i] = k[j]; k[j] = x; } }
    IFX_FN void insertion(int* k, int n) const { for (int i = 0; i < n - 1; i++) { int t2 = k[i + 1]; int j = i; while (j >= 0 && cmp(t2, k[j]) < 0) { k[j + 1] = k[j]; j--; } k[j + 1] = t2; } }
    IFX_FN void down_heap(int* k, int i, int n) const { int d = k[i - 1]; while (i <= n / 2) { int ch = 2 * i; if (ch < n && cmp(k[ch - 1], k[ch]) < 0) ch++; if (!(cmp(d, k[ch - 1]) < 0)) break; k[i - 1] = k[ch - 1]; i = ch; } k[i - 1] = d; }
    IFX_FN void heap_sort(int* k, int n) const { for (int i = n >> 1; i >= 1; i--) down_heap(k, i, n); for (int i = n; i > 1; i--) { int x = k[0]; k[0] = k[i - 1]; k[i - 1] = x; down_heap(k, 1, i - 1); } }
    IFX_FN int partition(int* k, int n) const {
        int hi = n - 1, mid = hi >> 1;
        swap_if_greater(k, 0, mid); swap_if_greater(k, 0, hi); swap_if_greater(k, mid, hi);
        int pivot = k[mid]; { int x = k[mid]; k[mid] = k[hi - 1]; k[hi - 1] = x; }
        int left = 0, right = hi - 1;
        while (left < right) {
            while (cmp(k[++left], pivot) < 0) {}
            while (cmp(pivot, k[--right]) < 0) {}
            if (left >= right) break;
            int x = k[left]; k[left] = k[right]; k[right] = x;
        }
        if (left != hi - 1) { int x = k[left]; k[left] = k[hi - 1]; k[hi - 1] = x; }
        return left;
    }
    IFX_FN void sort(int* keys, int n) const {
        if (n < 2) return;
        int lg = 0; for (unsigned v = (unsigned)n; v >>= 1;) lg++;
        // explicit stack instead of recursion: (start, length, depth)
        int st_s[64], st_n[64], st_d[64]; int sp = 0; st_s[0] = 0; st_n[0] = n; st_d[0] = 2 * (lg + 1); sp = 1;
        while (sp > 0) {
            sp--; int* k = keys + st_s[sp]; int len = st_n[sp], depth = st_d[sp]; int s0 = st_s[sp];
            while (len > 1) {
                if (len <= 16) { if (len == 2) swap_if_greater(k, 0, 1); else if (len == 3) { swap_if_greater(k, 0, 1); swap_if_greater(k, 0, 2); swap_if_greater(k, 1, 2); } else insertion(k, len); break; }
                if (depth == 0) { heap_sort(k, len); break; }
                depth--;
                int p = partition(k, len);
                // reference recurses into the right part first, then loops on the left part; the two parts are disjoint so order is irrelevant
                st_s[sp] = s0 + p + 1; st_n[sp] = len - (p + 1); st_d[sp] = depth; sp++;
                len = p;
            }
        }
    }
};

// .NET PriorityQueue<int,float> (4-ary min-heap) on shared arrays -- Bm25Scorer.UpdateTopK (Bm25Scorer.cs:654-670)
IFX_FN float kv_score(unsigned long long kv) {
#ifdef IFX_EMU
    unsigned u = (unsigned)kv; float f; memcpy(&f, &u, 4); return f;
#else
    return __uint_as_float((unsigned)kv);
#endif
}
IFX_FN unsigned long long kv_pack(int doc, float pr) {
#ifdef IFX_EMU
    unsigned u; memcpy(&u, &pr, 4); return ((unsigned long long)(unsigned)doc << 32) | u;
#else
    return ((unsigned long long)(unsigned)doc << 32) | __float_as_uint(pr);
#endif
}
#define IFX_HP(i) heap_pr[(i) + 3]
#define IFX_HD(i) heap_doc[(i) + 3]
IFX_FN void heap_move_up(S1Shared& sh, int doc, float pr, int idx) {
    while (idx > 0) { int parent = (idx - 1) >> 2; float pp = sh.IFX_HP(parent); if (pr < pp) { sh.IFX_HP(idx) = pp; sh.IFX_HD(idx) = sh.IFX_HD(parent); idx = parent; } else break; }
    sh.IFX_HP(idx) = pr; sh.IFX_HD(idx) = doc;
}
// PriorityQueue.DequeueEnqueue on a full heap (the root is replaced and sifted down); returns the new root priority so the caller
// can keep the threshold in a register. PriorityQueue.MoveDown picks the first strictly-smallest of the (up to) four children; here as
// a two-level tournament with the same winner (ties keep the lower index at both levels). Only the priorities are on the
// dependent chain; the document id of a moved node follows with one load/store off it.
IFX_FN float heap_replace_root(S1Shared& sh, int doc, float pr, int sz) {
    int idx = 0, i; float root = pr;
    while ((i = 4 * idx + 1) < sz) {
#ifdef IFX_EMU
        const float p0 = sh.IFX_HP(i), p1 = sh.IFX_HP(i + 1), p2 = sh.IFX_HP(i + 2), p3 = sh.IFX_HP(i + 3);
#else
        const float4 v = *reinterpret_cast<const float4*>(&sh.heap_pr[i + 3]); const float p0 = v.x, p1 = v.y, p2 = v.z, p3 = v.w;
#endif
        const bool b01 = p1 < p0, b23 = p3 < p2; const float pa = b01 ? p1 : p0, pb = b23 ? p3 : p2;
        const bool bb = pb < pa; const float mp = bb ? pb : pa; const int mi = i + (bb ? (b23 ? 3 : 2) : (b01 ? 1 : 0));
        if (!(mp < pr)) break;
        sh.IFX_HP(idx) = mp; sh.IFX_HD(idx) = sh.IFX_HD(mi); if (idx == 0) root = mp;
        idx = mi;
    }
    sh.IFX_HP(idx) = pr; sh.IFX_HD(idx) = doc;
    return root;
}

// Exclusive scan over the worker threads [hw, nthreads) only (named barrier 1); every worker must call it.
IFX_FN int worker_excl_scan(const Ctx& c, int v, ScanTmp& tmp, int hw, int ntw) {
#ifdef IFX_EMU
    (void)c; (void)v; (void)tmp; (void)hw; (void)ntw; return 0;
#else
    int incl = v;
    for (int d = 1; d < 32; d <<= 1) { int o = __shfl_up_sync(0xffffffffu, incl, d); if (c.lane() >= d) incl += o; }
    if (c.lane() == 31) tmp.w[c.warp()] = incl;
    c.sync_workers(ntw);
    int base = 0; const int w0 = hw >> 5;
    for (int i = w0; i < c.warp(); i++) base += tmp.w[i];
    c.sync_workers(ntw);
    return base + incl - v;
#endif
}

// tf of ONE term for every candidate slot of the chunk into tfb[0..cnt) (slots without a match are left untouched, i.e. 0),
// executed by the `ntw` threads numbered wt. `may_stream`: the chunk's container-local bitmap (sh.cbits / sh.cpref) exists.
IFX_FN void stage1_lookup_term(S1Shared& sh, const TermS& tm, uint8_t* tfb, int cnt, int wt, int ntw, bool may_stream) {
    const int64_t sublen = tm.s1 - tm.s0;
    if (tm.bm && sublen > 2LL * cnt) {          // dense term, sparse chunk: O(1) bitmap probe per candidate (doc -> posting index -> tf)
        for (int jb = wt; jb < cnt; jb += 4 * ntw) {
            unsigned wv[4]; int rk[4]; int dd[4];
            for (int u = 0; u < 4; u++) { int j = jb + u * ntw; dd[u] = j < cnt ? sh.cand_s[j] : -1; if (dd[u] >= 0) { wv[u] = tm.bm[dd[u] >> 5]; rk[u] = tm.bmr[dd[u] >> 5]; } }
            for (int u = 0; u < 4; u++) if (dd[u] >= 0) { unsigned bit = 1u << (dd[u] & 31); if (wv[u] & bit) tfb[jb + u * ntw] = tm.tf[rk[u] + popc(wv[u] & (bit - 1))]; }
        }
    } else if (may_stream && sublen <= 16LL * cnt) {   // stream the posting sub-range (coalesced), O(1) slot lookup per posting
        for (int64_t i0 = tm.s0 + wt; i0 < tm.s1; i0 += 4LL * ntw) {      // four independent loads in flight per thread
            int dd[4]; uint8_t tv[4];                                          // tf fetched alongside the id (one latency, not two)
            for (int u = 0; u < 4; u++) { int64_t i = i0 + (int64_t)u * ntw; bool in = i < tm.s1; dd[u] = in ? tm.docs[i] : -1; tv[u] = (in && tm.tf) ? tm.tf[i] : (uint8_t)1; }
            for (int u = 0; u < 4; u++) if (dd[u] >= 0) {
                int d = dd[u] & 0xFFFF; unsigned wv = sh.cbits[d >> 5], bit = 1u << (d & 31);
                if (wv & bit) tfb[sh.cpref[d >> 5] + popc(wv & (bit - 1))] = tv[u];
            }
        }
    } else {                                    // sparse candidates: level-synchronous binary searches, four per thread at a time
        for (int jb = wt; jb < cnt; jb += 4 * ntw) {
            int64_t lo[4], hi[4]; int32_t d[4];
            for (int u = 0; u < 4; u++) { int j = jb + u * ntw; d[u] = j < cnt ? sh.cand_s[j] : 0x7fffffff; lo[u] = tm.s0; hi[u] = j < cnt ? tm.s1 : tm.s0; }
            for (int64_t span = sublen; span > 0; span >>= 1) {
                int32_t v[4]; int64_t mid[4];
                for (int u = 0; u < 4; u++) { mid[u] = lo[u] + ((hi[u] - lo[u]) >> 1); v[u] = lo[u] < hi[u] ? tm.docs[mid[u]] : 0; }
                for (int u = 0; u < 4; u++) if (lo[u] < hi[u]) { if (v[u] < d[u]) lo[u] = mid[u] + 1; else hi[u] = mid[u]; }
            }
            for (int u = 0; u < 4; u++) { int j = jb + u * ntw; if (j < cnt && lo[u] < tm.s1 && tm.docs[lo[u]] == d[u]) tfb[j] = tm.tf ? tm.tf[lo[u]] : (uint8_t)1; }
        }
    }
}

// Phase A of one term tile: tf of every (term, candidate slot) pair of the chunk into sh.tfm (0 = no match). Independent of the
// scores and of the threshold, so it runs on the worker threads (wt of ntw) while the heap warp is still draining the last chunk.
IFX_FN void stage1_phase_a(const Ctx& c, S1Shared& sh, int t0, int T, int cnt, int wt, int ntw) {
    (void)c;
    const int tile = T - t0 < S1_TILE ? T - t0 : S1_TILE; const bool use_bitmap = sh.bcast[5] != 0;
    for (int tt = 0; tt < tile; tt++) {
        const TermS& tm = sh.terms[t0 + tt];
        if (tm.idf <= 0.f || tm.s1 == tm.s0) continue;   // uniform
        stage1_lookup_term(sh, tm, sh.tfm[tt], cnt, wt, ntw, use_bitmap);
    }
}

// Bm25Scorer.cs:395-433 (Vector256 lanes) and :643-652 (scalar remainder); must not be contracted into FMAs.
// The document-length part of both forms depends only on the candidate, so it is evaluated once per chunk and slot:
//   vector form  norm = K1 * ((1 - B) + (B / avgdl) * dl)        (the scalar form, needed for < 8 matches per term and chunk, is recomputed)
IFX_FN float bm25_norm_vector(float dl, float avgdl) { const float K1 = 1.2f, B = 0.75f; float bdiv = B / avgdl; return K1 * ((1.f - B) + bdiv * dl); }
IFX_FN float bm25_from_norm_vector(float tf, float norm, float idf) { const float K1 = 1.2f, Delta = 1.0f; float denom = tf + norm; float core = (tf * (K1 + 1.0f)) / denom; return idf * (core + Delta); }
IFX_FN float bm25_scalar(float tf, float dl, float avgdl, float idf) {
    const float K1 = 1.2f, B = 0.75f, Delta = 1.0f;
    if (dl <= 0.f) dl = 1.f;
    float norm = K1 * (1.f - B + B * (dl / avgdl)); float denom = tf + norm;
    if (denom <= 0.f) return 0.f;
    float core = (tf * (K1 + 1.f)) / denom; return idf * (core + Delta);
}

struct Stage1Out { int64_t* key; int32_t* doc; float* score; int32_t* n; long long* dbg; };   // dbg: [n_cand, n_terms, selection ns, path] or null   // row pointers for this query (cap = depth)

// ---------------------------------------------------------------------------------------------------------------
// LD1 expansion of one unknown word: first 1024 trie-order matches (Myers bit-vector, search variant), union of
// their posting lists -> ascending unique doc list appended to the fuzzy pool.
IFX_FN void expand_fuzzy(const Ctx& c, const DevIndex& ix, QueryPlan& p, int fslot, S1Workspace& ws, S1Shared& sh,
                         int32_t* pool, unsigned long long pool_cap, BatchCounters* bc, const uint8_t* sorted_len, int32_t* matches /* [LD1_CAP] global or shared */) {
    const FuzzyReq fr = p.fuzzy[fslot];
    const uint16_t* q = p.ttext + fr.off; const int m = fr.len;
    const uint64_t maskM = 1ULL << (m - 1);
    int64_t T = ix.terms.n; int total = 0;
    // a term within (search-variant) edit distance 1 of the word lacks at most one of the word's distinct characters
    const unsigned long long qsig = char_sig(q, m);
    for (int ch = c.tid(); ch < 128; ch += c.nthreads()) { unsigned long long pm = 0; for (int j = 0; j < m; j++) if (q[j] == ch) pm |= 1ULL << j; sh.peq[ch] = pm; }
    c.sync();
    // Myers bit-vector (search variant, FstIndex.cs:316-335) along one dictionary term (ordinal `ord`, length L)
    auto myers_hit = [&](int ord, int L) -> bool {
        const uint16_t* s = ix.terms.chars + ix.terms.off[ord];
        uint64_t vp = ~0ULL, vn = 0ULL; int score = m;
        for (int k = 0; k < L; k++) {
            uint16_t ch = s[k]; uint64_t pm;
            if (ch < 128) pm = sh.peq[ch]; else { pm = 0; for (int j = 0; j < m; j++) if (q[j] == ch) pm |= 1ULL << j; }
            uint64_t x = pm | vn; uint64_t d0 = ((vp + (x & vp)) ^ vp) | x; uint64_t hn = vp & d0; uint64_t hp = vn | ~(vp | d0);
            uint64_t nvp = (hn << 1) | ~(d0 | (hp << 1)); uint64_t nvn = d0 & (hp << 1);
            if (hp & maskM) score++; if (hn & maskM) score--;
            vp = nvp; vn = nvn;
        }
        return score <= 1;
    };
    {   // Fast path: only dictionary terms of length m-1..m+1 can match, and the dictionary is also stored grouped by length
        // (ix.len_ptr / len_sig / len_ord), so the scan is a coalesced stream over those three groups. Matches are appended in
        // arbitrary order: the union below is order-free as long as all of them fit (<= LD1_CAP, the usual case).
        if (c.tid() == 0) sh.bcast[6] = 0;
        c.sync();
        const int l0 = m - 1 < 0 ? 0 : m - 1, l1 = m + 1 > 254 ? 254 : m + 1;
        const int64_t gb = ix.len_ptr[l0], ge = l1 >= l0 ? ix.len_ptr[l1 + 1] : gb; const int64_t g1 = ix.len_ptr[l0 + 1], g2 = l0 + 2 <= 255 ? ix.len_ptr[l0 + 2] : ge;
        const int64_t NT4 = 4LL * c.nthreads();
        for (int64_t base0 = gb; base0 < ge; base0 += NT4) {       // uniform trip count (warp votes inside); four signature loads in flight per thread
            const int64_t i0 = base0 + c.tid(); unsigned long long sg[4];
            for (int u = 0; u < 4; u++) { int64_t i = i0 + (int64_t)u * c.nthreads(); sg[u] = i < ge ? ix.len_sig[i] : ~0ULL; }
            for (int u = 0; u < 4; u++) {
                int64_t i = i0 + (int64_t)u * c.nthreads(); bool hit = false; int ord = 0;
                if (i < ge && popc64(qsig & ~sg[u]) <= 1) { ord = ix.len_ord[i]; int L = l0 + (i >= g1 ? 1 : 0) + (i >= g2 ? 1 : 0); hit = myers_hit(ord, L); }
                unsigned bm = c.ballot(hit);
                if (bm) { int leader = ffs32(bm) - 1; int base = 0; if (c.lane() == leader) base = atomic_add(&sh.bcast[6], popc(bm)); base = c.shfl(base, leader);
                          int at = base + popc(bm & c.lanemask_lt()); if (hit && at < LD1_CAP) matches[at] = ord; }
            }
        }
        c.sync();
        total = sh.bcast[6];
        c.sync();
        if (total > LD1_CAP) {
            // more matches than VectorModel.cs:662 keeps: the reference takes the first LD1_CAP in trie DFS order, so redo the scan
            // over the lexicographically sorted dictionary with an ordered compaction and stop there
            total = 0;
            for (int64_t base = 0; base < T && total < LD1_CAP; base += c.nthreads()) {
                int64_t i = base + c.tid(); bool hit = false;
                if (i < T) { int L = sorted_len[i]; hit = L >= m - 1 && L <= m + 1 && L < 255 && popc64(qsig & ~ix.term_sig[i]) <= 1 && myers_hit(ix.term_sorted[i], L); }
                int t2; int o2 = block_excl_scan(c, hit ? 1 : 0, sh.scan, t2);
                if (hit && total + o2 < LD1_CAP) matches[total + o2] = ix.term_sorted[i];
                total += t2;
            }
        }
    }
    c.sync();
    const int nm = total < LD1_CAP ? total : LD1_CAP;
    // Union of the matches' posting lists into the CTA's bitset: short lists one warp each (no block barrier per list), long
    // lists block-wide afterwards.
    const int BIG = 4096, BIGQ = 128; int fresh = 0;
    if (c.tid() == 0) sh.bcast[5] = 0;
    c.sync();
    for (int k = c.warp(); k < nm; k += c.nwarps()) {
        int ord = matches[k]; if (ix.df[ord] <= 0) continue;
        int64_t r0 = ix.row_ptr[ord], r1 = ix.row_ptr[ord + 1];
        if (r1 - r0 >= BIG) { int slot = BIGQ; if (c.lane() == 0) slot = atomic_add(&sh.bcast[5], 1); slot = c.shfl(slot, 0); if (slot < BIGQ) { if (c.lane() == 0) sh.bprefix[slot] = ord; continue; } }
        for (int64_t i = r0 + c.lane(); i < r1; i += Ctx::WS) {
            int d = ix.post_doc[i]; unsigned bit = 1u << (d & 31);
            unsigned old = atomic_or(&ws.bits[d >> 5], bit);
            if (!(old & bit)) fresh++;
            sh.dirty[d >> 16] = 1;
        }
    }
    c.sync();
    int df = block_sum(c, fresh, sh.scan);
    const int nbig = sh.bcast[5] < BIGQ ? sh.bcast[5] : BIGQ;
    for (int k = 0; k < nbig; k++) { int ord = sh.bprefix[k]; int64_t r0 = ix.row_ptr[ord], r1 = ix.row_ptr[ord + 1]; df += or_list_into_bits(c, ix.post_doc + r0, r1 - r0, ws, sh); }
    c.sync();
    QTerm& t = p.terms[fr.term_slot];
    if (df == 0) { if (c.tid() == 0) { t.df = 0; t.list_len = 0; } c.sync(); return; }
    if (c.tid() == 0) { unsigned long long b = atomic_add64(&bc->fuzzy_pool_used, (unsigned long long)df); sh.bcast64[0] = (long long)b; }
    c.sync();
    unsigned long long b = (unsigned long long)sh.bcast64[0]; bool ovf = false;
    int64_t cap = b + (unsigned long long)df <= pool_cap ? df : 0;
    int64_t n = compact_bits(c, ix, ws, sh, pool + b, cap, ovf);
    if (c.tid() == 0) {
        if (ovf || n != df) { p.status |= 4; atomic_add(&bc->overflow, 1); t.df = 0; t.list_len = 0; }
        else { float avgdl = ix.avgdl > 0.f ? ix.avgdl : 1.f; t.df = df; t.list_len = df; t.list_off = (int64_t)b; t.idf = compute_idf(ix, df); t.max_score = max_term_score(t.idf, avgdl); }
    }
    c.sync();
}

// Scoring + flush of a small chunk (cnt <= SMALL_CHUNK) by a team of SMALL_TEAM warps: team thread i owns the consecutive slots
// [i*PER, (i+1)*PER), ranks come from a warp scan plus the team's per-warp totals (named barrier 2), nothing synchronises with the
// rest of the block. Same arithmetic and the same order as the tiled path below (MaxScore test, rank -> Vector256 / scalar form,
// accumulation, eligibility, survivors in candidate order); tf bytes are read from the tile buffer re-cut as [term][SMALL_CHUNK]
// (filled by the workers) and zeroed again after use. `tw0`: first warp of the team.
#ifdef IFX_EMU
constexpr int SMALL_TEAM = 1;
#else
#ifndef IFX_SMALL_TEAM
#define IFX_SMALL_TEAM 8
#endif
constexpr int SMALL_TEAM = IFX_SMALL_TEAM;
#endif
IFX_FN void stage1_small_chunk(const Ctx& c, const DevIndex& ix, S1Shared& sh, int T, int cnt, int K, float avgdl, int tw0) {
    constexpr int PER = SMALL_CHUNK / (SMALL_TEAM * Ctx::WS);     // 4 slots per thread on the GPU
    const int wi = c.warp() - tw0, ti = wi * Ctx::WS + c.lane();
    const int j0 = ti * PER < cnt ? ti * PER : cnt, j1 = j0 + PER < cnt ? j0 + PER : cnt, nj = j1 - j0;
    const float thr = sh.thr; uint8_t* tfs = &sh.tfm[0][0]; int nscan = 0;
    auto team_excl = [&](int mine, int& total) -> int {           // exclusive prefix over the team's threads + team total
        int incl = mine;
        for (int d = 1; d < Ctx::WS; d <<= 1) { int o = c.shfl(incl, c.lane() >= d ? c.lane() - d : 0); if (c.lane() >= d) incl += o; }
        const int buf = nscan & 1; nscan++;
        if (c.lane() == Ctx::WS - 1) sh.scan3[wi][buf] = (unsigned)incl;
        c.sync_team(SMALL_TEAM * Ctx::WS);
        int base = 0, tot = 0;
        for (int i = 0; i < SMALL_TEAM; i++) { int x = (int)sh.scan3[i][buf]; if (i < wi) base += x; tot += x; }
        total = tot;
        return base + incl - mine;
    };
    for (int t = 0; t < T; t++) {
        const TermS& tm = sh.terms[t];
        if (tm.idf <= 0.f || (tm.term_id < 0 && tm.s1 == tm.s0)) continue;          // uniform (dictionary terms came through the forward index: no sub-range)
        uint8_t* tfb = tfs + t * SMALL_CHUNK; const float tbound = tm.max_score, tsuffix = tm.suffix_after;
        uint8_t tfv[PER]; bool alive[PER]; int mine = 0;
        for (int k = 0; k < PER; k++) { tfv[k] = k < nj ? tfb[j0 + k] : (uint8_t)0; alive[k] = tfv[k] != 0 && !(sh.score[j0 + k] + tbound + tsuffix <= thr); mine += alive[k] ? 1 : 0; }
        int m; int rank = team_excl(mine, m);
        const int vec_end = m - (m & 7);
        for (int k = 0; k < PER; k++) {
            if (alive[k]) {
                const float tf = (float)tfv[k];
                const float add = rank < vec_end ? bm25_from_norm_vector(tf, sh.nv_s[j0 + k], tm.idf) : bm25_scalar(tf, ix.doc_len[sh.cand_s[j0 + k]], avgdl, tm.idf);
                sh.score[j0 + k] += add; rank++;
            }
            if (tfv[k] != 0) tfb[j0 + k] = 0;
        }
    }
    // flush, part 1 (see the tiled path): eligibility against the chunk-start threshold, survivors compacted in candidate order
    const bool full = sh.heap_size >= K; int mine = 0; bool el[PER];
    for (int k = 0; k < PER; k++) { el[k] = k < nj && sh.score[j0 + k] > 0.f && (!full || sh.score[j0 + k] > thr) && !ix.deleted[sh.cand_s[j0 + k]]; mine += el[k] ? 1 : 0; }
    int total; int off = team_excl(mine, total);
    for (int k = 0; k < PER; k++) if (el[k]) sh.surv[off++] = kv_pack(sh.cand_s[j0 + k], sh.score[j0 + k]);
    if (ti == 0) sh.bcast[6] = total;
}

// Prefix precedence (TieredCandidateSelector.TrySelectPrefixCandidates): the candidates are the doc set of the query's first 1-3
// characters when that set is small enough. Returns its range in ix.prefix.doc_id.
IFX_FN bool prefix_shortcut(const DevIndex& ix, const QueryPlan& p, int K, int64_t& r0, int64_t& pop) {
    int maxl = p.tlen < 3 ? p.tlen : 3;
    for (int len = maxl; len >= 1; len--) {
        int k = dict_lookup(ix.prefix.keys, p.ttext, len); if (k < 0) continue;
        r0 = ix.prefix.row_ptr[k]; pop = ix.prefix.row_ptr[k + 1] - r0;
        if (pop == 0) continue;
        if (pop > (int64_t)K * 20) continue;
        if (pop <= (int64_t)K * 10) { int lim = K * 2 < 100 ? K * 2 : 100; return pop >= lim; }
    }
    return false;
}

// ---------------------------------------------------------------------------------------------------------------
IFX_FN void stage1_query(const Ctx& c, const DevIndex& ix, const QueryPlan& p, const int32_t* pool, S1Workspace& ws, S1Shared& sh,
                         Stage1Out out, BatchCounters* bc) {
    const int K = p.depth; const int NT = c.nthreads();
    if (c.tid() == 0) {
        int n = 0;
        for (int i = 0; i < p.n_terms; i++) {
            const QTerm& q = p.terms[i];
            if (q.df <= 0 || q.df > ix.stop_term_limit) continue;      // VectorModel.cs:521
            TermS& t = sh.terms[n]; t.len = q.list_len; t.df = q.df; t.idf = q.idf; t.max_score = q.max_score; t.cursor = 0; t.term_id = q.term_id;
            if (q.term_id >= 0) { t.docs = ix.post_doc + q.list_off; t.tf = ix.post_tf + q.list_off; int sk = ix.skip_id[q.term_id]; t.skip = sk >= 0 ? ix.skip_ptr + (size_t)sk * (ix.n_cont + 1) : nullptr;
                int bi = ix.bm_id[q.term_id]; t.bm = bi >= 0 ? ix.bm_bits + (size_t)bi * ix.bm_words : nullptr; t.bmr = bi >= 0 ? ix.bm_rank + (size_t)bi * ix.bm_words : nullptr; }
            else { t.docs = pool + q.list_off; t.tf = nullptr; t.skip = nullptr; t.bm = nullptr; t.bmr = nullptr; }
            n++;
        }
        float suf = 0.f; for (int i = n - 1; i >= 0; i--) { sh.terms[i].suffix_after = suf; suf = suf + sh.terms[i].max_score; }   // ComputeSuffixSums
        for (int i = 0; i < QH_SIZE; i++) sh.qh_key[i] = -1;
        for (int i = 0; i < n; i++) { const TermS& t = sh.terms[i]; if (t.term_id < 0 || t.idf <= 0.f) continue;      // (ids are unique within a query)
            unsigned h = qh_hash(t.term_id); while (sh.qh_key[h] >= 0) h = (h + 1) & (QH_SIZE - 1); sh.qh_key[h] = t.term_id; sh.qh_slot[h] = (uint8_t)i; }
        sh.n_terms = n; sh.heap_size = 0; sh.thr = 0.f; sh.streamed_mask[0] = sh.streamed_mask[1] = 0;
        for (int i = 0; i < MAX_K + 8; i++) { sh.heap_pr[i] = 3.0e38f; sh.heap_doc[i] = 0; }    // sentinels (any real BM25 score is far smaller)
        out.n[0] = 0;
    }
    c.sync();
    const int T = sh.n_terms;
    if (T == 0 || ix.n_live == 0 || p.status != 0) return;
    const float avgdl = ix.avgdl > 0.f ? ix.avgdl : 1.f;

    // ---- candidate selection (TieredCandidateSelector.SelectCandidates)
#if !defined(IFX_EMU) && defined(IFX_S1_TIMERS)
    long long smark = 0; if (c.tid() == 0) { asm volatile("mov.u64 %0, %%clock64;" : "=l"(smark) :: "memory"); if (out.dbg) for (int k = 20; k < 24; k++) out.dbg[k] = 0; }
#define IFX_STICK(k) do { c.sync(); if (c.tid() == 0 && out.dbg) { long long now_; asm volatile("mov.u64 %0, %%clock64;" : "=l"(now_) : "r"(sh.bcast[0]) : "memory"); out.dbg[20 + (k)] += now_ - smark; smark = now_; } } while (0)
#else
#define IFX_STICK(k) do { } while (0)
#endif
    const int32_t* cand = nullptr; int64_t n_cand = 0;
    if (c.tid() == 0) { sh.bcast64[0] = -1; sh.bcast64[1] = 0; int64_t r0, pop; if (prefix_shortcut(ix, p, K, r0, pop)) { sh.bcast64[0] = r0; sh.bcast64[1] = pop; } }
    c.sync();
    unsigned long long algo = 0;
    if (sh.bcast64[0] >= 0) { cand = ix.prefix.doc_id + sh.bcast64[0]; n_cand = sh.bcast64[1]; algo += 4ULL * (unsigned long long)n_cand; }
    else {
        if (c.tid() == 0) {
            bool typo = false; float max_idf = 0.f;
            for (int i = 0; i < T; i++) { if (sh.terms[i].df < 10) typo = true; if (sh.terms[i].idf > max_idf) max_idf = sh.terms[i].idf; sh.order[i] = i; }
            IdfSorter srt{sh.terms}; srt.sort(sh.order, T);
            sh.bcast[0] = (typo || T == 1) ? 1 : 0; ((float*)sh.bcast)[1] = max_idf;
        }
        c.sync();
        const bool disjunctive = sh.bcast[0] != 0; const float max_idf = ((float*)sh.bcast)[1];
        int64_t g = 0;
        IFX_STICK(0);   // prefix shortcut + idf sort
        if (disjunctive) {   // SelectCandidatesDisjunctive
            bool selective = false;
            for (int oi = 0; oi < T; oi++) {
                const TermS& t = sh.terms[sh.order[oi]];
                bool lowq = t.idf < (max_idf * 0.2f);
                if (T > 1 && lowq && selective) continue;
                g += or_list_into_bits(c, t.docs, t.len, ws, sh);
                if (c.tid() == 0) sh.streamed_mask[sh.order[oi] >> 6] |= 1ULL << (sh.order[oi] & 63);
                if (!lowq && g > 0) selective = true;
                if (g >= (int64_t)K * 100) break;
            }
        } else {
            if (c.tid() == 0) for (int i = 0; i < T; i++) sh.streamed_mask[i >> 6] |= 1ULL << (i & 63);   // every list of the AND tier
            const int32_t* r0 = nullptr; int64_t n0 = intersect_terms(c, ix, ws, sh, T, r0);
            if (n0 < 0) { if (c.tid() == 0) out.n[0] = -1; return; }
            g += or_list_into_bits(c, r0, n0, ws, sh);
            IFX_STICK(1);   // AND tier 0
            if (g < (int64_t)K * 2) {
                if (T >= 3 && g < (int64_t)K * 3) { const int32_t* r1 = nullptr; int64_t n1 = intersect_terms(c, ix, ws, sh, T - 1, r1); if (n1 > 0) g += or_list_into_bits(c, r1, n1, ws, sh); }
                IFX_STICK(2);   // AND tier 1
                if (g < (int64_t)K * 5) {
                    int sel[2]; int ns = 0; float cutoff = max_idf * 0.3f; int capn = T < 2 ? T : 2;
                    for (int oi = 0; oi < T && ns < capn; oi++) { const TermS& t = sh.terms[sh.order[oi]]; if (t.idf <= 0.f) continue; if (t.idf < cutoff) continue; sel[ns++] = sh.order[oi]; }
                    for (int si = 0; si < ns; si++) { const TermS& t = sh.terms[sel[si]]; g += or_list_into_bits(c, t.docs, t.len, ws, sh); if (g >= (int64_t)K * 10) break; }
                }
            }
        }
        bool ovf = false;
        IFX_STICK(1);   // list unions (disjunctive: everything; AND path: the top-idf lists after the tiers)
        n_cand = compact_bits(c, ix, ws, sh, ws.cand, ws.cand_cap, ovf);
        IFX_STICK(3);   // bitset -> sorted candidate array
        if (ovf) { if (c.tid() == 0) out.n[0] = -1; return; }
        cand = ws.cand;
        algo += 2ULL * (unsigned long long)((ix.n_docs + 7) / 8);
    }
    if (c.tid() == 0 && out.dbg) { out.dbg[0] = n_cand; out.dbg[1] = T; out.dbg[3] = sh.bcast64[0] >= 0 ? 1 : (sh.bcast[0] ? 2 : 3);
#ifndef IFX_EMU
        unsigned long long tn; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tn)); out.dbg[2] = (long long)tn;
#endif
    }
    // ---- roofline accounting: full-stream lists 5 B/posting, probe-only lists min(5 df, 32 |C|), 4 B doc_len per candidate
    if (c.tid() == 0) {
        for (int i = 0; i < T; i++) {
            unsigned long long full = (sh.terms[i].tf ? 5ULL : 4ULL) * (unsigned long long)sh.terms[i].len;
            bool streamed = (sh.streamed_mask[i >> 6] >> (i & 63)) & 1ULL;
            unsigned long long probe = 32ULL * (unsigned long long)n_cand;
            algo += streamed ? full : (full < probe ? full : probe);
        }
        algo += 4ULL * (unsigned long long)n_cand;
        atomic_add64(&bc->algo_bytes, algo);
    }

    // ---- BM25 scoring over chunks (ProcessBlockedCandidates / ProcessChunk / ScoreBlockStruct)
    // The top-K heap is inherently sequential (one thread replays .NET's PriorityQueue), everything else is block-parallel.
    // To keep both busy the chunk loop is software-pipelined: the survivors of chunk i are compacted into `surv`, and while
    // thread 0 (warp 0 = "heap warp") drains them into the heap, warps 1.. ("workers") already run chunk i+1's set-up and the
    // score-independent membership lookups (phase A). The two sides meet at the barrier in front of phase B, which is the first
    // place chunk i+1 needs the threshold chunk i produced. Workers synchronise among themselves on named barrier 1.
    const int NW = c.nwarps();
    const int hw = NT > Ctx::WS ? Ctx::WS : 0;                 // threads of the heap warp (0: single-warp build, everything sequential)
    const bool worker = c.tid() >= hw; const int wt = c.tid() - hw, NTW = NT - hw;
    int pend = 0;                                              // survivors of the previous chunk waiting in sh.surv (uniform)
#if !defined(IFX_EMU) && defined(IFX_S1_TIMERS)
    long long tph[6] = {0, 0, 0, 0, 0, 0}; long long tmark = 0; long long wph[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long wmark = 0;
#define IFX_COUNT(x) tph[5] += (x)
#else
#define IFX_COUNT(x) do { } while (0)
#endif
#if !defined(IFX_EMU) && defined(IFX_S1_TIMERS)
    // phase timers (debug builds with -DIFX_S1_TIMERS only; they cost registers in every thread): thread 0 = heap warp's view, thread `hw` = the workers' view. The "memory" clobber keeps
    // the clock reads from being scheduled across the barriers they bracket.
    // and the dependence on a shared-memory load issued after the barrier keeps ptxas from hoisting them above it
    auto rdclock = [&]() -> long long { long long t_; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t_) : "r"(*(volatile int*)&sh.bcast[2]) : "memory"); return t_; };
#define IFX_TICK(k) do { if (c.tid() == 0) { long long now_ = rdclock(); tph[k] += now_ - tmark; tmark = now_; } } while (0)
#define IFX_WTICK(k) do { if (c.tid() == hw) { long long now_ = rdclock(); wph[k] += now_ - wmark; wmark = now_; } } while (0)
    if (c.tid() == 0) tmark = rdclock();
    if (c.tid() == hw) wmark = rdclock();
#else
#define IFX_TICK(k) do { } while (0)
#define IFX_WTICK(k) do { } while (0)
#endif
    auto drain = [&]() {                                       // Bm25Scorer.cs:316-329 + UpdateTopK (:654-670) over the compacted survivors, in candidate order
        float thr_r = sh.thr; int hs = sh.heap_size;           // threshold and size live in registers for the whole drain
        auto one = [&](unsigned long long kv) {
            const float s = kv_score(kv); IFX_COUNT(1);
            if (hs < K) { heap_move_up(sh, (int)(kv >> 32), s, hs); hs++; if (hs == K) thr_r = sh.IFX_HP(0); IFX_COUNT(1 << 20); }
            else if (s > thr_r) { thr_r = heap_replace_root(sh, (int)(kv >> 32), s, hs); IFX_COUNT(1 << 20); }
        };
        if (pend <= SURV_CAP) { for (int i = 0; i < pend; i++) one(sh.surv[i]); }
        else for (int i = 0; i < pend; i += 8) {               // global staging: eight independent loads in flight, then the sequential updates
            unsigned long long kv[8];
            for (int u = 0; u < 8; u++) kv[u] = i + u < pend ? ws.surv_g[i + u] : 0ULL;
            for (int u = 0; u < 8; u++) if (i + u < pend) one(kv[u]);
        }
        sh.thr = thr_r; sh.heap_size = hs;
    };
    for (int64_t pos = 0; pos < n_cand;) {
        if (c.tid() == 0 && pend) drain();
        IFX_TICK(4);   // heap drain of the previous chunk (overlapped with the workers below)
        if (worker) {
            IFX_WTICK(7);  // everything after the join (phase B, eligibility, compaction)
            // container run: candidates sharing id >> 16, cut into sub-chunks of 4096
            if (c.warp() == hw / Ctx::WS) {
                int hb = cand[pos] >> 16; int64_t lim = ((int64_t)hb + 1) << 16;
                int64_t ce = lim > 0x7fffffffLL ? n_cand : warp_lower_bound(c, cand, pos, n_cand, (int32_t)lim);
                if (c.lane() == 0) {
                    int64_t cnt = ce - pos; sh.bcast[4] = (cnt <= CHUNK && (pos == 0 || (cand[pos - 1] >> 16) != hb)) ? 1 : 0;   // chunk == all candidates of the container
                    if (cnt > CHUNK) cnt = CHUNK; sh.bcast[2] = (int)cnt; sh.bcast[5] = 0;
                }
            }
            c.sync_workers(NTW);
            const int cnt = sh.bcast[2]; const bool whole_container = sh.bcast[4] != 0;
            for (int jb = wt; jb < cnt; jb += 4 * NTW) {             // four candidates per thread in flight (id -> length is a dependent load)
                int d[4]; float dl[4];
                for (int u = 0; u < 4; u++) { int j = jb + u * NTW; d[u] = j < cnt ? cand[pos + j] : -1; }
                for (int u = 0; u < 4; u++) dl[u] = d[u] >= 0 ? ix.doc_len[d[u]] : 0.f;
                for (int u = 0; u < 4; u++) if (d[u] >= 0) { int j = jb + u * NTW; sh.cand_s[j] = d[u]; sh.nv_s[j] = bm25_norm_vector(dl[u], avgdl); sh.score[j] = 0.f; }
            }
            c.sync_workers(NTW);
            IFX_WTICK(0);  // container run + candidate ids, lengths, norms
            const int32_t first = sh.cand_s[0], last = sh.cand_s[cnt - 1];
            const bool small = cnt <= SMALL_CHUNK && T <= SMALL_TERMS;
            for (int t = c.warp() - hw / Ctx::WS; t < T; t += NW - hw / Ctx::WS) {   // posting sub-range of every term for this chunk (monotone cursors); one warp per term, 32-way searches
                TermS& tm = sh.terms[t];
                if (small && tm.term_id >= 0) continue;          // small chunks reach dictionary terms through the forward index
                int64_t lo = tm.cursor, hi = tm.len;
                if (tm.skip) { int cc = first >> 16; int64_t b0 = tm.skip[cc], b1 = tm.skip[cc + 1]; if (b0 > lo) lo = b0; hi = b1; if (lo > hi) lo = hi; }   // window = this container's postings
                int64_t s0, s1;
                if (tm.skip && whole_container) { s0 = lo; s1 = hi; }
                else { s0 = warp_lower_bound(c, tm.docs, lo, hi, first); s1 = last == 0x7fffffff ? hi : warp_lower_bound(c, tm.docs, s0, hi, last + 1); }
#ifndef IFX_EMU
                __syncwarp();                                    // every lane has read tm.cursor / tm.len before lane 0 overwrites them (racecheck)
#endif
                if (c.lane() == 0) {
                    tm.s0 = s0; tm.s1 = s1; tm.cursor = s1;
                    if (s1 > s0 && s1 - s0 <= 16LL * cnt) sh.bcast[5] = 1;     // benign race: every writer stores 1
                }
            }
            c.sync_workers(NTW);
            IFX_WTICK(1);  // posting sub-range bounds
            if (small) {
                // Small chunk: tf of ALL terms now, for the single-warp scorer. Dictionary terms: one warp per candidate walks the doc's
                // forward list (a few dozen (term, tf) pairs, one coalesced read) and keeps the pairs whose term is in the query's hash --
                // one memory latency per candidate instead of a binary search per (candidate, term). Fuzzy unions have no term id: they
                // are searched in their pool list, one warp per term.
                uint8_t* tfs = &sh.tfm[0][0];
                for (int j = c.warp() - hw / Ctx::WS; j < cnt; j += NW - hw / Ctx::WS) {
                    const int d = sh.cand_s[j]; const int64_t r0 = ix.fwd_ptr[d], r1 = ix.fwd_ptr[d + 1];
                    for (int64_t i = r0 + c.lane(); i < r1; i += Ctx::WS) {
                        const int32_t tid = ix.fwd_term[i]; unsigned h = qh_hash(tid);
                        for (;;) { const int32_t k = sh.qh_key[h]; if (k == tid) { tfs[(int)sh.qh_slot[h] * SMALL_CHUNK + j] = ix.fwd_tf[i]; break; } if (k < 0) break; h = (h + 1) & (QH_SIZE - 1); }
                    }
                }
                for (int t = c.warp() - hw / Ctx::WS; t < T; t += NW - hw / Ctx::WS) {
                    const TermS& tm = sh.terms[t];
                    if (tm.term_id >= 0 || tm.idf <= 0.f || tm.s1 == tm.s0) continue;
                    stage1_lookup_term(sh, tm, tfs + t * SMALL_CHUNK, cnt, c.lane(), Ctx::WS, false);
                }
            } else {
            // Container-local bitmap of the chunk's candidates + per-word rank directory: posting -> candidate slot in O(1)
            // (all candidates of a chunk share id >> 16). Built only when some term streams its posting sub-range.
            if (sh.bcast[5] != 0) {
                for (int w = wt; w < 2048; w += NTW) sh.cbits[w] = 0;
                c.sync_workers(NTW);
                for (int j = wt; j < cnt; j += NTW) { int d = sh.cand_s[j] & 0xFFFF; atomic_or(&sh.cbits[d >> 5], 1u << (d & 31)); }
                c.sync_workers(NTW);
                int per = (2048 + NTW - 1) / NTW; int w0 = wt * per < 2048 ? wt * per : 2048, w1 = w0 + per < 2048 ? w0 + per : 2048; int mine = 0;
                for (int w = w0; w < w1; w++) mine += popc(sh.cbits[w]);
                int run = worker_excl_scan(c, mine, sh.scan, hw, NTW);
                for (int w = w0; w < w1; w++) { sh.cpref[w] = (uint16_t)run; run += popc(sh.cbits[w]); }
                c.sync_workers(NTW);
            }
            IFX_WTICK(2);  // candidate bitmap + rank directory
            stage1_phase_a(c, sh, 0, T, cnt, wt, NTW);
            }
            IFX_WTICK(3);  // phase A, tile 0 (this thread's share) / small-chunk lookups
        }
        c.sync();      // join: heap drained, chunk staged, tile 0 looked up
        IFX_WTICK(4);  // waiting at the join (slower workers / the heap drain)
        IFX_TICK(0);   // heap warp waiting for the workers (set-up + phase A beyond the drain)
        const int cnt = sh.bcast[2];
        if (cnt <= SMALL_CHUNK && T <= SMALL_TERMS) {
            if (c.warp() >= hw / Ctx::WS && c.warp() < hw / Ctx::WS + SMALL_TEAM) stage1_small_chunk(c, ix, sh, T, cnt, K, avgdl, hw / Ctx::WS);
            c.sync();
            pend = sh.bcast[6]; pos += cnt;
            IFX_TICK(3);
            continue;
        }
        const float thr = sh.thr; const int rounds = (cnt + NT - 1) / NT;
        const int per_thread = (CHUNK + NT - 1) / NT; const int j0 = c.tid() * per_thread < cnt ? c.tid() * per_thread : cnt; const int j1 = j0 + per_thread < cnt ? j0 + per_thread : cnt;
        // Terms are processed in tiles: the membership (tf) lookups of a whole tile are issued back to back with no barrier in
        // between (independent of the scores), then the order-dependent part -- MaxScore skip, rank within the chunk, formula
        // choice, accumulation -- runs term by term with a single barrier each.
        for (int t0 = 0; t0 < T; t0 += S1_TILE) {
            const int tile = T - t0 < S1_TILE ? T - t0 : S1_TILE;
            if (t0 > 0) { IFX_WTICK(7); if (worker) stage1_phase_a(c, sh, t0, T, cnt, wt, NTW); IFX_WTICK(5); c.sync(); IFX_WTICK(6); }
            IFX_TICK(2);   // phase A of the later tiles
            // Each thread owns the consecutive candidate slots [j0, j1). A (candidate, term) pair counts as a match only if the
            // MaxScore test keeps it (Bm25Scorer.cs:354); its rank among the chunk's matches of this term selects the formula
            // (first 8*floor(m/8) matches: Vector256 form, the rest: scalar form; Bm25Scorer.cs:395-444).
            int nscan = 0;
#ifndef IFX_EMU
            const bool fast = per_thread == 8 && j1 - j0 == 8;      // the 8 owned slots live in registers for the whole term
            // A term whose own bound plus the bounds of the terms after it already exceeds the threshold can never be skipped
            // (scores are >= 0 and float addition is monotone), so its matches are exactly the non-zero tf slots, known before any
            // score exists: the ranks of all such terms of the tile come from ONE packed block scan (16-bit fields, <= 4096 each).
            unsigned uns = 0, act = 0, ex01 = 0, ex23 = 0, ex45 = 0, m01 = 0, m23 = 0, m45 = 0;
            {   // lane tt classifies term tt of the tile; the votes give every thread the (block-uniform) masks
                bool a = false, u = false;
                if (c.lane() < tile) { const TermS& tm = sh.terms[t0 + c.lane()]; a = tm.idf > 0.f && tm.s1 != tm.s0; u = a && !((0.f + tm.max_score) + tm.suffix_after <= thr); }
                act = __ballot_sync(0xffffffffu, a); uns = __ballot_sync(0xffffffffu, u);
            }
            if (__popc(uns) >= 2) {
                unsigned pk[3] = {0u, 0u, 0u};
#pragma unroll
                for (int tt = 0; tt < S1_TILE; tt++) if ((uns >> tt) & 1u) {
                    const uint8_t* tfb = sh.tfm[tt]; unsigned n = 0;
                    if (fast) { unsigned long long v = *reinterpret_cast<const unsigned long long*>(tfb + j0); v |= v >> 4; v |= v >> 2; v |= v >> 1; n = (unsigned)__popcll(v & 0x0101010101010101ULL); }
                    else for (int j = j0; j < j1; j++) n += tfb[j] != 0;
                    pk[tt >> 1] |= n << (16 * (tt & 1));
                }
                unsigned in0 = pk[0], in1 = pk[1], in2 = pk[2];
                for (int d = 1; d < 32; d <<= 1) {
                    unsigned o0 = __shfl_up_sync(0xffffffffu, in0, d), o1 = __shfl_up_sync(0xffffffffu, in1, d), o2 = __shfl_up_sync(0xffffffffu, in2, d);
                    if (c.lane() >= d) { in0 += o0; in1 += o1; in2 += o2; }
                }
                if (c.lane() == 31) { sh.scan3[c.warp()][0] = in0; sh.scan3[c.warp()][1] = in1; sh.scan3[c.warp()][2] = in2; }
                c.sync();
                // lane i holds warp i's totals; a warp scan over them yields this warp's base (lane warp-1) and the block totals (lane NW-1)
                uint4 x = make_uint4(0u, 0u, 0u, 0u); if (c.lane() < NW) x = *reinterpret_cast<const uint4*>(sh.scan3[c.lane()]);
                for (int d = 1; d < NW; d <<= 1) {
                    unsigned o0 = __shfl_up_sync(0xffffffffu, x.x, d), o1 = __shfl_up_sync(0xffffffffu, x.y, d), o2 = __shfl_up_sync(0xffffffffu, x.z, d);
                    if (c.lane() >= d) { x.x += o0; x.y += o1; x.z += o2; }
                }
                const int src = c.warp() > 0 ? c.warp() - 1 : 0;
                unsigned b0 = __shfl_sync(0xffffffffu, x.x, src), b1 = __shfl_sync(0xffffffffu, x.y, src), b2 = __shfl_sync(0xffffffffu, x.z, src);
                if (c.warp() == 0) { b0 = 0; b1 = 0; b2 = 0; }
                m01 = __shfl_sync(0xffffffffu, x.x, NW - 1); m23 = __shfl_sync(0xffffffffu, x.y, NW - 1); m45 = __shfl_sync(0xffffffffu, x.z, NW - 1);
                ex01 = b0 + in0 - pk[0]; ex23 = b1 + in1 - pk[1]; ex45 = b2 + in2 - pk[2];
            } else uns = 0;
#else
            const bool fast = false; const unsigned uns = 0; unsigned act = 0;
            for (int tt = 0; tt < tile; tt++) { const TermS& tm = sh.terms[t0 + tt]; if (tm.idf > 0.f && tm.s1 != tm.s0) act |= 1u << tt; }
#endif
            for (int tt = 0; tt < tile; tt++) {
                if (!((act >> tt) & 1u)) continue;                 // no idf or no postings inside this chunk's id range
                const TermS& tm = sh.terms[t0 + tt];
                uint8_t* tfb = sh.tfm[tt];
                int mine = 0; const float tbound = tm.max_score; const float tsuffix = tm.suffix_after;
                const bool ranked = (uns >> tt) & 1u;              // uniform: rank and match count already known, every non-zero tf is a match
                // NOTE: the block scan below contains a barrier and full-mask shuffles, so it sits at ONE call site under a block-uniform
                // condition; only the per-thread counting / accumulation around it may diverge.
#ifndef IFX_EMU
                unsigned long long tf8 = 0ULL; unsigned alive = 0; float sc8[8];
                if (fast) {
                    tf8 = *reinterpret_cast<const unsigned long long*>(tfb + j0);
                    if (tf8 != 0ULL) {
                        float4 sa = *reinterpret_cast<const float4*>(&sh.score[j0]), sb = *reinterpret_cast<const float4*>(&sh.score[j0 + 4]);
                        sc8[0] = sa.x; sc8[1] = sa.y; sc8[2] = sa.z; sc8[3] = sa.w; sc8[4] = sb.x; sc8[5] = sb.y; sc8[6] = sb.z; sc8[7] = sb.w;
#pragma unroll
                        for (int k = 0; k < 8; k++) { unsigned tfv = (unsigned)(tf8 >> (8 * k)) & 0xFFu; if (tfv != 0 && (ranked || !(sc8[k] + tbound + tsuffix <= thr))) alive |= 1u << k; }
                        mine = __popc(alive);
                    }
                } else
#endif
                { for (int j = j0; j < j1; j++) if (tfb[j] != 0 && (ranked || !(sh.score[j] + tbound + tsuffix <= thr))) mine++; }
                int m, rank;
#ifndef IFX_EMU
                if (ranked) { const unsigned e = tt < 2 ? ex01 : (tt < 4 ? ex23 : ex45), tm_ = tt < 2 ? m01 : (tt < 4 ? m23 : m45); rank = (int)((e >> (16 * (tt & 1))) & 0xFFFFu); m = (int)((tm_ >> (16 * (tt & 1))) & 0xFFFFu); }
                else
#endif
                { rank = block_excl_scan_1b(c, mine, sh.scan2[nscan & 1], m); nscan++; }
                const int vec_end = m - (m & 7);
#ifndef IFX_EMU
                if (fast) {
                    if (alive) {
                        float4 da = *reinterpret_cast<const float4*>(&sh.nv_s[j0]), db = *reinterpret_cast<const float4*>(&sh.nv_s[j0 + 4]);
                        float nv8[8] = {da.x, da.y, da.z, da.w, db.x, db.y, db.z, db.w};
#pragma unroll
                        for (int k = 0; k < 8; k++) if (alive & (1u << k)) {
                            float tf = (float)((unsigned)(tf8 >> (8 * k)) & 0xFFu);
                                                        float add = rank < vec_end ? bm25_from_norm_vector(tf, nv8[k], tm.idf) : bm25_scalar(tf, ix.doc_len[sh.cand_s[j0 + k]], avgdl, tm.idf);
                            sc8[k] += add; rank++;
                        }
                        *reinterpret_cast<float4*>(&sh.score[j0]) = make_float4(sc8[0], sc8[1], sc8[2], sc8[3]);
                        *reinterpret_cast<float4*>(&sh.score[j0 + 4]) = make_float4(sc8[4], sc8[5], sc8[6], sc8[7]);
                    }
                    if (tf8 != 0ULL) *reinterpret_cast<unsigned long long*>(tfb + j0) = 0ULL;
                } else
#endif
                for (int j = j0; j < j1; j++) {
                    const uint8_t tfv = tfb[j];
                    if (tfv != 0) {
                        if (ranked || !(sh.score[j] + tbound + tsuffix <= thr)) {
                            float tf = (float)tfv;
                            float sc = rank < vec_end ? bm25_from_norm_vector(tf, sh.nv_s[j], tm.idf) : bm25_scalar(tf, ix.doc_len[sh.cand_s[j]], avgdl, tm.idf);
                            sh.score[j] += sc; rank++;
                        }
                        tfb[j] = 0;
                    }
                }
                // no further barrier: score[j] / tfm[.][j] of these slots are private to this thread throughout the tile
            }
            c.sync();                                    // tile buffers are rewritten by arbitrary threads in the next tile
            IFX_TICK(3);   // phase B (ranks + accumulation)
        }
        {   // flush, part 1 (Bm25Scorer.cs:316-329): eligibility in parallel (the threshold only rises during a flush, so anything not
            // above the chunk-start threshold can never enter), survivors compacted in candidate order. Part 2 -- the exact sequential
            // emulation of .NET's 4-ary PriorityQueue over the survivors -- is `drain`, deferred into the next iteration.
            // (A set-based top-K was tried: it is only equivalent when no documents tied at the final threshold straddle the cut, and
            // on real corpora such ties are the norm -- identical tf pattern and length -- so the heap layout, which decides which of
            // them survive, has to be reproduced.)
            const bool full = sh.heap_size >= K;
            for (int r = 0; r < rounds; r++) {
                int j = r * NT + c.tid();
                bool e = j < cnt && sh.score[j] > 0.f && (!full || sh.score[j] > thr) && !ix.deleted[sh.cand_s[j]];
                unsigned b = c.ballot(e);
                if (c.lane() == 0) sh.ballots[0][r * NW + c.warp()] = b;
            }
            c.sync();
            const int slots = rounds * NW;
            if (c.warp() == 0) {     // exclusive prefix of the ballot popcounts (slot order == candidate order)
                int per = (slots + Ctx::WS - 1) / Ctx::WS; int s0 = c.lane() * per < slots ? c.lane() * per : slots, s1 = s0 + per < slots ? s0 + per : slots; int mine = 0;
                for (int sl = s0; sl < s1; sl++) mine += popc(sh.ballots[0][sl]);
                int incl = mine;
                for (int d = 1; d < Ctx::WS; d <<= 1) { int o = c.shfl(incl, c.lane() >= d ? c.lane() - d : 0); if (c.lane() >= d) incl += o; }
                int run = incl - mine;
                for (int sl = s0; sl < s1; sl++) { sh.bprefix[sl] = run; run += popc(sh.ballots[0][sl]); }
                if (c.lane() == Ctx::WS - 1) sh.bcast[6] = incl;
            }
            c.sync();
            const int n_surv = sh.bcast[6];
            unsigned long long* dst = n_surv <= SURV_CAP ? sh.surv : ws.surv_g;     // the rare big sets (heap still filling) go through global memory
            for (int r = 0; r < rounds; r++) {
                int j = r * NT + c.tid(); int sl = r * NW + c.warp(); unsigned bm = sh.ballots[0][sl];
                if ((bm >> c.lane()) & 1u) dst[sh.bprefix[sl] + popc(bm & c.lanemask_lt())] = kv_pack(sh.cand_s[j], sh.score[j]);
            }
            pend = n_surv;
        }
        c.sync();
        IFX_TICK(1);   // eligibility + compaction (+ in-place drain while the heap fills)
        pos += cnt;
    }
    if (c.tid() == 0 && pend) drain();
    c.sync();
#if !defined(IFX_EMU) && defined(IFX_S1_TIMERS)
    if (c.tid() == 0 && out.dbg) for (int k = 0; k < 6; k++) out.dbg[6 + k] = tph[k];
    if (c.tid() == hw && out.dbg) for (int k = 0; k < 8; k++) out.dbg[12 + k] = wph[k];
#endif
    for (int i = c.tid(); i < MAX_CONTAINERS; i += NT) sh.dirty[i] = 0;   // `cbits` aliased the dirty flags during scoring
    c.sync();
    // ---- PopulateResultHeapFromPruning + TopKHeap.GetTopK + ConsolidateSegments: order by (score desc, key asc)
    const int n = sh.heap_size; int n2 = 1; while (n2 < n) n2 <<= 1;
    float* ks = sh.score; int32_t* kd = sh.cand_s;            // reuse chunk arrays (CHUNK >= MAX_K)
    for (int i = c.tid(); i < n2; i += NT) { if (i < n) { ks[i] = sh.IFX_HP(i); kd[i] = sh.IFX_HD(i); } else { ks[i] = -1.f; kd[i] = 0x7fffffff; } }
    c.sync();
    auto before = [&](int a, int b) -> bool {   // a ranks before b
        if (ks[a] != ks[b]) return ks[a] > ks[b];
        if (kd[a] == 0x7fffffff || kd[b] == 0x7fffffff) return kd[a] < kd[b];
        return ix.doc_key[kd[a]] < ix.doc_key[kd[b]];
    };
    for (int k = 2; k <= n2; k <<= 1) for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = c.tid(); i < n2; i += NT) { int l = i ^ j; if (l > i) { bool up = (i & k) == 0; bool sw = up ? before(l, i) : before(i, l); if (sw) { float x = ks[i]; ks[i] = ks[l]; ks[l] = x; int y = kd[i]; kd[i] = kd[l]; kd[l] = y; } } }
        c.sync();
    }
    for (int i = c.tid(); i < n; i += NT) { out.doc[i] = kd[i]; out.score[i] = ks[i]; out.key[i] = ix.doc_key[kd[i]]; }
    if (c.tid() == 0) out.n[0] = n;
    c.sync();
}

}  // namespace ifx
