// infidex_b200 -- implementation of the C-ABI (include/infidex_gpu.h).
// Included by ifx_api.cu (CUDA product) and by tests/emu/emu_api.cpp (IFX_EMU host emulation of the kernels, tests only).
#include "../../include/infidex_gpu.h"
#include "ifx_stage1_score.h"
#include "ifx_short.h"
#include "ifx_stage2.h"
#include "ifx_build.h"
#include <thread>
#include <functional>
#include <vector>
#include <string>
#include <algorithm>
#include <numeric>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cmath>
#include <mutex>
#include <unordered_map>

namespace {
#include "chartables.inc"
}

using namespace ifx;

static thread_local std::string g_err;
static int fail(int code, const std::string& m) { g_err = m; return code; }

// ---- device memory / launch shims ---------------------------------------------------------------------------------
#ifdef IFX_EMU
static bool dev_ok() { return true; }
static void* dev_alloc(size_t n) { return calloc(n ? n : 1, 1); }
static void dev_free(void* p) { free(p); }
static void h2d(void* d, const void* s, size_t n) { if (n) memcpy(d, s, n); }
static void d2h(void* d, const void* s, size_t n) { if (n) memcpy(d, s, n); }
static void dev_zero(void* d, size_t n) { if (n) memset(d, 0, n); }
struct Timer { void start() {} float stop() { return 0.f; } };
struct DeviceGuard { explicit DeviceGuard(int) {} };
#else
#define CUDA_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) throw std::string(#x ": ") + cudaGetErrorString(e_); } while (0)
static bool dev_ok() { int n = 0; return cudaGetDeviceCount(&n) == cudaSuccess && n > 0; }
static void* dev_alloc(size_t n) { void* p = nullptr; CUDA_TRY(cudaMalloc(&p, n ? n : 1)); return p; }
static void dev_free(void* p) { if (p) cudaFree(p); }
static void h2d(void* d, const void* s, size_t n) { if (n) CUDA_TRY(cudaMemcpy(d, s, n, cudaMemcpyHostToDevice)); }
static void d2h(void* d, const void* s, size_t n) { if (n) CUDA_TRY(cudaMemcpy(d, s, n, cudaMemcpyDeviceToHost)); }
static void dev_zero(void* d, size_t n) { if (n) CUDA_TRY(cudaMemset(d, 0, n)); }
struct Timer { cudaEvent_t a = nullptr, b = nullptr; cudaStream_t s = 0;
    void start() { if (!a) { cudaEventCreate(&a); cudaEventCreate(&b); } cudaEventRecord(a, s); }
    float stop() { cudaEventRecord(b, s); cudaEventSynchronize(b); float ms = 0; cudaEventElapsedTime(&ms, a, b); return ms; } };
// The CUDA current device is per host thread: every entry point that touches the device selects the index's device for its own
// duration (callers may come from any thread of the C# pool) and restores the caller's.
struct DeviceGuard { int prev = -1; explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) CUDA_TRY(cudaSetDevice(dev)); else prev = -1; } ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); } };
#ifndef IFX_EXPAND_THREADS
#define IFX_EXPAND_THREADS 512
#endif
#ifndef IFX_S1_THREADS
#define IFX_S1_THREADS 512
#endif
#endif

// ---- host-side handle ----------------------------------------------------------------------------------------------

struct ifx_index {
    DevIndex v{};                       // device pointers
    std::vector<void*> allocs;
    uint8_t* d_sorted_len = nullptr;
    int n_ctas = 1, n_ctas_sel = 1;       // resident CTAs of the 512-thread persistent kernels / of the selection + lookup kernel (one workspace each)
    std::vector<S1Workspace> ws; S1Workspace* d_ws = nullptr;
    int32_t* d_pool = nullptr; unsigned long long pool_cap = 0;
    unsigned char* d_spool = nullptr; unsigned long long spool_cap = 0;   // Stage-1 staging pool of a batch (candidates, lengths, chunk tables, tf matrices)
    int max_batch = 16384;
    SqScratch sq{}; std::vector<uint16_t> h_champ_chars; std::vector<int32_t> h_champ_off; bool attr_sq = false;     // short-query path: lazily allocated scratch, host copy of the champion directory
    int device = 0; uint8_t* d_flush = nullptr; bool attr_s1 = false, attr_s2 = false;   // kernel attributes are per device: tracked per index (guarded by mu)
    std::vector<FilterProg> h_filters; FilterProg* d_filters = nullptr;
    std::vector<Column> h_columns; std::vector<std::u16string> column_names;
    std::mutex mu; std::mutex call_mu; struct ifx_batch* cached = nullptr;   // batch workspace reused by ifx_search_batch
    ~ifx_index();
    size_t bytes = 0;                   // device bytes held by this index (reported by IFX_CREATE_TIMING)
    template <class T> T* up(const T* src, size_t n) { T* d = (T*)dev_alloc(n * sizeof(T)); allocs.push_back(d); bytes += n * sizeof(T); if (src && n) h2d(d, src, n * sizeof(T)); return d; }
    template <class T> T* alloc(size_t n) { T* d = (T*)dev_alloc(n * sizeof(T)); allocs.push_back(d); bytes += n * sizeof(T); return d; }
};

struct ifx_batch {
    ifx_index* idx = nullptr; int nq = 0; int depth_max = 0; int cap_max = 0; size_t text_cap = 0;
    int s1_stride = 0;                  // row length of the Stage-1 list arrays: max(depth, max_results) -- the single-character short-query path returns max_results entries whatever the depth
    std::vector<void*> allocs;
    uint16_t* d_text = nullptr; int64_t* d_off = nullptr; int32_t* d_par = nullptr;   // par: [nq][5] max_results, depth, enable_cov, filter_id, enable_facets
    QueryPlan* d_plans = nullptr; FuzzyItem* d_items = nullptr; BatchCounters* d_bc = nullptr; int* d_work = nullptr;
    int32_t* d_short_kind = nullptr; int32_t* d_s1_total = nullptr; bool use_gcnt = false; int32_t* d_sel_cnt = nullptr; int32_t* d_sel_done = nullptr; S1Rec* d_recs = nullptr; int32_t* d_light = nullptr; int32_t* d_mid = nullptr; int32_t* d_heavy = nullptr;
    int* d_order = nullptr; long long* d_qdbg = nullptr;
    int64_t* d_s1_key = nullptr; int32_t* d_s1_doc = nullptr; float* d_s1_score = nullptr; int32_t* d_s1_n = nullptr;
    Stage2Buffers s2{};                 // WordMatcher + coverage outputs
    FinalOut fin{}; int fcap = 0; int32_t* d_g_di = nullptr; int32_t* d_shard_info = nullptr; int64_t* d_shard_dkey = nullptr;
    bool ran = false;
    ~ifx_batch() { for (void* p : allocs) dev_free(p); }
    template <class T> T* alloc(size_t n) { T* d = (T*)dev_alloc(n * sizeof(T)); allocs.push_back(d); return d; }
};

ifx_index::~ifx_index() { delete cached; for (void* p : allocs) dev_free(p); }

static uint64_t hash_host(const uint16_t* s, int n) {
    uint64_t h = 0xcbf29ce484222325ULL ^ (uint64_t)n;
    for (int i = 0; i < n; i++) { h ^= s[i]; h *= 0x100000001b3ULL; }
    h ^= h >> 29; h *= 0xbf58476d1ce4e5b9ULL; h ^= h >> 32;
    return h | 1ULL;
}

// `with_hash = false`: a plain indexed string array (never looked up by key; its entries may repeat, which would degrade the
// open-addressing build to quadratic probing)
struct HostHash { std::vector<uint64_t> hk; std::vector<int32_t> hv; uint32_t cap = 0; };
static void build_hash(const ifx_strings& s, HostHash& H) {      // host only: the big dictionaries are hashed on parallel host threads, then uploaded
    uint32_t cap = 8; while (cap < (uint32_t)s.n * 2u + 2u) cap <<= 1;
    H.cap = cap; H.hk.assign(cap, 0); H.hv.assign(cap, -1);
    for (int i = 0; i < s.n; i++) {
        uint64_t h = hash_host(s.chars + s.off[i], (int)(s.off[i + 1] - s.off[i])); uint32_t slot = (uint32_t)(h >> 7) & (cap - 1);
        while (H.hk[slot] != 0) slot = (slot + 1) & (cap - 1);
        H.hk[slot] = h; H.hv[slot] = i;
    }
}
static StrDict upload_dict(ifx_index* ix, const ifx_strings& s, bool with_hash = true, const HostHash* pre = nullptr) {
    StrDict d{}; d.n = s.n;
    size_t nchars = s.n ? s.off[s.n] : 0;
    d.chars = ix->up(s.chars, nchars ? nchars : 1); d.off = ix->up(s.off, (size_t)s.n + 1);
    if (!with_hash) { static const uint64_t z64 = 0; static const int32_t z32 = 0; d.hkeys = ix->up(&z64, 1); d.hvals = ix->up(&z32, 1); d.hmask = 0; return d; }
    HostHash local; if (!pre) { build_hash(s, local); pre = &local; }
    d.hkeys = ix->up(pre->hk.data(), pre->cap); d.hvals = ix->up(pre->hv.data(), pre->cap); d.hmask = pre->cap - 1;
    return d;
}
static DocsetDict upload_docset(ifx_index* ix, const ifx_docset_dict& s, const HostHash* pre = nullptr) {
    DocsetDict d{}; d.keys = upload_dict(ix, s.keys, true, pre);
    size_t np = s.keys.n ? (size_t)s.row_ptr[s.keys.n] : 0;
    static const int64_t zero = 0;
    d.row_ptr = ix->up(s.keys.n ? s.row_ptr : &zero, (size_t)s.keys.n + 1); d.doc_id = ix->up(s.doc_id, np ? np : 1);
    return d;
}

static bool host_try_parse_double(const uint16_t* s, int n, double& out);   // defined with the filter code below

extern "C" void ifx_params_default(ifx_params* p) { p->stop_term_limit = 1250000; p->device = 0; p->max_batch = 16384; p->reserved = 0; }
extern "C" const char* ifx_last_error(void) { return g_err.c_str(); }
extern "C" int ifx_device_count(void) {
#ifdef IFX_EMU
    return 1;
#else
    int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) return 0; return n;
#endif
}

#ifndef IFX_EMU
// document token table (ifx_cov.h): one thread per document, count pass then fill pass
__global__ void __launch_bounds__(256) k_tok_count(DevIndex ix, unsigned* cnt) { const int d = blockIdx.x * blockDim.x + threadIdx.x; if (d < ix.n_docs) cnt[d] = (unsigned)doc_tokens_emit(ix, d, nullptr); }
__global__ void __launch_bounds__(256) k_tok_fill(DevIndex ix, const int64_t* ptr, uint32_t* tab) { const int d = blockIdx.x * blockDim.x + threadIdx.x; if (d < ix.n_docs) doc_tokens_emit(ix, d, tab + ptr[d]); }
#endif

extern "C" int ifx_index_create(const ifx_index_image* img, const ifx_params* pp, ifx_index** out) {
    if (!img || !out) return fail(IFX_ERR_INVALID, "null argument");
    if (!dev_ok()) return fail(IFX_ERR_NO_DEVICE, "no CUDA device available (infidex_b200 has no CPU fallback)");
    ifx_params P; if (pp) P = *pp; else ifx_params_default(&P);
    if (img->n_docs < 0 || ((int64_t)img->n_docs + 65535) / 65536 > MAX_CONTAINERS) return fail(IFX_ERR_INVALID, "n_docs out of range");
    ifx_index* ix = new ifx_index();
    const bool timing = getenv("IFX_CREATE_TIMING") != nullptr; auto t_last = std::chrono::steady_clock::now();
    auto stage = [&](const char* what) { if (!timing) return; auto now = std::chrono::steady_clock::now(); fprintf(stderr, "[ifx_index_create] %-28s %.2f s\n", what, std::chrono::duration<double>(now - t_last).count()); t_last = now; };
    try {
#ifndef IFX_EMU
        CUDA_TRY(cudaSetDevice(P.device)); ix->device = P.device;
#endif
        DevIndex& v = ix->v; const int N = img->n_docs; const int T = img->terms.n;
        // ---- host-side preparation of the dictionaries on parallel threads (hash tables, the two sort orders of the term and the
        //      affix dictionaries) while this thread uploads the bulk arrays and derives the posting-side structures on the device
        HostHash hh_terms, hh_words, hh_prefix, hh_exact, hh_ld1;
        std::vector<int32_t> order(T); std::vector<uint8_t> slen(T); std::vector<unsigned long long> sig(std::max(T, 1), 0ULL);
        std::vector<int32_t> lp(257, 0), lord(std::max(T, 1), 0); std::vector<unsigned long long> lsig(std::max(T, 1), 0ULL);
        auto term_sv = [&](int i) { return std::u16string_view((const char16_t*)img->terms.chars + img->terms.off[i], img->terms.off[i + 1] - img->terms.off[i]); };
        const int A = img->affix_words.n;
        std::vector<uint16_t> achars; std::vector<uint32_t> aoff(A + 1, 0); std::vector<int32_t> fdoc(std::max(A, 1)), ro(std::max(A, 1)), rdoc(std::max(A, 1)); HostHash hh_affix;
        std::vector<std::function<void()>> jobs;
        jobs.push_back([&] { build_hash(img->terms, hh_terms); });
        jobs.push_back([&] { build_hash(img->words, hh_words); });
        jobs.push_back([&] { build_hash(img->prefix.keys, hh_prefix); });
        jobs.push_back([&] { build_hash(img->wm_exact.keys, hh_exact); });
        jobs.push_back([&] { build_hash(img->wm_ld1.keys, hh_ld1); });
        jobs.push_back([&] {   // trie DFS order == ordinal-lexicographic order of the term texts (FstBuilder.CompactTrie sorts arcs by label)
            std::iota(order.begin(), order.end(), 0);
            std::sort(order.begin(), order.end(), [&](int x, int y) { return term_sv(x) < term_sv(y); });
            for (int i = 0; i < T; i++) { auto sv = term_sv(order[i]); size_t l = sv.size(); slen[i] = (uint8_t)(l > 255 ? 255 : l); unsigned long long g = 0; for (char16_t ch : sv) g |= 1ULL << (((uint16_t)ch * 0x9E37u >> 4) & 63); sig[i] = g; }
            for (int i = 0; i < T; i++) lp[slen[i] + 1]++; for (int l = 0; l < 256; l++) lp[l + 1] += lp[l];       // counting sort of the sorted positions by length
            std::vector<int32_t> cur(lp.begin(), lp.end() - 1);
            for (int i = 0; i < T; i++) { int at = cur[slen[i]]++; lord[at] = order[i]; lsig[at] = sig[i]; } });
        jobs.push_back([&] {   // affix words: forward (prefix) order and reverse-string (suffix) order, each with the doc its trie output resolves to
            auto aw = [&](int i) { return std::u16string_view((const char16_t*)img->affix_words.chars + img->affix_words.off[i], img->affix_words.off[i + 1] - img->affix_words.off[i]); };
            std::vector<int32_t> fo(A); std::iota(fo.begin(), fo.end(), 0); std::sort(fo.begin(), fo.end(), [&](int x, int y) { return aw(x) < aw(y); });
            for (int i = 0; i < A; i++) { auto w = aw(fo[i]); achars.insert(achars.end(), w.begin(), w.end()); aoff[i + 1] = (uint32_t)achars.size(); fdoc[i] = img->affix_last_doc[fo[i]]; }
            if (achars.empty()) achars.push_back(0);
            auto rev_less = [&](int x2, int y2) { auto x = std::u16string_view((const char16_t*)achars.data() + aoff[x2], aoff[x2 + 1] - aoff[x2]), y = std::u16string_view((const char16_t*)achars.data() + aoff[y2], aoff[y2 + 1] - aoff[y2]);
                size_t n = std::min(x.size(), y.size()); for (size_t k = 0; k < n; k++) { char16_t p = x[x.size() - 1 - k], q = y[y.size() - 1 - k]; if (p != q) return p < q; } return x.size() < y.size(); };
            std::iota(ro.begin(), ro.begin() + A, 0); std::sort(ro.begin(), ro.begin() + A, rev_less);
            for (int i = 0; i < A; i++) rdoc[i] = fdoc[ro[i]];
            ifx_strings ss{achars.data(), aoff.data(), A}; build_hash(ss, hh_affix); });
        std::vector<std::thread> workers; for (auto& j : jobs) workers.emplace_back(j);
        struct Joiner { std::vector<std::thread>& w; ~Joiner() { for (auto& t : w) if (t.joinable()) t.join(); } } joiner{workers};

        v.n_docs = N; v.n_live = img->n_live; v.avgdl = img->avgdl; v.stop_term_limit = P.stop_term_limit;
        v.doc_key = ix->up(img->doc_key, N); v.deleted = ix->up(img->deleted, N); v.doc_len = ix->up(img->doc_len, N);
        {   // duplicate DocumentKeys (segments of one document): first live document per key
            std::vector<int64_t> ks(img->doc_key, img->doc_key + N); std::sort(ks.begin(), ks.end()); const bool dup = std::adjacent_find(ks.begin(), ks.end()) != ks.end(); v.key_first = nullptr;
            if (dup) { std::unordered_map<int64_t, int32_t> first; first.reserve((size_t)N * 2); std::vector<int32_t> kf(N);
                for (int d = 0; d < N; d++) if (!img->deleted[d]) first.emplace(img->doc_key[d], d);
                for (int d = 0; d < N; d++) { auto it = first.find(img->doc_key[d]); kf[d] = it == first.end() ? d : it->second; }
                v.key_first = ix->up(kf.data(), N); } }
        size_t ntext = N ? (size_t)img->text_off[N] : 0;
        v.text = ix->up(img->text_chars, ntext ? ntext : 1); v.text_off = ix->up(img->text_off, (size_t)N + 1);
        v.first_token = upload_dict(ix, img->first_token, false); v.token_count = ix->up(img->token_count, N);
        stage("docs + text");
        v.df = ix->up(img->df, T ? T : 1); v.row_ptr = ix->up(img->row_ptr, (size_t)T + 1);
        size_t P_ = T ? (size_t)img->row_ptr[T] : 0;
        v.post_doc = ix->up(img->post_doc, P_ ? P_ : 1); v.post_tf = ix->up(img->post_tf, P_ ? P_ : 1);
        stage("postings");
        const int ncont = (N + 65535) >> 16; const int64_t SKIP_MIN = 512; v.n_cont = ncont;
        const int bw = (N + 31) / 32; v.bm_words = bw; const int64_t dense = std::max<int64_t>(1024, N / 32);
        std::vector<int32_t> sid(std::max(T, 1), -1), bid(std::max(T, 1), -1), skip_terms, bm_terms;
        for (int t = 0; t < T; t++) { const int64_t len = img->row_ptr[t + 1] - img->row_ptr[t]; if (len >= SKIP_MIN) { sid[t] = (int)skip_terms.size(); skip_terms.push_back(t); } if (len >= dense) { bid[t] = (int)bm_terms.size(); bm_terms.push_back(t); } }
        const int ns = (int)skip_terms.size(), nb = (int)bm_terms.size();
        v.skip_id = ix->up(sid.data(), sid.size()); v.bm_id = ix->up(bid.data(), bid.size());
#ifdef IFX_EMU
        {   // forward index: CSR transpose of the live posting lists (counting sort by doc; entries of a doc end up in term order)
            std::vector<int64_t> fp((size_t)N + 2, 0);
            for (int t = 0; t < T; t++) { if (img->df[t] <= 0) continue; for (int64_t i = img->row_ptr[t]; i < img->row_ptr[t + 1]; i++) fp[(size_t)img->post_doc[i] + 2]++; }
            for (int d = 0; d < N; d++) fp[(size_t)d + 2] += fp[(size_t)d + 1];
            const int64_t FP = fp[(size_t)N + 1]; std::vector<int32_t> ft((size_t)std::max<int64_t>(FP, 1)); std::vector<uint8_t> fw((size_t)std::max<int64_t>(FP, 1));
            for (int t = 0; t < T; t++) { if (img->df[t] <= 0) continue; for (int64_t i = img->row_ptr[t]; i < img->row_ptr[t + 1]; i++) { int64_t at = fp[(size_t)img->post_doc[i] + 1]++; ft[(size_t)at] = t; fw[(size_t)at] = img->post_tf[i]; } }
            v.fwd_ptr = ix->up(fp.data(), (size_t)N + 1); v.fwd_term = ix->up(ft.data(), ft.size()); v.fwd_tf = ix->up(fw.data(), fw.size());
            v.fwd_avg_bytes = (int32_t)(N > 0 ? 5 * FP / N + 16 : 16);
        }
        {   // container skip table for long posting lists: turns the per-chunk sub-range search of the scorer into a lookup
            std::vector<int32_t> sp((size_t)std::max(ns, 0) * (ncont + 1));
            for (int r = 0; r < ns; r++) { const int t = skip_terms[r]; int64_t r0 = img->row_ptr[t], len = img->row_ptr[t + 1] - r0; int64_t i = 0;
                for (int c = 0; c <= ncont; c++) { int64_t lim = (int64_t)c << 16; while (i < len && img->post_doc[r0 + i] < lim) i++; sp[(size_t)r * (ncont + 1) + c] = (int32_t)i; } }
            if (sp.empty()) sp.push_back(0);
            v.skip_ptr = ix->up(sp.data(), sp.size());
        }
        {   // dense terms additionally get a membership bitmap + rank directory: O(1) probes (doc -> posting index -> tf) in the scorer
            std::vector<unsigned> bits((size_t)std::max(nb, 1) * std::max(bw, 1), 0u); std::vector<int32_t> rank((size_t)std::max(nb, 1) * std::max(bw, 1), 0);
            for (int k = 0; k < nb; k++) { const int t = bm_terms[k]; unsigned* bb = bits.data() + (size_t)k * bw; int32_t* r = rank.data() + (size_t)k * bw;
                for (int64_t i = img->row_ptr[t]; i < img->row_ptr[t + 1]; i++) { int d = img->post_doc[i]; bb[d >> 5] |= 1u << (d & 31); }
                int run = 0; for (int w = 0; w < bw; w++) { r[w] = run; run += __builtin_popcount(bb[w]); } }
            v.bm_bits = ix->up(bits.data(), bits.size()); v.bm_rank = ix->up(rank.data(), rank.size());
        }
#else
        {   // the same three structures derived on the device from the uploaded CSR (csrc/ifx_build.h)
            const int sms = [&] { cudaDeviceProp prop; CUDA_TRY(cudaGetDeviceProperties(&prop, P.device)); return prop.multiProcessorCount; }();
            const int grid = sms * 8;
            unsigned* cnt = (unsigned*)dev_alloc(((size_t)N + 1) * 4); dev_zero(cnt, ((size_t)N + 1) * 4);
            if (T > 0) k_fwd_count<<<grid, 256>>>(v.row_ptr, v.df, T, v.post_doc, cnt);
            const int64_t n_tiles = ((int64_t)N + SCAN_TILE - 1) / SCAN_TILE;
            unsigned long long* tile = (unsigned long long*)dev_alloc((size_t)std::max<int64_t>(n_tiles, 1) * 8);
            int64_t* fwd_ptr = ix->alloc<int64_t>((size_t)N + 1);
            if (N > 0) { k_scan_tile_sums<<<(unsigned)n_tiles, SCAN_THREADS>>>(cnt, N, tile); k_scan_tiles<<<1, 1024>>>(tile, n_tiles); k_scan_final<<<(unsigned)n_tiles, SCAN_THREADS>>>(cnt, N, tile, fwd_ptr); }
            else dev_zero(fwd_ptr, 8);
            int64_t FP = 0; d2h(&FP, fwd_ptr + N, 8); v.fwd_avg_bytes = (int32_t)(N > 0 ? 5 * FP / N + 16 : 16);
            int32_t* fwd_term = ix->alloc<int32_t>((size_t)std::max<int64_t>(FP, 1)); uint8_t* fwd_tf = ix->alloc<uint8_t>((size_t)std::max<int64_t>(FP, 1));
            dev_zero(cnt, ((size_t)N + 1) * 4);
            if (T > 0) k_fwd_scatter<<<grid, 256>>>(v.row_ptr, v.df, T, v.post_doc, v.post_tf, fwd_ptr, cnt, fwd_term, fwd_tf);
            v.fwd_ptr = fwd_ptr; v.fwd_term = fwd_term; v.fwd_tf = fwd_tf;
            int32_t* d_skip_terms = (int32_t*)dev_alloc((size_t)std::max(ns, 1) * 4); h2d(d_skip_terms, skip_terms.data(), (size_t)ns * 4);
            int32_t* skip_ptr = ix->alloc<int32_t>(std::max<size_t>((size_t)ns * (ncont + 1), 1));
            if (ns > 0) { const int64_t tot = (int64_t)ns * (ncont + 1); k_skip_table<<<(unsigned)((tot + 255) / 256), 256>>>(v.row_ptr, d_skip_terms, ns, ncont, v.post_doc, skip_ptr); }
            v.skip_ptr = skip_ptr;
            int32_t* d_bm_terms = (int32_t*)dev_alloc((size_t)std::max(nb, 1) * 4); h2d(d_bm_terms, bm_terms.data(), (size_t)nb * 4);
            const size_t bmn = (size_t)std::max(nb, 1) * std::max(bw, 1);
            unsigned* bm_bits = ix->alloc<unsigned>(bmn); int32_t* bm_rank = ix->alloc<int32_t>(bmn); dev_zero(bm_bits, bmn * 4);
            if (nb > 0) { k_bitmap_fill<<<nb, 512>>>(v.row_ptr, d_bm_terms, bw, v.post_doc, bm_bits); k_bitmap_rank<<<nb, 512>>>(bw, bm_bits, bm_rank); } else dev_zero(bm_rank, bmn * 4);
            v.bm_bits = bm_bits; v.bm_rank = bm_rank;
            CUDA_TRY(cudaGetLastError()); CUDA_TRY(cudaDeviceSynchronize());
            dev_free(cnt); dev_free(tile); dev_free(d_skip_terms); dev_free(d_bm_terms);
        }
#endif
        stage("forward index + skip tables + dense bitmaps");
        for (auto& t : workers) t.join();
        stage("waiting for the host dictionary threads");
        v.terms = upload_dict(ix, img->terms, true, &hh_terms);
        v.term_sorted = ix->up(order.data(), T ? T : 1); ix->d_sorted_len = ix->up(slen.data(), T ? T : 1);
        v.term_sig = ix->up(sig.data(), sig.size());
        v.len_ptr = ix->up(lp.data(), lp.size()); v.len_sig = ix->up(lsig.data(), lsig.size()); v.len_ord = ix->up(lord.data(), lord.size());
        v.words = upload_dict(ix, img->words, true, &hh_words); v.word_idf = ix->up(img->word_idf, img->words.n ? img->words.n : 1);
        v.prefix = upload_docset(ix, img->prefix, &hh_prefix); v.prefix_gcard = img->prefix_global_card ? ix->up(img->prefix_global_card, std::max(img->prefix.keys.n, 1)) : nullptr; v.wm_exact = upload_docset(ix, img->wm_exact, &hh_exact); v.wm_ld1 = upload_docset(ix, img->wm_ld1, &hh_ld1);
        {   ifx_strings ss{achars.data(), aoff.data(), A};
            v.affix = upload_dict(ix, ss, true, &hh_affix); v.affix_fwd_doc = ix->up(fdoc.data(), A ? A : 1);
            v.affix_rev = ix->up(ro.data(), A ? A : 1); v.affix_rev_doc = ix->up(rdoc.data(), A ? A : 1); }
        stage("dictionaries");
        {   // short-query structures
            const int nc = img->n_champ_chars; v.n_champ = nc; static const uint16_t z16 = 0; static const int32_t z32[2] = {0, 0}; static const float zf = 0.f; static const int64_t z64[2] = {0, 0}; static const int32_t m1 = -1;
            v.champ_chars = ix->up(nc ? img->champ_chars : &z16, std::max(nc, 1)); v.champ_off = ix->up(nc ? img->champ_off : z32, (size_t)nc + 1);
            const size_t ne = nc ? (size_t)img->champ_off[nc] : 0; v.champ_doc = ix->up(ne ? img->champ_doc : z32, std::max<size_t>(ne, 1)); v.champ_score = ix->up(ne ? img->champ_score : &zf, std::max<size_t>(ne, 1));
            if (nc) { ix->h_champ_chars.assign(img->champ_chars, img->champ_chars + nc); ix->h_champ_off.assign(img->champ_off, img->champ_off + nc + 1); }
            const int nr = img->n_raw; v.n_raw = nr; v.raw_doc = ix->up(nr ? img->raw_doc : &m1, std::max(nr, 1)); v.raw_off = ix->up(nr ? img->raw_off : z64, (size_t)nr + 1);
            const size_t rc = nr ? (size_t)img->raw_off[nr] : 0; v.raw_chars = ix->up(rc ? img->raw_chars : &z16, std::max<size_t>(rc, 1));
        }
        {   // character tables (tools/gen_chartables.py) + host-evaluated MathF.Log2(len + 1)
            std::vector<uint16_t> lo(65536), upv(65536); std::vector<uint8_t> fl(65536, 0);
            for (int i = 0; i < 65536; i++) lo[i] = upv[i] = (uint16_t)i;
            for (int i = 0; i < IFX_LOWER_PAIRS_N; i++) lo[IFX_LOWER_PAIRS[i][0]] = IFX_LOWER_PAIRS[i][1];
            for (int i = 0; i < IFX_UPPER_PAIRS_N; i++) upv[IFX_UPPER_PAIRS[i][0]] = IFX_UPPER_PAIRS[i][1];
            for (int i = 0; i < IFX_LETTER_RANGES_N; i++) for (int c = IFX_LETTER_RANGES[i][0]; c <= IFX_LETTER_RANGES[i][1]; c++) fl[c] |= 1;
            for (int i = 0; i < IFX_SPACE_LIST_N; i++) fl[IFX_SPACE_LIST[i]] |= 2;
            const uint16_t dl[] = {' ', '-', '/', '.', ',', ':', ';', '\'', '`', 0x2013, 0x2014, '*', '&', '\\', '_', '(', ')', '{', '}', '[', ']', '\t'};
            for (uint16_t c : dl) fl[c] |= 4;
            v.lower = ix->up(lo.data(), 65536); v.upper = ix->up(upv.data(), 65536); v.cflags = ix->up(fl.data(), 65536);
            for (int k = 0; k < 4; k++) v.delim_ascii[k] = 0u; for (int ch = 0; ch < 128; ch++) if (fl[ch] & 4) v.delim_ascii[ch >> 5] |= 1u << (ch & 31);
            std::vector<float> l2(1024); for (int i = 0; i < 1024; i++) l2[i] = std::log2((float)(i + 1));
            v.log2_len = ix->up(l2.data(), 1024);
            int tn = std::max(1, std::min(img->n_live, P.stop_term_limit)) + 1; std::vector<float> idf(tn, 0.f);
            for (int df = 1; df < tn; df++) { float d = (float)df, Nf = (float)img->n_live; float ratio = (Nf - d + 0.5f) / (d + 0.5f); idf[df] = ratio <= 0.f ? 0.f : std::log(ratio + 1.f); }
            v.idf_table = ix->up(idf.data(), tn); v.idf_table_n = tn;
        }
        {   // document token table: needs the text and the character tables above
#ifdef IFX_EMU
            std::vector<int64_t> tp((size_t)N + 1, 0); for (int d = 0; d < N; d++) tp[(size_t)d + 1] = tp[d] + doc_tokens_emit(v, d, nullptr);
            std::vector<uint32_t> tab((size_t)std::max<int64_t>(tp[N], 1)); for (int d = 0; d < N; d++) doc_tokens_emit(v, d, tab.data() + tp[d]);
            v.tok_ptr = ix->up(tp.data(), tp.size()); v.tok_tab = ix->up(tab.data(), tab.size());
#else
            unsigned* cnt = (unsigned*)dev_alloc(((size_t)N + 1) * 4); const unsigned g = (unsigned)((N + 255) / 256);
            if (N > 0) k_tok_count<<<g, 256>>>(v, cnt);
            const int64_t n_tiles = ((int64_t)N + SCAN_TILE - 1) / SCAN_TILE;
            unsigned long long* tile = (unsigned long long*)dev_alloc((size_t)std::max<int64_t>(n_tiles, 1) * 8);
            int64_t* tok_ptr = ix->alloc<int64_t>((size_t)N + 1);
            if (N > 0) { k_scan_tile_sums<<<(unsigned)n_tiles, SCAN_THREADS>>>(cnt, N, tile); k_scan_tiles<<<1, 1024>>>(tile, n_tiles); k_scan_final<<<(unsigned)n_tiles, SCAN_THREADS>>>(cnt, N, tile, tok_ptr); }
            else dev_zero(tok_ptr, 8);
            int64_t TW = 0; d2h(&TW, tok_ptr + N, 8);
            uint32_t* tab = ix->alloc<uint32_t>((size_t)std::max<int64_t>(TW, 1));
            if (N > 0) k_tok_fill<<<g, 256>>>(v, tok_ptr, tab);
            CUDA_TRY(cudaGetLastError()); CUDA_TRY(cudaDeviceSynchronize());
            dev_free(cnt); dev_free(tile);
            v.tok_ptr = tok_ptr; v.tok_tab = tab;
#endif
            stage("document token table");
        }
        {   // filter / facet columns: ToString() dictionary + parsed numeric view of every dictionary entry
            ix->h_columns.resize(img->n_columns);
            for (int c = 0; c < img->n_columns; c++) {
                const ifx_column& ic = img->columns[c]; Column& col = ix->h_columns[c];
                col.value_id = ix->up(ic.value_id, N); col.dict = upload_dict(ix, ic.dict); col.flags = ic.flags;
                std::vector<double> num(std::max(ic.dict.n, 1), 0.0); std::vector<uint8_t> isn(std::max(ic.dict.n, 1), 0);
                for (int i = 0; i < ic.dict.n; i++) { double d; if (host_try_parse_double(ic.dict.chars + ic.dict.off[i], (int)(ic.dict.off[i + 1] - ic.dict.off[i]), d)) { num[i] = d; isn[i] = 1; } }
                col.dict_num = ix->up(num.data(), num.size()); col.dict_is_num = ix->up(isn.data(), isn.size());
                col.name_const_hash_lo = (int32_t)hash_host(ic.name, ic.name_len);
            }
            v.n_columns = img->n_columns; v.columns = ix->up(ix->h_columns.data(), std::max(img->n_columns, 1));
            for (int c = 0; c < img->n_columns; c++) ix->column_names.emplace_back((const char16_t*)img->columns[c].name, (size_t)img->columns[c].name_len);
        }
        // persistent-CTA workspaces
        int64_t max_list = 1; for (int t = 0; t < T; t++) max_list = std::max<int64_t>(max_list, img->row_ptr[t + 1] - img->row_ptr[t]);
        max_list = std::max<int64_t>(max_list, std::min<int64_t>(N, P.stop_term_limit));
        int64_t nwords = ((int64_t)N + 31) / 32 + 2048;
        // candidate array: the selector's sets are bounded by the lists it unions (each <= stop_term_limit; the disjunctive loop stops at
        // 100 K docs, the AND path adds at most two top-idf lists to its tiers) -- never sized by N for large shards
        const int64_t cand_cap = std::min<int64_t>(N, 3LL * std::min<int64_t>(N, P.stop_term_limit) + 100LL * MAX_K + 4096);
        (void)cand_cap;
        size_t per_cta = (size_t)nwords * 16 + (size_t)(nwords + ncont + 2) * 4 + 6 * (size_t)(ncont + 2) * 4 + 2 * (size_t)max_list * 4 + CHUNK * 8;
#ifdef IFX_EMU
        ix->n_ctas = 1; ix->n_ctas_sel = 1;
#else
        { cudaDeviceProp prop; CUDA_TRY(cudaGetDeviceProperties(&prop, P.device));
          size_t free_b = 0, total_b = 0; CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
          int want = prop.multiProcessorCount * 2;          // = resident CTAs of the persistent kernels (__launch_bounds__(.., 2)): one workspace each
          size_t budget = free_b / 2;
          ix->n_ctas_sel = (int)std::max<size_t>(1, std::min<size_t>((size_t)prop.multiProcessorCount * 4, budget / std::max<size_t>(per_cta, 1)));
          ix->n_ctas = std::min(want, ix->n_ctas_sel); }
#endif
        const size_t index_bytes = ix->bytes;
        ix->ws.resize(std::max(ix->n_ctas, ix->n_ctas_sel));
        for (auto& w : ix->ws) { w.bits = ix->alloc<unsigned>(nwords); dev_zero(w.bits, (size_t)nwords * 4); w.bits2 = ix->alloc<unsigned>(nwords); dev_zero(w.bits2, (size_t)nwords * 4); w.cand = nullptr; w.cand_cap = 0; w.rank = ix->alloc<int32_t>((size_t)nwords + ncont + 2); w.probe = ix->alloc<unsigned long long>((size_t)nwords); dev_zero(w.probe, (size_t)nwords * 8); w.cstart = ix->alloc<int32_t>((size_t)ncont + 2); w.cfirst = ix->alloc<int32_t>((size_t)ncont + 2); w.ctab = ix->alloc<int32_t>(4 * ((size_t)ncont + 2)); w.buf_a = ix->alloc<int32_t>(max_list); w.buf_b = ix->alloc<int32_t>(max_list); w.buf_cap = max_list; w.surv_g = ix->alloc<unsigned long long>(CHUNK); }
        ix->d_ws = ix->up(ix->ws.data(), ix->ws.size());
        ix->pool_cap = (unsigned long long)std::max<int64_t>(1 << 20, std::min<int64_t>((int64_t)N * 64, (int64_t)1 << 31));
        ix->d_pool = ix->alloc<int32_t>(ix->pool_cap);
        {   // Stage-1 staging pool: per candidate 8 bytes + one tf byte per scored term; a batch that outgrows it is finished in waves
            unsigned long long want = (unsigned long long)std::max<int64_t>(64LL << 20, std::min<int64_t>((int64_t)N * 1600, 24LL << 30));
            if (const char* e = getenv("IFX_S1_POOL_MB")) want = (unsigned long long)atoll(e) << 20;
#ifndef IFX_EMU
            size_t free_b = 0, total_b = 0; CUDA_TRY(cudaMemGetInfo(&free_b, &total_b)); want = std::min<unsigned long long>(want, free_b / 2);
#endif
            ix->spool_cap = want; ix->d_spool = ix->alloc<unsigned char>(want); }
        ix->max_batch = P.max_batch > 0 ? P.max_batch : 16384;
        stage("workspaces");
        if (timing) fprintf(stderr, "[ifx_index_create] device memory: index %.1f MB, workspaces %d x %.1f MB, fuzzy pool %.1f MB\n", index_bytes / 1e6, (int)ix->ws.size(),
                            (ix->bytes - index_bytes - (size_t)ix->pool_cap * 4) / 1e6 / ix->ws.size(), ix->pool_cap * 4 / 1e6);
    } catch (const std::string& e) { delete ix; return fail(IFX_ERR_CUDA, e); }
    catch (const std::bad_alloc&) { delete ix; return fail(IFX_ERR_OOM, "host allocation failed"); }
    *out = ix; return IFX_OK;
}

extern "C" void ifx_index_destroy(ifx_index* idx) { if (!idx) return; try { DeviceGuard dg(idx->device); delete idx; } catch (...) { } }

#include "ifx_launch.inl"
