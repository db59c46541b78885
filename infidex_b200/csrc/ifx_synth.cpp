// infidex_b200 -- synthetic benchmark corpus generator (SURVEY.md 8d / BASELINE.json configs[1..4]), part of libinfidex_host.so.
//
// Benchmark / test data only (nothing of the search path): documents whose words follow a rank-frequency Zipf(1.07) law over a
// given vocabulary. Every document draws from its own SplitMix64 stream seeded by (seed, global document index), so the corpus is
// identical for any thread count, chunking or shard split -- rank r of an N-GPU run generates exactly its own doc-id range.
//   title       = 1 + Poisson(2.7) words (cap 12)          description = 8 + Poisson(7) words (multi-field corpora only)
//   year  ~ U{1950..2024}     rating ~ U{1.0..10.0} step 0.1     genre ~ Zipf(1.07) over 20 categories
#include <stdint.h>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>
#include <algorithm>

namespace {
struct Sm64 {
    uint64_t s;
    uint64_t next() { uint64_t z = (s += 0x9E3779B97F4A7C15ULL); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; return z ^ (z >> 31); }
    double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    int poisson(double L) { int k = 0; double p = 1.0; do { k++; p *= uni(); } while (p > L); return k - 1; }   // Knuth; L = exp(-lambda)
};
inline Sm64 doc_stream(uint64_t seed, int64_t doc) { Sm64 r{seed ^ (0xD1B54A32D192ED03ULL * (uint64_t)(doc + 1))}; r.next(); return r; }
inline int zipf_pick(const double* cdf, int V, double u) { int i = (int)(std::lower_bound(cdf, cdf + V, u) - cdf); return i < V ? i : V - 1; }

template <class F> void par_for(int64_t n, int threads, F f) {
    threads = (int)std::max<int64_t>(1, std::min<int64_t>(threads, n));
    std::vector<std::thread> ts;
    for (int t = 0; t < threads; t++) ts.emplace_back([=] { f(n * t / threads, n * (t + 1) / threads); });
    for (auto& th : ts) th.join();
}
const double L_TITLE = std::exp(-2.7), L_DESC = std::exp(-7.0);
}  // namespace

extern "C" {

// Pass 1: per-document word counts and UTF-16 lengths, turned into offsets (all arrays n + 1 entries, entry 0 = 0).
int ifx_synth_sizes(int64_t n, int64_t start, uint64_t seed, int multi, int threads, const int32_t* vlens, const double* vcdf, int V,
                    int64_t* title_woff, int64_t* title_coff, int64_t* desc_woff, int64_t* desc_coff) {
    par_for(n, threads, [=](int64_t a, int64_t b) {
        for (int64_t d = a; d < b; d++) {
            Sm64 r = doc_stream(seed, start + d);
            int tc = std::min(1 + r.poisson(L_TITLE), 12); int64_t ch = 0;
            for (int k = 0; k < tc; k++) ch += vlens[zipf_pick(vcdf, V, r.uni())];
            title_woff[d + 1] = tc; title_coff[d + 1] = ch + tc - 1;
            if (multi) { int dc = 8 + r.poisson(L_DESC); int64_t c2 = 0; for (int k = 0; k < dc; k++) c2 += vlens[zipf_pick(vcdf, V, r.uni())]; desc_woff[d + 1] = dc; desc_coff[d + 1] = c2 + dc - 1; }
        }
    });
    title_woff[0] = title_coff[0] = 0; if (multi) desc_woff[0] = desc_coff[0] = 0;
    for (int64_t d = 0; d < n; d++) { title_woff[d + 1] += title_woff[d]; title_coff[d + 1] += title_coff[d]; if (multi) { desc_woff[d + 1] += desc_woff[d]; desc_coff[d + 1] += desc_coff[d]; } }
    return 0;
}

// Pass 2: the same streams again, now writing word ids, texts (words joined by single spaces) and the scalar columns.
int ifx_synth_fill(int64_t n, int64_t start, uint64_t seed, int multi, int threads, const uint16_t* vblob, const int64_t* voffs, const int32_t* vlens,
                   const double* vcdf, int V, const int64_t* title_woff, const int64_t* title_coff, const int64_t* desc_woff, const int64_t* desc_coff,
                   int32_t* title_ids, uint16_t* title_blob, uint16_t* desc_blob, int64_t* year, double* rating, int64_t* genre) {
    double gc[20]; { double s = 0; for (int i = 0; i < 20; i++) { s += std::pow((double)(i + 1), -1.07); gc[i] = s; } for (int i = 0; i < 20; i++) gc[i] /= s; }
    par_for(n, threads, [=, &gc](int64_t a, int64_t b) {
        auto put = [&](uint16_t* out, int w, bool last) { std::memcpy(out, vblob + voffs[w], (size_t)vlens[w] * 2); if (!last) out[vlens[w]] = u' '; return vlens[w] + (last ? 0 : 1); };
        for (int64_t d = a; d < b; d++) {
            Sm64 r = doc_stream(seed, start + d);
            int tc = std::min(1 + r.poisson(L_TITLE), 12); uint16_t* o = title_blob + title_coff[d]; int32_t* ids = title_ids + title_woff[d];
            for (int k = 0; k < tc; k++) { int w = zipf_pick(vcdf, V, r.uni()); ids[k] = w; o += put(o, w, k + 1 == tc); }
            if (multi) { int dc = 8 + r.poisson(L_DESC); uint16_t* o2 = desc_blob + desc_coff[d]; for (int k = 0; k < dc; k++) { int w = zipf_pick(vcdf, V, r.uni()); o2 += put(o2, w, k + 1 == dc); } }
            year[d] = 1950 + (int64_t)(r.next() % 75);
            rating[d] = (double)(10 + (int64_t)(r.next() % 91)) / 10.0;
            genre[d] = zipf_pick(gc, 20, r.uni());
        }
    });
    return 0;
}

// The title word ids of ONE document of the corpus (at most 12), without materialising anything else: query sampling over a corpus
// that is sharded across ranks, or too large to keep, regenerates just the documents it draws.
int ifx_synth_title_ids(int64_t doc, uint64_t seed, const double* vcdf, int V, int32_t* out) {
    Sm64 r = doc_stream(seed, doc); int tc = std::min(1 + r.poisson(L_TITLE), 12);
    for (int k = 0; k < tc; k++) out[k] = zipf_pick(vcdf, V, r.uni());
    return tc;
}

}  // extern "C"
