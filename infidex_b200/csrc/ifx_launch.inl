// infidex_b200 -- kernels (thin wrappers over the Ctx-based routines) and the batch entry points.

#ifndef IFX_EMU
__global__ void k_prepare(DevIndex ix, const uint16_t* text, const int64_t* off, const int32_t* par, int nq, QueryPlan* plans,
                          FuzzyItem* items, int items_cap, BatchCounters* bc, int32_t* short_kind) {
    int q = blockIdx.x * blockDim.x + threadIdx.x; if (q >= nq) return;
    prepare_query(ix, text + off[q], (int)(off[q + 1] - off[q]), par[q * 5 + 1], par[q * 5 + 0], par[q * 5 + 2], par[q * 5 + 3], par[q * 5 + 4], plans[q], items, items_cap, bc, q);
    short_kind[q] = plans[q].short_kind; if (plans[q].short_kind) atomicAdd(&bc->n_short, 1);
}
// ---- short-query path (ifx_short.h): grid-wide launches per query ---------------------------------------------------------------------------
__global__ void k_sq_char(DevIndex ix, SqScratch S, uint16_t ch) {
    for (int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; d < ix.n_docs; d += (int64_t)gridDim.x * blockDim.x) S.vf[d] = ix.deleted[d] ? 0.f : sq_single_char_score(ix, (int)d, ch);
}
__global__ void k_sq_terms(DevIndex ix, const QueryPlan* plan, SqScratch S) { if (threadIdx.x == 0) { SqPatterns P; sq_build_patterns(plan->qtext, plan->qlen, P); sq_collect_pattern_terms(ix, P, S); } }
__global__ void k_sq_accum(DevIndex ix, SqScratch S) {
    Ctx c; const int n = S.counters[1]; const int64_t w0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t t = w0; t < n; t += nw) sq_process_term(c, ix, S.terms[t], S.tmult[t], S);
}
__global__ void k_sq_fuzzy(DevIndex ix, const QueryPlan* plan, SqScratch S) {
    __shared__ SqPatterns P; Ctx c; if (threadIdx.x == 0) sq_build_patterns(plan->qtext, plan->qlen, P); __syncthreads();
    const int64_t w0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t t = w0; t < ix.terms.n; t += nw) { int w = 0; if (c.lane() == 0) w = ix.df[t] > 0 ? sq_fuzzy_weight(ix, (int)t, P, plan->qtext, plan->qlen) : 0; w = __shfl_sync(0xffffffffu, w, 0); if (w) sq_process_term(c, ix, (int)t, w, S); }
}
__global__ void k_sq_max(DevIndex ix, SqScratch S) { int m = 0; for (int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; d < ix.n_docs; d += (int64_t)gridDim.x * blockDim.x) m = S.vi[d] > m ? S.vi[d] : m; if (m > 0) atomicMax(&S.counters[2], m); }
__global__ void k_sq_final(DevIndex ix, const QueryPlan* plan, SqScratch S) {
    const int vmax = S.counters[2];
    for (int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; d < ix.n_docs; d += (int64_t)gridDim.x * blockDim.x) { const int v = S.vi[d]; S.vf[d] = (v > 0 && !ix.deleted[d]) ? sq_final_score(ix, (int)d, v, vmax, plan->qtext, plan->qlen) : 0.f; }
}
__global__ void __launch_bounds__(256) k_sq_top1(DevIndex ix, SqScratch S) { extern __shared__ __align__(16) unsigned char smem_raw[]; Ctx c; sq_topk_stage1(c, ix, S, *reinterpret_cast<SqTopShared*>(smem_raw), blockIdx.x, gridDim.x); }
__global__ void __launch_bounds__(256) k_sq_top2(SqScratch S, int nblk, int keep, int64_t* key, int32_t* doc, float* score, int32_t* n, int32_t* total) {
    extern __shared__ __align__(16) unsigned char smem_raw[]; Ctx c; sq_topk_stage2(c, S, *reinterpret_cast<SqTopShared*>(smem_raw), nblk, keep, key, doc, score, n, total); }
__global__ void __launch_bounds__(256) k_sq_champ(DevIndex ix, int ci, int m, int64_t* key, int32_t* doc, float* score, int32_t* n, int32_t* total) {
    extern __shared__ __align__(16) unsigned char smem_raw[]; Ctx c; sq_champions(c, ix, *reinterpret_cast<SqTopShared*>(smem_raw), ci, m, key, doc, score, n, total); }
// Longest-processing-time-first order: queries bucketed by their estimated work (posting volume, or the candidate count when the
// prefix shortcut will apply; log2 scale, eight steps per octave), heaviest
// bucket first, so the long sequential chunk chains of heavy queries start at once instead of forming the tail of the launch.
__global__ void k_order(DevIndex ix, const QueryPlan* plans, int nq, int* order) {
    __shared__ int cnt[512]; __shared__ int base[512];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) cnt[i] = 0;
    __syncthreads();
    auto bucket = [&](int q) { const QueryPlan& p = plans[q]; long long c = 8; int64_t r0, pop;
                               if (p.status == 0 && p.n_terms > 0 && prefix_shortcut(ix, p, p.depth, r0, pop)) c += pop * 16;     // few candidates whatever the lists' lengths
                               else for (int i = 0; i < p.n_terms; i++) c += p.terms[i].list_len;
                               int msb = 63 - __clzll(c); return msb * 8 + (int)((c >> (msb - 3)) & 7); };
    for (int q = threadIdx.x; q < nq; q += blockDim.x) atomicAdd(&cnt[bucket(q)], 1);
    __syncthreads();
    if (threadIdx.x == 0) { int run = 0; for (int b = 511; b >= 0; b--) { base[b] = run; run += cnt[b]; } }
    __syncthreads();
    for (int q = threadIdx.x; q < nq; q += blockDim.x) { int pos = atomicAdd(&base[bucket(q)], 1); order[pos] = q; }
}
__global__ void __launch_bounds__(IFX_EXPAND_THREADS, 2) k_expand(DevIndex ix, QueryPlan* plans, const FuzzyItem* items, BatchCounters* bc, S1Workspace* wss,
                                                int32_t* pool, unsigned long long pool_cap, const uint8_t* sorted_len, int* work, int items_cap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    S1Shared& sh = *reinterpret_cast<S1Shared*>(smem_raw); Ctx c; S1Workspace ws = wss[blockIdx.x];
    for (int i = threadIdx.x; i < MAX_CONTAINERS; i += blockDim.x) sh.dirty[i] = 0;
    __syncthreads();
    const int n_items = bc->n_fuzzy_items < items_cap ? bc->n_fuzzy_items : items_cap;   // prepare_query counts past the cap (and flags those queries)
    for (;;) {
        if (threadIdx.x == 0) sh.bcast[7] = atomicAdd(work, 1);
        __syncthreads();
        int it = sh.bcast[7]; __syncthreads();
        if (it >= n_items) break;
        FuzzyItem fi = items[it];
        expand_fuzzy(c, ix, plans[fi.query], fi.slot, ws, sh, pool, pool_cap, bc, sorted_len, sh.cand_s);
    }
}
// Stage 1, kernel 1: candidate selection + tf lookups of one query per CTA (persistent, LPT order). `wave` > 0: only the queries an earlier
// wave deferred because the staging pool was full.
#ifndef IFX_SEL_THREADS
#define IFX_SEL_THREADS 256
#endif
#ifndef IFX_SEL_CTAS
#define IFX_SEL_CTAS 4
#endif
__global__ void __launch_bounds__(IFX_SEL_THREADS, IFX_SEL_CTAS) k_select_lookup(DevIndex ix, const QueryPlan* plans, int nq, const int32_t* pool, S1Workspace* wss, BatchCounters* bc,
                                                int32_t* s1_n, int* work, const int* order, long long* qdbg, S1Rec* recs, unsigned char* spool, unsigned long long spool_cap,
                                                S1Queues queues, int wave, int force_mode, int smode, int32_t* sel_cnt, int32_t* sel_done) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    S1SelShared& sh = *reinterpret_cast<S1SelShared*>(smem_raw); Ctx c; S1Workspace ws = wss[blockIdx.x];
    for (int i = threadIdx.x; i < MAX_CONTAINERS; i += blockDim.x) sh.dirty[i] = 0;
    __syncthreads();
    for (;;) {
        if (threadIdx.x == 0) sh.bcast[7] = atomicAdd(work, 1);
        __syncthreads();
        int qi = sh.bcast[7]; __syncthreads();
        if (qi >= nq) break;
        const int q = order[qi];
        if (wave > 0 && recs[q].state != 2) continue;
        Stage1Out o{nullptr, nullptr, nullptr, s1_n + q, qdbg + (size_t)q * IFX_QDBG};
        unsigned long long t0 = 0; if (threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        int32_t* cq = sel_cnt ? sel_cnt + (size_t)q * SEL_CNT : nullptr;
        if (smode == 2 && sel_done[q] == 1 && recs[q].state != 2) continue;   // its candidate set was already final in the count pass
        const int path = stage1_select(c, ix, plans[q], pool, ws, sh, o, smode, cq);
        if (smode == 1) {      // count pass (doc-id-range shards): the cardinalities leave; a query whose decisions this shard could take alone is finished now
            __syncthreads();
            if (path > 0 && cq[4] != 1) { stage1_clear_bits(c, ix, ws, sh); continue; }
            if (threadIdx.x == 0) sel_done[q] = 1;
        }
        stage1_lookup(c, ix, plans[q], path, q, ws, sh, recs, spool, spool_cap, queues, bc, o, ix.fwd_avg_bytes, force_mode);
        __syncthreads();
        if (threadIdx.x == 0) { unsigned long long t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); atomicAdd(&bc->s1_ns_sum, t1 - t0); atomicMax(&bc->s1_ns_max, t1 - t0); qdbg[(size_t)q * IFX_QDBG + 4] = (long long)(t1 - t0); qdbg[(size_t)q * IFX_QDBG + 2] -= (long long)t0; qdbg[(size_t)q * IFX_QDBG + 5] = blockIdx.x; }
    }
}
// Stage 1, kernel 2a: one warp per light / mid query (persistent warps pulling from the class's queue).
#ifndef IFX_SW_WARPS
#define IFX_SW_WARPS 16
#endif
#ifndef IFX_SW_WARPS_MID
#define IFX_SW_WARPS_MID 16
#endif
template <int CAPW, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) k_score_warp(DevIndex ix, const S1Rec* recs, const unsigned char* spool, const int32_t* queue, const int32_t* n_queue, int* work,
                                                              int32_t* s1_doc, float* s1_score, int32_t* s1_n, int K, long long* qdbg) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Ctx c; WarpScoreShared& sh = reinterpret_cast<WarpScoreShared*>(smem_raw)[threadIdx.x >> 5];
    const int n = *n_queue;
    for (;;) {
        int qi = 0; if (c.lane() == 0) qi = atomicAdd(work, 1); qi = __shfl_sync(0xffffffffu, qi, 0);
        if (qi >= n) break;
        const int q = queue[qi];
        unsigned long long t0 = 0; if (c.lane() == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        score_warp<CAPW>(c, ix.avgdl, recs[q], spool, sh, s1_doc + (size_t)q * K, s1_score + (size_t)q * K, s1_n + q);
        if (c.lane() == 0) { unsigned long long t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); qdbg[(size_t)q * IFX_QDBG + 10] = (long long)(t1 - t0); }
    }
}
// Stage 1, kernel 2b: one CTA per heavy query.
__global__ void __launch_bounds__(IFX_S1_THREADS, 2) k_score_cta(DevIndex ix, const S1Rec* recs, const unsigned char* spool, const int32_t* heavy, const BatchCounters* bc, int* work, S1Workspace* wss,
                                                                 int32_t* s1_doc, float* s1_score, int32_t* s1_n, int K, long long* qdbg) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    S1Shared& sh = *reinterpret_cast<S1Shared*>(smem_raw); Ctx c; S1Workspace ws = wss[blockIdx.x];
    const int n = bc->s1_n_heavy;
    for (;;) {
        if (threadIdx.x == 0) sh.bcast[7] = atomicAdd(work, 1);
        __syncthreads();
        int qi = sh.bcast[7]; __syncthreads();
        if (qi >= n) break;
        const int q = heavy[qi];
        unsigned long long t0 = 0; if (threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        score_cta(c, ix.avgdl, recs[q], spool, ws, sh, s1_doc + (size_t)q * K, s1_score + (size_t)q * K, s1_n + q);
        if (threadIdx.x == 0) { unsigned long long t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); qdbg[(size_t)q * IFX_QDBG + 10] = (long long)(t1 - t0); }
    }
}
// Stage 1, kernel 3: final order of every query scored in this wave.
__global__ void __launch_bounds__(256) k_s1_finish(DevIndex ix, const BatchCounters* bc, S1Queues queues, int64_t* s1_key, int32_t* s1_doc, float* s1_score, int32_t* s1_n, int K) {
    __shared__ FinishShared sh; Ctx c; const int nl = bc->s1_n_light, nm = bc->s1_n_mid, nh = bc->s1_n_heavy; const int b = blockIdx.x;
    if (b >= nl + nm + nh) return;
    const int q = b < nl ? queues.light[b] : (b < nl + nm ? queues.mid[b - nl] : queues.heavy[b - nl - nm]);
    s1_finish(c, ix, sh, s1_key + (size_t)q * K, s1_doc + (size_t)q * K, s1_score + (size_t)q * K, s1_n + q);
}
#endif

// Stage 1 of the queries without a word of >= 3 characters (ifx_short.h): a handful of grid-wide launches per such query. Returns how many ran.
static int run_short_queries(ifx_batch* b, ifx_stats* st) {
    ifx_index* ix = b->idx; const int nq = b->nq; const int K = b->s1_stride;
    dev_zero(b->d_s1_total, (size_t)nq * 4);
    BatchCounters bc; d2h(&bc, b->d_bc, sizeof(bc)); if (bc.n_short == 0) return 0;
    std::vector<int32_t> kinds(nq); d2h(kinds.data(), b->d_short_kind, (size_t)nq * 4);
    const int N = ix->v.n_docs; const int nblk = std::max(1, std::min(ix->n_ctas, (N + 4095) / 4096));
    if (!ix->sq.vi) { SqScratch& S = ix->sq; S.vi = ix->alloc<int32_t>((size_t)N + 1); S.vf = ix->alloc<float>((size_t)N + 1); S.terms = ix->alloc<int32_t>(SQ_PATTERNS * SQ_TERMS); S.tmult = ix->alloc<int32_t>(SQ_PATTERNS * SQ_TERMS); S.counters = ix->alloc<int32_t>(8);
        const size_t nb = (size_t)std::max(ix->n_ctas, 1) * SQ_KB; S.cand_score = ix->alloc<float>(nb); S.cand_key = ix->alloc<int64_t>(nb); S.cand_doc = ix->alloc<int32_t>(nb); S.cand_n = ix->alloc<int32_t>(std::max(ix->n_ctas, 1)); }
    SqScratch S = ix->sq; int ran = 0; QueryPlan plan;
#ifdef IFX_EMU
    static SqTopShared* tsh = new SqTopShared(); Ctx c;
#else
    if (!ix->attr_sq) { CUDA_TRY(cudaFuncSetAttribute(k_sq_top1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SqTopShared))); CUDA_TRY(cudaFuncSetAttribute(k_sq_top2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SqTopShared))); CUDA_TRY(cudaFuncSetAttribute(k_sq_champ, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SqTopShared))); ix->attr_sq = true; }
    const int grid = ix->n_ctas * 2;
#endif
    for (int q = 0; q < nq; q++) {
        if (!kinds[q]) continue;
        d2h(&plan, b->d_plans + q, sizeof(QueryPlan)); if (plan.status != 0) continue;
        int64_t* o_key = b->d_s1_key + (size_t)q * K; int32_t* o_doc = b->d_s1_doc + (size_t)q * K; float* o_score = b->d_s1_score + (size_t)q * K; int32_t* o_n = b->d_s1_n + q; int32_t* o_tot = b->d_s1_total + q;
        const int keep = std::min(K, SQ_KB); ran++;       // K = row length of the list arrays (>= depth and >= max_results up to 1024)
        if (plan.short_kind == 1) {
            const uint16_t ch = plan.qtext[0]; int ci = -1; int have = 0;      // (the host normalised and lower-cased the query)
            for (int i = 0; i < ix->h_champ_chars.size(); i++) if (ix->h_champ_chars[i] == ch) { ci = i; have = ix->h_champ_off[i + 1] - ix->h_champ_off[i]; }
            const int m = plan.max_results;
            if (ci >= 0 && m > 0 && have >= m && m <= keep) {      // ShortQueryResolver.TryGetChampions
#ifdef IFX_EMU
                sq_champions(c, ix->v, *tsh, ci, m, o_key, o_doc, o_score, o_n, o_tot);
#else
                k_sq_champ<<<1, 256, sizeof(SqTopShared)>>>(ix->v, ci, m, o_key, o_doc, o_score, o_n, o_tot);
#endif
                continue;
            }
#ifdef IFX_EMU
            for (int d = 0; d < N; d++) S.vf[d] = ix->v.deleted[d] ? 0.f : sq_single_char_score(ix->v, d, ch);
#else
            k_sq_char<<<grid, 256>>>(ix->v, S, ch);
#endif
        } else {
            dev_zero(S.vi, (size_t)N * 4); dev_zero(S.counters, 32);
#ifdef IFX_EMU
            { SqPatterns P; sq_build_patterns(plan.qtext, plan.qlen, P); sq_collect_pattern_terms(ix->v, P, S);
              for (int t = 0; t < S.counters[1]; t++) sq_process_term(c, ix->v, S.terms[t], S.tmult[t], S);
              if (S.counters[0] < 100) for (int t = 0; t < ix->v.terms.n; t++) { int w = ix->v.df[t] > 0 ? sq_fuzzy_weight(ix->v, t, P, plan.qtext, plan.qlen) : 0; if (w) sq_process_term(c, ix->v, t, w, S); }
              int vmax = 0; for (int d = 0; d < N; d++) vmax = std::max(vmax, S.vi[d]); S.counters[2] = vmax;
              for (int d = 0; d < N; d++) { const int v = S.vi[d]; S.vf[d] = (v > 0 && !ix->v.deleted[d]) ? sq_final_score(ix->v, d, v, vmax, plan.qtext, plan.qlen) : 0.f; } }
#else
            k_sq_terms<<<1, 32>>>(ix->v, b->d_plans + q, S);
            k_sq_accum<<<grid, 256>>>(ix->v, S);
            int32_t cnt[4]; d2h(cnt, S.counters, 16);
            if (cnt[0] < 100) k_sq_fuzzy<<<grid, 256>>>(ix->v, b->d_plans + q, S);      // ProcessFuzzyFallback: every term of the dictionary
            k_sq_max<<<grid, 256>>>(ix->v, S);
            k_sq_final<<<grid, 256>>>(ix->v, b->d_plans + q, S);
#endif
        }
        // the query's Stage-1 list: exact top-`keep` by (score descending, key ascending) + the number of matched documents
#ifdef IFX_EMU
        S.counters[3] = 0; for (int bl = 0; bl < nblk; bl++) sq_topk_stage1(c, ix->v, S, *tsh, bl, nblk);
        int keep_q = plan.short_kind == 1 ? std::min(keep, plan.max_results) : keep;
        sq_topk_stage2(c, S, *tsh, nblk, keep_q, o_key, o_doc, o_score, o_n, o_tot);
        if (plan.short_kind == 1) *o_tot = *o_n;
#else
        CUDA_TRY(cudaMemsetAsync(S.counters + 3, 0, 4));
        k_sq_top1<<<nblk, 256, sizeof(SqTopShared)>>>(ix->v, S);
        const int keep_q = plan.short_kind == 1 ? std::min(keep, plan.max_results) : keep;
        k_sq_top2<<<1, 256, sizeof(SqTopShared)>>>(S, nblk, keep_q, o_key, o_doc, o_score, o_n, o_tot);
        if (plan.short_kind == 1) CUDA_TRY(cudaMemcpyAsync(o_tot, o_n, 4, cudaMemcpyDeviceToDevice));      // SearchSingleCharacter cuts its list to max_results before anything counts it
        CUDA_TRY(cudaGetLastError());
#endif
    }
    if (st) st->kernel_launches += 6 * ran;
    return ran;
}

static int s1_force_mode() { const char* e = getenv("IFX_S1_LOOKUP"); return e ? atoi(e) : 0; }      // 1 forward-index lookups, 2 streamed lists (tests / A-B runs)

// `part`: 1 = query preparation + LD1 expansion, 2 = selection / lookups / scoring / final order, 3 = both (the unsharded call)
//         4 = selection count pass only (doc-id-range shards; the hosts sum b->d_sel_cnt over the shards before part 2 runs with smode 2)
static void run_stage1_phase(ifx_batch* b, ifx_stats* st, int part = 3) {
    ifx_index* ix = b->idx; const int nq = b->nq; const int K = b->s1_stride;
    if (part & 1) { BatchCounters zero{}; h2d(b->d_bc, &zero, sizeof(zero)); }
    const int items_cap = nq * MAX_FUZZY;       // every query may carry MAX_FUZZY unknown words: the item list can never overflow
    const int force_mode = s1_force_mode(); S1Queues queues{b->d_light, b->d_mid, b->d_heavy};
    Timer t;
#ifdef IFX_EMU
    std::vector<int64_t> off(nq + 1); d2h(off.data(), b->d_off, (nq + 1) * 8);
    if (part & 1) for (int q = 0; q < nq; q++) { prepare_query(ix->v, b->d_text + off[q], (int)(off[q + 1] - off[q]), b->d_par[q * 5 + 1], b->d_par[q * 5 + 0], b->d_par[q * 5 + 2], b->d_par[q * 5 + 3], b->d_par[q * 5 + 4], b->d_plans[q], b->d_items, items_cap, b->d_bc, q); b->d_short_kind[q] = b->d_plans[q].short_kind; if (b->d_plans[q].short_kind) b->d_bc->n_short++; }
    static S1Shared* sh = new S1Shared(); memset(sh->dirty, 0, sizeof(sh->dirty)); static WarpScoreShared* wsh = new WarpScoreShared(); static FinishShared* fsh = new FinishShared();
    Ctx c; int nit = std::min(b->d_bc->n_fuzzy_items, items_cap);
    if (part & 1) for (int i = 0; i < nit; i++) expand_fuzzy(c, ix->v, b->d_plans[b->d_items[i].query], b->d_items[i].slot, ix->ws[0], *sh, ix->d_pool, ix->pool_cap, b->d_bc, ix->d_sorted_len, sh->cand_s);
    const int smode = b->use_gcnt ? 2 : 0;
    if (part & 4) for (int q = 0; q < nq; q++) { Stage1Out o{nullptr, nullptr, nullptr, b->d_s1_n + q, nullptr}; for (int k = 0; k < SEL_CNT; k++) b->d_sel_cnt[(size_t)q * SEL_CNT + k] = 0; b->d_sel_done[q] = 0;
        int32_t* cq = b->d_sel_cnt + (size_t)q * SEL_CNT;
        const int path = stage1_select(c, ix->v, b->d_plans[q], ix->d_pool, ix->ws[0], *sh, o, 1, cq);
        if (path > 0 && cq[4] != 1) { stage1_clear_bits(c, ix->v, ix->ws[0], *sh); continue; }
        b->d_sel_done[q] = 1; stage1_lookup(c, ix->v, b->d_plans[q], path, q, ix->ws[0], *sh, b->d_recs, ix->d_spool, ix->spool_cap, queues, b->d_bc, o, ix->v.fwd_avg_bytes, force_mode); }
    const bool no_warp = getenv("IFX_S1_NO_WARP") != nullptr;      // tests: force every query through the block-wide scorer
    for (int wave = 0; wave < 64 && (part & 2); wave++) {
        if (wave > 0 || smode == 0) { b->d_bc->s1_pool_used = 0; b->d_bc->s1_deferred = 0; b->d_bc->s1_n_light = 0; b->d_bc->s1_n_mid = 0; b->d_bc->s1_n_heavy = 0; }      // (shards: the count pass already queued the queries it could finish)
        b->d_bc->s1_wave = wave;
        for (int q = 0; q < nq; q++) {
            if (wave > 0 && b->d_recs[q].state != 2) continue;
            if (smode == 2 && b->d_sel_done[q] == 1 && b->d_recs[q].state != 2) continue;
            Stage1Out o{nullptr, nullptr, nullptr, b->d_s1_n + q, nullptr};
            const int path = stage1_select(c, ix->v, b->d_plans[q], ix->d_pool, ix->ws[0], *sh, o, smode, b->d_sel_cnt + (size_t)q * SEL_CNT);
            stage1_lookup(c, ix->v, b->d_plans[q], path, q, ix->ws[0], *sh, b->d_recs, ix->d_spool, ix->spool_cap, queues, b->d_bc, o, ix->v.fwd_avg_bytes, force_mode);
        }
        for (int i = 0; i < b->d_bc->s1_n_light; i++) { const int q = b->d_light[i];
            if (no_warp) score_cta(c, ix->v.avgdl, b->d_recs[q], ix->d_spool, ix->ws[0], *sh, b->d_s1_doc + (size_t)q * K, b->d_s1_score + (size_t)q * K, b->d_s1_n + q);
            else score_warp<W_CAP>(c, ix->v.avgdl, b->d_recs[q], ix->d_spool, *wsh, b->d_s1_doc + (size_t)q * K, b->d_s1_score + (size_t)q * K, b->d_s1_n + q); }
        for (int i = 0; i < b->d_bc->s1_n_mid; i++) { const int q = b->d_mid[i];
            if (no_warp) score_cta(c, ix->v.avgdl, b->d_recs[q], ix->d_spool, ix->ws[0], *sh, b->d_s1_doc + (size_t)q * K, b->d_s1_score + (size_t)q * K, b->d_s1_n + q);
            else score_warp<W_CAP_MID>(c, ix->v.avgdl, b->d_recs[q], ix->d_spool, *wsh, b->d_s1_doc + (size_t)q * K, b->d_s1_score + (size_t)q * K, b->d_s1_n + q); }
        for (int i = 0; i < b->d_bc->s1_n_heavy; i++) { const int q = b->d_heavy[i]; score_cta(c, ix->v.avgdl, b->d_recs[q], ix->d_spool, ix->ws[0], *sh, b->d_s1_doc + (size_t)q * K, b->d_s1_score + (size_t)q * K, b->d_s1_n + q); }
        const int nl_ = b->d_bc->s1_n_light, nm_ = b->d_bc->s1_n_mid, nh_ = b->d_bc->s1_n_heavy;
        for (int i = 0; i < nl_ + nm_ + nh_; i++) { const int q = i < nl_ ? b->d_light[i] : (i < nl_ + nm_ ? b->d_mid[i - nl_] : b->d_heavy[i - nl_ - nm_]);
            s1_finish(c, ix->v, *fsh, b->d_s1_key + (size_t)q * K, b->d_s1_doc + (size_t)q * K, b->d_s1_score + (size_t)q * K, b->d_s1_n + q); }
        if (b->d_bc->s1_deferred == 0) break;
    }
    (void)t;
#else
    const size_t smem = sizeof(S1Shared), smem_sel = sizeof(S1SelShared), smem_w = sizeof(WarpScoreShared) * IFX_SW_WARPS, smem_m = sizeof(WarpScoreShared) * IFX_SW_WARPS_MID;
    auto k_light = k_score_warp<W_CAP, IFX_SW_WARPS>; auto k_mid = k_score_warp<W_CAP_MID, IFX_SW_WARPS_MID>;
    if (!ix->attr_s1) { CUDA_TRY(cudaFuncSetAttribute(k_expand, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); CUDA_TRY(cudaFuncSetAttribute(k_select_lookup, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_sel));
        CUDA_TRY(cudaFuncSetAttribute(k_score_cta, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); CUDA_TRY(cudaFuncSetAttribute(k_light, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w)); CUDA_TRY(cudaFuncSetAttribute(k_mid, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_m)); ix->attr_s1 = true; }
    float ms_prep = 0.f, ms_exp = 0.f; int launches = 0;
    if (part & 1) {
        t.start();
        k_prepare<<<(nq + 127) / 128, 128>>>(ix->v, b->d_text, b->d_off, b->d_par, nq, b->d_plans, b->d_items, items_cap, b->d_bc, b->d_short_kind);
        ms_prep = t.stop();
        t.start();
        CUDA_TRY(cudaMemsetAsync(b->d_work, 0, 8 * sizeof(int)));
        k_expand<<<ix->n_ctas, IFX_EXPAND_THREADS, smem>>>(ix->v, b->d_plans, b->d_items, b->d_bc, ix->d_ws, ix->d_pool, ix->pool_cap, ix->d_sorted_len, b->d_work, items_cap);
        ms_exp = t.stop(); launches += 2;
    }
    float ms_sel = 0.f, ms_sw = 0.f, ms_sc = 0.f, ms_fin = 0.f; const int sms = ix->n_ctas / 2 > 0 ? ix->n_ctas / 2 : 1; const int smode = b->use_gcnt ? 2 : 0;
    if (part & 6) { k_order<<<1, 1024>>>(ix->v, b->d_plans, nq, b->d_order); launches++; }
    if (part & 4) {      // count pass of the selection
        t.start();
        CUDA_TRY(cudaMemsetAsync(b->d_sel_cnt, 0, (size_t)nq * SEL_CNT * 4)); CUDA_TRY(cudaMemsetAsync(b->d_sel_done, 0, (size_t)nq * 4)); CUDA_TRY(cudaMemsetAsync(b->d_work + 1, 0, sizeof(int)));
        k_select_lookup<<<std::min(ix->n_ctas_sel, nq), IFX_SEL_THREADS, smem_sel>>>(ix->v, b->d_plans, nq, ix->d_pool, ix->d_ws, b->d_bc, b->d_s1_n, b->d_work + 1, b->d_order, b->d_qdbg, b->d_recs, ix->d_spool, ix->spool_cap, queues, 0, force_mode, 1, b->d_sel_cnt, b->d_sel_done);
        ms_sel += t.stop(); launches++;
        CUDA_TRY(cudaMemsetAsync(b->d_work + 1, 0, sizeof(int)));
    }
    for (int wave = 0; wave < 64 && (part & 2); wave++) {
        if (wave > 0) { BatchCounters bc; d2h(&bc, b->d_bc, sizeof(bc)); bc.s1_pool_used = 0; bc.s1_deferred = 0; bc.s1_n_light = 0; bc.s1_n_mid = 0; bc.s1_n_heavy = 0; bc.s1_wave = wave; h2d(b->d_bc, &bc, sizeof(bc)); CUDA_TRY(cudaMemsetAsync(b->d_work + 1, 0, 5 * sizeof(int))); }
        t.start();
        k_select_lookup<<<std::min(ix->n_ctas_sel, nq), IFX_SEL_THREADS, smem_sel>>>(ix->v, b->d_plans, nq, ix->d_pool, ix->d_ws, b->d_bc, b->d_s1_n, b->d_work + 1, b->d_order, b->d_qdbg, b->d_recs, ix->d_spool, ix->spool_cap, queues, wave, force_mode, smode, b->d_sel_cnt, b->d_sel_done);
        ms_sel += t.stop(); t.start();
        k_score_cta<<<std::min(ix->n_ctas, nq), IFX_S1_THREADS, smem>>>(ix->v, b->d_recs, ix->d_spool, b->d_heavy, b->d_bc, b->d_work + 2, ix->d_ws, b->d_s1_doc, b->d_s1_score, b->d_s1_n, K, b->d_qdbg);
        ms_sc += t.stop(); t.start();
        k_mid<<<sms, IFX_SW_WARPS_MID * 32, smem_m>>>(ix->v, b->d_recs, ix->d_spool, b->d_mid, &b->d_bc->s1_n_mid, b->d_work + 4, b->d_s1_doc, b->d_s1_score, b->d_s1_n, K, b->d_qdbg);
        k_light<<<sms, IFX_SW_WARPS * 32, smem_w>>>(ix->v, b->d_recs, ix->d_spool, b->d_light, &b->d_bc->s1_n_light, b->d_work + 3, b->d_s1_doc, b->d_s1_score, b->d_s1_n, K, b->d_qdbg);
        ms_sw += t.stop(); t.start();
        k_s1_finish<<<nq, 256>>>(ix->v, b->d_bc, queues, b->d_s1_key, b->d_s1_doc, b->d_s1_score, b->d_s1_n, K);
        ms_fin += t.stop(); launches += 5;
        int deferred = 0; d2h(&deferred, &b->d_bc->s1_deferred, sizeof(int));
        if (deferred == 0) break;
    }
    CUDA_TRY(cudaGetLastError());
    if (st) { st->ms_prepare += ms_prep; st->ms_expand += ms_exp; st->ms_stage1 += ms_sel + ms_sc + ms_sw + ms_fin; st->ms_s1_select += ms_sel; st->ms_s1_score_cta += ms_sc; st->ms_s1_score_warp += ms_sw; st->ms_s1_finish += ms_fin; st->kernel_launches += launches; }
#endif
    if (st) { BatchCounters bc; d2h(&bc, b->d_bc, sizeof(bc)); st->algo_bytes_stage1 += (int64_t)bc.algo_bytes; st->s1_query_ms_max = (float)(bc.s1_ns_max * 1e-6); st->s1_query_ms_sum = (float)(bc.s1_ns_sum * 1e-6);
        st->s1_light = bc.s1_n_light; st->s1_mid = bc.s1_n_mid; st->s1_heavy = bc.s1_n_heavy; st->s1_waves = bc.s1_wave + 1; st->s1_pool_bytes = (int64_t)bc.s1_pool_used; }
}

// (re)fill the per-batch inputs; allocates on first use or when the batch outgrows its buffers
static int fill_batch(ifx_batch* b, const ifx_query* q, int nq) {
    std::vector<int64_t> off(nq + 1, 0); std::vector<int32_t> par((size_t)nq * 5);
    int depth_max = 0, cap_max = 0;
    for (int i = 0; i < nq; i++) { off[i + 1] = off[i] + std::max(q[i].len, 0); par[i * 5 + 0] = q[i].max_results; par[i * 5 + 1] = q[i].coverage_depth; par[i * 5 + 2] = q[i].enable_coverage; par[i * 5 + 3] = q[i].filter_id; par[i * 5 + 4] = q[i].enable_facets;
        depth_max = std::max(depth_max, q[i].coverage_depth); cap_max = std::max(cap_max, q[i].max_results); }
    if (depth_max < 1 || depth_max > MAX_K) return fail(IFX_ERR_INVALID, "coverage_depth must be in [1,1024]");
    std::vector<uint16_t> text((size_t)off[nq] + 1);
    for (int i = 0; i < nq; i++) if (q[i].len > 0) memcpy(text.data() + off[i], q[i].text, (size_t)q[i].len * 2);
    const bool fresh = b->d_plans == nullptr;
    if (!fresh && (nq != b->nq || depth_max != b->depth_max || cap_max != b->cap_max)) return fail(IFX_ERR_INVALID, "batch shape changed");
    if (fresh) {
        b->nq = nq; b->depth_max = depth_max; b->cap_max = cap_max; b->text_cap = std::max<size_t>(text.size(), (size_t)nq * 64);
        b->d_text = b->alloc<uint16_t>(b->text_cap); b->d_off = b->alloc<int64_t>(nq + 1); b->d_par = b->alloc<int32_t>(par.size());
        b->d_plans = b->alloc<QueryPlan>(nq); b->d_items = b->alloc<FuzzyItem>((size_t)nq * MAX_FUZZY); b->d_bc = b->alloc<BatchCounters>(1); b->d_work = b->alloc<int>(8); dev_zero(b->d_work, 8 * sizeof(int));
        b->s1_stride = std::max(depth_max, std::min(cap_max, (int)MAX_K)); size_t K = b->s1_stride;
        b->d_recs = b->alloc<S1Rec>(nq); dev_zero(b->d_recs, sizeof(S1Rec) * (size_t)nq); b->d_short_kind = b->alloc<int32_t>(nq); b->d_s1_total = b->alloc<int32_t>(nq); dev_zero(b->d_s1_total, (size_t)nq * 4); b->d_sel_cnt = b->alloc<int32_t>((size_t)nq * SEL_CNT); b->d_sel_done = b->alloc<int32_t>(nq); dev_zero(b->d_sel_done, (size_t)nq * 4); b->d_light = b->alloc<int32_t>(nq); b->d_mid = b->alloc<int32_t>(nq); b->d_heavy = b->alloc<int32_t>(nq);
        b->d_s1_key = b->alloc<int64_t>(nq * K); b->d_s1_doc = b->alloc<int32_t>(nq * K); b->d_s1_score = b->alloc<float>(nq * K); b->d_s1_n = b->alloc<int32_t>(nq); b->d_order = b->alloc<int>(nq); b->d_qdbg = b->alloc<long long>((size_t)nq * IFX_QDBG); dev_zero(b->d_qdbg, (size_t)nq * IFX_QDBG * 8);
    } else if (text.size() > b->text_cap) return fail(IFX_ERR_INVALID, "batch text outgrew its buffer");
    h2d(b->d_text, text.data(), text.size() * 2); h2d(b->d_off, off.data(), (nq + 1) * 8); h2d(b->d_par, par.data(), par.size() * 4);
    b->ran = false;
    return IFX_OK;
}

extern "C" int ifx_batch_upload(ifx_index* idx, const ifx_query* q, int nq, ifx_batch** out) {
    if (!idx || !q || nq <= 0 || !out) return fail(IFX_ERR_INVALID, "bad batch arguments");
    ifx_batch* b = new ifx_batch(); b->idx = idx;
    try { DeviceGuard dg(idx->device); int rc = fill_batch(b, q, nq); if (rc) { delete b; return rc; } }
    catch (const std::string& e) { delete b; return fail(IFX_ERR_CUDA, e); }
    *out = b; return IFX_OK;
}
extern "C" int ifx_batch_refill(ifx_batch* b, const ifx_query* q, int nq) {
    if (!b || !q || nq <= 0) return fail(IFX_ERR_INVALID, "bad batch arguments");
    try { DeviceGuard dg(b->idx->device); return fill_batch(b, q, nq); } catch (const std::string& e) { return fail(IFX_ERR_CUDA, e); }
}
extern "C" void ifx_batch_free(ifx_batch* b) { if (!b) return; try { DeviceGuard dg(b->idx->device); delete b; } catch (...) { } }

extern "C" int ifx_stage1_batch(ifx_index* idx, const ifx_query* q, int nq, int depth, int64_t* doc_key, float* score, int32_t* n, int32_t* status, ifx_stats* st) {
    if (!dev_ok()) return fail(IFX_ERR_NO_DEVICE, "no CUDA device available (infidex_b200 has no CPU fallback)");
    std::vector<ifx_query> qq(q, q + nq); for (auto& x : qq) x.coverage_depth = depth;
    ifx_batch* b = nullptr; int rc = ifx_batch_upload(idx, qq.data(), nq, &b); if (rc) return rc;
    if (st) memset(st, 0, sizeof(*st));
    try {
        std::lock_guard<std::mutex> lk(idx->mu); DeviceGuard dg(idx->device);
        run_stage1_phase(b, st); run_short_queries(b, st);
        d2h(n, b->d_s1_n, (size_t)nq * 4);
        if (b->s1_stride == depth) { d2h(doc_key, b->d_s1_key, (size_t)nq * depth * 8); d2h(score, b->d_s1_score, (size_t)nq * depth * 4); }
        else { std::vector<int64_t> k((size_t)nq * b->s1_stride); std::vector<float> sc((size_t)nq * b->s1_stride); d2h(k.data(), b->d_s1_key, k.size() * 8); d2h(sc.data(), b->d_s1_score, sc.size() * 4);
            for (int i = 0; i < nq; i++) { memcpy(doc_key + (size_t)i * depth, k.data() + (size_t)i * b->s1_stride, (size_t)depth * 8); memcpy(score + (size_t)i * depth, sc.data() + (size_t)i * b->s1_stride, (size_t)depth * 4); if (n[i] > depth) n[i] = depth; } }
        if (status) { std::vector<QueryPlan> pl(nq); d2h(pl.data(), b->d_plans, sizeof(QueryPlan) * (size_t)nq); for (int i = 0; i < nq; i++) { status[i] = pl[i].status; if (n[i] < 0) { status[i] |= IFX_Q_OVERFLOW; n[i] = 0; } } }
    } catch (const std::string& e) { delete b; return fail(IFX_ERR_CUDA, e); }
    delete b; return IFX_OK;
}

#include "ifx_search.inl"

// debugging aid: per-query [n_cand, n_terms, selection_ns, path, total_ns, cta] of the last run's k_stage1
extern "C" int ifx_debug_stage1_queries(ifx_batch* b, long long* out) { try { DeviceGuard dg(b->idx->device); d2h(out, b->d_qdbg, (size_t)b->nq * IFX_QDBG * 8); } catch (const std::string& e) { return fail(IFX_ERR_CUDA, e); } return IFX_OK; }
