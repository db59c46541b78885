// infidex_b200 -- kernels (thin wrappers over the Ctx-based routines) and the batch entry points.

#ifndef IFX_EMU
__global__ void k_prepare(DevIndex ix, const uint16_t* text, const int64_t* off, const int32_t* par, int nq, QueryPlan* plans,
                          FuzzyItem* items, int items_cap, BatchCounters* bc) {
    int q = blockIdx.x * blockDim.x + threadIdx.x; if (q >= nq) return;
    prepare_query(ix, text + off[q], (int)(off[q + 1] - off[q]), par[q * 5 + 1], par[q * 5 + 0], par[q * 5 + 2], par[q * 5 + 3], par[q * 5 + 4], plans[q], items, items_cap, bc, q);
}
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
__global__ void __launch_bounds__(IFX_S1_THREADS, 2) k_stage1(DevIndex ix, const QueryPlan* plans, int nq, const int32_t* pool, S1Workspace* wss, BatchCounters* bc,
                                                int64_t* s1_key, int32_t* s1_doc, float* s1_score, int32_t* s1_n, int K, int* work, const int* order, long long* qdbg) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    S1Shared& sh = *reinterpret_cast<S1Shared*>(smem_raw); Ctx c; S1Workspace ws = wss[blockIdx.x];
    for (int i = threadIdx.x; i < MAX_CONTAINERS; i += blockDim.x) sh.dirty[i] = 0;
    for (int i = threadIdx.x; i < S1_TILE * CHUNK; i += blockDim.x) (&sh.tfm[0][0])[i] = 0;
    __syncthreads();
    for (;;) {
        if (threadIdx.x == 0) sh.bcast[7] = atomicAdd(work, 1);
        __syncthreads();
        int qi = sh.bcast[7]; __syncthreads();
        if (qi >= nq) break;
        const int q = order[qi];
        Stage1Out o{s1_key + (size_t)q * K, s1_doc + (size_t)q * K, s1_score + (size_t)q * K, s1_n + q, qdbg + (size_t)q * IFX_QDBG};
        unsigned long long t0 = 0; if (threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        stage1_query(c, ix, plans[q], pool, ws, sh, o, bc);
        __syncthreads();
        if (threadIdx.x == 0) { unsigned long long t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); atomicAdd(&bc->s1_ns_sum, t1 - t0); atomicMax(&bc->s1_ns_max, t1 - t0); qdbg[(size_t)q * IFX_QDBG + 4] = (long long)(t1 - t0); qdbg[(size_t)q * IFX_QDBG + 2] -= (long long)t0; qdbg[(size_t)q * IFX_QDBG + 5] = blockIdx.x; }
    }
}
#endif

static void run_stage1_phase(ifx_batch* b, ifx_stats* st) {
    ifx_index* ix = b->idx; const int nq = b->nq; const int K = b->depth_max;
    BatchCounters zero{}; h2d(b->d_bc, &zero, sizeof(zero));
    const int items_cap = nq * MAX_FUZZY;       // every query may carry MAX_FUZZY unknown words: the item list can never overflow
    Timer t;
#ifdef IFX_EMU
    std::vector<int64_t> off(nq + 1); d2h(off.data(), b->d_off, (nq + 1) * 8);
    for (int q = 0; q < nq; q++) prepare_query(ix->v, b->d_text + off[q], (int)(off[q + 1] - off[q]), b->d_par[q * 5 + 1], b->d_par[q * 5 + 0], b->d_par[q * 5 + 2], b->d_par[q * 5 + 3], b->d_par[q * 5 + 4], b->d_plans[q], b->d_items, items_cap, b->d_bc, q);
    static S1Shared* sh = new S1Shared(); memset(sh->dirty, 0, sizeof(sh->dirty));
    Ctx c; int nit = std::min(b->d_bc->n_fuzzy_items, items_cap);
    for (int i = 0; i < nit; i++) expand_fuzzy(c, ix->v, b->d_plans[b->d_items[i].query], b->d_items[i].slot, ix->ws[0], *sh, ix->d_pool, ix->pool_cap, b->d_bc, ix->d_sorted_len, sh->cand_s);
    for (int q = 0; q < nq; q++) { Stage1Out o{b->d_s1_key + (size_t)q * K, b->d_s1_doc + (size_t)q * K, b->d_s1_score + (size_t)q * K, b->d_s1_n + q, nullptr}; stage1_query(c, ix->v, b->d_plans[q], ix->d_pool, ix->ws[0], *sh, o, b->d_bc); }
    (void)t;
#else
    size_t smem = sizeof(S1Shared);
    if (!ix->attr_s1) { CUDA_TRY(cudaFuncSetAttribute(k_expand, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); CUDA_TRY(cudaFuncSetAttribute(k_stage1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); ix->attr_s1 = true; }
    t.start();
    k_prepare<<<(nq + 127) / 128, 128>>>(ix->v, b->d_text, b->d_off, b->d_par, nq, b->d_plans, b->d_items, items_cap, b->d_bc);
    float ms_prep = t.stop();
    t.start();
    CUDA_TRY(cudaMemsetAsync(b->d_work, 0, 2 * sizeof(int)));
    k_expand<<<ix->n_ctas, IFX_EXPAND_THREADS, smem>>>(ix->v, b->d_plans, b->d_items, b->d_bc, ix->d_ws, ix->d_pool, ix->pool_cap, ix->d_sorted_len, b->d_work, items_cap);
    float ms_exp = t.stop();
    t.start();
    k_order<<<1, 1024>>>(ix->v, b->d_plans, nq, b->d_order);
    k_stage1<<<std::min(ix->n_ctas, nq), IFX_S1_THREADS, smem>>>(ix->v, b->d_plans, nq, ix->d_pool, ix->d_ws, b->d_bc, b->d_s1_key, b->d_s1_doc, b->d_s1_score, b->d_s1_n, K, b->d_work + 1, b->d_order, b->d_qdbg);
    float ms_s1 = t.stop();
    CUDA_TRY(cudaGetLastError());
    if (st) { st->ms_prepare += ms_prep; st->ms_expand += ms_exp; st->ms_stage1 += ms_s1; st->kernel_launches += 4; }
#endif
    if (st) { BatchCounters bc; d2h(&bc, b->d_bc, sizeof(bc)); st->algo_bytes_stage1 += (int64_t)bc.algo_bytes; st->s1_query_ms_max = (float)(bc.s1_ns_max * 1e-6); st->s1_query_ms_sum = (float)(bc.s1_ns_sum * 1e-6); }
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
        size_t K = b->depth_max;
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
extern "C" void ifx_batch_free(ifx_batch* b) { if (!b) return; try { DeviceGuard dg(b->idx->device); delete b; } catch (...) { } }

extern "C" int ifx_stage1_batch(ifx_index* idx, const ifx_query* q, int nq, int depth, int64_t* doc_key, float* score, int32_t* n, int32_t* status, ifx_stats* st) {
    if (!dev_ok()) return fail(IFX_ERR_NO_DEVICE, "no CUDA device available (infidex_b200 has no CPU fallback)");
    std::vector<ifx_query> qq(q, q + nq); for (auto& x : qq) x.coverage_depth = depth;
    ifx_batch* b = nullptr; int rc = ifx_batch_upload(idx, qq.data(), nq, &b); if (rc) return rc;
    if (st) memset(st, 0, sizeof(*st));
    try {
        std::lock_guard<std::mutex> lk(idx->mu); DeviceGuard dg(idx->device);
        run_stage1_phase(b, st);
        d2h(doc_key, b->d_s1_key, (size_t)nq * depth * 8); d2h(score, b->d_s1_score, (size_t)nq * depth * 4); d2h(n, b->d_s1_n, (size_t)nq * 4);
        if (status) { std::vector<QueryPlan> pl(nq); d2h(pl.data(), b->d_plans, sizeof(QueryPlan) * (size_t)nq); for (int i = 0; i < nq; i++) { status[i] = pl[i].status; if (n[i] < 0) { status[i] |= IFX_Q_OVERFLOW; n[i] = 0; } } }
    } catch (const std::string& e) { delete b; return fail(IFX_ERR_CUDA, e); }
    delete b; return IFX_OK;
}

#include "ifx_search.inl"

// debugging aid: per-query [n_cand, n_terms, selection_ns, path, total_ns, cta] of the last run's k_stage1
extern "C" int ifx_debug_stage1_queries(ifx_batch* b, long long* out) { try { DeviceGuard dg(b->idx->device); d2h(out, b->d_qdbg, (size_t)b->nq * IFX_QDBG * 8); } catch (const std::string& e) { return fail(IFX_ERR_CUDA, e); } return IFX_OK; }
