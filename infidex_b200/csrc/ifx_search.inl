// infidex_b200 -- full search entry points: Stage 2 launches, Infiscript registration, batch run / download.

static bool host_try_parse_double(const uint16_t* s, int n, double& out) {   // double.TryParse (Float | AllowThousands, invariant)
    int b = 0, e = n; while (b < e && (s[b] == ' ' || (s[b] >= 9 && s[b] <= 13))) b++; while (e > b && (s[e - 1] == ' ' || (s[e - 1] >= 9 && s[e - 1] <= 13))) e--;
    if (b >= e) return false;
    std::string a; bool digits = false; int i = b;
    if (s[i] == '+' || s[i] == '-') { a.push_back((char)s[i]); i++; }
    while (i < e && ((s[i] >= '0' && s[i] <= '9') || s[i] == ',')) { if (s[i] != ',') { a.push_back((char)s[i]); digits = true; } i++; }
    if (i < e && s[i] == '.') { a.push_back('.'); i++; while (i < e && s[i] >= '0' && s[i] <= '9') { a.push_back((char)s[i]); digits = true; i++; } }
    if (!digits) return false;
    if (i < e && (s[i] == 'e' || s[i] == 'E')) { int j = i + 1; std::string ex = "e"; if (j < e && (s[j] == '+' || s[j] == '-')) { ex.push_back((char)s[j]); j++; } bool ed = false; while (j < e && s[j] >= '0' && s[j] <= '9') { ex.push_back((char)s[j]); ed = true; j++; } if (!ed) return false; a += ex; i = j; }
    if (i != e) return false;
    out = strtod(a.c_str(), nullptr); return true;
}

// ---- INFISCRIPT-V1 (src/Infidex/Filtering/BytecodeSerializer.cs:16-117, ConstantPool.cs:75-163) -> device program ------------
static bool op_has_operand(uint8_t op) { return op == 0x01 || op == 0x02 || op == 0x60 || op == 0x61 || op == 0x62; }
static bool op_valid(uint8_t op) {
    switch (op) { case 0x01: case 0x02: case 0x03: case 0x04: case 0x10: case 0x11: case 0x12: case 0x13: case 0x14: case 0x15: case 0x20: case 0x21: case 0x22: case 0x30: case 0x31: case 0x32:
        case 0x33: case 0x34: case 0x40: case 0x41: case 0x50: case 0x51: case 0x60: case 0x61: case 0x62: case 0xFF: return true; default: return false; }
}
static std::u16string utf8_to_u16(const uint8_t* p, size_t n) {
    std::u16string r; size_t i = 0;
    while (i < n) { uint32_t c = p[i]; int extra = 0; if (c < 0x80) extra = 0; else if ((c >> 5) == 6) { c &= 0x1F; extra = 1; } else if ((c >> 4) == 14) { c &= 0x0F; extra = 2; } else { c &= 0x07; extra = 3; }
        i++; for (int k = 0; k < extra && i < n; k++, i++) c = (c << 6) | (p[i] & 0x3F);
        if (c >= 0x10000) { c -= 0x10000; r.push_back((char16_t)(0xD800 + (c >> 10))); r.push_back((char16_t)(0xDC00 + (c & 0x3FF))); } else r.push_back((char16_t)c); }
    return r;
}

static int filter_register_impl(ifx_index* idx, const uint8_t* data, size_t len, int* out_id) {
    size_t p = 0; bool trunc = false; auto need = [&](size_t n) { return n <= len && p <= len - n; };
    auto rd_i32 = [&]() { int32_t v = 0; if (!need(4)) { trunc = true; p = len; return v; } memcpy(&v, data + p, 4); p += 4; return v; };
    auto rd_str = [&]() { uint32_t n = 0; int sh = 0; bool done = false; while (p < len && sh < 35) { uint8_t b = data[p++]; n |= (uint32_t)(b & 0x7F) << sh; if (!(b & 0x80)) { done = true; break; } sh += 7; }
                          if (!done || !need(n)) { trunc = true; p = len; return std::u16string(); } std::u16string s = utf8_to_u16(data + p, n); p += n; return s; };
    if (len < 23 || memcmp(data, "INFISCRIPT-V1", 13) != 0) return fail(IFX_ERR_INVALID, "bad INFISCRIPT magic");
    p = 13; uint16_t ver; memcpy(&ver, data + p, 2); p += 2; if (ver != 1) return fail(IFX_ERR_INVALID, "unsupported INFISCRIPT version");
    const int pool_size = rd_i32(); if (pool_size < 4 || !need((size_t)pool_size)) return fail(IFX_ERR_INVALID, "truncated constant pool");
    const size_t pool_end = p + (size_t)pool_size;
    const int cnt = rd_i32(); if (cnt < 0 || cnt > pool_size) return fail(IFX_ERR_INVALID, "constant count out of range");   // every constant takes >= 1 byte
    std::vector<FConst> consts(cnt); std::vector<uint16_t> chars; std::vector<FConst> extra;
    std::vector<std::vector<std::u16string>> arrays(cnt);
    auto add_str = [&](FConst& k, const std::u16string& s) { k.kind = 1; k.off = (int32_t)chars.size(); k.len = (int32_t)s.size(); chars.insert(chars.end(), s.begin(), s.end()); double d = 0; k.is_num = host_try_parse_double((const uint16_t*)s.data(), (int)s.size(), d) ? 1 : 0; k.num = d; k.col = -1; k.arr_start = k.arr_len = 0;
        for (size_t c = 0; c < idx->column_names.size(); c++) if (idx->column_names[c] == s) { k.col = (int32_t)c; break; } };
    for (int i = 0; i < cnt; i++) {
        if (!need(1) || p >= pool_end) return fail(IFX_ERR_INVALID, "truncated constant"); int kind = data[p++]; FConst& k = consts[i]; memset(&k, 0, sizeof(k)); k.col = -1;
        if (kind == 1) add_str(k, rd_str());
        else if (kind == 2) { if (!need(8)) return fail(IFX_ERR_INVALID, "truncated number constant"); k.kind = 2; memcpy(&k.num, data + p, 8); p += 8; k.is_num = 1; }
        else if (kind == 3) { k.kind = 3; const int n = rd_i32(); if (n < 0 || (size_t)n > len - p) return fail(IFX_ERR_INVALID, "array constant length out of range");   // every element takes >= 1 byte
            for (int j = 0; j < n && !trunc; j++) arrays[i].push_back(rd_str()); }
        else return fail(IFX_ERR_INVALID, "unknown constant type");
        if (trunc) return fail(IFX_ERR_INVALID, "truncated constant");
    }
    for (int i = 0; i < cnt; i++) if (consts[i].kind == 3) { consts[i].arr_start = cnt + (int)extra.size(); consts[i].arr_len = (int)arrays[i].size(); for (auto& s : arrays[i]) { FConst e; memset(&e, 0, sizeof(e)); add_str(e, s); extra.push_back(e); } }
    consts.insert(consts.end(), extra.begin(), extra.end());
    p = pool_end; if (!need(4)) return fail(IFX_ERR_INVALID, "truncated code"); const int ic = rd_i32();
    if (ic < 0 || (size_t)ic > len - p) return fail(IFX_ERR_INVALID, "instruction count out of range");
    std::vector<FInstr> code;
    for (int i = 0; i < ic; i++) { if (!need(1)) return fail(IFX_ERR_INVALID, "truncated code"); FInstr in; in.op = data[p++]; in.a = 0;
        if (!op_valid((uint8_t)in.op)) return fail(IFX_ERR_INVALID, "unknown opcode");
        if (op_has_operand((uint8_t)in.op)) { if (!need(4)) return fail(IFX_ERR_INVALID, "truncated operand"); in.a = rd_i32(); if (p < len && !op_valid(data[p]) && need(4)) rd_i32(); }
        if ((in.op == 0x01 || in.op == 0x02) && (in.a < 0 || in.a >= cnt)) return fail(IFX_ERR_INVALID, "constant index out of range");
        if (in.op == 0x01 && consts[in.a].kind != 1) return fail(IFX_ERR_INVALID, "PUSH_FIELD operand is not a string");
        if ((in.op == 0x60 || in.op == 0x61 || in.op == 0x62) && (in.a < 0 || in.a > ic)) return fail(IFX_ERR_INVALID, "jump target out of range");
        code.push_back(in); }
    std::lock_guard<std::mutex> lk(idx->mu);
    DeviceGuard dg(idx->device);
    FilterProg fp{}; if (chars.empty()) chars.push_back(0);
    fp.consts = idx->up(consts.data(), std::max<size_t>(consts.size(), 1)); fp.code = idx->up(code.data(), std::max<size_t>(code.size(), 1)); fp.chars = idx->up(chars.data(), chars.size());
    fp.n_consts = (int)consts.size(); fp.n_code = (int)code.size();
    idx->h_filters.push_back(fp);
    idx->d_filters = idx->up(idx->h_filters.data(), idx->h_filters.size());   // small; previous copies are released with the index
    *out_id = (int)idx->h_filters.size() - 1;
    return IFX_OK;
}

// Externally supplied bytecode: every size is validated against the remaining bytes before it is trusted, and nothing may unwind
// through the C boundary.
extern "C" int ifx_filter_register(ifx_index* idx, const uint8_t* data, size_t len, int* out_id) {
    if (!idx || !data || !out_id) return fail(IFX_ERR_INVALID, "null argument");
    try { return filter_register_impl(idx, data, len, out_id); }
    catch (const std::string& e) { return fail(IFX_ERR_CUDA, e); }
    catch (const std::bad_alloc&) { return fail(IFX_ERR_OOM, "host allocation failed while parsing the filter"); }
    catch (...) { return fail(IFX_ERR_INVALID, "malformed INFISCRIPT bytecode"); }
}

// ---- kernels ---------------------------------------------------------------------------------------------------------------
#ifndef IFX_EMU
__global__ void __launch_bounds__(256, 4) k_wm(DevIndex ix, const QueryPlan* plans, int nq, const int32_t* s1_doc, const float* s1_score, const int32_t* s1_n, int K,
                                            S1Workspace* wss, Stage2Buffers B, int* work) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WmShared& sh = *reinterpret_cast<WmShared*>(smem_raw); Ctx c; S1Workspace ws = wss[blockIdx.x];
    for (int i = threadIdx.x; i < MAX_CONTAINERS; i += blockDim.x) sh.dirty[i] = 0;
    __syncthreads();
    for (;;) {
        if (threadIdx.x == 0) sh.bcast[7] = atomicAdd(work, 1);
        __syncthreads();
        int q = sh.bcast[7]; __syncthreads();
        if (q >= nq) break;
        wm_query(c, ix, plans[q], s1_doc + (size_t)q * K, s1_score + (size_t)q * K, s1_n[q], ws, sh, B, q);
        __syncthreads();
    }
}
__global__ void k_cov_prepare(DevIndex ix, const QueryPlan* plans, int nq, Stage2Buffers B) {
    int q = blockIdx.x * blockDim.x + threadIdx.x; if (q >= nq) return;
    if (B.mode[q] == 0) prepare_cov_query(ix, plans[q].qtext, plans[q].qlen, B.covq[q]);
}
#ifndef IFX_COV_THREADS
#define IFX_COV_THREADS 512
#endif
__global__ void __launch_bounds__(IFX_COV_THREADS) k_cov_eval(DevIndex ix, const QueryPlan* plans, int nq, Stage2Buffers B) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; int q = (int)(i / B.ent_cap), e = (int)(i % B.ent_cap);
    if (q >= nq || B.mode[q] != 0 || e >= B.ent_n[q]) return;
    cov_eval_entry(ix, plans[q], B, q, e);
}
__global__ void __launch_bounds__(256) k_finalize(DevIndex ix, const QueryPlan* plans, int nq, const int32_t* s1_doc, const float* s1_score, const int32_t* s1_n, int K,
                                                  Stage2Buffers B, const FilterProg* filters, int n_filters, FinalOut O) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FinShared& sh = *reinterpret_cast<FinShared*>(smem_raw); Ctx c; int q = blockIdx.x;
    finalize_query(c, ix, plans[q], s1_doc + (size_t)q * K, s1_score + (size_t)q * K, s1_n[q], B, filters, n_filters, sh, O, q);
}
#endif

static void run_stage2_phase(ifx_batch* b, ifx_stats* st, int part = 3) {      // part: 1 = WordMatcher lookups, 2 = coverage / fusion / finalize, 3 = both
    ifx_index* ix = b->idx; const int nq = b->nq; const int K = b->s1_stride;
    Timer t;
#ifdef IFX_EMU
    static WmShared* wsh = new WmShared(); static FinShared* fsh = new FinShared(); memset(wsh->dirty, 0, sizeof(wsh->dirty));
    Ctx c;
    if (part & 1) for (int q = 0; q < nq; q++) wm_query(c, ix->v, b->d_plans[q], b->d_s1_doc + (size_t)q * K, b->d_s1_score + (size_t)q * K, b->d_s1_n[q], ix->ws[0], *wsh, b->s2, q);
    if (part & 2) for (int q = 0; q < nq; q++) if (b->s2.mode[q] == 0) { prepare_cov_query(ix->v, b->d_plans[q].qtext, b->d_plans[q].qlen, b->s2.covq[q]); for (int e = 0; e < b->s2.ent_n[q]; e++) cov_eval_entry(ix->v, b->d_plans[q], b->s2, q, e); }
    if (part & 2) for (int q = 0; q < nq; q++) finalize_query(c, ix->v, b->d_plans[q], b->d_s1_doc + (size_t)q * K, b->d_s1_score + (size_t)q * K, b->d_s1_n[q], b->s2, ix->d_filters, (int)ix->h_filters.size(), *fsh, b->fin, q);
    (void)t; (void)st;
#else
    if (!ix->attr_s2) { CUDA_TRY(cudaFuncSetAttribute(k_wm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WmShared))); CUDA_TRY(cudaFuncSetAttribute(k_finalize, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FinShared))); ix->attr_s2 = true; }
    float ms_wm = 0.f, ms_cov = 0.f, ms_fin = 0.f; int launches = 0;
    if (part & 1) {
    t.start();
    CUDA_TRY(cudaMemsetAsync(b->d_work + 2, 0, sizeof(int)));
    k_wm<<<std::min(ix->n_ctas_sel, nq), 256, sizeof(WmShared)>>>(ix->v, b->d_plans, nq, b->d_s1_doc, b->d_s1_score, b->d_s1_n, K, ix->d_ws, b->s2, b->d_work + 2);
    ms_wm = t.stop(); launches++;
    }
    if (part & 2) {
    t.start();
    k_cov_prepare<<<(nq + 63) / 64, 64>>>(ix->v, b->d_plans, nq, b->s2);
    long long total = (long long)nq * b->s2.ent_cap;
    k_cov_eval<<<(unsigned)((total + IFX_COV_THREADS - 1) / IFX_COV_THREADS), IFX_COV_THREADS>>>(ix->v, b->d_plans, nq, b->s2);
    ms_cov = t.stop();
    t.start();
    k_finalize<<<nq, 256, sizeof(FinShared)>>>(ix->v, b->d_plans, nq, b->d_s1_doc, b->d_s1_score, b->d_s1_n, K, b->s2, ix->d_filters, (int)ix->h_filters.size(), b->fin);
    ms_fin = t.stop(); launches += 3;
    }
    CUDA_TRY(cudaGetLastError());
    if (st) { st->ms_wordmatch += ms_wm; st->ms_stage2 += ms_cov; st->ms_final += ms_fin; st->kernel_launches += launches; }
#endif
}

static int alloc_stage2(ifx_batch* b, int cap, int fcap) {
    const size_t nq = b->nq; const size_t ec = 2 * (size_t)b->depth_max;
    Stage2Buffers& S = b->s2; S.ent_cap = (int)ec;
    S.ent_doc = b->alloc<int32_t>(nq * ec); S.ent_base = b->alloc<float>(nq * ec); S.ent_twin = b->alloc<int32_t>(nq * ec); S.ent_n = b->alloc<int32_t>(nq);
    S.ent_score = b->alloc<float>(nq * ec); S.ent_tie = b->alloc<uint8_t>(nq * ec); S.ent_hits = b->alloc<int32_t>(nq * ec); S.ent_lcs = b->alloc<uint8_t>(nq * ec);
    S.di_doc = b->alloc<int32_t>(nq * 2); S.wm_cnt = b->alloc<int32_t>(nq * 4); b->d_g_di = b->alloc<int32_t>(nq * 2); S.g_di = nullptr; S.wm_any = b->alloc<int32_t>(nq); S.mode = b->alloc<int32_t>(nq); S.covq = b->alloc<CovQuery>(nq);
    FinalOut& O = b->fin; O.cap = cap; O.fcap = fcap; size_t fc = std::max(fcap, 1);
    O.key = b->alloc<int64_t>(nq * cap); O.score = b->alloc<float>(nq * cap); O.tie = b->alloc<uint8_t>(nq * cap); O.n = b->alloc<int32_t>(nq); O.total = b->alloc<int32_t>(nq); O.status = b->alloc<int32_t>(nq);
    b->d_shard_info = b->alloc<int32_t>(nq * 8); b->d_shard_dkey = b->alloc<int64_t>(nq * 2); O.shard_info = nullptr; O.shard_dkey = nullptr;
    O.facet_col = b->alloc<int32_t>(nq * fc); O.facet_val = b->alloc<int32_t>(nq * fc); O.facet_cnt = b->alloc<int32_t>(nq * fc); O.n_facets = b->alloc<int32_t>(nq);
    return IFX_OK;
}

extern "C" int ifx_batch_run(ifx_batch* b, ifx_stats* st) {
    if (!b) return fail(IFX_ERR_INVALID, "null batch");
    if (!dev_ok()) return fail(IFX_ERR_NO_DEVICE, "no CUDA device available (infidex_b200 has no CPU fallback)");
    if (st) { int64_t h = st->h2d_bytes, d = st->d2h_bytes; memset(st, 0, sizeof(*st)); st->h2d_bytes = h; st->d2h_bytes = d; }
    try {
        std::lock_guard<std::mutex> lk(b->idx->mu); DeviceGuard dg(b->idx->device);
        if (!b->s2.ent_doc) alloc_stage2(b, std::max(b->cap_max, 1), b->fcap);
        b->s2.gmax = nullptr;
        Timer tt; tt.start();
        run_stage1_phase(b, st);
        run_short_queries(b, st); b->s2.s1_total = b->d_s1_total;
        run_stage2_phase(b, st);
        float ms = tt.stop(); if (st) st->ms_total = ms;
        b->ran = true;
    } catch (const std::string& e) { return fail(IFX_ERR_CUDA, e); }
    return IFX_OK;
}

// ---- doc-id-range shards: the batch run split at the points where the shards' hosts exchange data -------------------------
//   phase 1  query preparation + LD1 expansion            -> ifx_batch_fuzzy_df(get) -> all-reduce(sum) -> ifx_batch_fuzzy_df(set)
//   phase 2  selection, tf lookups, scoring, final order  -> ifx_batch_stage1_lists -> all-gather -> ifx_batch_stage1_restrict (global cut, global top score)
//   phase 3  WordMatcher, coverage / fusion, truncation, filter, facets -> ifx_batch_download -> all-gather of the shard's records -> merge on every host
// Buffers passed to the exchange entry points are DEVICE pointers (host pointers in the test-only emulation build).
#ifndef IFX_EMU
__global__ void k_fuzzy_df(DevIndex ix, QueryPlan* plans, int nq, int32_t* buf, int set) {
    int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= nq * MAX_FUZZY) return;
    const int q = i / MAX_FUZZY, f = i % MAX_FUZZY; QueryPlan& p = plans[q];
    if (f >= p.n_fuzzy) { if (!set) buf[i] = 0; return; }
    QTerm& t = p.terms[p.fuzzy[f].term_slot];
    if (!set) { buf[i] = t.df; return; }
    const float avgdl = ix.avgdl > 0.f ? ix.avgdl : 1.f; t.df = buf[i]; t.idf = compute_idf(ix, t.df); t.max_score = max_term_score(t.idf, avgdl);
}
__global__ void k_s1_restrict(int nq, int K, int64_t* key, int32_t* doc, float* score, int32_t* n, const uint8_t* keep, const int32_t* n_global, int32_t* g_di) {
    int q = blockIdx.x * blockDim.x + threadIdx.x; if (q >= nq) return;
    int m = 0; const int cnt = n[q] < 0 ? 0 : n[q]; int d0 = n_global[q] >= 2 ? -1 : -2, d1 = d0;
    for (int i = 0; i < cnt; i++) { const uint8_t k = keep[(size_t)q * K + i]; if (!k) continue; const size_t o = (size_t)q * K;
        if (k == 2 && d0 != -2) d0 = doc[o + i]; if (k == 3 && d1 != -2) d1 = doc[o + i];
        key[o + m] = key[o + i]; doc[o + m] = doc[o + i]; score[o + m] = score[o + i]; m++; }
    if (n[q] >= 0) n[q] = m;
    g_di[q * 2] = d0; g_di[q * 2 + 1] = d1;
}
__global__ void k_wm_apply(int nq, int cap, Stage2Buffers B, const int32_t* allowed, const int32_t* any) {
    int q = blockIdx.x * blockDim.x + threadIdx.x; if (q >= nq) return;
    if (B.mode[q] != 0) return;
    const int first = B.wm_cnt[q * 4 + 3], taken = B.wm_cnt[q * 4 + 1]; const int a = allowed[q] < taken ? (allowed[q] > 0 ? allowed[q] : 0) : taken;
    for (int i = first + a; i < first + taken && i < cap; i++) B.ent_twin[(size_t)q * cap + i] = -3;      // beyond this shard's part of the global WordMatcher quota
    B.wm_any[q] = any[q];
}
#endif
extern "C" int ifx_batch_run_phase(ifx_batch* b, int phase, ifx_stats* st) {
    if (!b || phase < 1 || phase > 5) return fail(IFX_ERR_INVALID, "bad phase");
    if (!dev_ok()) return fail(IFX_ERR_NO_DEVICE, "no CUDA device available (infidex_b200 has no CPU fallback)");
    try {
        std::lock_guard<std::mutex> lk(b->idx->mu); DeviceGuard dg(b->idx->device);
        if (!b->s2.ent_doc) alloc_stage2(b, std::max(b->cap_max, 1), b->fcap);
        if (phase == 1) { if (st) { int64_t h = st->h2d_bytes, d = st->d2h_bytes; memset(st, 0, sizeof(*st)); st->h2d_bytes = h; st->d2h_bytes = d; } b->s2.gmax = nullptr; b->s2.g_di = nullptr; b->fin.shard_info = nullptr; b->fin.shard_dkey = nullptr; b->use_gcnt = false; run_stage1_phase(b, st, 1); }
        else if (phase == 2) run_stage1_phase(b, st, 4);
        else if (phase == 3) { run_stage1_phase(b, st, 2); b->s2.s1_total = nullptr; }      // (short queries are not run on shards: they keep an empty list and IFX_Q_UNSUPPORTED_OP)
        else if (phase == 4) run_stage2_phase(b, st, 1);
        else { b->fin.shard_info = b->d_shard_info; b->fin.shard_dkey = b->d_shard_dkey; run_stage2_phase(b, st, 2); b->ran = true; }
    } catch (const std::string& e) { return fail(IFX_ERR_CUDA, e); }
    return IFX_OK;
}
extern "C" int ifx_batch_fuzzy_df(ifx_batch* b, int32_t* buf, int set) {
    if (!b || !buf) return fail(IFX_ERR_INVALID, "null argument");
    try { DeviceGuard dg(b->idx->device);
#ifdef IFX_EMU
        for (int i = 0; i < b->nq * MAX_FUZZY; i++) { const int q = i / MAX_FUZZY, f = i % MAX_FUZZY; QueryPlan& p = b->d_plans[q]; if (f >= p.n_fuzzy) { if (!set) buf[i] = 0; continue; }
            QTerm& t = p.terms[p.fuzzy[f].term_slot]; if (!set) { buf[i] = t.df; continue; }
            const float avgdl = b->idx->v.avgdl > 0.f ? b->idx->v.avgdl : 1.f; t.df = buf[i]; t.idf = compute_idf(b->idx->v, t.df); t.max_score = max_term_score(t.idf, avgdl); }
#else
        k_fuzzy_df<<<(b->nq * MAX_FUZZY + 255) / 256, 256>>>(b->idx->v, b->d_plans, b->nq, buf, set); CUDA_TRY(cudaGetLastError()); CUDA_TRY(cudaDeviceSynchronize());
#endif
    } catch (const std::string& e) { return fail(IFX_ERR_CUDA, e); }
    return IFX_OK;
}
extern "C" int ifx_batch_stage1_lists(ifx_batch* b, int64_t* key, float* score, int32_t* n) {      // [nq][depth] keys / scores, [nq] counts of this shard's Stage-1 lists
    if (!b || !key || !score || !n) return fail(IFX_ERR_INVALID, "null argument");
    try { DeviceGuard dg(b->idx->device); const size_t m = (size_t)b->nq * b->s1_stride;
#ifdef IFX_EMU
        memcpy(key, b->d_s1_key, m * 8); memcpy(score, b->d_s1_score, m * 4); memcpy(n, b->d_s1_n, (size_t)b->nq * 4);
#else
        CUDA_TRY(cudaMemcpy(key, b->d_s1_key, m * 8, cudaMemcpyDeviceToDevice)); CUDA_TRY(cudaMemcpy(score, b->d_s1_score, m * 4, cudaMemcpyDeviceToDevice)); CUDA_TRY(cudaMemcpy(n, b->d_s1_n, (size_t)b->nq * 4, cudaMemcpyDeviceToDevice));
#endif
    } catch (const std::string& e) { return fail(IFX_ERR_CUDA, e); }
    return IFX_OK;
}
// keep[q][i] != 0: entry i of this shard's list of query q is inside the global top-`depth`; gmax[q]: top Stage-1 score over all shards (borrowed
// until phase 3 has run).
extern "C" int ifx_batch_stage1_restrict(ifx_batch* b, const uint8_t* keep, const float* gmax, const int32_t* n_global) {
    if (!b || !keep || !gmax || !n_global) return fail(IFX_ERR_INVALID, "null argument");
    try { DeviceGuard dg(b->idx->device); const int K = b->s1_stride;
#ifdef IFX_EMU
        for (int q = 0; q < b->nq; q++) { int m = 0; const int cnt = b->d_s1_n[q] < 0 ? 0 : b->d_s1_n[q]; const size_t o = (size_t)q * K; int d0 = n_global[q] >= 2 ? -1 : -2, d1 = d0;
            for (int i = 0; i < cnt; i++) { const uint8_t k = keep[o + i]; if (!k) continue; if (k == 2 && d0 != -2) d0 = b->d_s1_doc[o + i]; if (k == 3 && d1 != -2) d1 = b->d_s1_doc[o + i];
                b->d_s1_key[o + m] = b->d_s1_key[o + i]; b->d_s1_doc[o + m] = b->d_s1_doc[o + i]; b->d_s1_score[o + m] = b->d_s1_score[o + i]; m++; }
            if (b->d_s1_n[q] >= 0) b->d_s1_n[q] = m; b->d_g_di[q * 2] = d0; b->d_g_di[q * 2 + 1] = d1; }
#else
        k_s1_restrict<<<(b->nq + 127) / 128, 128>>>(b->nq, K, b->d_s1_key, b->d_s1_doc, b->d_s1_score, b->d_s1_n, keep, n_global, b->d_g_di); CUDA_TRY(cudaGetLastError()); CUDA_TRY(cudaDeviceSynchronize());
#endif
        b->s2.gmax = gmax; b->s2.g_di = b->d_g_di;
    } catch (const std::string& e) { return fail(IFX_ERR_CUDA, e); }
    return IFX_OK;
}

// the selector's cardinalities (SEL_CNT int32 per query): get after phase 2, set the sums over all shards before phase 3
extern "C" int ifx_batch_select_counts(ifx_batch* b, int32_t* buf, int set) {
    if (!b || !buf) return fail(IFX_ERR_INVALID, "null argument");
    try { DeviceGuard dg(b->idx->device); const size_t n = (size_t)b->nq * SEL_CNT * 4;
#ifdef IFX_EMU
        if (set) memcpy(b->d_sel_cnt, buf, n); else memcpy(buf, b->d_sel_cnt, n);
#else
        if (set) CUDA_TRY(cudaMemcpy(b->d_sel_cnt, buf, n, cudaMemcpyDeviceToDevice)); else CUDA_TRY(cudaMemcpy(buf, b->d_sel_cnt, n, cudaMemcpyDeviceToDevice));
#endif
        if (set) b->use_gcnt = true;
    } catch (const std::string& e) { return fail(IFX_ERR_CUDA, e); }
    return IFX_OK;
}
// WordMatcher quota across shards: counts out after phase 4 (k_wm) -- [nq][4]: overlap, WordMatcher-only entries taken, union non-empty, -- ;
// `allowed`[q] = how many of this shard's WordMatcher-only entries fall inside the global quota, `any`[q] = union non-empty on any shard.
extern "C" int ifx_batch_wm_counts(ifx_batch* b, int32_t* buf) {
    if (!b || !buf) return fail(IFX_ERR_INVALID, "null argument");
    try { DeviceGuard dg(b->idx->device);
#ifdef IFX_EMU
        memcpy(buf, b->s2.wm_cnt, (size_t)b->nq * 16);
#else
        CUDA_TRY(cudaMemcpy(buf, b->s2.wm_cnt, (size_t)b->nq * 16, cudaMemcpyDeviceToDevice));
#endif
    } catch (const std::string& e) { return fail(IFX_ERR_CUDA, e); }
    return IFX_OK;
}
extern "C" int ifx_batch_wm_apply(ifx_batch* b, const int32_t* allowed, const int32_t* any) {
    if (!b || !allowed || !any) return fail(IFX_ERR_INVALID, "null argument");
    try { DeviceGuard dg(b->idx->device);
#ifdef IFX_EMU
        for (int q = 0; q < b->nq; q++) { if (b->s2.mode[q] != 0) continue; const int first = b->s2.wm_cnt[q * 4 + 3], taken = b->s2.wm_cnt[q * 4 + 1]; const int a = allowed[q] < taken ? (allowed[q] > 0 ? allowed[q] : 0) : taken;
            for (int i = first + a; i < first + taken && i < b->s2.ent_cap; i++) b->s2.ent_twin[(size_t)q * b->s2.ent_cap + i] = -3; b->s2.wm_any[q] = any[q]; }
#else
        k_wm_apply<<<(b->nq + 127) / 128, 128>>>(b->nq, b->s2.ent_cap, b->s2, allowed, any); CUDA_TRY(cudaGetLastError()); CUDA_TRY(cudaDeviceSynchronize());
#endif
    } catch (const std::string& e) { return fail(IFX_ERR_CUDA, e); }
    return IFX_OK;
}
// per query [max word hits, records with Score >= 254, word hits / lcs of the docIndex-0 and docIndex-1 documents when this shard owns them (-1 else)]
// and their keys: what the hosts need to evaluate ResultProcessor.CalculateTruncationIndex over the merged list (valid after phase 4)
extern "C" int ifx_batch_shard_info(ifx_batch* b, int32_t* info /* [nq][8] */, int64_t* dkey /* [nq][2] */) {
    if (!b || !info || !dkey) return fail(IFX_ERR_INVALID, "null argument");
    try { DeviceGuard dg(b->idx->device); d2h(info, b->d_shard_info, (size_t)b->nq * 32); d2h(dkey, b->d_shard_dkey, (size_t)b->nq * 16); } catch (const std::string& e) { return fail(IFX_ERR_CUDA, e); }
    return IFX_OK;
}

extern "C" int ifx_batch_download(ifx_batch* b, ifx_batch_result* out) {
    if (!b || !out || !b->ran) return fail(IFX_ERR_INVALID, "batch has not been run");
    if (out->cap < b->fin.cap) return fail(IFX_ERR_INVALID, "result capacity smaller than the batch's max_results");
    const size_t nq = b->nq; const int cap = b->fin.cap;
    try {
        DeviceGuard dg(b->idx->device);
        if (out->cap == cap) { d2h(out->doc_key, b->fin.key, nq * cap * 8); d2h(out->score, b->fin.score, nq * cap * 4); d2h(out->tie, b->fin.tie, nq * cap); }
        else { std::vector<int64_t> k(nq * cap); std::vector<float> s(nq * cap); std::vector<uint8_t> t(nq * cap); d2h(k.data(), b->fin.key, nq * cap * 8); d2h(s.data(), b->fin.score, nq * cap * 4); d2h(t.data(), b->fin.tie, nq * cap);
            for (size_t q = 0; q < nq; q++) { memcpy(out->doc_key + q * out->cap, k.data() + q * cap, cap * 8); memcpy(out->score + q * out->cap, s.data() + q * cap, cap * 4); memcpy(out->tie + q * out->cap, t.data() + q * cap, cap); } }
        d2h(out->n, b->fin.n, nq * 4); d2h(out->total_candidates, b->fin.total, nq * 4); d2h(out->status, b->fin.status, nq * 4);
        if (out->n_facets) { d2h(out->n_facets, b->fin.n_facets, nq * 4);
            if (out->facet_cap > 0 && b->fin.fcap > 0 && out->facet_column) {
                if (out->facet_cap == b->fin.fcap) { size_t n = nq * b->fin.fcap * 4; d2h(out->facet_column, b->fin.facet_col, n); d2h(out->facet_value, b->fin.facet_val, n); d2h(out->facet_count, b->fin.facet_cnt, n); }
                else return fail(IFX_ERR_INVALID, "facet_cap mismatch"); } }
    } catch (const std::string& e) { return fail(IFX_ERR_CUDA, e); }
    return IFX_OK;
}

extern "C" int ifx_search_batch(ifx_index* idx, const ifx_query* q, int nq, ifx_batch_result* out, ifx_stats* st) {
    if (!idx || !q || nq <= 0 || !out) return fail(IFX_ERR_INVALID, "bad arguments");
    if (!dev_ok()) return fail(IFX_ERR_NO_DEVICE, "no CUDA device available (infidex_b200 has no CPU fallback)");
    std::lock_guard<std::mutex> call_lock(idx->call_mu);      // one batch in flight per index (callers queue, like writers on the C# RW lock)
    int rc = IFX_OK;
    try {
        DeviceGuard dg(idx->device);
        ifx_batch* b = idx->cached;
        if (b) { rc = fill_batch(b, q, nq); if (rc || b->fcap != out->facet_cap) { delete b; b = nullptr; idx->cached = nullptr; } }
        if (!b) { b = new ifx_batch(); b->idx = idx; b->fcap = out->facet_cap; rc = fill_batch(b, q, nq); if (rc) { delete b; return rc; } idx->cached = b; }
        if (st) { memset(st, 0, sizeof(*st)); int64_t h = 0; for (int i = 0; i < nq; i++) h += 2LL * q[i].len; st->h2d_bytes = h + (int64_t)nq * 28 + 8; }
        rc = ifx_batch_run(b, st);
        if (!rc) rc = ifx_batch_download(b, out);
        if (!rc && st) st->d2h_bytes = (int64_t)nq * ((int64_t)b->fin.cap * 13 + 12 + (out->facet_cap > 0 ? 12LL * out->facet_cap + 4 : 0));
    } catch (const std::string& e) { return fail(IFX_ERR_CUDA, e); }
    return rc;
}

// Benchmark hygiene helper: evict the L2 (126 MB on B200) by overwriting a 256 MiB scratch buffer.
extern "C" int ifx_flush_l2(ifx_index* idx) {
    if (!idx) return fail(IFX_ERR_INVALID, "null index");
    try { DeviceGuard dg(idx->device); if (!idx->d_flush) idx->d_flush = idx->alloc<uint8_t>((size_t)256 << 20); dev_zero(idx->d_flush, (size_t)256 << 20);
#ifndef IFX_EMU
        CUDA_TRY(cudaDeviceSynchronize());
#endif
    } catch (const std::string& e) { return fail(IFX_ERR_CUDA, e); }
    return IFX_OK;
}
