// infidex_b200 -- full search (Stage 2 + filter + facets) entry points. (Filled in by the Stage-2 milestone.)
static bool host_try_parse_double(const uint16_t* s, int n, double& out) {
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
extern "C" int ifx_filter_register(ifx_index*, const uint8_t*, size_t, int*) { return fail(IFX_ERR_UNSUPPORTED, "filter VM not built yet"); }
extern "C" int ifx_batch_run(ifx_batch*, ifx_stats*) { return fail(IFX_ERR_UNSUPPORTED, "stage 2 not built yet"); }
extern "C" int ifx_batch_download(ifx_batch*, ifx_batch_result*) { return fail(IFX_ERR_UNSUPPORTED, "stage 2 not built yet"); }
extern "C" int ifx_search_batch(ifx_index*, const ifx_query*, int, ifx_batch_result*, ifx_stats*) { return fail(IFX_ERR_UNSUPPORTED, "stage 2 not built yet"); }
