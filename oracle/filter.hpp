// ORACLE (test infrastructure, NOT product code): CPU restatement of the Infiscript bytecode VM, its wire
// format, and facet counting.
//
// Follows (under /root/reference/src/Infidex):
//   Filtering/BytecodeInstruction.cs:6-49      opcodes
//   Filtering/BytecodeSerializer.cs:16-117     "INFISCRIPT-V1" wire format
//   Filtering/ConstantPool.cs:75-163           constant pool (type tags 1/2/3, 7-bit length-prefixed UTF-8)
//   Filtering/FilterVM.cs:26-357               stack VM semantics (peeking jumps, TryParse-then-string compare)
//   Scoring/ResultProcessor.cs:35-70           ApplyFilter over result records
//   Core/FacetBuilder.cs:19-105                facet counts (top 100 by count desc, key asc)
#pragma once
#include "stage2.hpp"

namespace ifxo {

struct Const { int kind = 1; str s; double d = 0; std::vector<str> arr; };   // 1 string, 2 number, 3 array
struct Instr { uint8_t op; int a = 0; };
struct CompiledFilter { std::vector<Const> consts; std::vector<Instr> code; bool ok = false; };

inline str utf8_to_utf16(const uint8_t* p, size_t n) {
    str r; size_t i = 0;
    while (i < n) {
        uint32_t c = p[i]; int extra = 0;
        if (c < 0x80) extra = 0; else if ((c >> 5) == 6) { c &= 0x1F; extra = 1; } else if ((c >> 4) == 14) { c &= 0x0F; extra = 2; } else { c &= 0x07; extra = 3; }
        i++; for (int k = 0; k < extra && i < n; k++, i++) c = (c << 6) | (p[i] & 0x3F);
        if (c >= 0x10000) { c -= 0x10000; r.push_back((char16_t)(0xD800 + (c >> 10))); r.push_back((char16_t)(0xDC00 + (c & 0x3FF))); } else r.push_back((char16_t)c);
    }
    return r;
}

inline bool op_has_operand(uint8_t op) { return op == 0x01 || op == 0x02 || op == 0x60 || op == 0x61 || op == 0x62; }
inline bool op_valid(uint8_t op) {
    switch (op) { case 0x01: case 0x02: case 0x03: case 0x04: case 0x10: case 0x11: case 0x12: case 0x13: case 0x14: case 0x15: case 0x20: case 0x21: case 0x22:
        case 0x30: case 0x31: case 0x32: case 0x33: case 0x34: case 0x40: case 0x41: case 0x50: case 0x51: case 0x60: case 0x61: case 0x62: case 0xFF: return true; default: return false; }
}

inline CompiledFilter deserialize_filter(const uint8_t* data, size_t len) {   // BytecodeSerializer.Deserialize
    CompiledFilter f; size_t p = 0;
    auto need = [&](size_t n) { return p + n <= len; };
    auto rd_i32 = [&]() { int32_t v; std::memcpy(&v, data + p, 4); p += 4; return v; };
    auto rd_str = [&]() { uint32_t n = 0; int sh = 0; while (p < len) { uint8_t b = data[p++]; n |= (uint32_t)(b & 0x7F) << sh; if (!(b & 0x80)) break; sh += 7; } str s = utf8_to_utf16(data + p, std::min<size_t>(n, len - p)); p += n; return s; };
    const char* magic = "INFISCRIPT-V1";
    if (len < 15 || std::memcmp(data, magic, 13) != 0) return f;
    p = 13; uint16_t ver; std::memcpy(&ver, data + p, 2); p += 2; if (ver != 1) return f;
    if (!need(4)) return f; int pool_size = rd_i32(); size_t pool_end = p + pool_size; if (pool_end > len) return f;
    int cnt = rd_i32();
    for (int i = 0; i < cnt; i++) {
        Const c; c.kind = data[p++];
        if (c.kind == 1) c.s = rd_str();
        else if (c.kind == 2) { std::memcpy(&c.d, data + p, 8); p += 8; }
        else if (c.kind == 3) { int n = rd_i32(); for (int j = 0; j < n; j++) c.arr.push_back(rd_str()); }
        else return f;
        f.consts.push_back(std::move(c));
    }
    p = pool_end;
    if (!need(4)) return f; int ic = rd_i32();
    for (int i = 0; i < ic; i++) {
        if (!need(1)) return f; Instr in; in.op = data[p++];
        if (op_has_operand(in.op)) { if (!need(4)) return f; in.a = rd_i32();
            if (p < len && !op_valid(data[p])) { if (need(4)) rd_i32(); }   // optional Operand2 (never emitted by the compiler)
        }
        f.code.push_back(in);
    }
    f.ok = true; return f;
}

// double.TryParse(string) -- NumberStyles.Float | AllowThousands, invariant culture
inline bool try_parse_double(sv s, double& out) {
    size_t b = 0, e = s.size();
    while (b < e && is_space(s[b])) b++; while (e > b && is_space(s[e - 1])) e--;
    if (b >= e) return false;
    std::string a; bool digits = false; size_t i = b;
    if (s[i] == u'+' || s[i] == u'-') { a.push_back((char)s[i]); i++; }
    while (i < e && ((s[i] >= u'0' && s[i] <= u'9') || s[i] == u',')) { if (s[i] != u',') { a.push_back((char)s[i]); digits = true; } i++; }
    if (i < e && s[i] == u'.') { a.push_back('.'); i++; while (i < e && s[i] >= u'0' && s[i] <= u'9') { a.push_back((char)s[i]); digits = true; i++; } }
    if (!digits) return false;
    if (i < e && (s[i] == u'e' || s[i] == u'E')) { size_t j = i + 1; std::string ex = "e"; if (j < e && (s[j] == u'+' || s[j] == u'-')) { ex.push_back((char)s[j]); j++; }
        bool ed = false; while (j < e && s[j] >= u'0' && s[j] <= u'9') { ex.push_back((char)s[j]); ed = true; j++; } if (!ed) return false; a += ex; i = j; }
    if (i != e) return false;
    out = std::strtod(a.c_str(), nullptr); return true;
}

struct FilterVM {
    struct V { int kind = 0; str s; double d = 0; bool b = false; long long i = 0; const std::vector<str>* arr = nullptr;   // 0 null 1 string 2 int 3 double 4 bool 5 array
        str to_string() const { Value v; v.kind = kind == 5 ? 0 : kind; v.s = s; v.d = d; v.b = b; v.i = i; return v.to_string(); } };
    bool unsupported = false;
    static bool are_equal(const V& l, const V& r) { if (l.kind == 0 && r.kind == 0) return true; if (l.kind == 0 || r.kind == 0) return false; return eq_ic(l.to_string(), r.to_string()); }
    static int compare(const V& l, const V& r) {
        if (l.kind == 0 && r.kind == 0) return 0; if (l.kind == 0) return -1; if (r.kind == 0) return 1;
        str ls = l.to_string(), rs = r.to_string(); double a, b;
        if (try_parse_double(ls, a) && try_parse_double(rs, b)) return a < b ? -1 : (a > b ? 1 : 0);
        return cmp_ic(ls, rs);
    }
    static bool like(sv text, sv pat) {   // ^escape(pat) with % -> .*, _ -> . $ , IgnoreCase
        size_t n = text.size(), m = pat.size(); std::vector<char> dp(m + 1, 0), nx(m + 1, 0);
        // dp over text positions: classic wildcard match
        std::vector<std::vector<char>> t(n + 1, std::vector<char>(m + 1, 0)); t[0][0] = 1;
        for (size_t j = 1; j <= m; j++) t[0][j] = t[0][j - 1] && pat[j - 1] == u'%';
        for (size_t i = 1; i <= n; i++) for (size_t j = 1; j <= m; j++) {
            char16_t pc = pat[j - 1];
            if (pc == u'%') t[i][j] = t[i][j - 1] || (t[i - 1][j] && text[i - 1] != u'\n');
            else if (pc == u'_') t[i][j] = t[i - 1][j - 1] && text[i - 1] != u'\n';
            else t[i][j] = t[i - 1][j - 1] && up(pc) == up(text[i - 1]);
        }
        return t[n][m];
    }
    bool execute(const CompiledFilter& f, const Index& ix, int doc) {
        std::vector<V> st; size_t ip = 0;
        auto field = [&](const str& name) { V v; for (size_t k = 0; k < ix.schema.size(); k++) if (ix.schema[k].name == name) { const Value& x = ix.docs[doc].values[k]; v.kind = x.kind; v.s = x.s; v.d = x.d; v.b = x.b; v.i = x.i; break; } return v; };
        auto pop = [&]() { V v = st.back(); st.pop_back(); return v; };
        auto as_bool = [](const V& v) { return v.kind == 4 && v.b; };
        auto push_b = [&](bool b) { V v; v.kind = 4; v.b = b; st.push_back(v); };
        auto sstr = [](const V& v) { return v.kind == 0 ? str() : v.to_string(); };
        while (ip < f.code.size()) {
            const Instr& in = f.code[ip];
            switch (in.op) {
                case 0x01: st.push_back(field(f.consts[in.a].s)); break;
                case 0x02: { const Const& c = f.consts[in.a]; V v; if (c.kind == 1) { v.kind = 1; v.s = c.s; } else if (c.kind == 2) { v.kind = 3; v.d = c.d; } else { v.kind = 5; v.arr = &c.arr; } st.push_back(v); break; }
                case 0x03: st.pop_back(); break;
                case 0x04: st.push_back(st.back()); break;
                case 0x10: { V r = pop(), l = pop(); push_b(are_equal(l, r)); break; }
                case 0x11: { V r = pop(), l = pop(); push_b(!are_equal(l, r)); break; }
                case 0x12: { V r = pop(), l = pop(); push_b(compare(l, r) < 0); break; }
                case 0x13: { V r = pop(), l = pop(); push_b(compare(l, r) <= 0); break; }
                case 0x14: { V r = pop(), l = pop(); push_b(compare(l, r) > 0); break; }
                case 0x15: { V r = pop(), l = pop(); push_b(compare(l, r) >= 0); break; }
                case 0x20: { bool r = as_bool(pop()), l = as_bool(pop()); push_b(l && r); break; }
                case 0x21: { bool r = as_bool(pop()), l = as_bool(pop()); push_b(l || r); break; }
                case 0x22: { bool v = as_bool(pop()); push_b(!v); break; }
                case 0x30: { str p = sstr(pop()), t = sstr(pop()); push_b(contains_ic(t, p)); break; }
                case 0x31: { str p = sstr(pop()), t = sstr(pop()); push_b(starts_ic(t, p)); break; }
                case 0x32: { str p = sstr(pop()), t = sstr(pop()); push_b(ends_ic(t, p)); break; }
                case 0x33: { str p = sstr(pop()), t = sstr(pop()); push_b(like(t, p)); break; }
                case 0x34: { pop(); pop(); unsupported = true; push_b(false); break; }   // MATCHES (regex): SURVEY 8(f) "next"
                case 0x40: { V a = pop(), v = pop(); bool found = false; if (a.kind == 5) for (auto& it : *a.arr) { V x; x.kind = 1; x.s = it; if (are_equal(v, x)) { found = true; break; } } push_b(found); break; }
                case 0x41: { V mx = pop(), mn = pop(), v = pop(); push_b(compare(v, mn) >= 0 && compare(v, mx) <= 0); break; }
                case 0x50: { V v = pop(); push_b(v.kind == 0 || (v.kind == 1 && v.s.empty())); break; }
                case 0x51: { V v = pop(); push_b(!(v.kind == 0 || (v.kind == 1 && v.s.empty()))); break; }
                case 0x60: ip = (size_t)in.a - 1; break;
                case 0x61: { const V& v = st.back(); if (v.kind == 4 && !v.b) ip = (size_t)in.a - 1; break; }
                case 0x62: { const V& v = st.back(); if (v.kind == 4 && v.b) ip = (size_t)in.a - 1; break; }
                case 0xFF: ip = f.code.size(); break;
                default: unsupported = true; return false;
            }
            ip++;
        }
        if (st.empty()) return false;
        return as_bool(st.back());
    }
};

struct FacetEntry { str field, value; int count; };

// Approximation of the culture-sensitive default string order used by `.ThenBy(kvp => kvp.Key)`:
// case-insensitive primary order, lower-case first on ties; exact for same-case ASCII keys.
inline bool facet_key_less(const str& a, const str& b) { int c = cmp_ic(a, b); if (c != 0) return c < 0; return a > b; }

inline std::vector<FacetEntry> build_facets(const Index& ix, const std::vector<ScoreEntry>& results, int max_per_field = 100) {
    std::vector<FacetEntry> out;
    if (results.empty()) return out;
    for (size_t k = 0; k < ix.schema.size(); k++) {
        if (!ix.schema[k].facetable) continue;
        std::vector<std::pair<str, int>> counts;   // insertion-ordered
        for (auto& r : results) {
            int id = ix.doc_by_key(r.key); if (id < 0) continue;
            const Value& v = ix.docs[id].values[k]; if (v.is_null()) continue;
            str s = v.to_string(); if (s.empty()) continue;
            bool found = false; for (auto& c : counts) if (c.first == s) { c.second++; found = true; break; }
            if (!found) counts.emplace_back(s, 1);
        }
        std::stable_sort(counts.begin(), counts.end(), [](auto& a, auto& b) { if (a.second != b.second) return a.second > b.second; return facet_key_less(a.first, b.first); });
        if ((int)counts.size() > max_per_field) counts.resize(max_per_field);
        for (auto& c : counts) out.push_back({ix.schema[k].name, c.first, c.second});
    }
    return out;
}

}  // namespace ifxo
