// ORACLE (test infrastructure, NOT product code): CPU restatement of lofcz/Infidex text handling.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
//
// Follows (reference file:line, under /root/reference/src/Infidex):
//   Tokenization/TextNormalizer.cs:120-200,216-299   normalize()
//   Tokenization/TokenizerSetup.cs:36-42              delimiters
//   Tokenization/Tokenizer.cs:89-139,144-200,276-327  index / search token enumeration
//   .NET char.ToLowerInvariant / ToUpperInvariant / OrdinalIgnoreCase semantics (simple 1:1 maps)
#pragma once
#include <cstdint>
#include <string>
#include <string_view>
#include <vector>
#include <algorithm>
#include <functional>

namespace ifxo {

using str = std::u16string;
using sv = std::u16string_view;

#include "chartables.inc"

struct CharTables {
    std::vector<char16_t> lower, upper, norm;
    std::vector<uint8_t> flags;  // bit0 letter, bit1 whitespace, bit2 delimiter
    CharTables() : lower(65536), upper(65536), norm(65536), flags(65536, 0) {
        for (int i = 0; i < 65536; i++) lower[i] = upper[i] = norm[i] = (char16_t)i;
        for (int i = 0; i < IFX_LOWER_PAIRS_N; i++) lower[IFX_LOWER_PAIRS[i][0]] = IFX_LOWER_PAIRS[i][1];
        for (int i = 0; i < IFX_UPPER_PAIRS_N; i++) upper[IFX_UPPER_PAIRS[i][0]] = IFX_UPPER_PAIRS[i][1];
        for (int i = 0; i < IFX_NORM_PAIRS_N; i++) norm[IFX_NORM_PAIRS[i][0]] = IFX_NORM_PAIRS[i][1];
        for (int i = 0; i < IFX_LETTER_RANGES_N; i++)
            for (int c = IFX_LETTER_RANGES[i][0]; c <= IFX_LETTER_RANGES[i][1]; c++) flags[c] |= 1;
        for (int i = 0; i < IFX_SPACE_LIST_N; i++) flags[IFX_SPACE_LIST[i]] |= 2;
        // TokenizerSetup.cs:36-42
        const char16_t d[] = {u' ', u'-', u'/', u'.', u',', u':', u';', u'\'', u'`', 0x2013, 0x2014,
                              u'*', u'&', u'\\', u'_', u'(', u')', u'{', u'}', u'[', u']', u'\t'};
        for (char16_t c : d) flags[c] |= 4;
    }
};
inline const CharTables& tables() { static CharTables t; return t; }

inline char16_t lo(char16_t c) { return tables().lower[c]; }
inline char16_t up(char16_t c) { return tables().upper[c]; }
inline bool is_letter(char16_t c) { return tables().flags[c] & 1; }
inline bool is_space(char16_t c) { return tables().flags[c] & 2; }
inline bool is_delim(char16_t c) { return tables().flags[c] & 4; }

inline str to_lower(sv s) { str r(s); for (auto& c : r) c = lo(c); return r; }

// TextNormalizer.NormalizeWithStandardWhitespace (TextNormalizer.cs:137-194)
inline str normalize(sv text) {
    str out; out.reserve(text.size());
    bool prev_space = false;
    const auto& t = tables();
    for (char16_t c : text) {
        char16_t m = (c == u'\t' || c == u'\n' || c == u'\r') ? u' ' : t.norm[c];
        bool sp = m == u' ';
        if (sp && prev_space) continue;
        out.push_back(m);
        prev_space = sp;
    }
    return out;
}

// --- OrdinalIgnoreCase helpers (upper-invariant fold per UTF-16 unit)
inline bool eq_ic(sv a, sv b) {
    if (a.size() != b.size()) return false;
    for (size_t i = 0; i < a.size(); i++) if (a[i] != b[i] && up(a[i]) != up(b[i])) return false;
    return true;
}
inline bool starts_ic(sv s, sv p) { return s.size() >= p.size() && eq_ic(s.substr(0, p.size()), p); }
inline bool ends_ic(sv s, sv p) { return s.size() >= p.size() && eq_ic(s.substr(s.size() - p.size()), p); }
inline int index_of_ic(sv s, sv p) {
    if (p.empty()) return 0;
    if (p.size() > s.size()) return -1;
    for (size_t i = 0; i + p.size() <= s.size(); i++) if (eq_ic(s.substr(i, p.size()), p)) return (int)i;
    return -1;
}
inline bool contains_ic(sv s, sv p) { return index_of_ic(s, p) >= 0; }
inline int cmp_ic(sv a, sv b) {  // string.Compare(.., OrdinalIgnoreCase)
    size_t n = std::min(a.size(), b.size());
    for (size_t i = 0; i < n; i++) { char16_t x = up(a[i]), y = up(b[i]); if (x != y) return x < y ? -1 : 1; }
    return a.size() == b.size() ? 0 : (a.size() < b.size() ? -1 : 1);
}

// string.Split(delimiters, RemoveEmptyEntries)
inline std::vector<sv> split_words(sv s) {
    std::vector<sv> r; size_t i = 0;
    while (i < s.size()) {
        while (i < s.size() && is_delim(s[i])) i++;
        size_t b = i;
        while (i < s.size() && !is_delim(s[i])) i++;
        if (i > b) r.push_back(s.substr(b, i - b));
    }
    return r;
}

inline bool is_blank(sv s) { for (char16_t c : s) if (!is_space(c)) return false; return true; }
inline sv trim(sv s) {
    size_t b = 0, e = s.size();
    while (b < e && is_space(s[b])) b++;
    while (e > b && is_space(s[e - 1])) e--;
    return s.substr(b, e - b);
}

constexpr char16_t PAD = 0xFFFF;   // Tokenizer.START_PAD_CHAR
constexpr int NGRAM = 3;           // config 400: IndexSizes=[3]
constexpr int START_PAD = 2;       // StartPadSize

// Tokenizer.EnumerateTokensForIndexing (Tokenizer.cs:89-139): 3-grams of PAD PAD + text (skip all-pad), then
// every delimiter-split word with len >= 3; `pos` is the offset in the padded text.
template <class F> inline void tokens_for_indexing(sv text_in, F&& visit) {
    if (text_in.empty()) return;
    str text = normalize(text_in);
    str padded(START_PAD, PAD); padded += text;
    if ((int)padded.size() >= NGRAM)
        for (size_t i = 0; i + NGRAM <= padded.size(); i++) {
            sv g(padded.data() + i, NGRAM);
            if (g[0] == PAD && g[1] == PAD && g[2] == PAD) continue;
            visit(g, (int)i);
        }
    sv span(text); size_t i = 0;
    while (i < span.size()) {
        while (i < span.size() && is_delim(span[i])) i++;
        if (i >= span.size()) break;
        size_t b = i;
        while (i < span.size() && !is_delim(span[i])) i++;
        if ((int)(i - b) >= NGRAM) visit(span.substr(b, i - b), START_PAD + (int)b);
    }
}

// Tokenizer.EnumerateShinglesForSearch (Tokenizer.cs:144-200): words (len>=3) first, then padded 3-grams.
template <class F> inline void shingles_for_search(sv text_in, F&& visit) {
    str text = normalize(text_in);
    sv span(text); size_t i = 0;
    while (i < span.size()) {
        while (i < span.size() && is_delim(span[i])) i++;
        if (i >= span.size()) break;
        size_t b = i;
        while (i < span.size() && !is_delim(span[i])) i++;
        if ((int)(i - b) >= NGRAM) visit(span.substr(b, i - b));
    }
    str padded(START_PAD, PAD); padded += text;
    // GenerateShinglesToVisitor: single index size -> ExtractNGrams(text, 3)
    if ((int)padded.size() >= NGRAM)
        for (size_t k = 0; k + NGRAM <= padded.size(); k++) {
            sv g(padded.data() + k, NGRAM);
            if (g[0] == PAD && g[1] == PAD && g[2] == PAD) continue;
            visit(g);
        }
}

// .NET ArraySortHelper<T>.IntrospectiveSort(keys, Comparison<T>) -- used by List<T>.Sort(Comparison) and
// Array.Sort(T[], Comparison). Unstable; reproduced because idf ties decide union order
// (TieredCandidateSelector.cs:128,253; SURVEY App.B Q7).
template <class T, class Cmp> struct DotnetSort {
    Cmp cmp;
    void swap_if_greater(T* k, int i, int j) { if (cmp(k[i], k[j]) > 0) std::swap(k[i], k[j]); }
    void insertion(T* k, int n) {
        for (int i = 0; i < n - 1; i++) {
            T t = k[i + 1]; int j = i;
            while (j >= 0 && cmp(t, k[j]) < 0) { k[j + 1] = k[j]; j--; }
            k[j + 1] = t;
        }
    }
    void down_heap(T* k, int i, int n) {
        T d = k[i - 1];
        while (i <= n / 2) {
            int child = 2 * i;
            if (child < n && cmp(k[child - 1], k[child]) < 0) child++;
            if (!(cmp(d, k[child - 1]) < 0)) break;
            k[i - 1] = k[child - 1]; i = child;
        }
        k[i - 1] = d;
    }
    void heap_sort(T* k, int n) {
        for (int i = n >> 1; i >= 1; i--) down_heap(k, i, n);
        for (int i = n; i > 1; i--) { std::swap(k[0], k[i - 1]); down_heap(k, 1, i - 1); }
    }
    int partition(T* k, int n) {
        int hi = n - 1, mid = hi >> 1;
        swap_if_greater(k, 0, mid); swap_if_greater(k, 0, hi); swap_if_greater(k, mid, hi);
        T pivot = k[mid];
        std::swap(k[mid], k[hi - 1]);
        int left = 0, right = hi - 1;
        while (left < right) {
            while (cmp(k[++left], pivot) < 0) {}
            while (cmp(pivot, k[--right]) < 0) {}
            if (left >= right) break;
            std::swap(k[left], k[right]);
        }
        if (left != hi - 1) std::swap(k[left], k[hi - 1]);
        return left;
    }
    void intro(T* k, int n, int depth) {
        while (n > 1) {
            if (n <= 16) {
                if (n == 2) { swap_if_greater(k, 0, 1); return; }
                if (n == 3) { swap_if_greater(k, 0, 1); swap_if_greater(k, 0, 2); swap_if_greater(k, 1, 2); return; }
                insertion(k, n); return;
            }
            if (depth == 0) { heap_sort(k, n); return; }
            depth--;
            int p = partition(k, n);
            intro(k + p + 1, n - (p + 1), depth);
            n = p;
        }
    }
    void sort(T* k, int n) {
        if (n > 1) { int lg = 0; for (unsigned v = (unsigned)n; v >>= 1;) lg++; intro(k, n, 2 * (lg + 1)); }
    }
};
template <class T, class Cmp> inline void dotnet_sort(std::vector<T>& v, Cmp cmp) {
    DotnetSort<T, Cmp> s{cmp}; s.sort(v.data(), (int)v.size());
}

// .NET PriorityQueue<TElement,TPriority>: 4-ary min-heap (System.Collections.Generic.PriorityQueue).
template <class E, class P, class Less> struct DotnetPQ {
    std::vector<std::pair<E, P>> n; Less less;
    int size() const { return (int)n.size(); }
    void move_up(std::pair<E, P> node, int idx) {
        while (idx > 0) {
            int parent = (idx - 1) >> 2;
            if (less(node.second, n[parent].second)) { n[idx] = n[parent]; idx = parent; } else break;
        }
        n[idx] = node;
    }
    void move_down(std::pair<E, P> node, int idx) {
        int sz = size(), i;
        while ((i = 4 * idx + 1) < sz) {
            int mi = i; int upper = std::min(i + 4, sz);
            while (++i < upper) if (less(n[i].second, n[mi].second)) mi = i;
            if (!less(n[mi].second, node.second)) break;   // node <= minChild
            n[idx] = n[mi]; idx = mi;
        }
        n[idx] = node;
    }
    void enqueue(E e, P p) { n.emplace_back(e, p); move_up(n.back(), size() - 1); }
    const std::pair<E, P>& peek() const { return n[0]; }
    std::pair<E, P> dequeue() {
        auto root = n[0]; auto last = n.back(); n.pop_back();
        if (!n.empty()) move_down(last, 0);
        return root;
    }
    void enqueue_dequeue(E e, P p) {   // replace root iff p > root priority
        if (!n.empty() && less(n[0].second, p)) move_down({e, p}, 0);
    }
};

}  // namespace ifxo
