// ORACLE (test infrastructure, NOT product code): CPU restatement of the lofcz/Infidex in-memory index build.
//
// Follows (under /root/reference/src/Infidex):
//   Api/DocumentFields.cs:73-80,124-170      GetSearchableTexts ('§' joined, High->Low)
//   Indexing/VectorModel.cs:73-128           IndexDocument / DetermineFieldWeight
//   Core/TermCollection.cs:75-138, Core/Term.cs:71-145   df / tf-weight accumulation, stop terms
//   Indexing/VectorModel.cs:130-220,864-908  doc lengths, avgdl, word-idf cache
//   Indexing/VectorModel.cs:250-313          document metadata (first token, token count)
//   Indexing/Fst/FstBuilder.cs:54-168, FstIndex.cs  trie ("FST") -- forward + reverse
//   Indexing/ShortQuery/PositionalPrefixIndex.cs:55-119, PrefixPosting.cs:109-137  prefix docsets
//   WordMatcher/WordMatcher.cs:82-196        exact / LD1-deletion / affix dictionaries (incl. quirk Q4)
//   Indexing/Bm25Scorer.cs:686-695           ComputeIdf
#pragma once
#include "text.hpp"
#include <unordered_map>
#include <map>
#include <cmath>
#include <charconv>
#include <cstring>
#include <thread>

namespace ifxo {

struct Value {           // boxed field value
    int kind = 0;        // 0 null, 1 string, 2 int64, 3 double, 4 bool
    str s; long long i = 0; double d = 0; bool b = false;
    bool is_null() const { return kind == 0; }
    str to_string() const {   // .NET object.ToString() (invariant culture)
        char buf[64];
        switch (kind) {
            case 1: return s;
            case 2: { auto r = std::to_chars(buf, buf + 64, i); return str(buf, r.ptr); }
            case 3: { auto r = std::to_chars(buf, buf + 64, d); return str(buf, r.ptr); }   // shortest round-trip
            case 4: return b ? u"True" : u"False";
            default: return u"";
        }
    }
};

struct FieldDef { str name; int weight = 1; bool indexable = true, filterable = false, facetable = false; };

struct Doc {
    long long key = 0; int seg = 0; bool deleted = false;
    str indexed_text;                 // IndexedText (original case, '§' joined)
    std::vector<Value> values;        // per schema field
};

struct Term { str text; int df = 0; std::vector<int> docs; std::vector<uint8_t> w; };

inline float compute_idf(int total, int df) {     // Bm25Scorer.ComputeIdf
    if (df <= 0 || total <= 0) return 0.f;
    float d = (float)df, N = (float)total;
    float ratio = (N - d + 0.5f) / (d + 0.5f);
    return ratio <= 0.f ? 0.f : std::log(ratio + 1.f);
}

// Trie with contiguous, label-sorted arcs (FstBuilder.CompactTrie + FstIndex).
struct Trie {
    struct Node { int arc_start = 0; int arc_count = 0; bool final_ = false; int out = -1; };
    struct Arc { char16_t label; int target; };
    std::vector<Node> nodes; std::vector<Arc> arcs;
    bool empty() const { return nodes.empty(); }
    // items must be sorted ordinally with unique keys
    void build(const std::vector<std::pair<str, int>>& items) {
        nodes.clear(); arcs.clear();
        if (items.empty()) return;
        nodes.emplace_back();
        struct Job { int node, lo, hi, depth; };
        std::vector<Job> st; st.push_back({0, 0, (int)items.size(), 0});
        while (!st.empty()) {
            Job j = st.back(); st.pop_back();
            int lo = j.lo;
            if ((int)items[lo].first.size() == j.depth) { nodes[j.node].final_ = true; nodes[j.node].out = items[lo].second; lo++; }
            int astart = (int)arcs.size(); int cnt = 0;
            for (int a = lo; a < j.hi;) {
                char16_t c = items[a].first[j.depth]; int b = a;
                while (b < j.hi && items[b].first[j.depth] == c) b++;
                arcs.push_back({c, -(a + 1)});  // temp: remember range start
                st.push_back({-1, a, b, j.depth + 1});   // node id filled below
                a = b; cnt++;
            }
            nodes[j.node].arc_start = astart; nodes[j.node].arc_count = cnt;
            // allocate child nodes and patch jobs
            for (int k = 0; k < cnt; k++) {
                int id = (int)nodes.size(); nodes.emplace_back();
                arcs[astart + k].target = id;
                st[st.size() - cnt + k].node = id;
            }
        }
    }
    int find_arc(int node, char16_t c) const {
        const Node& n = nodes[node];
        int lo = 0, hi = n.arc_count - 1;
        while (lo <= hi) { int mid = (lo + hi) >> 1; char16_t l = arcs[n.arc_start + mid].label;
            if (l == c) return n.arc_start + mid; if (l < c) lo = mid + 1; else hi = mid - 1; }
        return -1;
    }
    int walk(sv s) const {
        if (nodes.empty()) return -1;
        int n = 0;
        for (char16_t c : s) { int a = find_arc(n, c); if (a < 0) return -1; n = arcs[a].target; }
        return n;
    }
    int get_exact(sv s) const {   // FstIndex.GetExact
        if (s.empty()) return -1;
        int n = walk(s); if (n < 0) return -1;
        return nodes[n].final_ ? nodes[n].out : -1;
    }
    int count_outputs(int start) const {   // CountOutputs
        if (start < 0) return 0;
        int c = 0; std::vector<int> st{start};
        while (!st.empty()) { int n = st.back(); st.pop_back(); if (nodes[n].final_) c++;
            for (int i = 0; i < nodes[n].arc_count; i++) st.push_back(arcs[nodes[n].arc_start + i].target); }
        return c;
    }
    // CollectOutputs: DFS pre-order, ascending labels, stop at `cap`
    int collect(int start, int cap, std::vector<int>& out) const {
        if (start < 0 || cap <= 0) return 0;
        int c = 0; std::vector<int> st{start};
        while (!st.empty()) {
            int n = st.back(); st.pop_back();
            if (nodes[n].final_ && nodes[n].out >= 0) { out.push_back(nodes[n].out); if (++c >= cap) return c; }
            for (int i = nodes[n].arc_count - 1; i >= 0; i--) st.push_back(arcs[nodes[n].arc_start + i].target);
        }
        return c;
    }
    // FstIndex.MatchWithinEditDistance1 (FstIndex.cs:202-352): Myers bit-vector over the trie, "search" variant
    // (no +1 carried into row 0), DFS ascending labels, prune at depth >= m+1. Returns total count; fills <= cap.
    int match_ld1(sv q, int cap, std::vector<int>& out) const {
        if (nodes.empty()) return 0;
        int m = (int)q.size(), count = 0;
        if (m == 0) {
            if (nodes[0].final_ && nodes[0].out >= 0) { if (count < cap) out.push_back(nodes[0].out); count++; }
            for (int i = 0; i < nodes[0].arc_count; i++) { const Node& t = nodes[arcs[nodes[0].arc_start + i].target];
                if (t.final_ && t.out >= 0) { if (count < cap) out.push_back(t.out); count++; } }
            return count;
        }
        if (m > 64) return match_ld1_slow(q, cap, out);
        struct Frame { int node; uint64_t vp, vn; int score, depth; };
        std::vector<Frame> st; st.push_back({0, ~0ULL, 0ULL, m, 0});
        uint64_t maskM = 1ULL << (m - 1);
        while (!st.empty()) {
            Frame f = st.back(); st.pop_back();
            const Node& nd = nodes[f.node];
            if (f.score <= 1 && nd.final_ && nd.out >= 0) { if (count < cap) out.push_back(nd.out); count++; }
            if (f.depth >= m + 1) continue;
            for (int i = nd.arc_count - 1; i >= 0; i--) {
                const Arc& a = arcs[nd.arc_start + i];
                uint64_t pm = 0; for (int k = 0; k < m; k++) if (q[k] == a.label) pm |= 1ULL << k;
                uint64_t x = pm | f.vn;
                uint64_t d0 = ((f.vp + (x & f.vp)) ^ f.vp) | x;
                uint64_t hn = f.vp & d0;
                uint64_t hp = f.vn | ~(f.vp | d0);
                uint64_t nvp = (hn << 1) | ~(d0 | (hp << 1));
                uint64_t nvn = d0 & (hp << 1);
                int ns = f.score; if (hp & maskM) ns++; if (hn & maskM) ns--;
                st.push_back({a.target, nvp, nvn, ns, f.depth + 1});
            }
        }
        return count;
    }
    int match_ld1_slow(sv q, int cap, std::vector<int>& out) const {  // FstIndex.MatchEditDistance1Slow
        int m = (int)q.size(), count = 0;
        std::vector<std::pair<int, std::vector<int>>> st;
        std::vector<int> r0(m + 1); for (int i = 0; i <= m; i++) r0[i] = i;
        st.emplace_back(0, r0);
        while (!st.empty()) {
            auto [ni, row] = st.back(); st.pop_back();
            const Node& nd = nodes[ni];
            if (row[m] <= 1 && nd.final_ && nd.out >= 0) { if (count < cap) { out.push_back(nd.out); count++; } if (count >= cap) return count; }
            int mn = row[0]; for (int i = 1; i <= m; i++) mn = std::min(mn, row[i]);
            if (mn > 1) continue;
            for (int k = 0; k < nd.arc_count; k++) {
                const Arc& a = arcs[nd.arc_start + k];
                std::vector<int> nr(m + 1); nr[0] = row[0] + 1;
                for (int i = 1; i <= m; i++) { int cost = q[i - 1] == a.label ? 0 : 1;
                    nr[i] = std::min(std::min(nr[i - 1] + 1, row[i] + 1), row[i - 1] + cost); }
                st.emplace_back(a.target, std::move(nr));
            }
        }
        return count;
    }
};

struct StrHash { size_t operator()(const str& s) const { return std::hash<sv>()(sv(s)); } };
template <class V> using StrMap = std::unordered_map<str, V, StrHash>;

struct Index {
    // config 400 (ConfigurationParameters.cs:101-124)
    int stop_term_limit = 1250000;
    float field_weights[3] = {1.5f, 1.25f, 1.0f};
    std::vector<FieldDef> schema;                 // DocumentFields of the first doc (facet schema, Q14)
    std::vector<Doc> docs;
    int live_count = 0;                           // Documents.Count
    StrMap<int> term_ids; std::vector<Term> terms;
    std::vector<float> doc_len; float avgdl = 0.f;
    StrMap<float> word_idf;                       // keys are lower-case words
    std::vector<str> first_token; std::vector<uint16_t> token_count;   // DocumentMetadataCache
    Trie term_trie;
    StrMap<std::vector<int>> prefix_docs;         // PositionalPrefixIndex DocSet per 1..3-char prefix
    struct PrefixPost { int doc; uint16_t pos; };
    StrMap<std::vector<PrefixPost>> prefix_post; // PositionalPrefixIndex postings (doc, token index), ascending (doc, pos); every posting is a word start
    StrMap<std::vector<int>> wm_exact, wm_ld1;    // WordMatcher exact / deletion-variant dictionaries
    StrMap<int> wm_affix_last;                    // word -> last doc containing it (Q4)
    Trie wm_fwd, wm_rev;
    std::unordered_map<long long, std::vector<int>> key_to_ids;
    bool built = false;
    bool from_image = false;                      // loaded from a flattened image (load_image): no token positions -> no short-query champion lists

    int doc_by_key(long long key) const {         // GetDocumentByPublicKey: first non-deleted
        auto it = key_to_ids.find(key); if (it == key_to_ids.end()) return -1;
        for (int id : it->second) if (!docs[id].deleted) return id;
        return -1;
    }

    static uint8_t round_w(float w) {             // (byte)Math.Min(Math.Round(w), 255) -- banker's rounding
        double r = std::nearbyint((double)w);     // default FE_TONEAREST = half-to-even
        return (uint8_t)std::min(r, 255.0);
    }

    void add_document(long long key, const std::vector<Value>& values) {
        int id = (int)docs.size();
        docs.emplace_back(); Doc& d = docs.back(); d.key = key; d.values = values;
        live_count++;
        key_to_ids[key].push_back(id);
        // GetSearchableTexts: indexable fields stably ordered by Weight (High=0..Low=2), joined by '§'
        std::vector<int> order;
        for (int w = 0; w < 3; w++) for (size_t f = 0; f < schema.size(); f++) if (schema[f].indexable && schema[f].weight == w) order.push_back((int)f);
        std::vector<std::pair<int, int>> bounds; str text;
        for (size_t k = 0; k < order.size(); k++) {
            bounds.emplace_back((int)(uint16_t)text.size(), schema[order[k]].weight);
            text += values[order[k]].to_string();
            if (k + 1 < order.size()) text.push_back(u'§');
        }
        std::stable_sort(bounds.begin(), bounds.end(), [](auto& a, auto& b) { return a.first < b.first; });
        d.indexed_text = text;
        str index_text = to_lower(normalize(text));
        tokens_for_indexing(index_text, [&](sv tok, int pos) {
            float fw = 1.0f;
            if (!bounds.empty()) { int wi = 0; for (auto& b : bounds) { if (b.first <= pos) wi = b.second; else break; } fw = wi < 3 ? field_weights[wi] : 1.0f; }
            str key_s(tok);
            auto it = term_ids.find(key_s); int tid;
            if (it == term_ids.end()) { tid = (int)terms.size(); term_ids.emplace(key_s, tid); terms.emplace_back(); terms[tid].text = key_s; }
            else tid = it->second;
            Term& t = terms[tid];
            // IncrementTermUsageCounter
            if (t.df != -1) { t.df++; if (t.df > stop_term_limit) t.df = -1; }
            // FirstCycleAdd
            if (t.df < 0) return;
            if ((int)t.w.size() < stop_term_limit) {
                if (t.docs.empty() || t.docs.back() != id) { t.w.push_back(round_w(fw)); t.docs.push_back(id); }
                else { float nw = (float)t.w.back() + fw; if (nw <= 255.f) { t.w.back() = (uint8_t)std::nearbyint((double)nw); t.df--; } }
            } else { t.df = -1; t.w.clear(); t.docs.clear(); }
        });
        // PositionalPrefixIndex.IndexDocument: prefixes (len 1..3) of every word
        {   int token_index = 0;
            for (sv w : split_words(index_text)) {
                int ml = std::min<int>(3, (int)w.size());
                for (int l = 1; l <= ml; l++) { str k(w.substr(0, l)); auto& v = prefix_docs[k]; if (v.empty() || v.back() != id) v.push_back(id); prefix_post[k].push_back({id, (uint16_t)token_index}); }
                token_index++;
            }
        }
        // WordMatcher.Load (lower first, then normalize)
        str wm_text = normalize(to_lower(text));
        for (sv w : split_words(wm_text)) {
            int len = (int)w.size();
            auto add = [&](StrMap<std::vector<int>>& m, const str& k) { auto& v = m[k]; if (v.empty() || v.back() != id) v.push_back(id); };
            if (len >= 2 && len <= 8) add(wm_exact, str(w));
            if (len >= 3 && len <= 8) for (int i = 0; i < len; i++) { str v(w); v.erase(i, 1); add(wm_ld1, v); }
            if (len >= 3) wm_affix_last[str(w)] = id;
        }
    }

    void build() {
        size_t N = docs.size();
        // BuildInvertedLists: doc lengths (integer-valued partial sums => order independent), avgdl sequential float sum
        doc_len.assign(N, 0.f);
        for (auto& t : terms) { if (t.df <= 0) continue; for (size_t i = 0; i < t.docs.size(); i++) doc_len[t.docs[i]] += (float)t.w[i]; }
        float total = 0.f; for (size_t i = 0; i < N; i++) total += doc_len[i];
        avgdl = live_count > 0 ? total / (float)live_count : 0.f;
        // BuildWordIdfCache
        StrMap<int> wdf;
        for (size_t d = 0; d < N; d++) {
            if (docs[d].deleted || docs[d].indexed_text.empty()) continue;
            str nrm = normalize(to_lower(docs[d].indexed_text));
            std::vector<str> uniq;
            for (sv w : split_words(nrm)) { str lw = to_lower(w); bool dup = false; for (auto& u : uniq) if (eq_ic(u, lw)) { dup = true; break; } if (!dup) uniq.push_back(lw); }
            for (auto& u : uniq) wdf[u]++;
        }
        word_idf.clear();
        for (auto& kv : wdf) if (kv.second > 0 && kv.second <= live_count) word_idf[kv.first] = compute_idf(live_count, kv.second);
        // BuildOptimizedIndexes: full trie over all terms (ordinal = TermCollection index)
        std::vector<std::pair<str, int>> items; items.reserve(terms.size());
        for (size_t i = 0; i < terms.size(); i++) items.emplace_back(terms[i].text, (int)i);
        std::sort(items.begin(), items.end());
        term_trie.build(items);
        // BuildDocumentMetadataCache
        first_token.assign(N, str()); token_count.assign(N, 0);
        for (size_t d = 0; d < N; d++) {
            if (docs[d].deleted || docs[d].indexed_text.empty()) continue;
            str t = normalize(to_lower(docs[d].indexed_text));
            auto ws = split_words(t);
            if (!ws.empty()) first_token[d] = str(ws[0]);
            token_count[d] = (uint16_t)std::min<size_t>(ws.size(), 65535);
        }
        // WordMatcher.FinalizeIndex: forward + reverse tries over affix words, output = last doc (Q4)
        std::vector<std::pair<str, int>> fw, rv;
        for (auto& kv : wm_affix_last) { fw.emplace_back(kv.first, kv.second); str r(kv.first.rbegin(), kv.first.rend()); rv.emplace_back(r, kv.second); }
        std::sort(fw.begin(), fw.end()); std::sort(rv.begin(), rv.end());
        wm_fwd.build(fw); wm_rev.build(rv);
        built = true;
    }

    // Benchmark-scale shortcut (bench.py only): take the index STATE from a flattened image (include/infidex_gpu.h ifx_index_image)
    // instead of re-indexing the documents with add_document/build above, which is sequential and takes ~135 us per multi-field
    // document. tests/test_host_builder.py checks that the image of the product's host builder equals what add_document/build
    // produce (terms, df, postings, tf bytes, doc lengths, dictionaries), and tests/test_oracle_image.py that searches over a loaded
    // image equal searches over the oracle's own build. Every search-time structure (tries, maps, posting vectors) is still the
    // oracle's own. Not available from an image: token positions (short-query champion lists) and non-column field values.
    template <class Img> void load_image(const Img& im) {
        auto S = [](const auto& ss, int i) { return str((const char16_t*)ss.chars + ss.off[i], (size_t)(ss.off[i + 1] - ss.off[i])); };
        const int N = im.n_docs; docs.assign(N, Doc()); live_count = im.n_live; doc_len.assign(im.doc_len, im.doc_len + N); avgdl = im.avgdl;
        first_token.assign(N, str()); token_count.assign(im.token_count, im.token_count + N);
        std::vector<std::thread> ts;
        ts.emplace_back([&] {
            for (int d = 0; d < N; d++) { Doc& x = docs[d]; x.key = im.doc_key[d]; x.deleted = im.deleted[d] != 0; x.indexed_text = str((const char16_t*)im.text_chars + im.text_off[d], (size_t)(im.text_off[d + 1] - im.text_off[d]));
                x.values.assign(schema.size(), Value()); first_token[d] = S(im.first_token, d); }
            for (int c = 0; c < im.n_columns; c++) { const auto& col = im.columns[c]; str name((const char16_t*)col.name, (size_t)col.name_len); int f = -1;
                for (size_t k = 0; k < schema.size(); k++) if (schema[k].name == name) f = (int)k;
                if (f < 0) continue;
                std::vector<str> dict(col.dict.n); for (int i = 0; i < col.dict.n; i++) dict[i] = S(col.dict, i);
                for (int d = 0; d < N; d++) if (col.value_id[d] >= 0) { Value& v = docs[d].values[f]; v.kind = 1; v.s = dict[col.value_id[d]]; } }
            for (int d = 0; d < N; d++) key_to_ids[docs[d].key].push_back(d); });
        ts.emplace_back([&] {
            const int T = im.terms.n; terms.assign(T, Term());
            for (int t = 0; t < T; t++) { Term& x = terms[t]; x.text = S(im.terms, t); x.df = im.df[t]; if (x.df > 0) { x.docs.assign(im.post_doc + im.row_ptr[t], im.post_doc + im.row_ptr[t + 1]); x.w.assign(im.post_tf + im.row_ptr[t], im.post_tf + im.row_ptr[t + 1]); } term_ids.emplace(x.text, t); }
            std::vector<std::pair<str, int>> items; items.reserve(T); for (int t = 0; t < T; t++) items.emplace_back(terms[t].text, t);
            std::sort(items.begin(), items.end()); term_trie.build(items); });
        auto load_docsets = [&S](const auto& dd, StrMap<std::vector<int>>& m) { m.reserve((size_t)dd.keys.n * 2); for (int k = 0; k < dd.keys.n; k++) m.emplace(S(dd.keys, k), std::vector<int>(dd.doc_id + dd.row_ptr[k], dd.doc_id + dd.row_ptr[k + 1])); };
        ts.emplace_back([&] { load_docsets(im.prefix, prefix_docs); });
        ts.emplace_back([&] { load_docsets(im.wm_exact, wm_exact); });
        ts.emplace_back([&] { load_docsets(im.wm_ld1, wm_ld1); });
        ts.emplace_back([&] {
            for (int i = 0; i < im.words.n; i++) word_idf.emplace(S(im.words, i), im.word_idf[i]);
            std::vector<std::pair<str, int>> fw, rv;
            for (int i = 0; i < im.affix_words.n; i++) { str w = S(im.affix_words, i); wm_affix_last[w] = im.affix_last_doc[i]; fw.emplace_back(w, im.affix_last_doc[i]); rv.emplace_back(str(w.rbegin(), w.rend()), im.affix_last_doc[i]); }
            std::sort(fw.begin(), fw.end()); std::sort(rv.begin(), rv.end()); wm_fwd.build(fw); wm_rev.build(rv); });
        for (auto& t : ts) t.join();
        from_image = true; built = true;
    }
};

}  // namespace ifxo
