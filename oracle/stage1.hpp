// ORACLE (test infrastructure, NOT product code): CPU restatement of lofcz/Infidex Stage 1
// (term resolution, LD1 expansion, tiered candidate selection, BM25+ with MaxScore, top-K).
//
// Follows (under /root/reference/src/Infidex):
//   Scoring/QueryAnalyzer.cs:10-54
//   Indexing/VectorModel.cs:376-602 (SearchWithMaxScore), :604-743 (GatherTermInfo / ExpandMissingTerm)
//   Scoring/TieredCandidateSelector.cs:53-237,243-322,328-421,455-532
//   Indexing/Bm25Scorer.cs:56-193,195-330,332-445,524-533,589-695
//   Indexing/ArrayPostingsEnum.cs:50-135, RoaringPostingsEnum.cs   (Advance == membership for ascending targets)
//   Core/TopKHeap.cs, Core/ScoreEntry.cs:25-35, Scoring/SegmentProcessor.cs:15-37
#pragma once
#include "index.hpp"

namespace ifxo {

struct ScoreEntry {
    float score = 0; long long key = 0; uint8_t tie = 0;
    int compare(const ScoreEntry& o) const {     // ScoreEntry.CompareTo
        if (score != o.score) return score < o.score ? -1 : 1;
        if (tie != o.tie) return tie < o.tie ? -1 : 1;
        return o.key < key ? -1 : (o.key > key ? 1 : 0);
    }
};

struct TopKHeap {   // Core/TopKHeap.cs (PriorityQueue<ScoreEntry,ScoreEntry>)
    struct Less { bool operator()(const ScoreEntry& a, const ScoreEntry& b) const { return a.compare(b) < 0; } };
    DotnetPQ<ScoreEntry, ScoreEntry, Less> pq; int limit;
    explicit TopKHeap(int l) : limit(l) {}
    void add(const ScoreEntry& e) {
        if (pq.size() < limit) pq.enqueue(e, e);
        else if (e.compare(pq.peek().first) > 0) { pq.dequeue(); pq.enqueue(e, e); }
    }
    std::vector<ScoreEntry> get_top_k() {
        std::vector<ScoreEntry> r(pq.size());
        for (int i = 0, n = (int)r.size(); i < n; i++) r[n - 1 - i] = pq.dequeue().first;
        return r;
    }
};

// SegmentProcessor.ConsolidateSegments: best entry per key, sorted descending
inline std::vector<ScoreEntry> consolidate(const std::vector<ScoreEntry>& in) {
    std::unordered_map<long long, size_t> pos; std::vector<ScoreEntry> out;
    for (auto& e : in) {
        auto it = pos.find(e.key);
        if (it == pos.end()) { pos[e.key] = out.size(); out.push_back(e); }
        else if (e.compare(out[it->second]) > 0) out[it->second] = e;
    }
    std::sort(out.begin(), out.end(), [](const ScoreEntry& a, const ScoreEntry& b) { return b.compare(a) < 0; });
    return out;
}

struct QueryAnalysis { bool can_use_ngrams = false, mixed = false; str long_words; };
inline QueryAnalysis analyze_query(sv text) {    // QueryAnalyzer.Analyze
    QueryAnalysis a; a.long_words = str(text);
    auto words = split_words(text);
    if (words.empty()) { a.can_use_ngrams = (int)text.size() >= NGRAM; return a; }
    std::vector<sv> lw; int shortc = 0;
    for (sv w : words) { if ((int)w.size() >= NGRAM) lw.push_back(w); else shortc++; }
    if (!lw.empty()) { a.can_use_ngrams = true; str j; for (size_t i = 0; i < lw.size(); i++) { if (i) j.push_back(u' '); j += lw[i]; } a.long_words = j; }
    if (shortc > 0 && !lw.empty()) a.mixed = true;
    return a;
}

struct QTerm {
    int term_id = -1; bool fuzzy = false;
    std::vector<int> fuzzy_docs;                  // sorted unique (virtual term, tf == 1)
    int df = 0; float idf = 0, max_score = 0;
    const std::vector<int>* docs = nullptr; const std::vector<uint8_t>* w = nullptr;
    int cost() const { return (int)docs->size(); }
};

struct Stage1Stats {          // what the candidate selector did (for roofline accounting / debugging)
    int path = 0;             // 1 prefix shortcut, 2 disjunctive, 3 tiers, 4 full scan
    long long candidates = 0; long long streamed_postings = 0; int n_terms = 0; int n_fuzzy = 0;
    // optional per-chunk trace of the MaxScore chain (tools/speculation_study.py): [threshold at chunk start, candidates, matched pairs,
    // matched pairs the MaxScore test skipped, max over those of (score + bound + suffix), heap updates during the flush]
    std::vector<double>* trace = nullptr;
};

struct Stage1 {
    const Index& ix;
    explicit Stage1(const Index& i) : ix(i) {}

    static float term_score_scalar(float tf, float dl, float avgdl, float idf) {   // Bm25Scorer.ComputeTermScore
        const float K1 = 1.2f, B = 0.75f, Delta = 1.0f;
        float norm = K1 * (1.f - B + B * (dl / avgdl));
        float denom = tf + norm;
        if (denom <= 0.f) return 0.f;
        float core = (tf * (K1 + 1.f)) / denom;
        return idf * (core + Delta);
    }
    static float term_score_vector(float tf, float dl, float avgdl, float idf) {   // Bm25Scorer.cs:395-433 lanes
        const float K1 = 1.2f, B = 0.75f, Delta = 1.0f;
        float bdiv = B / avgdl;
        float norm = K1 * ((1.f - B) + bdiv * dl);
        float denom = tf + norm;
        float core = (tf * (K1 + 1.0f)) / denom;
        return idf * (core + Delta);
    }

    // VectorModel.SearchWithMaxScore term preparation. Returns the scored terms in reference order.
    std::vector<QTerm> resolve_terms(sv query, Stage1Stats* st) const {
        struct Raw { int id; str text; };
        std::vector<Raw> raw;
        shingles_for_search(query, [&](sv tok) {
            if (raw.size() >= 128) return;
            int id = ix.term_trie.get_exact(tok);
            if (id >= 0) raw.push_back({id, str()}); else raw.push_back({-1, str(tok)});
        });
        std::sort(raw.begin(), raw.end(), [](const Raw& a, const Raw& b) { return a.id != b.id ? a.id < b.id : a.text < b.text; });
        std::vector<QTerm> out;
        int N = ix.live_count;
        float avgdl = ix.avgdl > 0.f ? ix.avgdl : 1.f;
        for (size_t i = 0; i < raw.size(); i++) {
            if (i > 0 && raw[i].id == raw[i - 1].id && (raw[i].id >= 0 || raw[i].text == raw[i - 1].text)) continue;
            QTerm q; q.term_id = raw[i].id;
            if (q.term_id >= 0) { const Term& t = ix.terms[q.term_id]; q.df = t.df; q.docs = &t.docs; q.w = &t.w; }
            else if (raw[i].text.size() >= 4) {
                // ExpandMissingTerm: first 1024 LD1 trie matches (DFS order), union of their postings
                std::vector<int> ords; ix.term_trie.match_ld1(raw[i].text, 1024, ords);
                std::vector<int> all;
                for (int o : ords) { const Term& t = ix.terms[o]; if (t.df > 0) all.insert(all.end(), t.docs.begin(), t.docs.end()); }
                if (!all.empty()) {
                    std::sort(all.begin(), all.end()); all.erase(std::unique(all.begin(), all.end()), all.end());
                    q.fuzzy = true; q.fuzzy_docs = std::move(all); q.df = (int)q.fuzzy_docs.size();
                    if (st) st->n_fuzzy++;
                }
            }
            if (q.df <= 0 || q.df > ix.stop_term_limit) continue;
            q.idf = compute_idf(N, q.df);
            const float maxTf = 255.f, k1 = 1.2f, b = 0.75f, delta = 1.0f;
            float minDlNorm = 1.f - b + b * (1.f / avgdl);
            float core = (maxTf * (k1 + 1.f)) / (maxTf + k1 * minDlNorm);
            q.max_score = q.idf * (core + delta);
            out.push_back(std::move(q));
        }
        for (auto& q : out) if (q.fuzzy) q.docs = &q.fuzzy_docs;   // after moves are done
        if (st) st->n_terms = (int)out.size();
        return out;
    }

    static std::vector<int> set_union(const std::vector<int>& a, const std::vector<int>& b) {
        std::vector<int> r; r.reserve(a.size() + b.size());
        std::set_union(a.begin(), a.end(), b.begin(), b.end(), std::back_inserter(r)); return r;
    }
    static std::vector<int> intersect_terms(const std::vector<const QTerm*>& ts) {   // IntersectTerms: exact intersection
        if (ts.empty()) return {};
        std::vector<const QTerm*> e(ts); std::stable_sort(e.begin(), e.end(), [](auto a, auto b) { return a->cost() < b->cost(); });
        std::vector<int> cur = *e[0]->docs;
        for (size_t i = 1; i < e.size() && !cur.empty(); i++) {
            std::vector<int> nx; std::set_intersection(cur.begin(), cur.end(), e[i]->docs->begin(), e[i]->docs->end(), std::back_inserter(nx)); cur.swap(nx);
        }
        return cur;
    }

    // TieredCandidateSelector.SelectCandidates
    std::vector<int> select_candidates(const std::vector<QTerm>& qts, int K, sv original_query, Stage1Stats* st) const {
        if (qts.empty()) return {};
        // prefix precedence (TrySelectPrefixCandidates)
        if (!original_query.empty()) {
            str ql = to_lower(original_query);
            int maxl = std::min<int>((int)ql.size(), 3);
            const std::vector<int>* chosen = nullptr;
            for (int len = maxl; len >= 1; len--) {
                auto it = ix.prefix_docs.find(ql.substr(0, len));
                if (it == ix.prefix_docs.end() || it->second.empty()) continue;
                long long pop = (long long)it->second.size();
                if (pop > (long long)K * 20) continue;
                if (pop <= (long long)K * 10) { chosen = &it->second; break; }
            }
            if (chosen && (long long)chosen->size() >= std::min(K * 2, 100)) { if (st) st->path = 1; return *chosen; }
        }
        std::vector<const QTerm*> terms;
        for (auto& q : qts) if (q.df > 0) terms.push_back(&q);
        int missing = (int)qts.size() - (int)terms.size();
        if (terms.empty()) return {};
        bool typo = false; float max_idf = 0.f;
        for (auto t : terms) { if (t->df < 10) typo = true; if (t->idf > max_idf) max_idf = t->idf; }
        auto idf_desc = [](const QTerm* a, const QTerm* b) { return b->idf < a->idf ? -1 : (b->idf > a->idf ? 1 : 0); };  // b.Idf.CompareTo(a.Idf)
        if (typo || missing > 0 || qts.size() == 1) {
            // SelectCandidatesDisjunctive
            if (st) st->path = 2;
            dotnet_sort(terms, idf_desc);
            bool selective = false; long long local = 0;
            // per-thread reusable membership array (the reference rents its float[N] upperBounds from ArrayPool the same way)
            static thread_local std::vector<uint8_t> seen; if (seen.size() < ix.docs.size()) seen.assign(ix.docs.size(), 0);
            std::vector<int> result;
            for (auto t : terms) {
                bool lowq = t->idf < (max_idf * 0.2f);
                if (terms.size() > 1 && lowq && selective) continue;
                for (int d : *t->docs) if (!seen[d]) { seen[d] = 1; local++; result.push_back(d); }
                if (st) st->streamed_postings += (long long)t->docs->size();
                if (!lowq && local > 0) selective = true;
                if (local >= (long long)K * 100) break;
            }
            for (int d : result) seen[d] = 0;
            std::sort(result.begin(), result.end());
            return result;
        }
        if (st) st->path = 3;
        dotnet_sort(terms, idf_desc);
        std::vector<int> global;
        if (terms.size() >= 2) {
            global = intersect_terms(terms);
            if (st) for (auto t : terms) st->streamed_postings += t->cost();
            if ((long long)global.size() >= (long long)K * 2) return global;
        }
        if (terms.size() >= 3 && (long long)global.size() < (long long)K * 3) {
            std::vector<const QTerm*> t1(terms.begin(), terms.end() - 1);
            global = set_union(global, intersect_terms(t1));
        }
        if ((long long)global.size() < (long long)K * 5) {
            std::vector<const QTerm*> sel; float cutoff = max_idf * 0.3f; size_t cap = std::min<size_t>(2, terms.size());
            for (auto t : terms) { if (t->idf <= 0.f) continue; if (t->idf < cutoff) continue; sel.push_back(t); if (sel.size() == cap) break; }
            for (auto t : sel) {
                global = set_union(global, *t->docs);
                if (st) st->streamed_postings += t->cost();
                if ((long long)global.size() >= (long long)K * 10) break;
            }
        }
        return global;
    }

    struct FloatLess { bool operator()(float a, float b) const { return a < b; } };
    using PruneHeap = DotnetPQ<int, float, FloatLess>;
    static void update_topk(int id, float s, int K, PruneHeap& h, float& thr) {   // Bm25Scorer.UpdateTopK
        if (K == INT32_MAX) return;
        if (h.size() < K) { h.enqueue(id, s); if (h.size() == K) thr = h.peek().second; }
        else if (s > thr) { h.enqueue_dequeue(id, s); thr = h.peek().second; }
    }

    // Bm25Scorer.Search (array overload): returns TopKHeap content consolidated (desc)
    std::vector<ScoreEntry> search(sv query, int K, Stage1Stats* st = nullptr) const {
        std::vector<QTerm> terms = resolve_terms(query, st);
        TopKHeap result(K);
        int N = ix.live_count;
        if (terms.empty() || N == 0) return {};
        float avgdl = ix.avgdl > 0.f ? ix.avgdl : 1.f;
        std::vector<int> cands = select_candidates(terms, K, query, st);
        if (st) st->candidates = (long long)cands.size();
        size_t T = terms.size();
        std::vector<float> suffix(T + 1, 0.f);
        for (int i = (int)T - 1; i >= 0; i--) suffix[i] = suffix[i + 1] + terms[i].max_score;
        PruneHeap heap; float thr = 0.f;
        if (!cands.empty()) {
            std::vector<size_t> cursor(T, 0);
            std::vector<float> score; std::vector<int> midx; std::vector<float> mtf;
            size_t p = 0;
            while (p < cands.size()) {
                // container run (same id>>16), then sub-chunks of 4096
                size_t ce = p; int hb = cands[p] >> 16;
                while (ce < cands.size() && (cands[ce] >> 16) == hb) ce++;
                while (p < ce) {
                    size_t cnt = std::min<size_t>(4096, ce - p);
                    score.assign(cnt, 0.f);
                    const bool tracing = st && st->trace; double tr_matched = 0, tr_skipped = 0, tr_vmax = -1.0, tr_updates = 0; const float tr_thr0 = thr;
                    for (size_t t = 0; t < T; t++) {
                        const QTerm& q = terms[t];
                        if (q.idf <= 0.f) continue;
                        float rem = suffix[t + 1];
                        midx.clear(); mtf.clear();
                        const std::vector<int>& dl = *q.docs; size_t& cur = cursor[t];
                        for (size_t j = 0; j < cnt; j++) {
                            if (K < INT32_MAX && score[j] + q.max_score + rem <= thr) {
                                if (tracing && std::binary_search(dl.begin(), dl.end(), cands[p + j])) { tr_skipped++; tr_matched++; tr_vmax = std::max(tr_vmax, (double)(score[j] + q.max_score + rem)); }
                                continue;
                            }
                            int target = cands[p + j];
                            // Advance(target): first posting >= target from the current position
                            if (cur < dl.size() && dl[cur] < target) cur = std::lower_bound(dl.begin() + cur, dl.end(), target) - dl.begin();
                            if (cur < dl.size() && dl[cur] == target) { midx.push_back((int)j); mtf.push_back(q.fuzzy ? 1.f : (float)(*q.w)[cur]); }
                        }
                        size_t m = midx.size(), vec_end = m - (m % 8); tr_matched += (double)m;
                        for (size_t i = 0; i < m; i++) {
                            int j = midx[i]; float d = ix.doc_len[cands[p + j]];
                            if (i < vec_end) score[j] += term_score_vector(mtf[i], d, avgdl, q.idf);
                            else { if (d <= 0.f) d = 1.f; score[j] += term_score_scalar(mtf[i], d, avgdl, q.idf); }
                        }
                    }
                    for (size_t j = 0; j < cnt; j++) if (score[j] > 0.f) { int id = cands[p + j]; if (!ix.docs[id].deleted) { float before = thr; int hs = heap.size(); update_topk(id, score[j], K, heap, thr); if (thr != before || heap.size() != hs) tr_updates++; } }
                    if (tracing) { double rec[6] = {(double)tr_thr0, (double)cnt, tr_matched, tr_skipped, tr_vmax, tr_updates}; st->trace->insert(st->trace->end(), rec, rec + 6); }
                    p += cnt;
                }
            }
            if (st) for (size_t t = 0; t < T; t++) (void)t;
            while (heap.size() > 0) { auto e = heap.dequeue(); if (!ix.docs[e.first].deleted) result.add({e.second, ix.docs[e.first].key, 0}); }
        } else {
            // full scan (Bm25Scorer.cs:151-176,589-641) -- unreachable with >= 1 scored term, kept for fidelity
            if (st) st->path = 4;
            std::vector<float> ds(ix.docs.size(), 0.f);
            for (size_t t = 0; t < T; t++) {
                const QTerm& q = terms[t]; if (q.idf <= 0.f) continue; float rem = suffix[t + 1];
                for (size_t k = 0; k < q.docs->size(); k++) {
                    int id = (*q.docs)[k]; if ((unsigned)id >= (unsigned)N) continue;
                    float cs = ds[id];
                    if (K < INT32_MAX && heap.size() >= K && cs + q.max_score + rem <= thr) continue;
                    if (ix.docs[id].deleted) continue;
                    float tf = q.fuzzy ? 1.f : (float)(*q.w)[k]; if (tf <= 0.f) continue;
                    float d = ix.doc_len[id]; if (d <= 0.f) d = 1.f;
                    float ns = cs + term_score_scalar(tf, d, avgdl, q.idf); ds[id] = ns;
                    update_topk(id, ns, K, heap, thr);
                }
            }
            if (K < INT32_MAX && heap.size() > 0) while (heap.size() > 0) { auto e = heap.dequeue(); if (!ix.docs[e.first].deleted) result.add({e.second, ix.docs[e.first].key, 0}); }
            else for (size_t i = 0; i < ds.size(); i++) if (ds[i] > 0.f && !ix.docs[i].deleted) result.add({ds[i], ix.docs[i].key, 0});
        }
        return consolidate(result.get_top_k());
    }
};

}  // namespace ifxo
