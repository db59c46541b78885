// TEST INFRASTRUCTURE ONLY (see oracle/text.hpp header). CPU restatement of the reference's Stage 1 for queries without any word of
// >= 3 characters ("next" row): ShortQueryResolver champion lists (Indexing/ShortQuery/ShortQueryResolver.cs:83-323),
// ShortQueryProcessor.SearchSingleCharacter / SearchShortQuery (Scoring/ShortQueryProcessor.cs:19-435) and the branch of
// SearchPipeline.ExecuteRelevancyStage that picks between them (Scoring/SearchPipeline.cs:208-297).
// Parity status: pinned only by the reference's own short-query tests (MovieSearchParityTests.cs:557-622,1085-1140: top-1 / non-empty
// assertions). One known unpinned point: ProcessFuzzyFallback uses string.StartsWith(string), which is culture-sensitive in .NET; it is
// restated as an ordinal comparison here.
#pragma once
#include "stage1.hpp"

namespace ifxo {

struct ShortQuery {
    const Index& ix;
    mutable StrMap<std::vector<ScoreEntry>> champions; mutable bool champions_built = false;
    explicit ShortQuery(const Index& i) : ix(i) {}

    // ShortQueryResolver.CalculateFinalScore (:268-311): (precedence << 8) | base, every posting being a word start
    uint16_t resolver_score(sv query, const Doc& doc, int first_pos, int word_starts) const {
        int prec = 128; if (first_pos == 0) prec |= 64;
        str title = to_lower(doc.indexed_text); auto tokens = split_words(title);
        bool any = false, first = false;
        for (size_t i = 0; i < tokens.size(); i++) if (tokens[i] == query) { any = true; if (i == 0) first = true; break; }
        if (any) prec |= 32;
        if (first) prec |= 16;
        if (trim(title) == query) prec |= 8;
        if (tokens.size() <= 3) prec |= 32;
        int pos_c = 255 - std::min(first_pos * 16, 240), dens = std::min(word_starts * 8, 32);
        int base = std::max(0, std::min(255, pos_c + dens));
        return (uint16_t)((prec << 8) | base);
    }
    // BuildChampionLists (:83-160): per prefix, docs in posting (= ascending id) order, List.Sort by score descending (unstable
    // introsort, reproduced), first 64
    void build_champions() const {
        if (champions_built) return; champions_built = true;
        for (auto& kv : ix.prefix_post) {
            const auto& posts = kv.second; if (posts.empty() || kv.first.empty()) continue;
            std::vector<ScoreEntry> scores;
            for (size_t i = 0; i < posts.size();) {
                size_t j = i; int d = posts[i].doc; int first_pos = 0x7fffffff, n = 0;
                while (j < posts.size() && posts[j].doc == d) { n++; first_pos = std::min<int>(first_pos, posts[j].pos); j++; }
                i = j;
                const Doc& doc = ix.docs[d]; if (doc.deleted) continue;
                ScoreEntry e; e.score = (float)resolver_score(kv.first, doc, first_pos, n); e.key = doc.key; e.tie = 0; scores.push_back(e);
            }
            if (scores.empty()) continue;
            dotnet_sort(scores, [](const ScoreEntry& a, const ScoreEntry& b) { return b.score < a.score ? -1 : (b.score > a.score ? 1 : 0); });
            if (scores.size() > 64) scores.resize(64);
            champions[kv.first] = std::move(scores);
        }
    }
    // TryGetChampions (:233-260)
    bool try_champions(sv prefix, int max_results, std::vector<ScoreEntry>& out) const {
        build_champions();
        if (max_results <= 0 || prefix.empty() || prefix.size() > 3) return false;
        auto it = champions.find(str(prefix)); if (it == champions.end() || it->second.empty()) return false;
        if ((long long)it->second.size() < max_results) return false;
        out.assign(it->second.begin(), it->second.begin() + max_results);
        return true;
    }

    // ShortQueryProcessor.SearchSingleCharacter (:19-150): full scan over the lower-cased IndexedText
    std::vector<ScoreEntry> single_character(char16_t ch, int max_results) const {
        ch = lo(ch); std::vector<ScoreEntry> raw;
        for (const Doc& doc : ix.docs) {
            if (doc.deleted || doc.indexed_text.empty()) continue;
            str lower = to_lower(doc.indexed_text);
            int char_count = 0, first_char = -1;
            for (size_t i = 0; i < lower.size(); i++) if (lower[i] == ch) { char_count++; if (first_char == -1) first_char = (int)i; }
            if (char_count == 0) continue;
            auto words = split_words(lower);
            bool word_start = false; int first_word = 0x7fffffff, ws_count = 0;
            for (size_t i = 0; i < words.size(); i++) if (!words[i].empty() && words[i][0] == ch) { word_start = true; ws_count++; if ((int)i < first_word) first_word = (int)i; }
            bool any_exact = false, first_exact = false;
            if (!words.empty()) {
                first_exact = words[0].size() == 1 && words[0][0] == ch;
                if (first_exact) any_exact = true; else for (sv w : words) if (w.size() == 1 && w[0] == ch) { any_exact = true; break; }
            }
            bool title_eq = lower.size() == 1 && lower[0] == ch;
            int prec = 0;
            if (word_start) { prec |= 128; if (first_word == 0) prec |= 64; }
            if (any_exact) prec |= 32;
            if (first_exact) prec |= 16;
            if (title_eq) prec |= 8;
            if (words.size() <= 3) prec |= 32;
            float base;
            if (word_start) { int pc = 255 - std::min(first_word * 16, 240), dc = std::min(ws_count * 8, 32); int rawv = std::max(0, std::min(255, pc + dc)); base = (float)rawv / 255.f; }
            else { int pc = 200 - std::min(std::max(first_char, 0) * 4, 180), dc = std::min(char_count * 4, 40); int rawv = std::max(0, std::min(200, pc + dc)); base = (float)std::max(1, rawv) / 255.f; }
            ScoreEntry e; e.score = (float)prec + base; e.key = doc.key; e.tie = 0; raw.push_back(e);
        }
        std::vector<ScoreEntry> cons = consolidate(raw);
        if (max_results < INT32_MAX && (long long)cons.size() > max_results) cons.resize(max_results);
        return cons;
    }

    // ShortQueryProcessor.SearchShortQuery (:152-222) with BuildPrefixPatterns, ProcessTermMatches, ProcessFuzzyFallback,
    // BuildFinalScores, ComputePrecedence. Dictionary<long,int> / HashSet<long> enumerate in insertion order (no removals).
    std::vector<ScoreEntry> short_query(sv search_lower) const {
        std::vector<long long> order; std::unordered_map<long long, int> score; std::unordered_map<long long, char> first_prefix; size_t matched = 0;
        auto process_term = [&](const Term& t, int mult) {
            if (t.docs.empty() || t.w.size() != t.docs.size()) return;
            for (size_t i = 0; i < t.docs.size(); i++) {
                const Doc& d = ix.docs[t.docs[i]]; if (d.deleted) continue;
                int sc = (int)t.w[i] * mult;
                auto it = score.find(d.key);
                if (it != score.end()) it->second += sc; else { score.emplace(d.key, sc); order.push_back(d.key); matched++; }
                if (!first_prefix.count(d.key)) { str tl = to_lower(d.indexed_text); if (tl.size() >= search_lower.size() && sv(tl).substr(0, search_lower.size()) == search_lower) first_prefix[d.key] = 1; }
            }
        };
        // BuildPrefixPatterns(searchLower, minIndexSize = 3, startPadSize = 2)
        std::vector<str> patterns; const int pad = 2, mis = NGRAM;
        for (int i = 0; i < mis && i < pad + (int)search_lower.size(); i++) {
            int pc = std::max(0, pad - i), qc = std::min<int>((int)search_lower.size(), mis - pc);
            if (qc > 0) { str p((size_t)pc, PAD); p += str(search_lower.substr(0, qc)); patterns.push_back(p); }
        }
        { str p(1, u' '); p += str(search_lower); patterns.push_back(p); }
        for (const str& pat : patterns) {
            int node = ix.term_trie.walk(pat); int cnt = ix.term_trie.count_outputs(node); if (cnt == 0) continue;
            std::vector<int> ords; ix.term_trie.collect(node, std::min(cnt, 4096), ords);
            for (int o : ords) process_term(ix.terms[o], 10);
        }
        if (matched < 100) {     // ProcessFuzzyFallback: every term, TermCollection order
            for (const Term& t : ix.terms) {
                bool already = false; for (const str& pat : patterns) if (t.text.size() >= pat.size() && sv(t.text).substr(0, pat.size()) == pat) { already = true; break; }
                if (already) continue;
                bool boundary = false; int cm = 0;
                for (char16_t qc : search_lower) {
                    str wb(1, u' '); wb.push_back(qc);
                    if (t.text.find(wb) != str::npos) { boundary = true; cm++; } else if (t.text.find(qc) != str::npos) cm++;
                }
                if (boundary || cm > 0) process_term(t, boundary ? 2 : 1);
            }
        }
        // BuildFinalScores
        int max_score = 0; for (auto& kv : score) max_score = std::max(max_score, kv.second);
        auto q_tokens = split_words(search_lower);
        TopKHeap heap(INT32_MAX);
        for (long long key : order) {
            int id = ix.doc_by_key(key); if (id < 0) continue; const Doc& d = ix.docs[id]; if (d.deleted) continue;
            int v = score[key];
            float normalized = max_score > 0 ? (float)v / (float)max_score : (float)v / 255.f;
            str title = to_lower(d.indexed_text); sv trimmed = trim(title); auto words = split_words(title);
            int prec = 0;
            if (q_tokens.size() >= 2) {
                int tm = 0; for (sv qt : q_tokens) { bool any = false; for (sv w : words) if (w == qt) { any = true; break; } if (any) tm++; }
                bool all = !q_tokens.empty() && tm == (int)q_tokens.size();
                if (all) { prec |= 8; if (words.size() <= q_tokens.size() + 1) prec |= 2; } else if (tm > 0) prec |= 4;
            } else {
                bool any_exact = false, first_exact = false;
                if (!words.empty()) { first_exact = words[0] == search_lower; any_exact = first_exact; if (!any_exact) for (sv w : words) if (w == search_lower) { any_exact = true; break; } }
                bool title_eq = trimmed == search_lower;
                if (any_exact) prec |= 1;
                if (first_prefix.count(key)) prec |= 2;
                if (first_exact) prec |= 4;
                if (title_eq) prec |= 8;
            }
            ScoreEntry e; e.score = (float)prec + normalized; e.key = key; e.tie = 0; heap.add(e);
        }
        return heap.get_top_k();
    }

    // SearchPipeline.ExecuteRelevancyStage, !canUseNGrams branch (:222-262) + GetTopK + ConsolidateSegments (:83-85)
    std::vector<ScoreEntry> stage1(sv search, int max_results) const {
        if (search.size() == 1) {
            char16_t ch = lo(search[0]); std::vector<ScoreEntry> res; str key(1, ch);
            if (!(max_results < INT32_MAX && try_champions(key, max_results, res))) res = single_character(ch, max_results);
            TopKHeap heap(max_results); for (auto& e : res) heap.add(e);
            return consolidate(heap.get_top_k());
        }
        return consolidate(short_query(to_lower(search)));
    }
};

}  // namespace ifxo
