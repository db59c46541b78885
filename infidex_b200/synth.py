"""Deterministic synthetic corpus + query generator for the benchmark configs (SURVEY.md 8d / BASELINE.md 2).

Vocabulary: V lower-case ASCII pseudo-words (3-12 chars, syllable generator), rank-frequency Zipf(s=1.07).
Docs     : title = 1 + Poisson(2.7) words (cap 12); description = 8 + Poisson(7) words; year U{1950..2024};
           rating U{1.0..10.0} step 0.1; genre in 20 categories (Zipf).
Queries  : 1-3 consecutive title words of a random doc; p=0.30 one edit (sub/ins/del/transpose) in a word of len >= 4;
           p=0.20 last word cut to a prefix of >= 3 chars; p=0.05 a word absent from the doc prepended.
All draws come from numpy Generator(PCG64(seed)) with seed 0x1F1DEC5 (documented substitute for the survey's SplitMix64).
"""
import numpy as np

SEED = 0x1F1DEC5
GENRES = ["drama", "comedy", "action", "thriller", "romance", "horror", "documentary", "crime", "adventure", "family",
          "fantasy", "mystery", "animation", "biography", "history", "music", "war", "western", "sport", "musical"]
_CONS = "bcdfghjklmnprstvwz"
_VOW = "aeiou"


def make_vocab(V=400_000, seed=SEED):
    rng = np.random.Generator(np.random.PCG64(seed ^ 0xA5A5))
    words, seen = [], set()
    cons, vow = np.array(list(_CONS)), np.array(list(_VOW))
    while len(words) < V:
        n = (V - len(words)) * 2 + 1000
        nsyl = rng.integers(1, 5, n)
        c1 = cons[rng.integers(0, len(cons), (n, 4))]; v1 = vow[rng.integers(0, len(vow), (n, 4))]; c2 = cons[rng.integers(0, len(cons), (n, 4))]
        tail = rng.random((n, 4)) < 0.35
        for i in range(n):
            w = "".join(c1[i, k] + v1[i, k] + (c2[i, k] if tail[i, k] else "") for k in range(nsyl[i]))
            if 3 <= len(w) <= 12 and w not in seen:
                seen.add(w); words.append(w)
                if len(words) == V:
                    break
    lens = np.array([len(w) for w in words], np.int32)
    offs = np.zeros(V + 1, np.int64); np.cumsum(lens, out=offs[1:])
    blob = np.frombuffer("".join(words).encode("ascii"), np.uint8).astype(np.uint16)
    ranks = np.arange(1, V + 1, dtype=np.float64)
    cdf = np.cumsum(ranks ** -1.07); cdf /= cdf[-1]
    return {"words": words, "blob": blob, "offs": offs, "lens": lens, "cdf": cdf}


def _join_words(vocab, ids, doc_off):
    """UTF-16 blob + offsets of docs whose words (ids, ragged by doc_off) are joined by single spaces."""
    n = len(doc_off) - 1
    wl = vocab["lens"][ids].astype(np.int64)
    out_len = wl + 1                                   # word + trailing space
    last = np.zeros(len(ids), bool); last[doc_off[1:][doc_off[1:] > doc_off[:-1]] - 1] = True
    out_len[last] -= 1                                 # no space after a doc's last word
    wstart = np.zeros(len(ids) + 1, np.int64); np.cumsum(out_len, out=wstart[1:])
    total = int(wstart[-1])
    blob = np.full(total, 32, np.uint16)
    # char index gather: for every output char of every word, its source index in the vocab blob
    rep = np.repeat(np.arange(len(ids)), wl)
    within = np.arange(int(wl.sum()), dtype=np.int64) - np.repeat(np.cumsum(wl) - wl, wl)
    blob[wstart[:-1][rep] + within] = vocab["blob"][vocab["offs"][ids][rep] + within]
    offs = np.zeros(n + 1, np.int64); offs[:] = wstart[doc_off]
    return blob, offs


def gen_docs(n, vocab, seed=SEED, with_description=False, start=0):
    """Columns for docs [start, start+n). Deterministic per (seed, start)."""
    rng = np.random.Generator(np.random.PCG64([seed, start]))
    V = len(vocab["lens"])
    tcount = np.minimum(1 + rng.poisson(2.7, n), 12).astype(np.int64)
    toff = np.zeros(n + 1, np.int64); np.cumsum(tcount, out=toff[1:])
    tids = np.minimum(np.searchsorted(vocab["cdf"], rng.random(int(toff[-1]))), V - 1).astype(np.int64)
    title = _join_words(vocab, tids, toff)
    out = {"n": n, "keys": np.arange(start, start + n, dtype=np.int64), "title": title, "title_ids": tids, "title_off": toff}
    if with_description:
        dcount = (8 + rng.poisson(7.0, n)).astype(np.int64)
        doff = np.zeros(n + 1, np.int64); np.cumsum(dcount, out=doff[1:])
        dids = np.minimum(np.searchsorted(vocab["cdf"], rng.random(int(doff[-1]))), V - 1).astype(np.int64)
        out["description"] = _join_words(vocab, dids, doff)
    out["year"] = rng.integers(1950, 2025, n).astype(np.int64)
    out["rating"] = np.round(rng.integers(10, 101, n) / 10.0, 1).astype(np.float64)
    gr = np.arange(1, 21, dtype=np.float64) ** -1.07; gc = np.cumsum(gr) / gr.sum()
    out["genre_id"] = np.minimum(np.searchsorted(gc, rng.random(n)), 19)
    return out


def gen_queries(nq, docs, vocab, seed=SEED):
    """Query strings sampled from the titles of `docs` (a gen_docs result)."""
    rng = np.random.Generator(np.random.PCG64([seed, 0x51]))
    words = vocab["words"]; out = []
    alpha = "abcdefghijklmnopqrstuvwxyz"
    for _ in range(nq):
        d = int(rng.integers(0, docs["n"]))
        a, b = int(docs["title_off"][d]), int(docs["title_off"][d + 1])
        k = int(min(rng.integers(1, 4), b - a)); s = int(rng.integers(a, b - k + 1))
        ws = [words[i] for i in docs["title_ids"][s:s + k]]
        if rng.random() < 0.30:
            cand = [i for i, w in enumerate(ws) if len(w) >= 4]
            if cand:
                i = cand[int(rng.integers(0, len(cand)))]; w = ws[i]; p = int(rng.integers(0, len(w))); kind = int(rng.integers(0, 4))
                c = alpha[int(rng.integers(0, 26))]
                if kind == 0: w = w[:p] + c + w[p + 1:]
                elif kind == 1: w = w[:p] + c + w[p:]
                elif kind == 2: w = w[:p] + w[p + 1:]
                else:
                    p = min(p, len(w) - 2); w = w[:p] + w[p + 1] + w[p] + w[p + 2:]
                ws[i] = w
        if rng.random() < 0.20 and len(ws[-1]) > 3:
            ws[-1] = ws[-1][: int(rng.integers(3, len(ws[-1])))]
        if rng.random() < 0.05:
            title_set = set(docs["title_ids"][a:b].tolist())
            while True:
                x = int(min(np.searchsorted(vocab["cdf"], rng.random()), len(words) - 1))
                if x not in title_set:
                    break
            ws.insert(0, words[x])
        out.append(" ".join(ws))
    return out


def schema_and_columns(docs, multi_field):
    """(schema, columns) for SearchEngine.IndexColumns / the oracle, per BASELINE.json configs."""
    from .engine import Field, Weight
    if not multi_field:      # configs[1]: new Document(i, title) -> single field "content", Weight.Med
        return [Field("content", None, Weight.Med)], [docs["title"]]
    genre = [GENRES[g] for g in docs["genre_id"]]
    schema = [Field("title", None, Weight.High), Field("description", None, Weight.Low),
              Field("year", None, Weight.Med, indexable=False, filterable=True, facetable=True),
              Field("rating", None, Weight.Med, indexable=False, filterable=True),
              Field("genre", None, Weight.Med, indexable=False, filterable=True, facetable=True)]
    return schema, [docs["title"], docs["description"], docs["year"], docs["rating"], genre]


class ChunkedCorpus:
    """Streams a large synthetic corpus chunk by chunk (deterministic per chunk) and keeps only what query sampling needs."""

    def __init__(self, n_docs, vocab, multi_field, chunk=500_000, seed=SEED):
        self.n, self.vocab, self.multi, self.chunk, self.seed = n_docs, vocab, multi_field, chunk, seed
        self.title_ids, self.title_off = [], [np.zeros(1, np.int64)]
        self.schema = schema_and_columns({"title": None, "description": None, "year": None, "rating": None, "genre_id": np.zeros(0, np.int64)}, multi_field)[0]
        self.text_chars = 0

    def chunks(self, workers=1):
        """Yields (keys, columns) in document order; chunk generation (numpy, releases the GIL) runs `workers` chunks ahead."""
        from concurrent.futures import ThreadPoolExecutor
        starts = list(range(0, self.n, self.chunk))
        gen = lambda st: gen_docs(min(self.chunk, self.n - st), self.vocab, seed=self.seed, with_description=self.multi, start=st)
        with ThreadPoolExecutor(max_workers=max(1, workers)) as ex:
            pending = [ex.submit(gen, st) for st in starts[:workers]]; nxt = workers
            for _ in starts:
                d = pending.pop(0).result()
                if nxt < len(starts):
                    pending.append(ex.submit(gen, starts[nxt])); nxt += 1
                self.title_ids.append(d["title_ids"].astype(np.int32)); self.title_off.append(d["title_off"][1:] + self.title_off[-1][-1])
                self.text_chars += int(d["title"][1][-1]) + (int(d["description"][1][-1]) if self.multi else 0)
                yield d["keys"], schema_and_columns(d, self.multi)[1]

    def docs_for_queries(self):
        return {"n": self.n, "title_ids": np.concatenate(self.title_ids), "title_off": np.concatenate(self.title_off)}
