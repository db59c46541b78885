"""How often would speculative multi-chunk scoring be invalidated? (offline study on the CPU, oracle only -- development tool)

Chunk c of a query is scored with the MaxScore threshold thr_c left by chunk c-1. Scoring it with a stale threshold thr' <= thr_c gives
identical results iff no matched (candidate, term) pair that the true run skips has score+bound+suffix > thr' (then the skip sets are
equal). The oracle's per-chunk trace gives thr_c and the maximum of that quantity over the skipped matched pairs; this script replays
"speculate with the threshold of L chunks ago" and counts violations.   usage: speculation_study.py N_DOCS N_QUERIES
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import infidex_b200 as ib                       # noqa: E402
from infidex_b200 import synth                  # noqa: E402
from oracle import oracle as O                  # noqa: E402

N = int(sys.argv[1]); NQ = int(sys.argv[2])
vocab = synth.make_vocab(400_000); docs = synth.gen_docs(N, vocab, with_description=True)
schema, cols = synth.schema_and_columns(docs, True)
eng = ib.SearchEngine.__new__(ib.SearchEngine); eng._host = ib.engine._load_host(); eng._builder = None; eng._index = None; eng._gpu = None
eng.IndexColumns(docs["keys"], schema, cols, upload=False)
orc = O.OracleEngine([O.Field(f.Name, f.Weight, f.Indexable, f.Filterable, f.Facetable) for f in schema]); orc.load_image(eng.image_ptr())
qs = synth.gen_queries(NQ, docs, vocab)
lags = [1, 2, 4, 8, 16, 32, 64]
viol = {L: 0 for L in lags}; chunks = 0; full_chunks = 0; cand_tot = 0; skipped_tot = 0; matched_tot = 0; upd = []
per_q = []
for q in qs:
    a = O.u16(q); out = np.zeros((4096, 6)); st = np.zeros(5, np.int64)
    n = O.lib().ifxo_stage1_trace(orc.h, O._p(a), len(a), 500, O._p(out), 4096, O._p(st))
    tr = out[:n]; chunks += n; cand_tot += int(st[1]); per_q.append((int(st[1]), n, int(st[3]), int(st[0])))
    if n == 0:
        continue
    thr = tr[:, 0]; vmax = tr[:, 4]; matched_tot += tr[:, 2].sum(); skipped_tot += tr[:, 3].sum()
    full = np.nonzero(thr > 0)[0]              # chunks that start with a full heap
    full_chunks += len(full); upd += tr[full, 5].tolist()
    for L in lags:
        for c in full:
            spec = thr[c - L] if c - L >= 0 else 0.0
            if vmax[c] > spec:
                viol[L] += 1
pq = np.array(per_q)
print("docs %d queries %d: candidates/query mean %.0f median %.0f max %d; chunks/query mean %.1f; terms mean %.1f; paths %s" % (
    N, NQ, pq[:, 0].mean(), np.median(pq[:, 0]), pq[:, 0].max(), pq[:, 1].mean(), pq[:, 2].mean(), np.bincount(pq[:, 3]).tolist()))
print("matched pairs %.3g, skipped by MaxScore %.3g (%.1f%%); heap updates per full-heap chunk: mean %.2f median %.1f" % (matched_tot, skipped_tot, 100 * skipped_tot / max(matched_tot, 1), np.mean(upd), np.median(upd)))
for L in lags:
    print("lag %3d chunks: %6d of %d full-heap chunks violated (%.2f%%)" % (L, viol[L], full_chunks, 100.0 * viol[L] / max(full_chunks, 1)))
