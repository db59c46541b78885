"""Scratch: Stage-1 timing on the configs[1] corpus (1M single-field docs, 1k queries)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import infidex_b200 as ib
from infidex_b200 import synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
NQ = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
t0 = time.time(); vocab = synth.make_vocab(400_000); docs = synth.gen_docs(N, vocab); qs = synth.gen_queries(NQ, docs, vocab); print("gen %.1fs" % (time.time() - t0), flush=True)
schema, cols = synth.schema_and_columns(docs, False)
t0 = time.time(); e = ib.SearchEngine.CreateDefault(); e.IndexColumns(docs["keys"], schema, cols); print("index %.1fs" % (time.time() - t0), flush=True)
for r in range(reps):
    st = ib.Stats(); t0 = time.time(); k, s, n, status = e.Stage1Batch(qs, 500, st); dt = time.time() - t0
    print("rep", r, "wall %.1f ms" % (dt * 1e3), {k2: round(v, 3) if isinstance(v, float) else v for k2, v in st.as_dict().items()}, "mean n", n.mean(), "bad status", int((status != 0).sum()), flush=True)
queries = [ib.Query(q, 10) for q in qs]
for r in range(reps):
    st = ib.Stats(); t0 = time.time(); res = e.SearchBatch(queries, st); dt = time.time() - t0
    print("search rep", r, "wall %.1f ms" % (dt * 1e3), {k2: round(v, 3) if isinstance(v, float) else v for k2, v in st.as_dict().items()}, "bad status", sum(1 for x in res if x.Status & ~8), flush=True)
