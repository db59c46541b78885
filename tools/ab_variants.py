"""Scratch: A/B timing of several builds of libinfidex_gpu.so on the configs[1] corpus (1M single-field docs, 1k queries).
usage: ab_variants.py [lib.so ...]   ("default" = the in-tree library). Prints per-phase ms (best of reps) and a digest of the results
so that variants can be checked against each other."""
import sys, time, os, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import infidex_b200 as ib
from infidex_b200 import synth
libs = sys.argv[1:] or ["default"]
N, NQ, reps = int(os.environ.get("AB_N", 1_000_000)), int(os.environ.get("AB_NQ", 1000)), int(os.environ.get("AB_REPS", 5))
vocab = synth.make_vocab(400_000); docs = synth.gen_docs(N, vocab); qs = synth.gen_queries(NQ, docs, vocab)
schema, cols = synth.schema_and_columns(docs, False)
queries = [ib.Query(q, 10) for q in qs]
for lib in libs:
    e = ib.SearchEngine.CreateDefault(_gpu_lib=None if lib == "default" else os.path.abspath(lib)); e.IndexColumns(docs["keys"], schema, cols)
    best = None
    for r in range(reps):
        st = ib.Stats(); res = e.SearchBatch(queries, st); d = st.as_dict()
        if best is None or d["ms_total"] < best["ms_total"]: best = d
    h = hashlib.sha1()
    for x in res:
        h.update(np.array([s.DocumentId for s in x.Records], np.int64).tobytes()); h.update(np.array([s.Score for s in x.Records], np.float32).tobytes())
    print(lib, {k: round(v, 3) for k, v in best.items() if isinstance(v, float)}, "digest", h.hexdigest()[:12], "bad", sum(1 for x in res if x.Status & ~8), flush=True)
    del e
