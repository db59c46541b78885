"""Offline check (CPU, minutes): the kernel sources compiled as the host emulation against the oracle on the BENCHMARK corpus itself --
configs[1] (1 M single-field docs, the 1000 queries bench.py times) and a 400 k multi-field corpus with the configs[3] filter + facets.
Bit-exact Stage-1 lists and final records (tests/parity_util.py). Last run at the end of round 1: 0 mismatches."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import infidex_b200 as ib
from infidex_b200 import synth
from parity_util import build_pair, compare_search, compare_stage1, emu_lib

vocab = synth.make_vocab(400_000); docs = synth.gen_docs(1_000_000, vocab); qs = synth.gen_queries(1000, docs, vocab)
schema, cols = synth.schema_and_columns(docs, False)
eng, orc = build_pair(docs["keys"], schema, cols, gpu_lib=emu_lib())
total = 0
for a in range(0, 1000, 100):
    bad = compare_stage1(eng, orc, qs[a:a + 100]) + compare_search(eng, orc, qs[a:a + 100]); total += len(bad)
    print("configs[1] queries %d-%d: %d mismatches" % (a, a + 99, len(bad)), str(bad[:1])[:300], flush=True)
del eng, orc
docs = synth.gen_docs(400_000, vocab, with_description=True); qs = synth.gen_queries(300, docs, vocab)
schema, cols = synth.schema_and_columns(docs, True)
eng, orc = build_pair(docs["keys"], schema, cols, gpu_lib=emu_lib())
bad = compare_stage1(eng, orc, qs) + compare_search(eng, orc, qs) + compare_search(eng, orc, qs[:150], flt=ib.Filter.Parse("year >= 2000 AND rating > 7.0"), facets=True)
print("multi-field 400k (+ filter, facets): %d mismatches" % len(bad), str(bad[:1])[:300]); total += len(bad)
sys.exit(1 if total else 0)
