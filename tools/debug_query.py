import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from infidex_b200 import synth
from parity_util import build_pair, compare_stage1, compare_search
vocab = synth.make_vocab(100_000); docs = synth.gen_docs(300_000, vocab); qs = synth.gen_queries(600, docs, vocab)
schema, cols = synth.schema_and_columns(docs, False)
eng, orc = build_pair(docs["keys"], schema, cols)
for rep in range(3):
    bad = compare_stage1(eng, orc, qs)
    print("rep", rep, "stage1 bad", len(bad)); [print("  ", b) for b in bad[:6]]
bad = compare_search(eng, orc, qs); print("search bad", len(bad)); [print("  ", b) for b in bad[:6]]
