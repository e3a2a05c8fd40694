#!/usr/bin/env python3
"""Edge-span statistics of the bench's progressive-pass jobs, computed on the CPU with the oracle as the aligner
(test infrastructure: drives the oracle).  python tests/edge_span_stats.py [n_seq=256] [L=400] [pickle-out]"""
import sys, os, collections, pickle
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from prographmsa_b200 import model, synth, progressive
from prographmsa_b200.options import Options
from oracle import pyoracle as po

n_seq = int(sys.argv[1]) if len(sys.argv) > 1 else 256
L = int(sys.argv[2]) if len(sys.argv) > 2 else 400
opts = Options(mldist_flag=True)
mf = model.getDefault(opts)
seqs, root = synth.simulate_family(n_seq, L, mf, seed=2)
tree = progressive.from_synth(root)
jobs = []
def batch(js, o):
    jobs.extend(js)
    return [po.align_graphs(a, b, m, o) for a, b, m in js]
progressive.progressive_alignment(seqs, tree, mf, opts, batch_align_fn=batch)
span = collections.Counter(); cnts = collections.Counter()
per_job = []
for a, b, m in jobs:
    tot = 0
    for g in (a, b):
        for y in range(1, g.N - 1):
            ks = list(g.e_col[g.e_rowptr[y]:g.e_rowptr[y + 1]]); rs = list(g.r_col[g.r_rowptr[y]:g.r_rowptr[y + 1]])
            c = len(ks) + len(rs)
            cnts[c] += 1
            if c == 1 and ks and ks[0] == y - 1: continue
            tot += 1
            for p in ks + rs:
                if p != y - 1: span[min(y - p, 64)] += 1
    per_job.append((a.N, b.N, tot))
print("jobs", len(jobs), "pred-count histogram", sorted(cnts.items()))
tot = sum(span.values()); acc = 0
for s in sorted(span):
    acc += span[s]
    print("span %2d: %6d  cum %.3f" % (s, span[s], acc / tot))
if len(sys.argv) > 3:
    pickle.dump(jobs, open(sys.argv[3], "wb"))
