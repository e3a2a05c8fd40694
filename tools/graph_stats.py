#!/usr/bin/env python3
"""Predecessor statistics of the alignGraphs jobs of a progressive pass (what the fill kernels' slow paths see)."""
import sys, collections
import numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from prographmsa_b200 import model, synth, progressive, graphalign
from prographmsa_b200._lib import Context
from prographmsa_b200.options import Options

n_seq = int(sys.argv[1]) if len(sys.argv) > 1 else 256
L = int(sys.argv[2]) if len(sys.argv) > 2 else 400
opts = Options(mldist_flag=True)
mf = model.getDefault(opts)
seqs, root = synth.simulate_family(n_seq, L, mf, seed=12345 + 1)
tree = progressive.from_synth(root)
ctx = Context(0)
jobs = []
def batch(js, o):
    jobs.extend(js)
    return graphalign.alignGraphsBatch(js, o, ctx=ctx)
progressive.progressive_alignment(seqs, tree, mf, opts, batch_align_fn=batch)
def stats(g):
    non = 0; c3 = 0; far = 0; rep = 0
    for y in range(1, g.N - 1):
        ks = g.e_col[g.e_rowptr[y]:g.e_rowptr[y + 1]]; rs = g.r_col[g.r_rowptr[y]:g.r_rowptr[y + 1]]
        c = len(ks) + len(rs)
        if not (c == 1 and len(ks) == 1 and ks[0] == y - 1):
            non += 1
            if c > 2: c3 += 1
            if len(rs): rep += 1
            if len(ks) and y - min(ks) > 8: far += 1
    return non, c3, far, rep
rows = []
for a, b, m in jobs:
    sa, sb = stats(a), stats(b)
    rows.append((a.N * b.N, a.N, b.N, sa, sb))
rows.sort(reverse=True)
print("jobs", len(rows))
for cells, n1, n2, sa, sb in rows[:12] + rows[len(rows)//2:len(rows)//2+4]:
    print("N1 %4d N2 %4d  g1 non-simple %4d (%.1f%%) >2preds %d far %d rep %d | g2 non-simple %4d (%.1f%%) >2preds %d far %d" %
          (n1, n2, sa[0], 100.0 * sa[0] / n1, sa[1], sa[2], sa[3], sb[0], 100.0 * sb[0] / n2, sb[1], sb[2]))
