#!/usr/bin/env python3
"""Strip-kernel timing on synthetic leaf x leaf jobs (no progressive pass: the first launch is the one to profile).
  python tools/profile_strip.py [jobs=1184] [length=400] [profile_cols=0]    profile_cols=1: graph 2's columns are real profiles"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
os.environ["PGM_GRAPH_PIPELINE"] = "0"
import numpy as np
from prographmsa_b200 import model, synth, graphalign, _lib
from prographmsa_b200.graph import sequence_graph
from prographmsa_b200.options import Options

nj = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
L = int(sys.argv[2]) if len(sys.argv) > 2 else 400
prof = int(sys.argv[3]) if len(sys.argv) > 3 else 0
opts = Options(mldist_flag=True)
mf = model.getDefault(opts)
m = mf.getModel(0.5, 0.5)
base = []
for k in range(16):
    a, b = synth.random_pair(L, 0.5, mf, seed=1000 + k)
    g1, g2 = sequence_graph(a, 20), sequence_graph(b, 20)
    if prof:
        g2.sites[1:-1] = g2.sites[1:-1] @ m.P.T
    base.append((g1, g2, m))
jobs = [base[k % 16] for k in range(nj)]
ctx = _lib.default_context()
ctx.enable_timing(True)
gb = graphalign.GraphBatch(jobs, opts, 0, ctx)
acc = dict(emission_ms=0.0, graph_fill_ms=0.0, graph_trace_ms=0.0)
n = 0
for it in range(4):
    gb.run(); ctx.sync()
    if it >= 1:
        t = ctx.timing(); n += 1
        for k in acc: acc[k] += t[k]
cells = sum(g1.N * g2.N for g1, g2, _ in jobs)
print("%d jobs of %d x %d: emission %.3f fill %.3f ms (%.1f GCUPS) trace %.3f ms" % (
    nj, L, L, acc["emission_ms"] / n, acc["graph_fill_ms"] / n, cells / (acc["graph_fill_ms"] / n) / 1e6, acc["graph_trace_ms"] / n))
