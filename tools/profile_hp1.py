#!/usr/bin/env python3
"""Small driver for ncu captures of the HP-1 kernels: a batch of leaf x leaf alignGraphs jobs."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from prographmsa_b200 import model, synth, graphalign, _lib
from prographmsa_b200.graph import sequence_graph
from prographmsa_b200.options import Options

n_jobs = int(sys.argv[1]) if len(sys.argv) > 1 else 296
L = int(sys.argv[2]) if len(sys.argv) > 2 else 400
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
opts = Options(mldist_flag=True)
mf = model.getDefault(opts)
m = mf.getModel(0.5, 0.5)
jobs = []
for k in range(n_jobs):
    a, b = synth.random_pair(L, 0.5, mf, seed=k)
    jobs.append((sequence_graph(a, 20), sequence_graph(b, 20), m))
ctx = _lib.default_context()
ctx.enable_timing(True)
gb = graphalign.GraphBatch(jobs, opts, mode, ctx)
for _ in range(3):
    gb.run(); ctx.sync()
t = ctx.timing()
upd = t["graph_cell_updates"]
print("jobs %d L %d: emission %.3f ms fill %.3f ms trace %.3f ms -> fill %.1f GCUPS" % (
    n_jobs, L, t["emission_ms"], t["graph_fill_ms"], t["graph_trace_ms"], upd / t["graph_fill_ms"] / 1e6))
res = gb.fetch()
print("score checksum", sum(r["score"] for r in res))
