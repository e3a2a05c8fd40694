#!/usr/bin/env python3
"""Where the host time of an alignGraphsBatch call goes (ctypes marshalling / create / run / fetch)."""
import sys, time, ctypes as C
import numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from prographmsa_b200 import model, synth, progressive, graphalign, _lib
from prographmsa_b200._lib import Context
from prographmsa_b200.options import Options

n_seq = int(sys.argv[1]) if len(sys.argv) > 1 else 256
opts = Options(mldist_flag=True)
mf = model.getDefault(opts)
seqs, root = synth.simulate_family(n_seq, 400, mf, seed=12346)
tree = progressive.from_synth(root)
ctx = Context(0)
jobs = []
def batch(js, o):
    jobs.extend(js)
    return graphalign.alignGraphsBatch(js, o, ctx=ctx)
progressive.progressive_alignment(seqs, tree, mf, opts, batch_align_fn=batch)
jobs = jobs * 4
for it in range(3):
    t0 = time.perf_counter()
    gb = graphalign.GraphBatch(jobs, opts, graphalign.EMISSION_TENSOR, ctx)
    t1 = time.perf_counter()
    gb.run(); ctx.sync()
    t2 = time.perf_counter()
    res = gb.fetch()
    t3 = time.perf_counter()
    gb.close()
    t4 = time.perf_counter()
    print("jobs %d: create %.1f ms (incl. python marshalling)  run %.1f  fetch %.1f  destroy %.1f" %
          (len(jobs), 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3)))
# marshalling alone
t0 = time.perf_counter()
keep = []
G1 = (_lib.Graph * len(jobs))(); G2 = (_lib.Graph * len(jobs))(); MM = (_lib.DPModel * len(jobs))()
for k, (g1, g2, m) in enumerate(jobs):
    G1[k] = graphalign.c_graph(g1, keep); G2[k] = graphalign.c_graph(g2, keep); MM[k] = graphalign.c_model(m, keep)
print("python marshalling alone %.1f ms" % (1e3 * (time.perf_counter() - t0)))
