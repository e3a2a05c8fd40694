#!/usr/bin/env python3
"""The alignGraphs jobs of the top levels of a 5000-sequence progressive pass (the long, branchy profile graphs that run the
banded anti-diagonal kernel), run once more between cudaProfilerStart / Stop:

  ncu --profile-from-start off --set full --import-source on -k regex:graph_fill_diag -o out python tools/profile_root_job.py [n_seq=5000] [levels=1]"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from prographmsa_b200 import model, synth, graphalign, progressive, _lib
from prographmsa_b200.options import Options
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
top = int(sys.argv[2]) if len(sys.argv) > 2 else 1
opts = Options(mldist_flag=True)
mf = model.getDefault(opts)
seqs, root = synth.simulate_family(n, 500, mf, seed=50, forbid_first=10)
tree = progressive.parse_newick(root.newick(6) + ";")
ctx = _lib.default_context()
ctx.enable_timing(True)
levels = []
def batch(js, o):
    levels.append(js)
    return graphalign.alignGraphsBatch(js, o, ctx=ctx)
progressive.progressive_alignment(seqs, tree, mf, opts, batch_align_fn=batch, batch_merge_fn=lambda jobs: graphalign.merge_profiles_batch(jobs))
import torch
os.environ["PGM_GRAPH_PIPELINE"] = "0"
for js in levels[-top:]:
    gb = graphalign.GraphBatch(js, opts, 0, ctx)
    gb.run(); ctx.sync()
    torch.cuda.profiler.start()
    gb.run(); ctx.sync()
    torch.cuda.profiler.stop()
    t = ctx.timing()
    print("%d jobs, largest %d x %d: fill %.3f ms" % (len(js), max(a.N for a, b, m in js), max(b.N for a, b, m in js), t["graph_fill_ms"]), flush=True)
    gb.close()
