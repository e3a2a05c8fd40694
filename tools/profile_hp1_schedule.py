import sys, os, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from prographmsa_b200 import model, synth, graphalign, progressive, _lib
from prographmsa_b200.options import Options
opts = Options(mldist_flag=True); mf = model.getDefault(opts)
seqs, root = synth.simulate_family(256, 400, mf, seed=2)
tree = progressive.from_synth(root)
ctx = _lib.default_context(); jobs = []
def batch(js, o):
    jobs.extend(js); return graphalign.alignGraphsBatch(js, o, ctx=ctx)
progressive.progressive_alignment(seqs, tree, mf, opts, batch_align_fn=batch)
jobs = jobs * 4
for mode in (0, 1):
    gb = graphalign.GraphBatch(jobs, opts, mode, ctx)
    for pl in ("0", "1"):
        os.environ["PGM_GRAPH_PIPELINE"] = pl
        for _ in range(3): gb.run(); ctx.sync()
        ts = []
        for _ in range(8):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); a.record(); gb.run(); ctx.sync(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        print("mode %d pipeline %s: %.3f ms (min %.3f)" % (mode, pl, sum(ts) / len(ts), min(ts)))
    gb.close()
