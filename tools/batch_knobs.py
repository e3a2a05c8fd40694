#!/usr/bin/env python3
"""Wall time of the bench's 1020-job alignGraphs batch (pipelined schedule) under a list of environment settings.

  python tools/batch_knobs.py "PGM_STRIP_MAX_GENERIC=32" "PGM_STRIP_MAX_GENERIC=64,PGM_FILL_WAVE=0" ...
Each argument is one comma-separated setting; the default build runs first.  The plan is made when the batch is staged,
so every setting stages its own batch."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from prographmsa_b200 import model, synth, graphalign, progressive, _lib
from prographmsa_b200.options import Options
opts = Options(mldist_flag=True)
mf = model.getDefault(opts)
seqs, root = synth.simulate_family(256, 400, mf, seed=2)
tree = progressive.from_synth(root)
ctx = _lib.default_context()
jobs = []
def batch(js, o):
    jobs.extend(js)
    return graphalign.alignGraphsBatch(js, o, ctx=ctx)
progressive.progressive_alignment(seqs, tree, mf, opts, batch_align_fn=batch)
jobs = jobs * 4
ref = None
for setting in [""] + sys.argv[1:]:
    kv = [s.split("=", 1) for s in setting.split(",") if s]
    for k, v in kv: os.environ[k] = v
    try:
        gb = graphalign.GraphBatch(jobs, opts, 0, ctx)
        for _ in range(3):
            gb.run(); ctx.sync()
        ts = []
        for _ in range(10):
            t0 = time.perf_counter(); gb.run(); ctx.sync(); ts.append(time.perf_counter() - t0)
        res = gb.fetch()
        chk = (float(sum(r["score"] for r in res)), sum(int(r["mapping1"].sum()) + int(r["mapping2"].sum()) for r in res))
        if ref is None: ref = chk
        print("%-60s median %.3f ms  min %.3f ms  %s" % (setting or "(default)", sorted(ts)[5] * 1e3, min(ts) * 1e3, "same results" if chk == ref else "RESULTS DIFFER"), flush=True)
        gb.close()
    except Exception as e:
        print("%-60s failed: %s" % (setting, e), flush=True)
    for k, v in kv: os.environ.pop(k, None)
