#!/usr/bin/env python3
"""Per tree level of a progressive pass: fill time of the level's alignGraphs batch under a list of environment settings.

  python tools/levels_knobs.py n_seq L max_jobs "PGM_FILL_BAND_ROWS=256" "PGM_FILL_BAND_ROWS=512,PGM_FILL_RPT=2" ...
Only levels with at most max_jobs jobs are timed (the upper, latency-bound levels); the default build runs first."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from prographmsa_b200 import model, synth, graphalign, progressive, _lib
from prographmsa_b200.options import Options

n = int(sys.argv[1]); L = int(sys.argv[2]); max_jobs = int(sys.argv[3])
settings = [""] + sys.argv[4:]
opts = Options(mldist_flag=True)
mf = model.getDefault(opts)
seqs, root = synth.simulate_family(n, L, mf, seed=50, forbid_first=10)
tree = progressive.parse_newick(root.newick(6) + ";")
ctx = _lib.default_context()
ctx.enable_timing(True)
def nonchain(g):
    c = 0
    for y in range(1, g.N - 1):
        ks = g.e_col[g.e_rowptr[y]:g.e_rowptr[y + 1]]; rs = g.r_col[g.r_rowptr[y]:g.r_rowptr[y + 1]]
        if not (len(ks) == 1 and len(rs) == 0 and ks[0] == y - 1):
            c += 1
    return c
lvl = [0]
tot = {s: 0.0 for s in settings}
def batch(js, o):
    lvl[0] += 1
    if len(js) <= max_jobs:
        res = []
        ref = None
        for s in settings:
            kv = [q.split("=", 1) for q in s.split(",") if q]
            for k, v in kv: os.environ[k] = v
            os.environ["PGM_GRAPH_PIPELINE"] = "0"
            gb = graphalign.GraphBatch(js, o, 0, ctx)
            acc = 0.0
            for it in range(3):
                gb.run(); ctx.sync(); t = ctx.timing()
                if it: acc += t["graph_fill_ms"] / 2
            r = gb.fetch()
            chk = (sum(x["score"] for x in r), sum(int(x["mapping1"].sum()) for x in r))
            if ref is None: ref = chk
            gb.close()
            for k, v in kv: os.environ.pop(k, None)
            os.environ.pop("PGM_GRAPH_PIPELINE")
            res.append("%8.3f%s" % (acc, "" if chk == ref else "!"))
            tot[s] += acc
        st = [(a.N, b.N, nonchain(a), nonchain(b)) for a, b, m in js]
        big = max(st, key=lambda q: q[0] * q[1])
        print("level %2d: %3d jobs, largest %4d x %4d (%3d + %3d non-chain): fill ms %s" % (lvl[0], len(js), big[0], big[1], big[2], big[3], " ".join(res)), flush=True)
    return graphalign.alignGraphsBatch(js, o, ctx=ctx)
progressive.progressive_alignment(seqs, tree, mf, opts, batch_align_fn=batch, batch_merge_fn=lambda jobs: graphalign.merge_profiles_batch(jobs))
for s in settings: print("total %-50s %9.2f ms" % (s or "(default)", tot[s]))
