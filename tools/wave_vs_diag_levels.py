#!/usr/bin/env python3
"""Per tree level of a progressive pass: fill time of the anti-diagonal jobs with the wave kernel against the thread-per-row
kernel (PGM_FILL_WAVE=1 / 0), with the levels' non-chain node counts.  python tools/wave_vs_diag_levels.py [n_seq=1000] [L=500]"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from prographmsa_b200 import model, synth, graphalign, progressive, _lib
from prographmsa_b200.options import Options

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 500
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
def batch(js, o):
    lvl[0] += 1
    res = {}
    if len(js) <= 200:
        for w in ("1", "0"):
            os.environ["PGM_FILL_WAVE"] = w
            os.environ["PGM_GRAPH_PIPELINE"] = "0"
            gb = graphalign.GraphBatch(js, o, 0, ctx)
            acc = 0.0
            for it in range(3):
                gb.run(); ctx.sync(); t = ctx.timing()
                if it: acc += t["graph_fill_ms"] / 2
            gb.close()
            res[w] = acc
        os.environ.pop("PGM_FILL_WAVE"); os.environ.pop("PGM_GRAPH_PIPELINE")
        st = [(a.N, b.N, nonchain(a), nonchain(b)) for a, b, m in js]
        big = max(st, key=lambda s: s[0] * s[1])
        print("level %2d: %3d jobs, largest %4d x %4d with %3d + %3d non-chain; mean non-chain %5.1f + %5.1f: fill wave %7.3f ms, thread-per-row %7.3f ms" % (
            lvl[0], len(js), big[0], big[1], big[2], big[3], np.mean([s[2] for s in st]), np.mean([s[3] for s in st]), res["1"], res["0"]), flush=True)
    return graphalign.alignGraphsBatch(js, o, ctx=ctx)
progressive.progressive_alignment(seqs, tree, mf, opts, batch_align_fn=batch, batch_merge_fn=lambda jobs: graphalign.merge_profiles_batch(jobs))
