#!/usr/bin/env python3
"""Kernel timing of the bench's progressive-pass jobs, split by how many non-chain nodes the graphs have.

  python tools/profile_hp1_mix.py [n_seq=256] [replicas=4] [only]      (PGM_LIB_PATH=<other build> for A/B runs)
Phases run one after the other (PGM_GRAPH_PIPELINE=0), so fill_ms is the wall time of the fill section."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
os.environ["PGM_GRAPH_PIPELINE"] = "0"
import numpy as np
from prographmsa_b200 import model, synth, graphalign, progressive, _lib
from prographmsa_b200.options import Options

n_seq = int(sys.argv[1]) if len(sys.argv) > 1 else 256
rep = int(sys.argv[2]) if len(sys.argv) > 2 else 4
opts = Options(mldist_flag=True)
mf = model.getDefault(opts)
seqs, root = synth.simulate_family(n_seq, 400, mf, seed=2)
tree = progressive.from_synth(root)
ctx = _lib.default_context()
jobs = []
def batch(js, o):
    jobs.extend(js)
    return graphalign.alignGraphsBatch(js, o, ctx=ctx)
progressive.progressive_alignment(seqs, tree, mf, opts, batch_align_fn=batch)

def nonchain(g):
    n = 0
    for y in range(1, g.N - 1):
        ks = g.e_col[g.e_rowptr[y]:g.e_rowptr[y + 1]]; rs = g.r_col[g.r_rowptr[y]:g.r_rowptr[y + 1]]
        if not (len(ks) == 1 and len(rs) == 0 and ks[0] == y - 1):
            n += 1
    return n
info = [(nonchain(a), nonchain(b), a.N, b.N) for a, b, m in jobs]
ctx.enable_timing(True)
def run(sel, label, mode=0):
    js = [jobs[i] for i in sel] * rep
    if not js:
        return
    gb = graphalign.GraphBatch(js, opts, mode, ctx)
    acc = dict(emission_ms=0.0, graph_fill_ms=0.0, graph_trace_ms=0.0)
    for it in range(5):
        gb.run(); ctx.sync()
        if it >= 2:
            t = ctx.timing()
            for k in acc: acc[k] += t[k] / 3
    cells = sum(info[i][2] * info[i][3] for i in sel) * rep
    print("%-34s jobs %4d cells %.3g: emission %.3f fill %.3f ms (%.1f GCUPS) trace %.3f ms" % (
        label, len(js), cells, acc["emission_ms"], acc["graph_fill_ms"], cells / acc["graph_fill_ms"] / 1e6, acc["graph_trace_ms"]))
    gb.close()
idx = list(range(len(jobs)))
only = sys.argv[3] if len(sys.argv) > 3 else None
if only == "few":
    run([i for i in idx if 0 < max(info[i][0], info[i][1]) <= 8], "1..8 non-chain nodes"); sys.exit(0)
if only == "big":
    big = sorted(idx, key=lambda i: -info[i][2] * info[i][3])[:1]
    for dbgv in os.environ.get("DBG_LIST", "0").split(","):
        os.environ["PGM_STRIP_DEBUG"] = dbgv
        if dbgv != "0": os.environ["PGM_SKIP_TRACE"] = "1"
        run(big, "largest job dbg=" + dbgv)
    sys.exit(0)
if only == "chain":
    run([i for i in idx if info[i][0] == 0 and info[i][1] == 0], "pure chain x chain"); sys.exit(0)
if only == "branchy":
    run([i for i in idx if info[i][0] + info[i][1] > 40], ">40 non-chain nodes"); sys.exit(0)
lays = (None,) if only == "default" else (None, "block", "diag")
for lay in lays:
    if lay: os.environ["PGM_FILL_LAYOUT"] = lay
    print("--- PGM_FILL_LAYOUT =", lay)
    run(idx, "all")
    run(idx, "all (tensor emission)", 1)
    run([i for i in idx if info[i][0] == 0 and info[i][1] == 0], "pure chain x chain")
    run([i for i in idx if 0 < info[i][0] + info[i][1] <= 4], "1..4 non-chain nodes")
    run([i for i in idx if 4 < info[i][0] + info[i][1] <= 8], "5..8 non-chain nodes")
    run([i for i in idx if 8 < info[i][0] + info[i][1] <= 16], "9..16 non-chain nodes")
    run([i for i in idx if 16 < info[i][0] + info[i][1] <= 40], "17..40 non-chain nodes")
    run([i for i in idx if info[i][0] + info[i][1] > 40], ">40 non-chain nodes")
    big = sorted(idx, key=lambda i: -info[i][2] * info[i][3])[:1]
    print("largest job: %d x %d nodes, %d + %d non-chain" % (info[big[0]][2], info[big[0]][3], info[big[0]][0], info[big[0]][1]))
    os.environ["PGM_FILL_SMALL_BATCH"] = "0"
    rep_save = rep; rep = 1
    run(big, "largest job alone")
    rep = rep_save
    os.environ.pop("PGM_FILL_SMALL_BATCH")
os.environ.pop("PGM_FILL_LAYOUT", None)
