#!/usr/bin/env python3
"""Run a progressive pass and, level by level, compare the alignGraphs results of two kernel selections (environment
switches) job by job.  python tools/compare_flavours.py [n_seq=1000] [L=500]   (A: PGM_FILL_STRIP=1, B: default)"""
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
bad = [0]
def nonchain(g):
    c = 0
    for y in range(1, g.N - 1):
        ks = g.e_col[g.e_rowptr[y]:g.e_rowptr[y + 1]]; rs = g.r_col[g.r_rowptr[y]:g.r_rowptr[y + 1]]
        if not (len(ks) == 1 and len(rs) == 0 and ks[0] == y - 1):
            c += 1
    return c
def batch(js, o):
    os.environ["PGM_FILL_STRIP"] = "1"
    ra = graphalign.alignGraphsBatch(js, o, ctx=ctx)
    os.environ.pop("PGM_FILL_STRIP")
    st = []
    gb = graphalign.GraphBatch(js, o, 0, ctx)
    gb.run(); rb = gb.fetch(st); gb.close()
    failed = [k for k, v in enumerate(st) if v != 0]
    if failed:
        print("level with %d jobs: %d jobs failed in B: %s" % (len(js), len(failed), [(k, st[k], js[k][0].N, js[k][1].N, nonchain(js[k][0]), nonchain(js[k][1])) for k in failed[:6]]))
        sub = [js[k] for k in failed]
        st2 = []
        gb = graphalign.GraphBatch(sub, o, 0, ctx); gb.run(); gb.fetch(st2); gb.close()
        print("   the same jobs as their own batch: statuses", st2[:10])
    for k, (a, b) in enumerate(zip(ra, rb)):
        if st[k] != 0:
            continue
        same = a["score"] == b["score"] and np.array_equal(a["mapping1"], b["mapping1"]) and np.array_equal(a["mapping2"], b["mapping2"])
        if not same and bad[0] < 8:
            bad[0] += 1
            g1, g2, _ = js[k]
            print("MISMATCH: level of %d jobs, job %d: %d x %d nodes, non-chain %d + %d, scores %r %r" % (len(js), k, g1.N, g2.N, nonchain(g1), nonchain(g2), a["score"], b["score"]))
    return ra
progressive.progressive_alignment(seqs, tree, mf, opts, batch_align_fn=batch, batch_merge_fn=lambda jobs: graphalign.merge_profiles_batch(jobs))
print("done, mismatches:", bad[0])
