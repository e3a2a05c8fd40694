import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from prographmsa_b200 import model, graphalign
from prographmsa_b200.graph import Graph
from prographmsa_b200.options import Options
from oracle import pyoracle as po
d = np.load("ab/failjob.npz")
def mk(s, r, c, v):
    g = Graph.__new__(Graph); g._pgm_struct = None
    g.sites = np.ascontiguousarray(s); g.N, g.n = g.sites.shape
    g.e_rowptr, g.e_col, g.e_val = r.astype(np.int32), c.astype(np.int32), v.astype(np.float32)
    g.r_rowptr = np.zeros(g.N + 1, np.int32); g.r_col = np.zeros(0, np.int32); g.r_units = np.zeros(0, np.uint32)
    return g
g1, g2 = mk(d["s1"], d["r1"], d["c1"], d["v1"]), mk(d["s2"], d["r2"], d["c2"], d["v2"])
opts = Options(mldist_flag=True)
mf = model.getDefault(opts)
m = mf.getModel(0.3, 0.3)
ref = po.align_graphs(g1, g2, m, opts)
nc = [x for x in range(1, g2.N - 1) if not (g2.e_rowptr[x + 1] - g2.e_rowptr[x] == 1 and g2.e_col[g2.e_rowptr[x]] == x - 1)]
print("non-chain columns:", nc, [list(g2.e_col[g2.e_rowptr[x]:g2.e_rowptr[x + 1]]) for x in nc])
for env in ({}, {"PGM_STRIP_HIST_KB": "24"}, {"PGM_STRIP_HIST_KB": "24", "PGM_STRIP_C": "8"}, {"PGM_STRIP_HIST_KB": "6"}, {"PGM_FILL_STRIP": "1"}):
    for k, v in env.items(): os.environ[k] = v
    try:
        r = graphalign.alignGraphs(g1, g2, m, opts)
        ok = r["score"] == ref["score"] and np.array_equal(r["mapping1"], ref["mapping1"]) and np.array_equal(r["mapping2"], ref["mapping2"])
        print(env, "score", r["score"], ref["score"], "identical" if ok else "DIFFERENT")
    except Exception as e:
        print(env, "EXC", e)
    for k in env: os.environ.pop(k)
