import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from prographmsa_b200 import model, synth, graphalign, progressive, _lib
from prographmsa_b200.options import Options
nj = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
opts = Options(mldist_flag=True)
mf = model.getDefault(opts)
seqs, root = synth.simulate_family(300, 500, mf, seed=50, forbid_first=10)
tree = progressive.parse_newick(root.newick(6) + ";")
ctx = _lib.default_context()
levels = []
def batch(js, o):
    os.environ["PGM_FILL_STRIP"] = "1"
    r = graphalign.alignGraphsBatch(js, o, ctx=ctx)
    os.environ.pop("PGM_FILL_STRIP")
    levels.append((js, r))
    return r
progressive.progressive_alignment(seqs, tree, mf, opts, batch_align_fn=batch, batch_merge_fn=lambda jobs: graphalign.merge_profiles_batch(jobs))
js, ref = [], []
for l, r in levels[1:4]:
    js += l; ref += r
mult = int(sys.argv[3]) if len(sys.argv) > 3 else 1
js, ref = js[:nj] * mult, ref[:nj] * mult
for it in range(reps):
    st = []
    gb = graphalign.GraphBatch(js, opts, 0, ctx); gb.run(); res = gb.fetch(st); gb.close()
    bad = [k for k in range(len(js)) if st[k] != 0 or res[k]["score"] != ref[k]["score"] or not np.array_equal(res[k]["mapping1"], ref[k]["mapping1"])]
    print(len(js), "jobs, bad:", bad[:10])
