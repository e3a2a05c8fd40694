import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from prographmsa_b200 import model, synth, graphalign, progressive, _lib
from prographmsa_b200.options import Options
opts = Options(mldist_flag=True)
mf = model.getDefault(opts)
seqs, root = synth.simulate_family(5000, 500, mf, seed=50, forbid_first=10)
tree = progressive.parse_newick(root.newick(6) + ";")
ctx = _lib.default_context()
levels = []
def batch(js, o):
    os.environ["PGM_FILL_STRIP"] = "1"
    r = graphalign.alignGraphsBatch(js, o, ctx=ctx)
    os.environ.pop("PGM_FILL_STRIP")
    if len(levels) < 3: levels.append(js)
    return r
progressive.progressive_alignment(seqs, tree, mf, opts, batch_align_fn=batch, batch_merge_fn=lambda jobs: graphalign.merge_profiles_batch(jobs))
js = levels[1]
def run(env):
    for k, v in env.items(): os.environ[k] = v
    st = []
    gb = graphalign.GraphBatch(js, opts, 0, ctx); gb.run(); gb.fetch(st); gb.close()
    for k in env: os.environ.pop(k)
    return [k for k, v in enumerate(st) if v != 0]
for env in ({}, {}, {"PGM_GRAPH_PIPELINE": "0"}, {"PGM_GRAPH_PIPELINE": "0"}, {"PGM_SITE_CLASSES": "0"}, {"PGM_STRIP_C": "8"}, {"PGM_STRIP_HIST_KB": "100"}):
    f = run(env)
    print(env, len(js), "jobs:", len(f), "failed", f[:12])
