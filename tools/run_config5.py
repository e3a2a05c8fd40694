#!/usr/bin/env python3
"""BASELINE.json configs[4]: a whole doAlign (k-mer angle distances -> BioNJ -> progressive pass -> Prealigned ML
distances -> BioNJ -> ... , -m -F -i 2) over n synthetic 500-aa sequences with every hot path on the GPU and the Python
host glue around it; prints where the wall time goes.

  python tools/run_config5.py [n_seq=5000] [iters=2]
"""
import sys, os, time, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from prographmsa_b200 import pipeline, model, synth, alphabet, tree as _tree, progressive as _prog
from prographmsa_b200.options import Options

n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
mode = sys.argv[3] if len(sys.argv) > 3 else ""                     # "nw": -a -m (NW + ML initial distances); "w": -m -w (Gonnet); default -m -F
nw = mode == "nw"
opts = Options(mldist_flag=True, aafreqs_flag=(mode == ""), nwdist_flag=nw, darwin_flag=(mode == "w"))
mf0 = model.getDefault(Options(mldist_flag=True))
seqs, _ = synth.simulate_family(n, 500, mf0, seed=50, forbid_first=10)
strs = {k: "".join(alphabet.AA_LETTERS[c] for c in v) for k, v in seqs.items()}

T = {}
def timed(label, fn):
    def wrap(*a, **kw):
        t0 = time.perf_counter(); r = fn(*a, **kw); T[label] = T.get(label, 0.0) + time.perf_counter() - t0; return r
    return wrap

from prographmsa_b200 import _lib
CTX = _lib.default_context()
CTX.enable_timing(True)
GPU = dict(emission_ms=0.0, graph_fill_ms=0.0, graph_trace_ms=0.0, batches=0, jobs=0, cells=0)


class Backend(pipeline.GpuBackend):
    def __init__(self, mf, o):
        super().__init__(mf, o, ctx=CTX)

    def batch_align(self, jobs, o):
        r = super().batch_align(jobs, o)
        t = CTX.timing()                                            # device time of this level's kernels (per size-class pipeline, summed)
        for k in ("emission_ms", "graph_fill_ms", "graph_trace_ms"):
            GPU[k] += t[k]
        GPU["batches"] += 1; GPU["jobs"] += len(jobs); GPU["cells"] += t["graph_cells"]
        if os.environ.get("RUN5_LEVELS"):
            print("    level batch %3d: %5d jobs, largest graphs %5d x %5d nodes, %.3g cells: emission %6.2f fill %7.2f traceback %6.2f ms" % (
                GPU["batches"], len(jobs), max(a.N for a, _, _ in jobs), max(b.N for _, b, _ in jobs), t["graph_cells"],
                t["emission_ms"], t["graph_fill_ms"], t["graph_trace_ms"]))
        return r
for name, label in (("initial_distances", "initial distances: k-mer angle or NW + ML (GPU)"), ("prealigned_distances", "Prealigned ML distances (GPU)"),
                    ("batch_align", "alignGraphs batches (GPU, incl. host packing)"), ("batch_merge", "mergeGraphs profiles (GPU)"),
                    ("nj_joins", "BioNJ join loop (GPU)")):
    setattr(Backend, name, timed(label, getattr(Backend, name)))
_tree.midpoint_root = timed("midpoint rooting (host Python)", _tree.midpoint_root)

t0 = time.perf_counter()
out, trees = pipeline.do_align(strs, opts, lambda mf, o: Backend(mf, o), iters=iters, return_trees=True)
wall = time.perf_counter() - t0
cols = len(next(iter(out.values())))
h = hashlib.sha256("".join("%s\n%s\n" % (k, out[k]) for k in sorted(out)).encode()).hexdigest()
print(("config 5 shape: %d seqs x 500 aa, " + {"nw": "-a -m", "w": "-m -w"}.get(mode, "-m -F") + " -i %d: %.1f s wall, %d trees, %d columns, sha256 %s") % (n, iters, wall, len(trees), cols, h[:16]))
acc = 0.0
for k, v in sorted(T.items(), key=lambda kv: -kv[1]):
    print("  %-48s %7.2f s" % (k, v)); acc += v
print("  %-48s %7.2f s" % ("rest: Python host glue (graphs, merge edges, trees)", wall - acc))
print("  device time inside the %d alignGraphs batches (%d jobs, %.3g cells; sums over the size-class pipelines of every level): "
      "emission %.0f ms, fill %.0f ms, traceback %.0f ms" % (GPU["batches"], GPU["jobs"], GPU["cells"], GPU["emission_ms"], GPU["graph_fill_ms"], GPU["graph_trace_ms"]))
