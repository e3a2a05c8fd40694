#!/usr/bin/env python3
"""The single-job levels at the top of the first progressive pass of a config-5 run (k-mer angle distances -> BioNJ -> midpoint
root: a short graph against a profile of thousands of columns), each timed under a list of environment settings.

  python tools/top_jobs_knobs.py n_seq max_jobs "PGM_FILL_BAND_MIN_N1=400" ..."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from prographmsa_b200 import pipeline, model, synth, alphabet, graphalign, _lib
from prographmsa_b200.options import Options
n = int(sys.argv[1]); max_jobs = int(sys.argv[2])
settings = [""] + sys.argv[3:]
opts = Options(mldist_flag=True, aafreqs_flag=True)
mf0 = model.getDefault(Options(mldist_flag=True))
seqs, _ = synth.simulate_family(n, 500, mf0, seed=50, forbid_first=10)
strs = {k: "".join(alphabet.AA_LETTERS[c] for c in v) for k, v in seqs.items()}
CTX = _lib.default_context()
CTX.enable_timing(True)
tot = {s: 0.0 for s in settings}
class Backend(pipeline.GpuBackend):
    def __init__(self, mf, o):
        super().__init__(mf, o, ctx=CTX)
    def batch_align(self, jobs, o):
        if len(jobs) <= max_jobs:
            res = []; ref = None
            for s in settings:
                kv = [q.split("=", 1) for q in s.split(",") if q]
                for k, v in kv: os.environ[k] = v
                os.environ["PGM_GRAPH_PIPELINE"] = "0"
                gb = graphalign.GraphBatch(jobs, o, 0, CTX)
                acc = 0.0
                for it in range(3):
                    gb.run(); CTX.sync(); t = CTX.timing()
                    if it: acc += t["graph_fill_ms"] / 2
                r = gb.fetch()
                chk = (sum(x["score"] for x in r), sum(int(x["mapping1"].sum()) for x in r))
                if ref is None: ref = chk
                gb.close()
                for k, v in kv: os.environ.pop(k, None)
                os.environ.pop("PGM_GRAPH_PIPELINE")
                res.append("%8.3f%s" % (acc, "" if chk == ref else "!"))
                tot[s] += acc
            print("%3d jobs, largest %5d x %5d: fill ms %s" % (len(jobs), max(a.N for a, _, _ in jobs), max(b.N for _, b, _ in jobs), " ".join(res)), flush=True)
        return super().batch_align(jobs, o)
pipeline.do_align(strs, opts, lambda mf, o: Backend(mf, o), iters=0)
for s in settings: print("total %-50s %9.2f ms" % (s or "(default)", tot[s]))
