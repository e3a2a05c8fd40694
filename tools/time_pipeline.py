#!/usr/bin/env python3
"""End-to-end doAlign on BASELINE configs[0] (50 x 300 aa, -m -i 2): this repo's pipeline (GPU back-ends + Python host glue)
next to the reference binary on the same FASTA, same box."""
import sys, os, time, json, subprocess, tempfile, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prographmsa_b200 import pipeline
from prographmsa_b200.options import Options

runs = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "pipelines.json")))["runs"]
run = [r for r in runs if len(r["seqs"]) == 50][0]
opts = Options(mldist_flag=True, nwdist_flag=("-a" in run["flags"]))
fac = lambda mf, o: pipeline.GpuBackend(mf, o)
pipeline.do_align(run["seqs"], opts, fac)                      # warm-up (context, pools)
t0 = time.perf_counter(); got = pipeline.do_align(run["seqs"], opts, fac); t1 = time.perf_counter()
print("this repo, GPU back-ends + Python glue: %.2f s  (alignment identical to the reference: %s)" % (t1 - t0, got == run["aligned"]))
if len(sys.argv) > 1:
    pr = cProfile.Profile(); pr.enable(); pipeline.do_align(run["seqs"], opts, fac); pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
ref = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle", "_ref", "ProGraphMSA_64")
if os.path.exists(ref):
    with tempfile.NamedTemporaryFile("w", suffix=".fa", delete=False) as f:
        for k, v in run["seqs"].items():
            f.write(">%s\n%s\n" % (k, v))
    t0 = time.perf_counter()
    subprocess.run([ref] + run["flags"] + ["-o", f.name + ".out", f.name], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    print("reference binary, 1 core: %.2f s" % (time.perf_counter() - t0))
