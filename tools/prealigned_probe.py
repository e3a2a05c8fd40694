#!/usr/bin/env python3
"""Kernel time of the Prealigned all-pairs path on a synthetic sparse alignment: python tools/prealigned_probe.py [N=3000] [L=11000] [res=500]"""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from prographmsa_b200 import model, distance, _lib
from prographmsa_b200.options import Options
N = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 11000
R = int(sys.argv[3]) if len(sys.argv) > 3 else 500
rng = np.random.default_rng(1)
rows = np.full((N, L), 254, dtype=np.uint8)
base = np.sort(rng.choice(L, R, replace=False))
for i in range(N):
    cols = base.copy()
    k = rng.integers(0, R, 60); cols[k] = rng.integers(0, L, 60)          # some columns of its own
    cols = np.unique(cols)
    rows[i, cols] = rng.integers(0, 20, len(cols))
opts = Options(mldist_flag=True)
mf = model.getDefault(opts)
ctx = _lib.default_context(); ctx.enable_timing(True)
f = distance.DistanceFactoryPrealigned(mf, opts, ctx=ctx)
seqs = {"s%05d" % i: rows[i] for i in range(N)}
order = sorted(seqs)
for it in range(3):
    t0 = time.perf_counter(); dm = f.computePwDistances(seqs, order); t1 = time.perf_counter()
    print("N %d L %d: call %.1f ms, kernels %s, checksum %.6f" % (N, L, (t1 - t0) * 1e3, {k: round(v, 2) for k, v in ctx.timing().items() if k.endswith("_ms") and v}, float(dm.distances.sum())), flush=True)
