#!/usr/bin/env python3
"""HP-2 on the shapes of BASELINE configs 3 (DNA 1000 x 1500 nt) and 4 (codon 300 x 300): time, properties, oracle spot-check.
Test infrastructure (it calls the oracle as the checker), not collected by pytest:  python tests/measure_other_configs.py [dna|codon|both]"""
import sys, time
import numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from prographmsa_b200 import model, synth, alphabet, distance, _lib
from prographmsa_b200.options import Options
from prographmsa_b200.distance import nw_table
from oracle import pyoracle as po

def run(tag, alpha, mf, opts, N, L, seed):
    seqs, _ = synth.simulate_family(N, L, mf, seed=seed, indel_rate=0.01)
    names = sorted(seqs)
    ctx = _lib.Context(0); ctx.enable_timing(True)
    fac = distance.DistanceFactoryAlign(mf, opts, ctx)
    fac.computePwDistances(seqs, names)                      # warm-up (allocations)
    t0 = time.perf_counter(); dm = fac.computePwDistances(seqs, names); wall = time.perf_counter() - t0
    tm = ctx.timing()
    D, V = dm.distances, dm.variances
    assert np.array_equal(D, D.T) and not D.diagonal().any() and np.isfinite(D).all() and np.isfinite(V).all()
    Ls = np.array([len(seqs[k]) for k in names], dtype=np.int64)
    cells = sum(int(((Ls[i] + 1) * (Ls[i + 1:] + 1)).sum()) for i in range(N))
    rng = np.random.default_rng(0)
    tab, go, ge = nw_table(alpha)
    worst = 0.0
    for _ in range(12):
        i, j = sorted(rng.choice(N, 2, replace=False))
        ci, cj = alpha.nw_codes(seqs[names[i]]), alpha.nw_codes(seqs[names[j]])
        d_ref, v_ref = po.nw_ml_all_pairs([ci, cj], alpha.DIM, tab, go, ge, mf.Q, mf.freqs, opts)
        worst = max(worst, abs(D[i, j] - d_ref[0, 1]) / max(abs(d_ref[0, 1]), 1e-300))
    print("%s: N=%d L=%d pairs=%d cells=%.3e  wall %.1f ms  kernels fwd %.1f walk %.1f ml %.1f ms (%d fwd launches) -> %.0f GCUPS; "
          "oracle spot-check max rel err %.2e" % (tag, N, L, N * (N - 1) // 2, cells, 1e3 * wall, tm["nw_forward_ms"], tm["nw_walk_ms"],
                                                  tm["ml_ms"], tm["nw_forward_launches"], cells / wall / 1e9, worst))
    ctx.close()

which = sys.argv[1] if len(sys.argv) > 1 else "both"
if which in ("dna", "both"):
    opts = Options(mldist_flag=True, nwdist_flag=True)
    rng = np.random.default_rng(0)
    ex = rng.random(6) + 0.5; fr = rng.random(4) + 0.5
    mf = model.getDefault(opts, alphabet.DNA, custom_text=" ".join(map(str, ex)) + " " + " ".join(map(str, fr)))
    run("config 3 shape (DNA)", alphabet.DNA, mf, opts, int(sys.argv[2]) if len(sys.argv) > 2 else 1000, 1500, 3)
if which in ("codon", "both"):
    opts = Options(mldist_flag=True, nwdist_flag=True).for_codon()
    mf = model.getDefault(opts, alphabet.Codon)
    run("config 4 shape (codon)", alphabet.Codon, mf, opts, 300, 300, 4)
