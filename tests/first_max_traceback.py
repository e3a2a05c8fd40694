#!/usr/bin/env python3
"""Groundwork for storing the graph-DP traceback as direction words instead of the {M,X,Y,W} floats:
the reference's traceback picks, at every step, the candidate closest to the current cell's value (GraphAlign.h:283-521).
Because the fill computed that value as the maximum of the very same float expressions, "closest" should always be "the
first candidate that attains the maximum", which the fill could record in a few bits.  This script re-traces the oracle's
dumped matrices with the first-maximum rule (CPU only, pure Python) and compares with the oracle's own traceback.

  python tests/first_max_traceback.py [n_seq=40] [length=120] [seed=33]      (test infrastructure: it drives the oracle)
"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from prographmsa_b200 import model, synth, progressive
from prographmsa_b200.options import Options
from oracle import pyoracle as po

f32 = np.float32
INF = f32(np.inf)


def preds(g, sc):
    """per node: list of (from, value, is_repeat) in PredIterator order (regular edges ascending, then repeat edges)"""
    rinit, rext = f32(sc[8]), f32(sc[9])
    out = []
    for node in range(g.N):
        lst = []
        for e in range(g.e_rowptr[node], g.e_rowptr[node + 1]):
            c = f32(g.e_val[e])
            lst.append((int(g.e_col[e]), INF if c == 0 else f32(c + f32(10000.0)), False))
        for e in range(g.r_rowptr[node], g.r_rowptr[node + 1]):
            u = int(g.r_units[e])
            lst.append((int(g.r_col[e]), INF if u == 0 else f32(rinit + f32(rext * f32(u - 1))), True))
        out.append(lst)
    return out


def first_max_trace(g1, g2, D, sc):
    """-> list of (m1, m2) from the end to the start, WITHOUT the alternative-path entries of repeat edges (returns None
    as soon as a repeat edge is taken: those cases are left to the |delta| search here)"""
    M, X, Y, W, S = (D[k] for k in "MXYWS")
    gi, ge, end_match, end_gap, end_skip = (f32(sc[k]) for k in (0, 1, 3, 4, 5))
    P1, P2 = preds(g1, sc), preds(g2, sc)
    N1, N2 = g1.N, g2.N
    path = [(N1 - 1, N2 - 1)]
    best, pick = None, None
    for yp, ey, ry in P1[N1 - 1]:
        for xp, ex, rx in P2[N2 - 1]:
            cands = [(f32(f32(f32(M[yp, xp] + end_match) - ey) - ex), 0), (f32(f32(f32(Y[yp, xp] + end_gap) - ey) - ex), 2),
                     (f32(f32(f32(X[yp, xp] + end_gap) - ey) - ex), 1)]
            if xp == 0 and yp == 0:
                cands.append((f32(f32(end_skip - ey) - ex), 3))
            for v, st in cands:
                if best is None or v > best:
                    best, pick = v, (st, yp, xp, ry or rx)
    st, y, x, rep = pick
    if rep:
        return None
    if st == 3:
        st = 0
    def emit(st, y, x):
        if x != 0 or y != 0:
            path.append((y, x) if st == 0 else ((-1, x) if st == 1 else (y, -1)))
    emit(st, y, x)
    def resolve(y, x):
        if W[y, x] == M[y, x]: return 0
        if W[y, x] == Y[y, x]: return 2
        if W[y, x] == X[y, x]: return 1
        raise RuntimeError("W matches none")
    guard = 0
    while x != 0 or y != 0:
        guard += 1
        assert guard < 4 * (N1 + N2) + 16
        best, pick = None, None
        if st == 2:
            for yp, ey, ry in P1[y]:
                for v, land in ((f32(f32(Y[yp, x] + ge) - ey), False), (f32(f32(W[yp, x] + gi) - ey), True)):
                    if best is None or v > best:
                        best, pick = v, (yp, x, land, ry)
        elif st == 1:
            for xp, ex, rx in P2[x]:
                for v, land in ((f32(f32(X[y, xp] + ge) - ex), False), (f32(f32(W[y, xp] + gi) - ex), True)):
                    if best is None or v > best:
                        best, pick = v, (y, xp, land, rx)
        else:
            for yp, ey, ry in P1[y]:
                for xp, ex, rx in P2[x]:
                    v = f32(f32(f32(W[yp, xp] + S[y, x]) - ey) - ex)
                    if best is None or v > best:
                        best, pick = v, (yp, xp, True, ry or rx)
        ny, nx, land, rep = pick
        if rep:
            return None
        if land and (nx != 0 or ny != 0):
            st = resolve(ny, nx)
        y, x = ny, nx
        emit(st, y, x)
    path.append((0, 0))
    return path


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 120
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 33
    opts = Options(mldist_flag=True)
    mf = model.getDefault(opts)
    seqs, root = synth.simulate_family(n, L, mf, seed=seed, indel_rate=0.05)
    tree = progressive.from_synth(root)
    stats = dict(jobs=0, same=0, differ=0, skipped=0, steps=0)

    def batch(jobs, o):
        out = []
        for g1, g2, m in jobs:
            r = po.align_graphs(g1, g2, m, o, dump=True)
            sc = po.dynprog_scores(g1, g2, m, o)
            with np.errstate(invalid="ignore", over="ignore"):
                p = first_max_trace(g1, g2, r, sc)
            stats["jobs"] += 1
            if p is None:
                stats["skipped"] += 1
            else:
                ref = list(zip(r["mapping1"].tolist()[::-1], r["mapping2"].tolist()[::-1]))
                stats["steps"] += len(ref)
                if p == ref:
                    stats["same"] += 1
                else:
                    stats["differ"] += 1
                    k = next(i for i, (a, b) in enumerate(zip(p, ref)) if a != b) if len(p) == len(ref) or True else -1
                    print("job %d (%d x %d nodes): first difference at step %d of %d: first-max %s, oracle %s" % (
                        stats["jobs"], g1.N, g2.N, k, len(ref), p[k] if k < len(p) else None, ref[k] if k < len(ref) else None))
            out.append(r)
        return out
    progressive.progressive_alignment(seqs, tree, mf, opts, batch_align_fn=batch)
    print("%d alignGraphs jobs of a %d x %d family: first-maximum traceback identical in %d, different in %d, %d skipped (repeat edge on the path); "
          "%d path steps compared" % (stats["jobs"], n, L, stats["same"], stats["differ"], stats["skipped"], stats["steps"]))


if __name__ == "__main__":
    main()
