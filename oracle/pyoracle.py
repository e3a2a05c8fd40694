"""ctypes binding of the CPU parity oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never by the product package.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.orc_average_alignment_length.restype = C.c_double
    return _LIB


def ref_binary():
    p = os.path.join(_HERE, "_ref", "ProGraphMSA_64")
    return p if os.path.exists(p) else None


class MLParams(C.Structure):
    _fields_ = [("n", C.c_int), ("Q", C.c_void_p), ("pi", C.c_void_p),
                ("mldist_flag", C.c_int), ("mldist_gap_flag", C.c_int),
                ("indel_rate", C.c_double), ("min_dist", C.c_double), ("max_dist", C.c_double),
                ("cutoff_dist", C.c_double), ("DIST_MAX", C.c_double), ("VAR_MAX", C.c_double), ("VAR_MIN", C.c_double)]


class CGraph(C.Structure):
    _fields_ = [("N", C.c_int), ("n", C.c_int), ("sites", C.c_void_p),
                ("e_rowptr", C.c_void_p), ("e_col", C.c_void_p), ("e_val", C.c_void_p),
                ("r_rowptr", C.c_void_p), ("r_col", C.c_void_p), ("r_units", C.c_void_p)]


class DPModel(C.Structure):
    _fields_ = [("n", C.c_int), ("M", C.c_void_p), ("pi", C.c_void_p),
                ("distance", C.c_double), ("epsilon", C.c_double), ("delta", C.c_double)]


class DPOpts(C.Structure):
    _fields_ = [("indel_rate", C.c_double), ("end_indel_prob", C.c_double),
                ("repeat_rate", C.c_double), ("repeatext_prob", C.c_double)]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def ml_consts(n):
    """DistanceFactoryML.cpp:3-32"""
    return (5.2, 5e3, 1e-5) if n == 61 else (2.2, 1e3, 1e-5)


class _Keep:
    pass


def make_ml_params(Q, pi, opts, keep):
    n = len(pi)
    keep.Q = np.ascontiguousarray(Q, dtype=np.float64)
    keep.pi = np.ascontiguousarray(pi, dtype=np.float64)
    dmax, vmax, vmin = ml_consts(n)
    return MLParams(n, _p(keep.Q), _p(keep.pi), int(opts.mldist_flag), int(opts.mldist_gap_flag),
                    opts.indel_rate, opts.min_dist, opts.max_dist, opts.cutoff_dist, dmax, vmax, vmin)


def nw_align_pair(s1, s2, n, tab, go, ge):
    s1 = np.ascontiguousarray(s1, dtype=np.uint8); s2 = np.ascontiguousarray(s2, dtype=np.uint8)
    tab = np.ascontiguousarray(tab, dtype=np.int32)
    counts = np.zeros(n * n, dtype=np.int32)
    gaps = C.c_uint32(0); score = C.c_int32(0)
    rc = lib().orc_nw_align_pair(_p(s1), len(s1), _p(s2), len(s2), n, _p(tab), go, ge, _p(counts), C.byref(gaps), C.byref(score))
    if rc:
        raise RuntimeError("error while backtracking")
    return counts.reshape(n, n), gaps.value, score.value


def prealigned_counts(r1, r2, n):
    r1 = np.ascontiguousarray(r1, dtype=np.uint8); r2 = np.ascontiguousarray(r2, dtype=np.uint8)
    counts = np.zeros(n * n, dtype=np.int32)
    gaps = C.c_uint32(0)
    lib().orc_prealigned_counts(_p(r1), _p(r2), len(r1), n, _p(counts), C.byref(gaps))
    return counts.reshape(n, n), gaps.value


def compute_distance(counts, gaps, seqlen, Q, pi, opts):
    k = _Keep()
    prm = make_ml_params(Q, pi, opts, k)
    counts = np.ascontiguousarray(counts, dtype=np.int32)
    d = C.c_double(); v = C.c_double(); ne = C.c_int()
    lib().orc_compute_distance(_p(counts), C.c_uint32(gaps), C.c_double(seqlen), C.byref(prm), C.byref(d), C.byref(v), C.byref(ne))
    return d.value, v.value, ne.value


def nw_ml_all_pairs(code_seqs, n, tab, go, ge, Q, pi, opts):
    k = _Keep()
    prm = make_ml_params(Q, pi, opts, k)
    N = len(code_seqs)
    off = np.zeros(N + 1, dtype=np.uint32)
    off[1:] = np.cumsum([len(s) for s in code_seqs])
    codes = np.concatenate([np.asarray(s, dtype=np.uint8) for s in code_seqs]) if N else np.zeros(0, np.uint8)
    tab = np.ascontiguousarray(tab, dtype=np.int32)
    dist = np.zeros((N, N)); var = np.zeros((N, N))
    rc = lib().orc_nw_ml_all_pairs(_p(codes), _p(off), N, n, _p(tab), go, ge, C.byref(prm), _p(dist), _p(var))
    if rc:
        raise RuntimeError("error while backtracking")
    return dist, var


def prealigned_ml_all_pairs(rows, n, Q, pi, opts):
    k = _Keep()
    prm = make_ml_params(Q, pi, opts, k)
    rows = np.ascontiguousarray(rows, dtype=np.uint8)
    N, L = rows.shape
    dist = np.zeros((N, N)); var = np.zeros((N, N))
    lib().orc_prealigned_ml_all_pairs(_p(rows), N, L, n, C.byref(prm), _p(dist), _p(var))
    return dist, var


def angle_all_pairs(code_seqs, n, K, ml_flag):
    N = len(code_seqs)
    off = np.zeros(N + 1, dtype=np.uint32)
    off[1:] = np.cumsum([len(s) for s in code_seqs])
    codes = np.concatenate([np.asarray(s, dtype=np.uint8) for s in code_seqs]) if off[-1] else np.zeros(1, np.uint8)
    dist = np.zeros((N, N)); var = np.zeros((N, N))
    lib().orc_angle_all_pairs(_p(codes), _p(off), N, n, K, int(ml_flag), _p(dist), _p(var))
    return dist, var


def merge_profiles(sites1, sites2, P1, P2, rm1, rm2, matched):
    sites1 = np.ascontiguousarray(sites1, dtype=np.float64); sites2 = np.ascontiguousarray(sites2, dtype=np.float64)
    P1 = np.ascontiguousarray(P1, dtype=np.float64); P2 = np.ascontiguousarray(P2, dtype=np.float64)
    rm1 = np.ascontiguousarray(rm1, dtype=np.int32); rm2 = np.ascontiguousarray(rm2, dtype=np.int32)
    matched = np.ascontiguousarray(matched, dtype=np.uint8)
    n = sites1.shape[1]
    out = np.zeros((len(rm1), n))
    lib().orc_merge_profiles(n, _p(sites1), _p(sites2), _p(P1), _p(P2), len(rm1), _p(rm1), _p(rm2), _p(matched), _p(out))
    return out


def expm(Q, t):
    Q = np.ascontiguousarray(Q, dtype=np.float64)
    P = np.zeros_like(Q)
    lib().orc_expm(_p(Q), Q.shape[0], C.c_double(t), _p(P))
    return P


def cgraph(g, keep):
    keep.append(g)
    return CGraph(g.N, g.n, _p(g.sites), _p(g.e_rowptr), _p(g.e_col), _p(g.e_val),
                  _p(g.r_rowptr), _p(g.r_col), _p(g.r_units))


def dp_model(model, keep):
    M = np.ascontiguousarray(model.M, dtype=np.float64); pi = np.ascontiguousarray(model.pi, dtype=np.float64)
    keep += [M, pi]
    return DPModel(len(pi), _p(M), _p(pi), model.distance, model.epsilon, model.delta)


def dp_opts(opts):
    return DPOpts(opts.indel_rate, opts.end_indel_prob, opts.repeat_rate, opts.repeatext_prob)


SCORE_NAMES = ("gap_init", "gap_extend", "match_init", "end_match", "end_gap", "end_skip",
               "start_gap", "start_init", "repeat_init", "repeat_ext")


def dynprog_scores(g1, g2, model, opts):
    keep = []
    a, b, m, o = cgraph(g1, keep), cgraph(g2, keep), dp_model(model, keep), dp_opts(opts)
    out = np.zeros(10, dtype=np.float32)
    lib().orc_dynprog_scores(C.byref(a), C.byref(b), C.byref(m), C.byref(o), _p(out))
    return out


def precompute_scores(g1, g2, model, match_init):
    keep = []
    a, b, m = cgraph(g1, keep), cgraph(g2, keep), dp_model(model, keep)
    S = np.zeros((g2.N, g1.N), dtype=np.float32)          # column-major N1 x N2
    lib().orc_precompute_scores(C.byref(a), C.byref(b), C.byref(m), C.c_float(match_init), _p(S))
    return S.T                                             # S[y, x]


def align_graphs(g1, g2, model, opts, dump=False):
    """-> dict(score, n_tr_indels, mapping1, mapping2, cell_updates[, M, X, Y, W, S as [y, x]])"""
    keep = []
    a, b, m, o = cgraph(g1, keep), cgraph(g2, keep), dp_model(model, keep), dp_opts(opts)
    cap = 2 * (g1.N + g2.N) + 8
    m1 = np.zeros(cap, dtype=np.int32); m2 = np.zeros(cap, dtype=np.int32)
    score = C.c_float(); ntr = C.c_uint32(); ln = C.c_int(); upd = C.c_uint64()
    dumps = [np.zeros((g2.N, g1.N), dtype=np.float32) for _ in range(5)] if dump else [None] * 5
    rc = lib().orc_align_graphs(C.byref(a), C.byref(b), C.byref(m), C.byref(o), C.byref(score), C.byref(ntr),
                                _p(m1), _p(m2), cap, C.byref(ln),
                                *[(_p(d) if d is not None else None) for d in dumps], C.byref(upd))
    if rc:
        raise RuntimeError("backtracking failed (rc=%d)" % rc)
    out = dict(score=score.value, n_tr_indels=ntr.value, mapping1=m1[:ln.value].astype(np.int64),
               mapping2=m2[:ln.value].astype(np.int64), cell_updates=upd.value)
    if dump:
        for k, d in zip("MXYWS", dumps):
            out[k] = d.T
    return out
