"""Host-side mirror of the reference's distance-factory interface for HP-2, over the C ABI.

Same names, argument meaning and error behaviour as the reference:
  DistanceFactory<A>::computePwDistances(sequences, order) -> DistanceMatrix   (DistanceFactory.h:12-51)
  DistanceFactoryAlign<A>       (DistanceFactoryAlign.h:29-127)
  DistanceFactoryPrealigned<A>  (DistanceFactoryPrealigned.h:33-90)
  DistanceFactory<A>::getDefault(model_factory, prealigned)                     (DistanceFactory.cpp:8-55)
All arithmetic happens in libpgm_b200.so on the GPU; nothing here computes a distance.
"""
import ctypes as C
import numpy as np

from . import _lib
from .alphabet import AA, DNA, Codon, GAP
from .options import Options

_T = None


def _tables():
    global _T
    if _T is None:
        import os
        _T = np.load(os.path.join(os.path.dirname(__file__), "data", "reference_tables.npz"))
    return _T


def nw_table(alphabet):
    """initMatrix() of DistanceFactoryAlign<A> (DistanceFactoryAlign.cpp:5-34, 38-235, 240-248)."""
    t = _tables()
    key = {"aa": ("nw_aa", 0), "codon": ("nw_codon", 1), "dna": ("nw_dna", 2)}[alphabet.name]
    go, ge = (int(v) for v in t["nw_gaps"][key[1]])
    return np.ascontiguousarray(t[key[0]], dtype=np.int32), go, ge


def ml_constants(n):
    """DIST_MAX, VAR_MAX, VAR_MIN (DistanceFactoryML.cpp:3-32)."""
    return (5.2, 5e3, 1e-5) if n == 61 else (2.2, 1e3, 1e-5)


class DistanceMatrix:
    """struct DistanceMatrix (DistanceFactory.h:12-42): two N x N double matrices."""

    def __init__(self, dim):
        self.distances = np.zeros((dim, dim))
        self.variances = np.zeros((dim, dim))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def make_ml_params(model_factory, opts: Options, keep: list):
    n = model_factory.alphabet.DIM
    Q = np.ascontiguousarray(model_factory.Q, dtype=np.float64)
    pi = np.ascontiguousarray(model_factory.freqs, dtype=np.float64)
    keep += [Q, pi]
    dmax, vmax, vmin = ml_constants(n)
    return _lib.MLParams(n, _p(Q), _p(pi), int(opts.mldist_flag), int(opts.mldist_gap_flag), opts.indel_rate,
                         opts.min_dist, opts.max_dist, opts.cutoff_dist, dmax, vmax, vmin)


def pack_sequences(code_seqs):
    N = len(code_seqs)
    offsets = np.zeros(N + 1, dtype=np.uint32)
    if N:
        offsets[1:] = np.cumsum([len(s) for s in code_seqs])
    codes = (np.concatenate([np.asarray(s, dtype=np.uint8) for s in code_seqs]) if N and offsets[-1] else np.zeros(1, np.uint8))
    return np.ascontiguousarray(codes), offsets


class DistanceFactory:
    """Abstract base (DistanceFactory.h:44-51)."""

    def __init__(self, model_factory=None, opts: Options = None, ctx=None):
        self.model_factory = model_factory
        self.opts = opts or (model_factory.opts if model_factory is not None else Options())
        self.ctx = ctx or _lib.default_context()

    def computePwDistances(self, sequences, order) -> DistanceMatrix:
        raise NotImplementedError

    @staticmethod
    def getDefault(model_factory=None, prealigned=False, opts: Options = None, ctx=None):
        opts = opts or model_factory.opts
        if not prealigned:
            if opts.nwdist_flag:
                return DistanceFactoryAlign(model_factory, opts, ctx)
            return DistanceFactoryAngle(model_factory, opts, ctx)
        return DistanceFactoryPrealigned(model_factory, opts, ctx)


class DistanceFactoryAlign(DistanceFactory):
    """All-pairs Gotoh alignment + counts + computeDistance on the GPU."""

    def pair_range(self, code_seqs, pair_begin, pair_end, want_counts=False):
        alpha = self.model_factory.alphabet
        n = alpha.DIM
        keep = []
        prm = make_ml_params(self.model_factory, self.opts, keep)
        tab, go, ge = nw_table(alpha)
        codes, offsets = pack_sequences([alpha.nw_codes(np.asarray(s, dtype=np.uint8)) for s in code_seqs])
        npairs = pair_end - pair_begin
        dist = np.zeros(max(npairs, 1)); var = np.zeros(max(npairs, 1))
        counts = np.zeros((max(npairs, 1), n, n), dtype=np.int32) if want_counts else None
        gaps = np.zeros(max(npairs, 1), dtype=np.uint32) if want_counts else None
        scores = np.zeros(max(npairs, 1), dtype=np.int32) if want_counts else None
        _lib.check(_lib.lib().pgm_nw_ml_pair_range(
            self.ctx.handle, C.byref(prm), _p(codes), _p(offsets), len(code_seqs), _p(tab), go, ge,
            C.c_int64(pair_begin), C.c_int64(pair_end), _p(dist), _p(var),
            _p(counts) if want_counts else None, _p(gaps) if want_counts else None, _p(scores) if want_counts else None))
        if want_counts:
            return dist[:npairs], var[:npairs], counts[:npairs], gaps[:npairs], scores[:npairs]
        return dist[:npairs], var[:npairs]

    def computePwDistances(self, sequences, order):
        alpha = self.model_factory.alphabet
        seqs = []
        for name in order:
            if name not in sequences:
                raise KeyError(name)            # reference: assert(iii != sequences.end())
            seqs.append(alpha.nw_codes(np.asarray(sequences[name], dtype=np.uint8)))
        N = len(seqs)
        keep = []
        prm = make_ml_params(self.model_factory, self.opts, keep)
        tab, go, ge = nw_table(alpha)
        codes, offsets = pack_sequences(seqs)
        dm = DistanceMatrix(N)
        _lib.check(_lib.lib().pgm_nw_ml_all_pairs(self.ctx.handle, C.byref(prm), _p(codes), _p(offsets), N, _p(tab), go, ge,
                                                  _p(dm.distances), _p(dm.variances)))
        return dm


class DistanceFactoryPrealigned(DistanceFactory):
    """All-pairs counting over MSA rows + computeDistance on the GPU."""

    def pair_range(self, rows, pair_begin, pair_end, want_counts=False):
        n = self.model_factory.alphabet.DIM
        rows = np.ascontiguousarray(rows, dtype=np.uint8)
        N, L = rows.shape
        keep = []
        prm = make_ml_params(self.model_factory, self.opts, keep)
        npairs = pair_end - pair_begin
        dist = np.zeros(max(npairs, 1)); var = np.zeros(max(npairs, 1))
        counts = np.zeros((max(npairs, 1), n, n), dtype=np.int32) if want_counts else None
        gaps = np.zeros(max(npairs, 1), dtype=np.uint32) if want_counts else None
        _lib.check(_lib.lib().pgm_prealigned_ml_pair_range(
            self.ctx.handle, C.byref(prm), _p(rows), N, L, C.c_int64(pair_begin), C.c_int64(pair_end), _p(dist), _p(var),
            _p(counts) if want_counts else None, _p(gaps) if want_counts else None))
        if want_counts:
            return dist[:npairs], var[:npairs], counts[:npairs], gaps[:npairs]
        return dist[:npairs], var[:npairs]

    def computePwDistances(self, sequences, order):
        rows = []
        for name in order:
            if name not in sequences:
                raise KeyError(name)
            rows.append(np.asarray(sequences[name], dtype=np.uint8))
        N = len(rows)
        L = len(rows[0]) if N else 0
        for r in rows:
            if len(r) != L:
                raise ValueError("aligned rows differ in length")   # reference: assert(seq1.length() == seq2.length())
        arr = np.ascontiguousarray(np.stack(rows)) if N else np.zeros((0, 0), np.uint8)
        keep = []
        prm = make_ml_params(self.model_factory, self.opts, keep)
        dm = DistanceMatrix(N)
        _lib.check(_lib.lib().pgm_prealigned_ml_all_pairs(self.ctx.handle, C.byref(prm), _p(arr) if arr.size else None, N, L,
                                                          _p(dm.distances), _p(dm.variances)))
        return dm


class DistanceFactoryAngle(DistanceFactory):
    """k-mer angle distances on the GPU (DistanceFactoryAngle<A,K>, DistanceFactoryAngle.h:24-131);
    K = 2 for AA and codon, 6 for DNA (DistanceFactory.cpp:14,32,50)."""

    def computePwDistances(self, sequences, order):
        alpha = self.model_factory.alphabet
        K = 6 if alpha is DNA else 2
        seqs = []
        for name in order:
            if name not in sequences:
                raise KeyError(name)
            seqs.append(np.minimum(np.asarray(sequences[name], dtype=np.uint8), alpha.DIM))
        codes, offsets = pack_sequences(seqs)
        N = len(seqs)
        dm = DistanceMatrix(N)
        ml = int(self.opts.mldist_flag or self.opts.mldist_gap_flag)
        _lib.check(_lib.lib().pgm_angle_all_pairs(self.ctx.handle, alpha.DIM, K, _p(codes), _p(offsets), N, ml,
                                                  _p(dm.distances), _p(dm.variances)))
        return dm


def ml_distances(model_factory, opts, counts, gaps, seqlen, ctx=None):
    """DistanceFactoryML<A>::computeDistance for caller-supplied count matrices."""
    ctx = ctx or _lib.default_context()
    keep = []
    prm = make_ml_params(model_factory, opts, keep)
    counts = np.ascontiguousarray(counts, dtype=np.int32)
    gaps = np.ascontiguousarray(gaps, dtype=np.uint32)
    seqlen = np.ascontiguousarray(seqlen, dtype=np.float64)
    P = counts.shape[0]
    dist = np.zeros(max(P, 1)); var = np.zeros(max(P, 1))
    _lib.check(_lib.lib().pgm_ml_distances(ctx.handle, C.byref(prm), C.c_int64(P), _p(counts), _p(gaps), _p(seqlen), _p(dist), _p(var)))
    return dist[:P], var[:P]


def bionj_joins(D, V, ctx=None):
    """The BioNJ join loop on the GPU (pgm_bionj_joins): same contract as tree.nj_joins_host."""
    ctx = ctx or _lib.default_context()
    D = np.ascontiguousarray(D, dtype=np.float64); V = np.ascontiguousarray(V, dtype=np.float64)
    N = D.shape[0]
    nj = max(0, N - 3)
    j1 = np.zeros(max(1, nj), np.int32); j2 = np.zeros(max(1, nj), np.int32)
    d1 = np.zeros(max(1, nj)); d2 = np.zeros(max(1, nj))
    rest = np.zeros(9); rdim = C.c_int(0)
    _lib.check(_lib.lib().pgm_bionj_joins(ctx.handle, N, D.ctypes.data_as(C.c_void_p), V.ctypes.data_as(C.c_void_p),
                                          j1.ctypes.data_as(C.c_void_p), j2.ctypes.data_as(C.c_void_p),
                                          d1.ctypes.data_as(C.c_void_p), d2.ctypes.data_as(C.c_void_p),
                                          rest.ctypes.data_as(C.c_void_p), C.byref(rdim)))
    r = rdim.value
    return [(int(j1[k]), int(j2[k]), float(d1[k]), float(d2[k])) for k in range(nj)], rest[:r * r].reshape(r, r).copy()
