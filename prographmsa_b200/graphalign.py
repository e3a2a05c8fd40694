"""Host-side mirror of alignGraphs<A> (GraphAlign.h:6-23,201) over the C ABI.

  alignGraphs(g1, g2, model, opts) -> AlignmentResult{score, n_tr_indels, mapping1, mapping2}
  alignGraphsBatch(jobs, opts): all ready guide-tree nodes of one level in one call.
Raises RuntimeError("backtracking failed") where the reference calls error() (GraphAlign.h:410,449,482).
"""
import ctypes as C
import numpy as np

from . import _lib
from .graph import c_graph

EMISSION_EXACT = 0
EMISSION_TENSOR = 1


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _addr(a):
    return a.__array_interface__["data"][0]


def c_model(model, keep):
    c = getattr(model, "_pgm_struct", None)
    if c is None:
        M = np.ascontiguousarray(model.M, dtype=np.float64)
        pi = np.ascontiguousarray(model.pi, dtype=np.float64)
        c = (_lib.DPModel(_addr(M), _addr(pi), model.distance, model.epsilon, model.delta), M, pi)
        try:
            model._pgm_struct = c
        except AttributeError:
            pass
    keep.append(c)
    return c[0]


def c_opts(opts, emission_mode=EMISSION_EXACT):
    return _lib.DPOpts(opts.indel_rate, opts.end_indel_prob, opts.repeat_rate, opts.repeatext_prob, int(emission_mode))


# numpy mirror of pgm_alignment (include/pgm_b200.h): lets a batch's results be written and read as columns
_ALIGNMENT = np.dtype([("score", "f4"), ("n_tr_indels", "u4"), ("mapping1", "u8"), ("mapping2", "u8"), ("capacity", "i4"),
                       ("length", "i4"), ("status", "i4"), ("cell_updates", "u8")], align=True)
assert _ALIGNMENT.itemsize == C.sizeof(_lib.Alignment)


class GraphBatch:
    """A batch of alignGraphs jobs staged on the device (pgm_graph_batch_*): create once, run many
    times without host<->device traffic, fetch results."""

    def __init__(self, jobs, opts, emission_mode=EMISSION_EXACT, ctx=None):
        self.ctx = ctx or _lib.default_context()
        self.n_jobs = len(jobs)
        keep = []
        n = jobs[0][0].n if jobs else 20
        G1 = (_lib.Graph * max(1, self.n_jobs))(); G2 = (_lib.Graph * max(1, self.n_jobs))()
        MM = (_lib.DPModel * max(1, self.n_jobs))()
        for k, (g1, g2, model) in enumerate(jobs):
            G1[k] = c_graph(g1, keep); G2[k] = c_graph(g2, keep); MM[k] = c_model(model, keep)
        o = c_opts(opts, emission_mode)
        self._h = C.c_void_p()
        _lib.check(_lib.lib().pgm_graph_batch_create(self.ctx.handle, n, self.n_jobs, G1, G2, MM, C.byref(o), C.byref(self._h)))
        self.caps = [2 * (g1.N + g2.N) + 8 for g1, g2, _ in jobs]

    def run(self):
        _lib.check(_lib.lib().pgm_graph_batch_run(self.ctx.handle, self._h))

    def fetch(self, statuses=None):
        """statuses: optional list that receives every job's status instead of raising on the first failed job (tools)"""
        n = self.n_jobs
        out = np.zeros(max(1, n), dtype=_ALIGNMENT)                          # pgm_alignment[n], filled and read as columns
        offs = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(self.caps, out=offs[1:])
        buf = np.empty((2, max(1, int(offs[-1]))), dtype=np.int32)           # every job's two mappings, one allocation
        out["mapping1"][:n] = _addr(buf[0]) + 4 * offs[:-1]
        out["mapping2"][:n] = _addr(buf[1]) + 4 * offs[:-1]
        out["capacity"][:n] = self.caps
        rc = _lib.lib().pgm_graph_batch_fetch(self.ctx.handle, self._h, C.c_void_p(out.ctypes.data))
        if statuses is not None:
            statuses.extend(out["status"][:n].tolist())
            rc = 0
        if rc == -4:
            raise RuntimeError("backtracking failed")
        _lib.check(rc)
        lens = out["length"][:n].tolist(); starts = offs[:-1].tolist()
        scores = out["score"][:n].astype(np.float64).tolist(); ntr = out["n_tr_indels"][:n].tolist()
        upd = out["cell_updates"][:n].tolist()
        m1, m2 = buf[0].astype(np.int64), buf[1].astype(np.int64)          # one widening pass; the jobs' mappings are views of it
        return [dict(score=scores[k], n_tr_indels=ntr[k], mapping1=m1[o:o + ln], mapping2=m2[o:o + ln],
                     cell_updates=upd[k]) for k, (o, ln) in enumerate(zip(starts, lens))]

    def close(self):
        if self._h:
            _lib.lib().pgm_graph_batch_destroy(self.ctx.handle, self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def alignGraphsBatch(jobs, opts, emission_mode=EMISSION_EXACT, ctx=None):
    """jobs: list of (g1, g2, model)."""
    if not jobs:
        return []
    b = GraphBatch(jobs, opts, emission_mode, ctx)
    try:
        b.run()
        return b.fetch()
    finally:
        b.close()


def alignGraphs(g1, g2, model, opts, emission_mode=EMISSION_EXACT, ctx=None):
    return alignGraphsBatch([(g1, g2, model)], opts, emission_mode, ctx)[0]


def dynprog_scores(g1, g2, model, opts):
    keep = []
    a, b, m, o = c_graph(g1, keep), c_graph(g2, keep), c_model(model, keep), c_opts(opts)
    out = np.zeros(10, dtype=np.float32)
    _lib.check(_lib.lib().pgm_dynprog_scores(g1.n, C.byref(a), C.byref(b), C.byref(m), C.byref(o), _p(out)))
    return out


def emission_scores(g1, g2, model, match_init, emission_mode=EMISSION_EXACT, ctx=None):
    ctx = ctx or _lib.default_context()
    keep = []
    a, b, m = c_graph(g1, keep), c_graph(g2, keep), c_model(model, keep)
    S = np.zeros((g2.N, g1.N), dtype=np.float32)
    _lib.check(_lib.lib().pgm_emission_scores(ctx.handle, g1.n, C.byref(a), C.byref(b), C.byref(m), C.c_float(match_init),
                                              int(emission_mode), _p(S)))
    return S.T


def merge_profiles_batch(jobs, ctx=None):
    """Profile part of mergeGraphs (GraphAlign.h:569-620) for all merges of one tree level on the GPU.
    jobs: [(g1, g2, P1, P2, rm1, rm2, matched)] with the plan of graph.merge_plan -> [len x n float64 arrays]."""
    ctx = ctx or _lib.default_context()
    if not jobs:
        return []
    n = jobs[0][0].n
    J = (_lib.MergeJob * len(jobs))()
    keep, outs = [], []
    for k, (g1, g2, P1, P2, rm1, rm2, matched) in enumerate(jobs):
        P1 = np.ascontiguousarray(P1, dtype=np.float64); P2 = np.ascontiguousarray(P2, dtype=np.float64)
        a = np.ascontiguousarray(rm1, dtype=np.int32); b = np.ascontiguousarray(rm2, dtype=np.int32)
        m = np.ascontiguousarray(matched, dtype=np.uint8)
        out = np.zeros((len(a), n))
        keep += [g1, g2, P1, P2, a, b, m]
        outs.append(out)
        J[k] = _lib.MergeJob(g1.N, g2.N, _addr(g1.sites), _addr(g2.sites), _addr(P1), _addr(P2), len(a), _addr(a), _addr(b), _addr(m), _addr(out))
    _lib.check(_lib.lib().pgm_merge_profiles_batch(ctx.handle, n, len(jobs), J))
    return outs
