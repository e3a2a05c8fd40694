"""Host-side sequence-graph container and the graph surgery around the hot path.

Mirrors (behaviour, not code) Graph.h:18-248,458-469 (container, PredIterator order, repeat
edges), SequenceGraph.h:44-104 (leaf profile), CleanedGraph.h:38-160 (pruning + uncleanMapping)
and GraphAlign.h:539-727 (mergeGraphs).  The device consumes exactly these arrays
(include/pgm_b200.h `pgm_graph`): ``sites`` DIM x N column-major double, two CSR matrices whose
rows are *to*-nodes and whose entries are ascending *from*-nodes.
"""
import ctypes as C
import math
import os
import numpy as np

from . import _lib

MAX_EDGE_COST = np.float32(10000.0)
NOGAP = -1

# The graph surgery between two alignGraphs calls exists twice: the Python below is the readable restatement of the
# reference (and what the oracle-backed CPU tests pin against the reference binary); csrc/host_graph.cpp is the same
# logic in C++ behind the C ABI (pgm_host_*), used whenever the library is built.  tests/test_cpu_host.py compares the
# two array by array; PGM_HOST_GLUE=python forces the Python path.
NATIVE = os.environ.get("PGM_HOST_GLUE", "native") != "python" and os.path.exists(_lib.LIB_PATH)


def _addr(a):
    return a.__array_interface__["data"][0]


def c_graph(g, keep=None):
    """pgm_graph view of a Graph.  The struct is cached on the graph (its arrays never change after construction)."""
    if keep is not None:
        keep.append(g)
    c = getattr(g, "_pgm_struct", None)
    if c is None:
        has_rep = g.r_col.size > 0
        c = _lib.Graph(g.N, _addr(g.sites), _addr(g.e_rowptr), _addr(g.e_col), _addr(g.e_val),
                       _addr(g.r_rowptr), _addr(g.r_col) if has_rep else None, _addr(g.r_units) if has_rep else None)
        try:
            g._pgm_struct = c
        except AttributeError:
            pass
    return c


def _f32(x):
    return np.float32(x)


class Graph:
    def __init__(self, sites, edges, repeats=None):
        """sites: (N, DIM) float64 (row = node => column-major DIM x N);
        edges: dict {(to, from): float32 cost}; repeats: dict {(to, from): units}."""
        self.sites = np.ascontiguousarray(sites, dtype=np.float64)
        self.N, self.n = self.sites.shape
        assert self.N >= 2 and not self.sites[0].any() and not self.sites[-1].any()
        self._set_edges(edges)
        self._set_repeats(repeats or {})

    # Graph.h:81-90 setEdgesFromMap: stored = min(cost, 1e4) - 1e4 in float
    def _set_edges(self, edges):
        self._pgm_struct = None                     # cached C view (graphalign.c_graph) is stale once arrays change
        items = sorted(edges.items())
        rowptr = np.zeros(self.N + 1, dtype=np.int32)
        col = np.empty(len(items), dtype=np.int32)
        val = np.empty(len(items), dtype=np.float32)
        for k, ((to, frm), cost) in enumerate(items):
            rowptr[to + 1] += 1
            col[k] = frm
            val[k] = _f32(min(_f32(cost), MAX_EDGE_COST)) - MAX_EDGE_COST
        self.e_rowptr = np.cumsum(rowptr, dtype=np.int32)
        self.e_col, self.e_val = col, val

    def _set_stored_edges(self, triplets):
        """triplets: list of (to, from, stored float32) -- keeps stored values verbatim."""
        self._pgm_struct = None
        items = sorted(triplets)
        rowptr = np.zeros(self.N + 1, dtype=np.int32)
        for to, _, _ in items:
            rowptr[to + 1] += 1
        self.e_rowptr = np.cumsum(rowptr, dtype=np.int32)
        self.e_col = np.array([f for _, f, _ in items], dtype=np.int32)
        self.e_val = np.array([v for _, _, v in items], dtype=np.float32)

    def _set_repeats(self, repeats):
        self._pgm_struct = None
        items = sorted(repeats.items())
        rowptr = np.zeros(self.N + 1, dtype=np.int32)
        for (to, _), _u in items:
            rowptr[to + 1] += 1
        self.r_rowptr = np.cumsum(rowptr, dtype=np.int32)
        self.r_col = np.array([f for (_, f), _ in items], dtype=np.int32)
        self.r_units = np.array([u for _, u in items], dtype=np.uint32)

    def size(self):
        return self.N

    # PredIterator with (repeat_init, repeat_ext) = (0, 0): regular first, then repeats (Graph.h:172-248)
    def preds(self, node):
        """yield (from, value float32, is_repeat, units)"""
        for k in range(self.e_rowptr[node], self.e_rowptr[node + 1]):
            c = self.e_val[k]
            yield int(self.e_col[k]), (np.float32(np.inf) if c == 0 else _f32(c + MAX_EDGE_COST)), False, 0
        for k in range(self.r_rowptr[node], self.r_rowptr[node + 1]):
            u = int(self.r_units[k])
            yield int(self.r_col[k]), (np.float32(np.inf) if u == 0 else np.float32(0.0)), True, u

    # Graph.h:48-79 getRepeatEdges + :458-469 addRepeats
    def addRepeats(self, tr_homologies):
        rep = {}
        for tr in tr_homologies:
            L = len(tr)
            for frm in range(L):
                if tr[frm] < 0:
                    continue
                n_units = 0
                take_next = False
                for to in range(frm + 1, L):
                    if tr[to] < 0:
                        continue
                    if tr[to] <= tr[to - 1]:
                        n_units += 1
                    if take_next:
                        key = (to, frm)
                        rep[key] = min(rep[key], n_units) if key in rep else n_units
                        take_next = False
                    if tr[to] == tr[frm]:
                        take_next = True
        self._set_repeats(rep)

    def total_pred_pairs(self, other):
        """sum over interior cells of #preds(y) * #preds(x) (the GCUPS numerator)."""
        d1 = (np.diff(self.e_rowptr) + np.diff(self.r_rowptr))[1:self.N - 1].astype(np.int64)
        d2 = (np.diff(other.e_rowptr) + np.diff(other.r_rowptr))[1:other.N - 1].astype(np.int64)
        return int(d1.sum() * d2.sum())


def sequence_graph(codes, n) -> Graph:
    """SequenceGraph(seq) (SequenceGraph.h:44-104 + Graph.h:35-46): one-hot (or uniform 1/n for a
    non-valid state) profile, chain edges of cost 0."""
    codes = np.asarray(codes)
    L = len(codes)
    sites = np.zeros((L + 2, n), dtype=np.float64)
    valid = codes < n
    pos = np.nonzero(valid)[0]
    sites[pos + 1, codes[pos]] = 1.0
    sites[np.nonzero(~valid)[0] + 1, :] = 1.0 / n
    g = Graph.__new__(Graph)
    g._pgm_struct = None
    g.sites, g.N, g.n = sites, L + 2, n
    # chain edges (i+1 <- i) of cost 0: stored value min(0, 1e4) - 1e4 (Graph.h:85)
    g.e_rowptr = np.concatenate([[0], np.arange(L + 2)]).astype(np.int32)
    g.e_col = np.arange(L + 1, dtype=np.int32)
    g.e_val = np.full(L + 1, -MAX_EDGE_COST, dtype=np.float32)
    g.r_rowptr = np.zeros(L + 3, dtype=np.int32)
    g.r_col = np.zeros(0, dtype=np.int32); g.r_units = np.zeros(0, dtype=np.uint32)
    return g


class CleanedGraph(Graph):
    """CleanedGraph.h:38-146: keep nodes on a finite-cost start->end path; drop stored-0 edges and
    0-unit repeats; remember the index mapping for uncleanMapping (:153-160)."""

    def __init__(self, original: Graph, repeatext_prob: float):
        if NATIVE:
            self._init_native(original, repeatext_prob)
            return
        N = original.N
        repeat_ext = np.float32(np.inf) if repeatext_prob == 0 else np.float32(0.0)

        def pred_vals(to):
            for k in range(original.e_rowptr[to], original.e_rowptr[to + 1]):
                c = original.e_val[k]
                yield int(original.e_col[k]), (np.float32(np.inf) if c == 0 else _f32(c + MAX_EDGE_COST))
            for k in range(original.r_rowptr[to], original.r_rowptr[to + 1]):
                u = int(original.r_units[k])
                if u == 0:
                    v = np.float32(np.inf)
                else:
                    with np.errstate(invalid="ignore"):
                        v = np.float32(0.0) + repeat_ext * np.float32(u - 1)
                yield int(original.r_col[k]), v

        fw = np.zeros(N, dtype=bool); bw = np.zeros(N, dtype=bool)
        mapping = np.full(N, -1, dtype=np.int64)
        fw[0] = True; bw[N - 1] = True; mapping[0] = 0
        new_dim = 1
        for to in range(N - 1, -1, -1):
            if not bw[to]:
                continue
            for frm, c in pred_vals(to):
                if c != np.inf:           # NaN != inf is True, as in the reference
                    bw[frm] = True
        for to in range(1, N):
            for frm, c in pred_vals(to):
                if c != np.inf and fw[frm]:
                    fw[to] = True
                    if bw[to]:
                        mapping[to] = new_dim
                        new_dim += 1
                    break
        assert bw[0] and fw[N - 1]
        trip = []
        for to in range(N):
            for k in range(original.e_rowptr[to], original.e_rowptr[to + 1]):
                y, x = mapping[to], mapping[original.e_col[k]]
                if x != -1 and y != -1 and original.e_val[k] < 0:
                    trip.append((int(y), int(x), original.e_val[k]))
        reps = {}
        for to in range(N):
            for k in range(original.r_rowptr[to], original.r_rowptr[to + 1]):
                y, x = mapping[to], mapping[original.r_col[k]]
                if x != -1 and y != -1 and original.r_units[k] > 0:
                    reps[(int(y), int(x))] = int(original.r_units[k])
        keep = np.nonzero(mapping != -1)[0]
        self.sites = np.ascontiguousarray(original.sites[keep])
        self.N, self.n = self.sites.shape
        assert self.N == new_dim
        self._set_stored_edges(trip)
        self._set_repeats(reps)
        self.outmapping = keep.astype(np.int64)

    def _init_native(self, original, repeatext_prob):
        N, E, R = original.N, original.e_col.size, original.r_col.size
        keep = np.empty(N, dtype=np.int32)
        e_rowptr = np.empty(N + 1, dtype=np.int32); e_col = np.empty(E, dtype=np.int32); e_val = np.empty(E, dtype=np.float32)
        r_rowptr = np.empty(N + 1, dtype=np.int32); r_col = np.empty(R, dtype=np.int32); r_units = np.empty(R, dtype=np.uint32)
        new_N = C.c_int(0)
        g = c_graph(original)
        _lib.check(_lib.lib().pgm_host_clean_graph(C.byref(g), C.c_double(repeatext_prob), _addr(keep), C.byref(new_N),
                                                   _addr(e_rowptr), _addr(e_col), _addr(e_val),
                                                   _addr(r_rowptr), _addr(r_col), _addr(r_units)))
        n = new_N.value
        keep = keep[:n]
        self._pgm_struct = None
        self.sites = original.sites if n == N else np.ascontiguousarray(original.sites[keep])
        self.N, self.n = self.sites.shape
        self.e_rowptr = e_rowptr[:n + 1]; ne = int(e_rowptr[n])
        self.e_col, self.e_val = e_col[:ne], e_val[:ne]
        self.r_rowptr = r_rowptr[:n + 1]; nr = int(r_rowptr[n])
        self.r_col, self.r_units = r_col[:nr], r_units[:nr]
        self.outmapping = keep.astype(np.int64)

    def uncleanMapping(self, mapping):
        m = np.asarray(mapping, dtype=np.int64).copy()
        ok = m != NOGAP
        m[ok] = self.outmapping[m[ok]]
        return m


def _update_edge(d, frm, to, cost):
    key = (to, frm)
    if key in d:
        d[key] = min(d[key], cost)
    else:
        d[key] = cost


def _normalized(p):
    nrm = math.sqrt(float(np.dot(p, p)))
    return p if nrm == 0 else p / nrm


def merge_plan(mapping1, mapping2):
    """Which child nodes every node of the merged graph comes from (GraphAlign.h:569-620, the loop structure):
    -> (rm1, rm2, matched) lists; matched[i] is False for nodes the alignment path skipped."""
    if NATIVE:
        m1 = np.ascontiguousarray(mapping1, dtype=np.int64); m2 = np.ascontiguousarray(mapping2, dtype=np.int64)
        cap = 2 * len(m1) + 2 + int(max(m1.max(initial=0), 0)) + int(max(m2.max(initial=0), 0))
        rm1 = np.empty(cap, dtype=np.int32); rm2 = np.empty(cap, dtype=np.int32); matched = np.empty(cap, dtype=np.uint8)
        n = C.c_int(0)
        _lib.check(_lib.lib().pgm_host_merge_plan(_addr(m1), _addr(m2), len(m1), _addr(rm1), _addr(rm2), _addr(matched), C.byref(n)))
        return rm1[:n.value], rm2[:n.value], matched[:n.value]
    rm1, rm2, matched = [], [], []
    i1 = i2 = 0
    for k1, k2 in zip(mapping1, mapping2):
        k1, k2 = int(k1), int(k2)
        if k1 != NOGAP:
            while i1 != k1:
                rm1.append(i1); rm2.append(NOGAP); matched.append(False)
                i1 += 1
            i1 += 1
        if k2 != NOGAP:
            while i2 != k2:
                rm1.append(NOGAP); rm2.append(i2); matched.append(False)
                i2 += 1
            i2 += 1
        if k1 == NOGAP and k2 == NOGAP:
            raise AssertionError("error in mapping")
        rm1.append(k1); rm2.append(k2); matched.append(True)
    return rm1, rm2, matched


def merge_profiles_host(g1, g2, P1, P2, rm1, rm2, matched):
    """Profile of every merged node on the host (numpy).  Reference quirk kept: skipped g2 nodes use model1.P (:591)."""
    nodes = []
    for k1, k2, m in zip(rm1, rm2, matched):
        if not m:
            p = P1 @ (g1.sites[k1] if k1 != NOGAP else g2.sites[k2])
        elif k1 != NOGAP and k2 != NOGAP:
            p = (P1 @ g1.sites[k1]) * (P2 @ g2.sites[k2])
        elif k1 != NOGAP:
            p = P1 @ g1.sites[k1]
        else:
            p = P2 @ g2.sites[k2]
        nodes.append(_normalized(p))
    return nodes


def merge_graphs(g1: Graph, g2: Graph, mapping1, mapping2, model1, model2, support1, support2, altsplice_prob, nodes=None):
    """mergeGraphs (GraphAlign.h:551-727).  Returns (graph, mapping1, mapping2, is_matched).
    Reference quirks kept: unmatched g2 nodes use model1.P (:591); is_matched[] is indexed with
    child-graph indices (:692-695,711-714).  nodes: profiles computed elsewhere (pgm_merge_profiles_batch) for the
    plan merge_plan(mapping1, mapping2), else they are computed here with numpy."""
    rm1, rm2, matched = merge_plan(mapping1, mapping2)
    if nodes is None:
        nodes = merge_profiles_host(g1, g2, model1.P, model2.P, rm1, rm2, matched)
    assert len(nodes) == len(rm1)
    if NATIVE:
        return _merge_graphs_native(g1, g2, rm1, rm2, matched, nodes, support1, support2, altsplice_prob)

    edges, repeats = {}, {}
    last_xy = last_x = last_y = last_mapped = 0
    zero = np.float32(0.0)
    for i in range(1, len(nodes)):
        if not matched[i]:
            continue
        _update_edge(edges, last_mapped, i, zero)
        last_mapped = i
        if rm1[i] != NOGAP and rm2[i] != NOGAP:
            if last_xy != i - 1:
                _update_edge(edges, last_xy, i, zero)
            last_xy = i
        if rm1[i] != NOGAP:
            if last_y != i - 1:
                _update_edge(edges, last_y, i, zero)
            last_y = i
        if rm2[i] != NOGAP:
            if last_x != i - 1:
                _update_edge(edges, last_x, i, zero)
            last_x = i

    inv1 = np.zeros(g1.N, dtype=np.int64); inv2 = np.zeros(g2.N, dtype=np.int64)
    for i, m in enumerate(rm1):
        if m != NOGAP:
            inv1[m] = i
    for i, m in enumerate(rm2):
        if m != NOGAP:
            inv2[m] = i

    def LOG2(x):
        return math.log(x) / math.log(2)

    pen1 = np.float32(-LOG2(altsplice_prob + (1.0 - altsplice_prob) * (1.0 - support1)))
    pen2 = np.float32(-LOG2(altsplice_prob + (1.0 - altsplice_prob) * (1.0 - support2)))
    for g, inv, pen in ((g1, inv1, pen1), (g2, inv2, pen2)):
        half = np.float32(pen / np.float32(2))
        for to in range(g.N):
            for frm, val, is_rep, units in g.preds(to):
                y, x = int(inv[frm]), int(inv[to])
                if not is_rep:
                    mf, mt = matched[frm], matched[to]       # sic: child indices into is_matched
                    if mf and mt:
                        _update_edge(edges, y, x, np.float32(val + pen))
                    elif mf or mt:
                        _update_edge(edges, y, x, np.float32(val + half))
                    else:
                        _update_edge(edges, y, x, val)
                else:
                    _update_edge(repeats, y, x, units)
    graph = Graph(np.array(nodes), edges, repeats)
    return graph, np.array(rm1, dtype=np.int64), np.array(rm2, dtype=np.int64), matched


def _merge_graphs_native(g1, g2, rm1, rm2, matched, nodes, support1, support2, altsplice_prob):
    """Edge part of mergeGraphs through pgm_host_merge_edges; the merged Graph is assembled from the returned CSR."""
    rm1 = np.ascontiguousarray(rm1, dtype=np.int32); rm2 = np.ascontiguousarray(rm2, dtype=np.int32)
    mt = np.ascontiguousarray(matched, dtype=np.uint8)
    ln = len(rm1)
    ecap = g1.e_col.size + g2.e_col.size + 4 * ln + 4
    rcap = g1.r_col.size + g2.r_col.size + 1
    e_rowptr = np.empty(ln + 1, dtype=np.int32); e_col = np.empty(ecap, dtype=np.int32); e_val = np.empty(ecap, dtype=np.float32)
    r_rowptr = np.empty(ln + 1, dtype=np.int32); r_col = np.empty(rcap, dtype=np.int32); r_units = np.empty(rcap, dtype=np.uint32)
    ne, nr = C.c_int(0), C.c_int(0)
    a, b = c_graph(g1), c_graph(g2)
    _lib.check(_lib.lib().pgm_host_merge_edges(C.byref(a), C.byref(b), ln, _addr(rm1), _addr(rm2), _addr(mt),
                                               C.c_double(support1), C.c_double(support2), C.c_double(altsplice_prob),
                                               _addr(e_rowptr), _addr(e_col), _addr(e_val), C.byref(ne),
                                               _addr(r_rowptr), _addr(r_col), _addr(r_units), C.byref(nr)))
    g = Graph.__new__(Graph)
    g._pgm_struct = None
    g.sites = np.ascontiguousarray(nodes, dtype=np.float64)
    g.N, g.n = g.sites.shape
    assert g.N >= 2 and not g.sites[0].any() and not g.sites[-1].any()
    g.e_rowptr, g.e_col, g.e_val = e_rowptr, e_col[:ne.value], e_val[:ne.value]
    g.r_rowptr, g.r_col, g.r_units = r_rowptr, r_col[:nr.value], r_units[:nr.value]
    return g, rm1.astype(np.int64), rm2.astype(np.int64), mt.astype(bool)
