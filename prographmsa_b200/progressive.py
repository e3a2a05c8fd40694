"""Host-side caller of HP-1: the progressive driver (post-order over the guide tree).

Mirrors ProgressiveAlignment.h:59-113 (recursion), :245-287 (extend_alignment /
extend_tr_homologies) and :413-476 (align_progressive_results): three getModel calls, two
CleanedGraphs, alignGraphs (the hot path -- injected as ``align_fn`` so the same driver runs the
GPU path in the product and the oracle in tests), uncleanMapping, mergeGraphs, row extension.
Not mirrored: CS profiles (-c; library file absent), early refinement, ancestral output.

``batch_align_fn`` (optional) receives all jobs of one tree level at once -- independent subtrees
never interact before their parent, which is the batching the GPU path exploits
(SURVEY.md §3C).
"""
import re
import numpy as np

from .alphabet import GAP
from .graph import Graph, CleanedGraph, sequence_graph, merge_graphs, merge_plan, NOGAP


class TreeNode:
    __slots__ = ("children", "length", "support", "name")

    def __init__(self, name=None, length=0.0, support=1.0):
        self.children, self.length, self.support, self.name = [], length, support, name

    def is_leaf(self):
        return not self.children


def parse_newick(text):
    """Minimal Newick reader (newick.cpp:40-147 semantics: branch support is ignored -> 1)."""
    s = text.strip().rstrip(";")
    pos = 0

    def node():
        nonlocal pos
        nd = TreeNode()
        if s[pos] == "(":
            pos += 1
            while True:
                nd.children.append(node())
                if s[pos] == ",":
                    pos += 1
                    continue
                assert s[pos] == ")", "newick: expected ')' at %d" % pos
                pos += 1
                break
        m = re.match(r"[^:,()\[\];]*", s[pos:])
        label = m.group(0).strip()
        pos += m.end()
        if nd.is_leaf():
            nd.name = label
        if pos < len(s) and s[pos] == ":":
            m = re.match(r":\s*([-+0-9.eE]+)", s[pos:])
            nd.length = float(m.group(1))
            pos += m.end()
        return nd

    return node()


def from_synth(node):
    nd = TreeNode(name=node.name, length=float("%.6g" % node.length))
    nd.children = [from_synth(c) for c in node.children]
    return nd


class _RowTree:
    """The alignment rows of a subtree, unmaterialised.  The reference's extend_alignment (ProgressiveAlignment.h:245-264)
    re-scatters EVERY row of both children at EVERY internal node -- O(rows x columns) per node, a third of a 5000-sequence
    pass in numpy.  A node here only records where each child's columns land in its own columns; the rows are produced once,
    at whichever node they are asked for, by composing these column maps top-down (one gather per node) and one scatter
    per leaf.  Same rows, bit for bit: a column map is exactly the `mapping != NOGAP` selection extend_alignment uses."""
    __slots__ = ("name", "codes", "parts", "ncols")

    def __init__(self, name=None, codes=None, parts=None, ncols=0):
        self.name, self.codes, self.parts, self.ncols = name, codes, parts, ncols

    def rows(self):
        out = {}
        stack = [(self, None)]                       # (subtree, root column of each of its columns; None = identity)
        while stack:
            t, pos = stack.pop()
            if t.parts is None:
                if pos is None:
                    out[t.name] = t.codes
                else:
                    row = np.full(self.ncols, GAP, dtype=np.uint8)
                    row[pos] = t.codes
                    out[t.name] = row
                continue
            for child, cols in t.parts:
                stack.append((child, cols if pos is None else pos[cols]))
        return out


class Result:
    __slots__ = ("_rows", "_tree", "tr_homologies", "tr_source", "graph", "score", "n_tr_indels")

    def __init__(self):
        self._rows, self._tree = {}, None
        self.tr_homologies, self.tr_source = [], []
        self.graph, self.score, self.n_tr_indels = None, 0.0, 0

    @property
    def aligned(self):
        """name -> aligned row (uint8 codes, GAP where the sequence has no residue); materialised on first use"""
        if self._rows is None:
            self._rows = self._tree.rows()
        return self._rows

    def row_tree(self):
        if self._tree is None:                        # a leaf (or a result whose rows were given explicitly)
            (name, codes), = self._rows.items()
            self._tree = _RowTree(name=name, codes=codes, ncols=len(codes))
        return self._tree


def _leaf(name, sequences, repeats, n):
    r = Result()
    if name not in sequences:
        raise KeyError("unknown sequence name: %s" % name)
    codes = sequences[name]
    r.graph = sequence_graph(codes, n)
    r.aligned[name] = np.asarray(codes, dtype=np.uint8)
    for rep in (repeats or {}).get(name, []):      # repeat_t {start, tr_hom} (Repeat.h)
        tr = [-1] * r.graph.N
        th = list(rep["tr_hom"])
        tr[rep["start"] + 1:rep["start"] + 1 + len(th)] = th
        r.tr_homologies.append(tr)
        r.tr_source.append(name)
    if r.tr_homologies:
        r.graph.addRepeats(r.tr_homologies)
    return r


def _child_columns(size, mapping, child_cols):
    """extend_alignment (ProgressiveAlignment.h:245-264): column j of the child alignment lands in the column of the merged
    graph its node was mapped to -> the merged column of every child column (see _RowTree)."""
    sel = np.asarray(mapping[1:size - 1])
    ok = sel != NOGAP
    assert np.array_equal(sel[ok] - 1, np.arange(ok.sum())) and int(ok.sum()) == child_cols
    return np.flatnonzero(ok)


def _extend_tr(size, mapping, trs):
    """extend_tr_homologies (ProgressiveAlignment.h:267-287): vectors of length N_g-2 (sic)."""
    out = []
    for orig in trs:
        ext = [-1] * (size - 2)
        k = 0
        for j in range(1, size - 1):
            if mapping[j] != NOGAP:
                ext[j - 1] = orig[k]
                k += 1
        out.append(ext)
    return out


def prepare_job(r1, r2, d1, d2, model_factory, opts):
    """Everything align_progressive_results does before alignGraphs (ProgressiveAlignment.h:418-437)."""
    gap_distance = d1 + d2
    distance = d1 + d2
    model = model_factory.getModel(distance, gap_distance)
    model1 = model_factory.getModel(d1, d1)
    model2 = model_factory.getModel(d2, d2)
    cg1 = CleanedGraph(r1.graph, opts.repeatext_prob)
    cg2 = CleanedGraph(r2.graph, opts.repeatext_prob)
    return dict(model=model, model1=model1, model2=model2, cg1=cg1, cg2=cg2)


def merge_job(r1, r2, job, ali):
    """The mergeGraphs profile job of an aligned node: (g1, g2, P1, P2, rm1, rm2, matched) for merge_profiles_batch."""
    m1 = job["cg1"].uncleanMapping(ali["mapping1"])
    m2 = job["cg2"].uncleanMapping(ali["mapping2"])
    rm1, rm2, matched = merge_plan(m1, m2)
    return (r1.graph, r2.graph, job["model1"].P, job["model2"].P, rm1, rm2, matched)


def finish_job(r1, r2, job, ali, s1, s2, opts, nodes=None):
    """...and everything after it (ProgressiveAlignment.h:440-468).  nodes: merged-node profiles computed elsewhere."""
    res = Result()
    res.score = ali["score"]
    res.n_tr_indels = ali["n_tr_indels"] + r1.n_tr_indels + r2.n_tr_indels
    m1 = job["cg1"].uncleanMapping(ali["mapping1"])
    m2 = job["cg2"].uncleanMapping(ali["mapping2"])
    graph, am1, am2, _ = merge_graphs(r1.graph, r2.graph, m1, m2, job["model1"], job["model2"], s1, s2,
                                      opts.altsplice_prob, nodes=nodes)
    res.graph = graph
    t1, t2 = r1.row_tree(), r2.row_tree()
    res._rows = None
    res._tree = _RowTree(parts=[(t1, _child_columns(graph.N, am1, t1.ncols)), (t2, _child_columns(graph.N, am2, t2.ncols))], ncols=graph.N - 2)
    res.tr_homologies = _extend_tr(graph.N, am1, r1.tr_homologies) + _extend_tr(graph.N, am2, r2.tr_homologies)
    res.tr_source = r1.tr_source + r2.tr_source
    res.graph.addRepeats(res.tr_homologies)       # replaces whatever mergeGraphs mapped through (:468)
    return res


def progressive_alignment(sequences, tree, model_factory, opts, align_fn=None, repeats=None,
                          batch_align_fn=None, stats=None, batch_merge_fn=None):
    """-> Result for the root.  sequences: name -> uint8 codes."""
    n = model_factory.alphabet.DIM
    if batch_align_fn is None:
        def rec(nd):
            if nd.is_leaf():
                return _leaf(nd.name, sequences, repeats, n)
            if len(nd.children) != 2:
                raise ValueError("only bifurcating trees allowed")
            r1, r2 = rec(nd.children[0]), rec(nd.children[1])
            job = prepare_job(r1, r2, nd.children[0].length, nd.children[1].length, model_factory, opts)
            ali = align_fn(job["cg1"], job["cg2"], job["model"], opts)
            if stats is not None:
                stats.append(dict(N1=job["cg1"].N, N2=job["cg2"].N, score=ali["score"],
                                  cell_updates=job["cg1"].total_pred_pairs(job["cg2"])))
            return finish_job(r1, r2, job, ali, nd.children[0].support, nd.children[1].support, opts)
        return rec(tree)

    # level-batched schedule: height(node) = 1 + max(height(children)); all nodes of one height are independent
    height, order = {}, []

    def walk(nd):
        if nd.is_leaf():
            height[id(nd)] = 0
        else:
            if len(nd.children) != 2:
                raise ValueError("only bifurcating trees allowed")
            for c in nd.children:
                walk(c)
            height[id(nd)] = 1 + max(height[id(c)] for c in nd.children)
        order.append(nd)
    walk(tree)
    results = {id(nd): _leaf(nd.name, sequences, repeats, n) for nd in order if nd.is_leaf()}
    for h in range(1, height[id(tree)] + 1):
        nodes = [nd for nd in order if height[id(nd)] == h]
        jobs = []
        for nd in nodes:
            r1, r2 = results[id(nd.children[0])], results[id(nd.children[1])]
            jobs.append(prepare_job(r1, r2, nd.children[0].length, nd.children[1].length, model_factory, opts))
        alis = batch_align_fn([(j["cg1"], j["cg2"], j["model"]) for j in jobs], opts)
        profiles = [None] * len(jobs)
        if batch_merge_fn is not None:                      # mergeGraphs' profile part for the whole level in one call
            profiles = batch_merge_fn([merge_job(results[id(nd.children[0])], results[id(nd.children[1])], job, ali)
                                       for nd, job, ali in zip(nodes, jobs, alis)])
        for nd, job, ali, prof in zip(nodes, jobs, alis, profiles):
            r1, r2 = results.pop(id(nd.children[0])), results.pop(id(nd.children[1]))
            if stats is not None:
                stats.append(dict(N1=job["cg1"].N, N2=job["cg2"].N, score=ali["score"], level=h,
                                  cell_updates=job["cg1"].total_pred_pairs(job["cg2"])))
            results[id(nd)] = finish_job(r1, r2, job, ali, nd.children[0].support, nd.children[1].support, opts, nodes=prof)
    return results[id(tree)]
