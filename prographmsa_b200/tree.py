"""Host-side guide tree: BioNJ + midpoint rooting (reference TreeNJ.cpp:22-29,132-281, PhyTree.cpp:11-116,
PhyTree.h:40-54,227-229).  north_star keeps NJ tree-building on the host; this is plain numpy glue that
follows the reference's arithmetic order so that the branch lengths (which feed getModel at every tree
node) agree to rounding.  The O(N^3) join loop has a GPU twin (csrc/nj_bionj.cu) that reproduces
nj_joins_host bit for bit; tree objects, the last three taxa and midpoint rooting stay here."""
import math
import numpy as np

COLSUM = None
MIN_DIST = 1e-4
MIN_VAR = 1e-5


class PhyTree:
    __slots__ = ("name", "children", "parent", "length", "support")

    def __init__(self, name=""):
        self.name, self.children, self.parent, self.length, self.support = name, [], None, 0.0, 1.0

    def is_leaf(self):
        return not self.children

    def add_child(self, child, length=0.0, support=1.0):
        assert child.parent is None
        self.children.append(child)
        child.parent, child.length, child.support = self, length, support

    def pluck(self):
        p = self.parent
        p.children.remove(self)
        self.parent, self.length, self.support = None, 0.0, 1.0

    def newick(self):
        """PhyTree::formatNewick: default ostream precision (6 significant digits)."""
        def rec(t):
            if t.is_leaf():
                return t.name
            return "(" + ",".join("%s:%s" % (rec(c), _fmt(c.length)) for c in t.children) + ")"
        return rec(self) + ";"

    def leaf_order(self):
        if self.is_leaf():
            return [self.name]
        return [n for c in self.children for n in c.leaf_order()]


def _fmt(x):
    s = "%.6g" % x
    if "e" in s:                       # ostream prints e-05, printf prints e-05 as well; keep as is
        return s
    return s


def support(d, edge_halflife):
    s = 1.0 - math.exp(-math.log(2.0) * d / edge_halflife) if edge_halflife != 0 else float("nan")
    s = min(1.0, max(0.0, s))
    return 0.0 if math.isnan(s) else s


def nj_joins_host(D, V):
    """The join loop of buildNJTree (TreeNJ.cpp:132-235) in numpy: -> (joins, D_rest) with joins = [(i1, i2, d1, d2)] in
    CURRENT indices (i2 is removed, later indices shift down, the new cluster takes i1).  This is the specification the
    GPU join loop (pgm_bionj_joins) reproduces bit for bit; sums run sequentially in index order."""
    D = np.array(D, dtype=np.float64); V = np.array(V, dtype=np.float64)
    dim = D.shape[0]
    joins = []
    while dim > 3:
        D = np.maximum(D, MIN_DIST); V = np.maximum(V, MIN_VAR)
        np.fill_diagonal(D, 0.0); np.fill_diagonal(V, 0.0)
        sums = np.array([COLSUM(D[:, j], j * dim) for j in range(dim)]) if COLSUM else D.sum(axis=0)
        S = np.ones((dim, dim)) * sums[None, :]
        Q = 0.5 * D - (0.5 / (dim - 2.0)) * (S + S.T)
        np.fill_diagonal(Q, np.inf)
        flat = int(np.argmin(Q.T.reshape(-1)))          # Eigen minCoeff walks column-major, first minimum wins
        i2, i1 = flat % dim, flat // dim                  # (row, col) = (index2, index1)
        if i2 < i1:
            i1, i2 = i2, i1
        d1 = (D[i1, i2] + (sums[i1] - sums[i2]) / (dim - 2.0)) / 2.0
        d1 = min(max(d1, MIN_DIST), D[i1, i2])
        d2 = max(D[i2, i1] - d1, MIN_DIST)
        lam = 0.5 + float(np.cumsum(V[i2, :] - V[i1, :])[-1]) / (2 * (dim - 2) * V[i1, i2]) if V[i1, i2] != 0 else float("nan")
        lam = 0.5 if math.isnan(lam) else min(max(0.0, lam), 1.0)
        keep = [k for k in range(dim) if k != i2]
        new_d = lam * (D[i1, keep] - d1) + (1.0 - lam) * (D[i2, keep] - d2)
        new_v = lam * V[i1, keep] + (1.0 - lam) * V[i2, keep] - lam * (1.0 - lam) * V[i1, i2]
        D = D[np.ix_(keep, keep)].copy(); V = V[np.ix_(keep, keep)].copy()
        new_d[i1] = 0.0; new_v[i1] = 0.0
        D[i1, :] = new_d; D[:, i1] = new_d
        V[i1, :] = new_v; V[:, i1] = new_v
        joins.append((i1, i2, float(d1), float(d2)))
        dim -= 1
    return joins, D


def build_nj_tree(names, distances, variances, edge_halflife, joins_fn=None):
    """buildNJTree without a topology plan (TreeNJ.cpp:132-281).  joins_fn(D, V) -> (joins, D_rest): the O(N^3) join
    loop, nj_joins_host by default, the GPU's (distance.bionj_joins) when a device back-end passes it."""
    names = list(names)
    sub = [PhyTree(n) for n in names]
    joins, D = (joins_fn or nj_joins_host)(distances, variances)
    for i1, i2, d1, d2 in joins:
        name1, name2 = names[i1], names[i2]
        del names[i2]
        names[i1] = name1 + "," + name2
        t = PhyTree(names[i1])
        t.add_child(sub[i1], d1, support(d1, edge_halflife))
        t.add_child(sub[i2], d2, support(d2, edge_halflife))
        del sub[i2]
        sub[i1] = t
    root = PhyTree("root")
    if len(names) == 2:
        d = D[0, 1] / 2
        root.add_child(sub[0], d, support(d, edge_halflife))
        root.add_child(sub[1], d, support(d, edge_halflife))
    else:
        d0 = (D[0, 1] + D[0, 2] - D[1, 2]) / 2.0
        d0 = min(max(d0, MIN_DIST), min(D[1, 0], D[2, 0]))
        d1 = max(D[1, 0] - d0, MIN_DIST)
        d2 = max(D[2, 0] - d0, MIN_DIST)
        t2 = PhyTree("root2")
        t2.add_child(sub[0], d0, support(d0, edge_halflife))
        t2.add_child(sub[1], d1, support(d1, edge_halflife))
        root.add_child(sub[2], d2 / 2, support(d2, edge_halflife))
        root.add_child(t2, d2 / 2, support(d2, edge_halflife))
    return root


def _max_dist_pair(root):
    """the two most distant leaves (PhyTree.cpp:60-80): every pair of leaves is met once at its lowest common ancestor, in
    the order of the reference's double loop, and a strictly larger sum replaces the current best -- evaluated per node as one
    outer sum with a first-maximum argmax, which visits the pairs in that same order."""
    best = {"max": None, "dist": -math.inf}

    def rec(t):
        if t.is_leaf():
            return np.zeros(1), [t]
        d1, l1 = rec(t.children[0]); d2, l2 = rec(t.children[1])
        d1 = d1 + t.children[0].length
        d2 = d2 + t.children[1].length
        S = d1[:, None] + d2[None, :]
        flat = int(np.argmax(S))
        i, j = divmod(flat, len(d2))
        if S[i, j] > best["dist"]:
            best["dist"] = float(S[i, j])
            best["max"] = l1[i] if d1[i] > d2[j] else l2[j]
        return np.concatenate([d1, d2]), l1 + l2
    rec(root)
    return best["dist"], best["max"]


def midpoint_root(root):
    """midpointRoot (PhyTree.cpp:60-116)."""
    dist, mx = _max_dist_pair(root)
    cur = mx
    dist /= 2
    while cur is not root and dist - cur.length > 0:
        dist -= cur.length
        cur = cur.parent
    if cur is root:
        return root
    new_root = PhyTree("new_root")
    cur_dist = cur.length - dist
    cur_support = cur.support
    parent = cur.parent
    cur.pluck()
    new_root.add_child(cur, dist, cur_support)
    cur = new_root
    while parent is not root:
        nd, ns, np_ = parent.length, parent.support, parent.parent
        parent.pluck()
        cur.add_child(parent, cur_dist, cur_support)
        cur, parent, cur_dist, cur_support = parent, np_, nd, ns
    assert len(root.children) == 1
    other = root.children[0]
    cur_dist += other.length
    cur_support = max(cur_support, other.support)
    other.pluck()
    cur.add_child(other, cur_dist, cur_support)
    return new_root


def tree_nj(names, dm, edge_halflife, joins_fn=None):
    """TreeNJ<A>() after computePwDistances (TreeNJ.h:45-58): zero diagonals, BioNJ, midpoint root."""
    D = np.array(dm.distances, dtype=np.float64); V = np.array(dm.variances, dtype=np.float64)
    np.fill_diagonal(D, 0.0); np.fill_diagonal(V, 0.0)
    return midpoint_root(build_nj_tree(names, D, V, edge_halflife, joins_fn=joins_fn))


def to_progressive(t):
    from .progressive import TreeNode
    nd = TreeNode(name=t.name if t.is_leaf() else None, length=t.length, support=t.support)
    nd.children = [to_progressive(c) for c in t.children]
    return nd
