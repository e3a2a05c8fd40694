"""Synthetic sequence families: random rooted Yule tree, branch lengths
Exp(mean) capped, root i.i.d. from pi, substitutions sampled from P(t) of the model, indels at
the reference's default rate / geometric length (main.cpp:63-66).  Only valid states; AA families
never start with 'M' (the reference strips a leading M, main.cpp:333-353).
Deterministic for a given seed.  Used by tests, bench.py and tools/make_golden.py.
"""
import numpy as np


class Node:
    __slots__ = ("children", "length", "name", "seq")

    def __init__(self, name=None, length=0.0):
        self.children, self.length, self.name, self.seq = [], length, name, None

    def is_leaf(self):
        return not self.children

    def leaves(self):
        if self.is_leaf():
            return [self]
        return [l for c in self.children for l in c.leaves()]

    def newick(self, digits=6):
        if self.is_leaf():
            return self.name
        return "(" + ",".join("%s:%.*g" % (c.newick(digits), digits, c.length) for c in self.children) + ")"


def yule_tree(n_leaves, rng, mean_len=0.1, cap=0.5, prefix="s"):
    width = len(str(n_leaves - 1))
    active = [Node() for _ in range(1)]
    root = active[0]
    while len(active) < n_leaves:
        k = rng.integers(len(active))
        parent = active.pop(k)
        for _ in range(2):
            c = Node(length=min(cap, rng.exponential(mean_len)))
            parent.children.append(c)
            active.append(c)
    for i, leaf in enumerate(root.leaves()):
        leaf.name = "%s%0*d" % (prefix, width, i)
    return root


def _evolve(seq, t, P_of, rng, indel_rate, gapext, pi_cdf):
    n = len(pi_cdf)
    P = P_of(t)
    cdf = np.cumsum(P, axis=1)
    cdf[:, -1] = 1.0
    u = rng.random(len(seq))
    out = (u[:, None] > cdf[seq]).sum(axis=1).astype(np.uint8)
    out = np.minimum(out, n - 1)
    # indels: Poisson number of events, half insertions half deletions, geometric lengths
    n_events = rng.poisson(indel_rate * t * max(1, len(out)))
    for _ in range(n_events):
        ln = rng.geometric(1.0 - gapext)
        pos = rng.integers(0, len(out) + 1)
        if rng.random() < 0.5:
            ins = np.minimum((rng.random(ln)[:, None] > pi_cdf[None, :]).sum(axis=1), n - 1).astype(np.uint8)
            out = np.concatenate([out[:pos], ins, out[pos:]])
        elif len(out) > ln + 8:
            pos = min(pos, len(out) - ln)
            out = np.concatenate([out[:pos], out[pos + ln:]])
    return out


def simulate_family(n_seqs, length, model_factory, seed=1, mean_len=0.1, cap=0.5, indel_rate=0.0093359375,
                    gapext=0.6119140625, forbid_first=None):
    """-> (dict name -> uint8 state codes, tree root Node)."""
    rng = np.random.default_rng(seed)
    root = yule_tree(n_seqs, rng, mean_len, cap)
    pi = np.asarray(model_factory.freqs, dtype=np.float64)
    pi_cdf = np.cumsum(pi / pi.sum()); pi_cdf[-1] = 1.0
    n = len(pi)
    root.seq = np.minimum((rng.random(length)[:, None] > pi_cdf[None, :]).sum(axis=1), n - 1).astype(np.uint8)
    stack = [root]
    while stack:
        nd = stack.pop()
        for c in nd.children:
            c.seq = _evolve(nd.seq, c.length, model_factory.P_of, rng, indel_rate, gapext, pi_cdf)
            stack.append(c)
    seqs = {}
    for leaf in root.leaves():
        s = leaf.seq.copy()
        if forbid_first is not None and len(s) and s[0] == forbid_first:
            s[0] = (forbid_first + 1) % n
        seqs[leaf.name] = s
    return seqs, root


def random_pair(length, divergence_t, model_factory, seed):
    """two sequences separated by total distance divergence_t"""
    rng = np.random.default_rng(seed)
    pi = np.asarray(model_factory.freqs); pi_cdf = np.cumsum(pi / pi.sum()); pi_cdf[-1] = 1.0
    n = len(pi)
    a = np.minimum((rng.random(length)[:, None] > pi_cdf[None, :]).sum(axis=1), n - 1).astype(np.uint8)
    b = _evolve(a, divergence_t, model_factory.P_of, rng, 0.0093359375, 0.6119140625, pi_cdf)
    return a, b
