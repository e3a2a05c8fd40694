"""End-to-end driver: the caller of both hot paths, mirroring doAlign<A> (main.cpp:324-483).

  initial guide tree  : TreeNJ(seqs, prealigned=false)  -> DistanceFactoryAlign (-a) or DistanceFactoryAngle
  iterations (-i)     : progressive_alignment -> TreeNJ(aligned, prealigned=true) (DistanceFactoryPrealigned)
  final alignment     : progressive_alignment on the last tree
Host glue only (tree building, merging, I/O stay on the host, as north_star prescribes); every distance
matrix and every alignGraphs call goes through the `backend`, which is the GPU path in the product
(`GpuBackend`) and the oracle in the CPU tests.
"""
import numpy as np

from . import alphabet as _alpha
from . import model as _model
from . import progressive as _prog
from . import tree as _tree
from .alphabet import GAP
from .options import Options


class GpuBackend:
    """All three distance factories + batched alignGraphs on the GPU (C ABI)."""

    def __init__(self, model_factory, opts, ctx=None, emission_mode=0):
        from . import distance, graphalign
        self._d, self._g = distance, graphalign
        self.mf, self.opts, self.ctx, self.mode = model_factory, opts, ctx, emission_mode

    def initial_distances(self, seqs, names):
        if self.opts.nwdist_flag:
            return self._d.DistanceFactoryAlign(self.mf, self.opts, self.ctx).computePwDistances(seqs, names)
        return self._d.DistanceFactoryAngle(self.mf, self.opts, self.ctx).computePwDistances(seqs, names)

    def prealigned_distances(self, rows, names):
        return self._d.DistanceFactoryPrealigned(self.mf, self.opts, self.ctx).computePwDistances(rows, names)

    def batch_align(self, jobs, opts):
        return self._g.alignGraphsBatch(jobs, opts, self.mode, self.ctx)

    def batch_merge(self, jobs):
        return self._g.merge_profiles_batch(jobs, self.ctx)

    def nj_joins(self, D, V):
        return self._d.bionj_joins(D, V, self.ctx)


def do_align(seq_strings, opts: Options, backend_factory, alphabet=_alpha.AA, iters=2, return_trees=False):
    """seq_strings: name -> residue string.  -> (name -> aligned string in tree order, [newick of every tree])."""
    assert alphabet is _alpha.AA, "the end-to-end driver is wired for proteins (configs 1, 2, 5)"
    names = sorted(seq_strings)                                     # std::map order (TreeNJ.h:34-39)
    stripped, seqs = {}, {}
    for k in names:
        s = seq_strings[k]
        stripped[k] = len(s) > 0 and s[0] == "M"                    # main.cpp:333-353 (AA::stripStart = 'M')
        seqs[k] = alphabet.encode(s[1:] if stripped[k] else s)
    any_stripped = any(stripped.values())
    mf = _model.getDefault(opts, alphabet, code_seqs=[seqs[k] for k in names])
    be = backend_factory(mf, opts)

    trees = []
    tree = _tree.tree_nj(names, be.initial_distances(seqs, names), opts.edge_halflife, joins_fn=getattr(be, "nj_joins", None))
    trees.append(tree.newick())
    result = old = None
    for i in range(iters):
        result = _prog.progressive_alignment(seqs, _tree.to_progressive(tree), mf, opts, batch_align_fn=be.batch_align,
                                              batch_merge_fn=getattr(be, "batch_merge", None))
        aligned = {k: v for k, v in result.aligned.items() if not k.startswith("(")}
        if i > 0 and old is not None and all(np.array_equal(aligned[k], old[k]) for k in aligned):
            break
        anames = sorted(aligned)
        tree = _tree.tree_nj(anames, be.prealigned_distances(aligned, anames), opts.edge_halflife, joins_fn=getattr(be, "nj_joins", None))
        trees.append(tree.newick())
        old = aligned
    result = _prog.progressive_alignment(seqs, _tree.to_progressive(tree), mf, opts, batch_align_fn=be.batch_align,
                                              batch_merge_fn=getattr(be, "batch_merge", None))

    out = {}
    for k in tree.leaf_order():
        row = np.asarray(result.aligned[k])
        orig = seq_strings[k]
        j = 1 if (any_stripped and stripped[k]) else 0               # re-insert the start residue column (main.cpp:455-463)
        body = np.full(len(row), ord("-"), dtype=np.uint8)
        keep = row != GAP
        residues = np.frombuffer(orig[j:].encode("ascii"), dtype=np.uint8)
        assert int(keep.sum()) == len(residues)
        body[keep] = residues
        head = (orig[0] if stripped[k] else "-") if any_stripped else ""
        out[k] = head + body.tobytes().decode("ascii")
    return (out, trees) if return_trees else out
