"""The CPU oracle against the reference binary's own outputs (tests/golden, tools/make_golden.py).

These pin the oracle: HP-2 distances at the binary's 6 printed digits for every distance method,
HP-1 through the exact text of multi-sequence alignments under a fixed guide tree.
"""
import numpy as np
import pytest

from prographmsa_b200 import alphabet, model, progressive
from prographmsa_b200.options import Options
from oracle import pyoracle as po
from common import golden, six_digit_close, aa_codes, nw_table

PAIRS = golden("pairs.json")["pairs"]
FAMS = golden("alignments.json")["families"]
AA = alphabet.AA


def _nw_dist(rec, opts):
    mf = model.getDefault(opts)
    tab, go, ge = nw_table(AA)
    c0, c1 = AA.nw_codes(aa_codes(rec["s0"])), AA.nw_codes(aa_codes(rec["s1"]))
    counts, gaps, _ = po.nw_align_pair(c0, c1, 20, tab, go, ge)
    d, v, _ = po.compute_distance(counts, gaps, (len(c0) + len(c1)) / 2.0, mf.Q, mf.freqs, opts)
    return d


@pytest.mark.parametrize("mode,opts", [
    ("nw_ml", Options(mldist_flag=True, nwdist_flag=True)),
    ("nw_pdist", Options(nwdist_flag=True)),
    ("nw_ml_gap", Options(mldist_gap_flag=True, nwdist_flag=True)),
    ("nw_ml_darwin", Options(mldist_flag=True, nwdist_flag=True, darwin_flag=True)),
])
def test_nw_distance_matches_binary(mode, opts):
    for rec in PAIRS:
        d = _nw_dist(rec, opts)
        for b in rec[mode]:
            assert six_digit_close(d / 2.0, abs(b)), (mode, rec["t"], d / 2.0, b)


def test_prealigned_distance_matches_binary():
    opts = Options(mldist_flag=True)
    mf = model.getDefault(opts)
    for rec in PAIRS:
        r0, r1 = aa_codes(rec["alignment_i0"]["s0"]), aa_codes(rec["alignment_i0"]["s1"])
        counts, gaps = po.prealigned_counts(r0, r1, 20)
        d, v, _ = po.compute_distance(counts, gaps, len(r0), mf.Q, mf.freqs, opts)
        for b in rec["prealigned_ml"]:
            assert six_digit_close(d / 2.0, abs(b)), (rec["t"], d / 2.0, b)


def _opts_from_extra(extra):
    kw = dict(mldist_flag=True)
    it = iter(extra)
    for flag in it:
        if flag == "-w": kw["darwin_flag"] = True
        elif flag == "-E": kw["end_indel_prob"] = float(next(it))
        elif flag == "-g": kw["indel_rate"] = float(next(it))
        elif flag == "-s": kw["altsplice_prob"] = float(next(it))
        elif flag == "-l": kw["edge_halflife"] = float(next(it))
        else: raise ValueError(flag)
    return Options(**kw)


def _align_family(fam, align_fn):
    opts = _opts_from_extra(fam["extra"])
    mf = model.getDefault(opts)
    seqs = {k: aa_codes(v) for k, v in fam["seqs"].items()}
    tree = progressive.parse_newick(fam["newick"])
    res = progressive.progressive_alignment(seqs, tree, mf, opts, align_fn=align_fn)
    return {k: AA.decode(v) for k, v in res.aligned.items()}


@pytest.mark.parametrize("idx", range(len(FAMS)))
def test_two_leaf_and_family_alignments_match_binary(idx):
    fam = FAMS[idx]
    got = _align_family(fam, po.align_graphs)
    # the binary prints the ORIGINAL letters at non-gap positions (stringFromSequence with orig)
    assert got == fam["aligned"]


def test_pair_alignment_matches_binary():
    opts = Options(mldist_flag=True)
    mf = model.getDefault(opts)
    # "-m -i 0" on two sequences: the initial tree comes from k-mer angle distances (not on the hot
    # path, N1 in SURVEY §8f); its branch is recorded in the golden file, so feed it as the tree.
    for rec in PAIRS:
        b = rec["angle_ml"]
        if b[0] == 0:
            continue
        seqs = {"s0": aa_codes(rec["s0"]), "s1": aa_codes(rec["s1"])}
        tree = progressive.parse_newick("(s1:%r,s0:%r);" % (abs(b[0]), abs(b[1])))
        res = progressive.progressive_alignment(seqs, tree, mf, opts, align_fn=po.align_graphs)
        got = {k: AA.decode(v) for k, v in res.aligned.items()}
        assert got == rec["alignment_i0"], rec["t"]


def test_oracle_progressive_pass_1000_sequences_matches_reference_binary():
    """a whole progressive pass over 1000 x 500 aa under a fixed tree (999 alignGraphs + mergeGraphs nodes): the oracle
    reproduces the reference binary's alignment (SHA-256 in large.json).  This is the case that pinned the summation
    order of the two emission normalisers (oracle.cpp dot_gemv)."""
    import hashlib
    from prographmsa_b200 import synth, progressive
    from oracle import pyoracle as po
    rec = [r for r in golden("large.json")["runs"] if r["n"] == 1000][0]
    opts = Options(mldist_flag=True)
    mf = model.getDefault(opts)
    seqs, root = synth.simulate_family(rec["n"], rec["L"], mf, seed=rec["seed"], forbid_first=10)
    tree = progressive.parse_newick(root.newick(6) + ";")
    res = progressive.progressive_alignment(seqs, tree, mf, opts,
                                            batch_align_fn=lambda jobs, o: [po.align_graphs(a, b, m, o) for a, b, m in jobs])
    h = hashlib.sha256()
    for k in sorted(res.aligned):
        h.update(k.encode()); h.update(b"\n"); h.update(alphabet.AA.decode(res.aligned[k]).encode()); h.update(b"\n")
    assert h.hexdigest() == rec["sha256"]
