"""GPU parity, HP-2: the CUDA path called through the C ABI against the CPU oracle on the same
seeded inputs (bit-exact counts / gap openings / NW scores; distances and variances within 1e-9
relative, well inside north_star's 1e-6), and against the reference binary's golden outputs."""
import numpy as np
import pytest

from prographmsa_b200 import alphabet, model, synth
from prographmsa_b200.options import Options
from common import golden, six_digit_close, aa_codes, nw_table

pytestmark = pytest.mark.gpu

AA = alphabet.AA
RTOL = 1e-9


@pytest.fixture(scope="module")
def dist_mod():
    from prographmsa_b200 import distance
    return distance


@pytest.fixture(scope="module")
def po():
    from oracle import pyoracle
    return pyoracle


def _pairs(N):
    return [(i, j) for i in range(N) for j in range(i + 1, N)]


def _check_nw(dist_mod, po, seqs, mf, opts, alpha=AA):
    n = alpha.DIM
    fac = dist_mod.DistanceFactoryAlign(mf, opts)
    N = len(seqs)
    P = N * (N - 1) // 2
    d, v, counts, gaps, scores = fac.pair_range(seqs, 0, P, want_counts=True)
    tab, go, ge = nw_table(alpha)
    for p, (i, j) in enumerate(_pairs(N)):
        c_ref, g_ref, s_ref = po.nw_align_pair(alpha.nw_codes(seqs[i]), alpha.nw_codes(seqs[j]), n, tab, go, ge)
        assert np.array_equal(counts[p], c_ref), (i, j)
        assert gaps[p] == g_ref and scores[p] == s_ref, (i, j, gaps[p], g_ref, scores[p], s_ref)
        d_ref, v_ref, _ = po.compute_distance(c_ref, g_ref, (len(seqs[i]) + len(seqs[j])) / 2.0, mf.Q, mf.freqs, opts)
        assert d[p] == pytest.approx(d_ref, rel=RTOL, abs=1e-300), (i, j)
        assert v[p] == pytest.approx(v_ref, rel=1e-7, abs=1e-300), (i, j)
    return d, v


@pytest.mark.parametrize("opts", [
    Options(mldist_flag=True, nwdist_flag=True),
    Options(nwdist_flag=True),
    Options(mldist_gap_flag=True, nwdist_flag=True),
    Options(mldist_flag=True, nwdist_flag=True, darwin_flag=True),
], ids=["ml", "pdist", "ml_gap", "darwin"])
def test_nw_family_bit_exact(dist_mod, po, opts):
    mf = model.getDefault(opts)
    seqs, _ = synth.simulate_family(20, 150, mf, seed=5, indel_rate=0.03)
    _check_nw(dist_mod, po, list(seqs.values()), mf, opts)


def test_nw_ragged_lengths_and_bands(dist_mod, po):
    """empty, 1-residue, lane-boundary and multi-band (L1 > 512) sequences"""
    opts = Options(mldist_flag=True, nwdist_flag=True)
    mf = model.getDefault(opts)
    rng = np.random.default_rng(3)
    lens = [0, 1, 2, 31, 32, 33, 127, 128, 129, 500, 512, 513, 700, 1100]
    base = rng.integers(0, 20, size=1200).astype(np.uint8)
    seqs = []
    for L in lens:
        s = base[:L].copy()
        mut = rng.random(L) < 0.2
        s[mut] = rng.integers(0, 20, size=int(mut.sum()))
        seqs.append(s)
    seqs.append(np.full(40, 20, dtype=np.uint8))          # all-unknown (X) sequence: nothing is counted
    _check_nw(dist_mod, po, seqs, mf, opts)


def test_nw_small_scratch_chunks_identically(dist_mod, po):
    from prographmsa_b200 import _lib
    opts = Options(mldist_flag=True, nwdist_flag=True)
    mf = model.getDefault(opts)
    seqs, _ = synth.simulate_family(40, 300, mf, seed=8, indel_rate=0.02)
    seqs = list(seqs.values())
    fac = dist_mod.DistanceFactoryAlign(mf, opts)
    P = 40 * 39 // 2
    d0, v0 = fac.pair_range(seqs, 0, P)
    ctx = _lib.Context(0)
    ctx.set_scratch_bytes(64 << 20)
    fac2 = dist_mod.DistanceFactoryAlign(mf, opts, ctx)
    d1, v1 = fac2.pair_range(seqs, 0, P)
    assert np.array_equal(d0, d1) and np.array_equal(v0, v1)
    # sharded ranges (the multi-GPU partition) reproduce the full result
    cuts = [0, 1, 37, 400, P]
    parts = [fac.pair_range(seqs, a, b) for a, b in zip(cuts[:-1], cuts[1:])]
    assert np.array_equal(np.concatenate([p[0] for p in parts]), d0)
    assert np.array_equal(np.concatenate([p[1] for p in parts]), v0)
    ctx.close()


def test_nw_matches_reference_binary(dist_mod):
    for mode, opts in (("nw_ml", Options(mldist_flag=True, nwdist_flag=True)), ("nw_pdist", Options(nwdist_flag=True)),
                       ("nw_ml_gap", Options(mldist_gap_flag=True, nwdist_flag=True)),
                       ("nw_ml_darwin", Options(mldist_flag=True, nwdist_flag=True, darwin_flag=True))):
        mf = model.getDefault(opts)
        fac = dist_mod.DistanceFactoryAlign(mf, opts)
        for rec in golden("pairs.json")["pairs"]:
            dm = fac.computePwDistances({"s0": aa_codes(rec["s0"]), "s1": aa_codes(rec["s1"])}, ["s0", "s1"])
            assert dm.distances[0, 0] == 0 and dm.distances[0, 1] == dm.distances[1, 0]
            for b in rec[mode]:
                assert six_digit_close(dm.distances[0, 1] / 2.0, abs(b)), (mode, rec["t"])


def test_all_pairs_matrix_vs_oracle(dist_mod, po):
    opts = Options(mldist_flag=True, nwdist_flag=True)
    mf = model.getDefault(opts)
    seqs, _ = synth.simulate_family(30, 200, mf, seed=21, indel_rate=0.02)
    names = sorted(seqs)
    dm = dist_mod.DistanceFactoryAlign(mf, opts).computePwDistances(seqs, names)
    tab, go, ge = nw_table(AA)
    d_ref, v_ref = po.nw_ml_all_pairs([AA.nw_codes(seqs[k]) for k in names], 20, tab, go, ge, mf.Q, mf.freqs, opts)
    np.testing.assert_allclose(dm.distances, d_ref, rtol=RTOL, atol=0)
    np.testing.assert_allclose(dm.variances, v_ref, rtol=1e-7, atol=0)
    assert np.array_equal(dm.distances, dm.distances.T) and not dm.distances.diagonal().any()


@pytest.mark.parametrize("alpha_name", ["codon", "dna"])
def test_nw_other_alphabets(dist_mod, po, alpha_name):
    opts = Options(mldist_flag=True, nwdist_flag=True)
    if alpha_name == "codon":
        opts = opts.for_codon()
        mf = model.getDefault(opts, alphabet.Codon)
        alpha, L, N = alphabet.Codon, 120, 8
    else:
        rng = np.random.default_rng(0)
        ex = rng.random(6) + 0.5
        fr = rng.random(4) + 0.5
        mf = model.getDefault(opts, alphabet.DNA, custom_text=" ".join(map(str, ex)) + " " + " ".join(map(str, fr)))
        alpha, L, N = alphabet.DNA, 700, 8
    seqs, _ = synth.simulate_family(N, L, mf, seed=2, indel_rate=0.02)
    _check_nw(dist_mod, po, list(seqs.values()), mf, opts, alpha)


def _random_msa(rng, N, L, n, gap_frac=0.2):
    rows = rng.integers(0, n, size=(N, L)).astype(np.uint8)
    anc = rng.integers(0, n, size=L).astype(np.uint8)
    keep = rng.random((N, L)) < 0.7
    rows = np.where(keep, anc[None, :], rows).astype(np.uint8)
    # gap runs
    for i in range(N):
        pos = 0
        while pos < L:
            if rng.random() < gap_frac:
                ln = int(rng.geometric(0.3))
                rows[i, pos:pos + ln] = alphabet.GAP
                pos += ln
            pos += int(rng.geometric(0.15))
    rows[0, :5] = alphabet.GAP; rows[1, :3] = alphabet.GAP          # leading double-gap columns
    return rows


@pytest.mark.parametrize("n,L", [(20, 333), (20, 31), (61, 200), (4, 1000)])
def test_prealigned_vs_oracle(dist_mod, po, n, L):
    opts = Options(mldist_flag=True)
    if n == 61:
        opts = opts.for_codon(); mf = model.getDefault(opts, alphabet.Codon)
    elif n == 4:
        mf = model.getDefault(opts, alphabet.DNA, custom_text="1 2 1 1 2 1 0.3 0.2 0.2 0.3")
    else:
        mf = model.getDefault(opts)
    rng = np.random.default_rng(n + L)
    N = 12
    rows = _random_msa(rng, N, L, n)
    rows[2, 7] = n                                                   # an unknown state: a match column that is not counted
    fac = dist_mod.DistanceFactoryPrealigned(mf, opts)
    P = N * (N - 1) // 2
    d, v, counts, gaps = fac.pair_range(rows, 0, P, want_counts=True)
    for p, (i, j) in enumerate(_pairs(N)):
        c_ref, g_ref = po.prealigned_counts(rows[i], rows[j], n)
        assert np.array_equal(counts[p], c_ref) and gaps[p] == g_ref, (i, j, gaps[p], g_ref)
        d_ref, v_ref, _ = po.compute_distance(c_ref, g_ref, float(L), mf.Q, mf.freqs, opts)
        assert d[p] == pytest.approx(d_ref, rel=RTOL, abs=1e-300)
        assert v[p] == pytest.approx(v_ref, rel=1e-7, abs=1e-300)


def test_prealigned_flavours_and_all_pairs_matrix_agree(dist_mod, po, monkeypatch):
    """The list-merge column scan (default) against the round-1 scan fused into the Newton kernel (PGM_PREALIGNED_FUSED=1): identical
    counts, gap openings, distances and variances; and pgm_prealigned_ml_all_pairs (results scattered into the N x N matrices on the
    device) against the packed pair range.  Rows as a large alignment has them: mostly gaps, a few residues each, some rows empty."""
    opts = Options(mldist_flag=True)
    mf = model.getDefault(opts)
    rng = np.random.default_rng(77)
    N, L = 40, 2000
    rows = np.full((N, L), alphabet.GAP, dtype=np.uint8)
    shared = np.sort(rng.choice(L, 120, replace=False))
    for i in range(N):
        cols = np.unique(np.concatenate([shared[rng.random(120) < 0.8], rng.integers(0, L, 25)]))
        rows[i, cols] = rng.integers(0, 21, len(cols))              # 20 = an unknown state: resets the runs, is not counted
    rows[3, :] = alphabet.GAP                                        # a row without residues
    rows[5, :] = alphabet.GAP; rows[5, L - 1] = 4                    # one residue in the last column
    fac = dist_mod.DistanceFactoryPrealigned(mf, opts)
    P = N * (N - 1) // 2
    d, v, counts, gaps = fac.pair_range(rows, 0, P, want_counts=True)
    for p, (i, j) in list(enumerate(_pairs(N)))[::7]:
        c_ref, g_ref = po.prealigned_counts(rows[i], rows[j], 20)
        assert np.array_equal(counts[p], c_ref) and gaps[p] == g_ref, (i, j, gaps[p], g_ref)
    monkeypatch.setenv("PGM_PREALIGNED_FUSED", "1")
    d2, v2, counts2, gaps2 = fac.pair_range(rows, 0, P, want_counts=True)
    monkeypatch.delenv("PGM_PREALIGNED_FUSED")
    assert np.array_equal(counts, counts2) and np.array_equal(gaps, gaps2)
    assert np.array_equal(d, d2) and np.array_equal(v, v2)
    d3, v3 = fac.pair_range(rows, 100, 300)                          # a pair range that starts inside row 2
    assert np.array_equal(d3, d[100:300]) and np.array_equal(v3, v[100:300])
    seqs = {"r%03d" % i: rows[i] for i in range(N)}
    dm = fac.computePwDistances(seqs, sorted(seqs))
    for p, (i, j) in enumerate(_pairs(N)):
        assert dm.distances[i, j] == d[p] and dm.distances[j, i] == d[p] and dm.variances[i, j] == v[p] and dm.variances[j, i] == v[p]
    assert not dm.distances.diagonal().any() and not dm.variances.diagonal().any()


def test_prealigned_matches_reference_binary(dist_mod):
    opts = Options(mldist_flag=True)
    mf = model.getDefault(opts)
    fac = dist_mod.DistanceFactory.getDefault(mf, prealigned=True)
    for rec in golden("pairs.json")["pairs"]:
        rows = {k: aa_codes(v) for k, v in rec["alignment_i0"].items()}
        dm = fac.computePwDistances(rows, ["s0", "s1"])
        for b in rec["prealigned_ml"]:
            assert six_digit_close(dm.distances[0, 1] / 2.0, abs(b)), rec["t"]


def test_ml_from_counts_edge_cases(dist_mod, po):
    """empty counts, identical sequences, saturated divergence, clamped (< min_dist) distances"""
    opts = Options(mldist_flag=True)
    mf = model.getDefault(opts)
    rng = np.random.default_rng(9)
    cs = [np.zeros((20, 20), np.int32), np.diag(rng.integers(1, 9, 20)).astype(np.int32)]
    c = np.diag(np.full(20, 10)).astype(np.int32); c[0, 1] = 1; cs.append(c)      # one mismatch in 201: below min_dist
    cs.append(rng.integers(0, 3, (20, 20)).astype(np.int32))                       # saturated
    for _ in range(20):
        c = np.diag(rng.integers(0, 12, 20)).astype(np.int32) + (rng.random((20, 20)) < 0.05 * rng.integers(1, 8)).astype(np.int32)
        cs.append(c)
    counts = np.stack(cs)
    gaps = rng.integers(0, 5, len(cs)).astype(np.uint32)
    seqlen = counts.sum(axis=(1, 2)).astype(np.float64) + 3
    for o in (opts, Options(mldist_gap_flag=True), Options()):
        d, v = dist_mod.ml_distances(mf, o, counts, gaps, seqlen)
        for k in range(len(cs)):
            d_ref, v_ref, _ = po.compute_distance(counts[k], int(gaps[k]), float(seqlen[k]), mf.Q, mf.freqs, o)
            assert d[k] == pytest.approx(d_ref, rel=RTOL, abs=1e-300), (k, o)
            assert v[k] == pytest.approx(v_ref, rel=1e-7, abs=1e-300), (k, o)


def test_config2_size_properties(dist_mod):
    """BASELINE config 2 at full size (500 x 400 aa): size-independent properties -- symmetry, zero
    diagonal, clamps, and a shard-union checksum equal to the single-call checksum."""
    opts = Options(mldist_flag=True, nwdist_flag=True)
    mf = model.getDefault(opts)
    seqs, _ = synth.simulate_family(500, 400, mf, seed=1)
    names = sorted(seqs)
    lst = [seqs[k] for k in names]
    fac = dist_mod.DistanceFactoryAlign(mf, opts)
    dm = fac.computePwDistances(seqs, names)
    D, V = dm.distances, dm.variances
    assert np.array_equal(D, D.T) and np.array_equal(V, V.T) and not D.diagonal().any()
    iu = np.triu_indices(500, 1)
    assert np.all(D[iu] > 0) and np.all(D[iu] <= 2.2) and np.all(V[iu] >= 1e-5) and np.all(V[iu] <= 1e3)
    P = 500 * 499 // 2
    cuts = np.linspace(0, P, 9).astype(np.int64)
    parts = [fac.pair_range(lst, int(a), int(b))[0] for a, b in zip(cuts[:-1], cuts[1:])]
    assert np.array_equal(np.concatenate(parts), D[iu])


def test_nw_32bit_flavour_matches_16bit(dist_mod, po, monkeypatch):
    """the 32-bit forward kernel (long sequences / large scores) and the 16-bit two-pairs-per-warp kernel
    must produce identical counts, gap openings, scores and distances"""
    from prographmsa_b200 import _lib
    opts = Options(mldist_flag=True, nwdist_flag=True)
    mf = model.getDefault(opts)
    seqs, _ = synth.simulate_family(33, 260, mf, seed=77, indel_rate=0.03)     # odd count: a lone pair per task
    lst = list(seqs.values())
    P = 33 * 32 // 2
    a = dist_mod.DistanceFactoryAlign(mf, opts).pair_range(lst, 0, P, want_counts=True)
    monkeypatch.setenv("PGM_NW_FORCE_32BIT", "1")
    ctx = _lib.Context(0)
    b = dist_mod.DistanceFactoryAlign(mf, opts, ctx).pair_range(lst, 0, P, want_counts=True)
    ctx.close()
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


@pytest.mark.parametrize("N,seed", [(2, 0), (3, 1), (4, 2), (5, 3), (17, 4), (200, 5), (700, 6)])
def test_bionj_join_loop_bit_identical_to_host_specification(N, seed):
    """N3: the GPU join loop (pgm_bionj_joins) against tree.nj_joins_host: same joins in the same order, branch lengths
    and the remaining <= 3 x 3 matrix bit for bit; random symmetric matrices incl. values below the clamps and exact ties"""
    from prographmsa_b200 import distance, tree
    rng = np.random.default_rng(seed)
    A = rng.random((N, N)) * (2.0 if seed % 2 else 0.02); D = np.round((A + A.T) / 2, 3 if N <= 17 else 12)      # rounded: exact ties
    B = rng.random((N, N)) * 1e-3; V = (B + B.T) / 2
    if N > 4:
        V[1, 2] = V[2, 1] = 0.0                                                                                  # below VAR_MIN
    np.fill_diagonal(D, 0.0); np.fill_diagonal(V, 0.0)
    ref_j, ref_D = tree.nj_joins_host(D, V)
    got_j, got_D = distance.bionj_joins(D, V)
    assert [(a, b) for a, b, _, _ in got_j] == [(a, b) for a, b, _, _ in ref_j]
    assert np.array_equal(np.array([[c, d] for _, _, c, d in got_j]).view(np.uint64), np.array([[c, d] for _, _, c, d in ref_j]).view(np.uint64))
    assert np.array_equal(np.asarray(got_D).view(np.uint64), np.asarray(ref_D).view(np.uint64))


def test_bionj_tree_from_gpu_joins_equals_host_tree():
    """whole guide tree (join loop on the GPU, tree objects + midpoint root on the host) on real distances"""
    from prographmsa_b200 import distance, tree
    opts = Options(mldist_flag=True, nwdist_flag=True)
    mf = model.getDefault(opts)
    seqs, _ = synth.simulate_family(60, 120, mf, seed=21, indel_rate=0.03)
    names = sorted(seqs)
    dm = distance.DistanceFactoryAlign(mf, opts).computePwDistances(seqs, names)
    assert tree.tree_nj(names, dm, opts.edge_halflife, joins_fn=distance.bionj_joins).newick() == tree.tree_nj(names, dm, opts.edge_halflife).newick()
