"""GPU parity, HP-1: emission scores, DP score, traceback mappings and whole progressive alignments
from the CUDA path (through the C ABI) against the CPU oracle and the reference binary's goldens.
PGM_EMISSION_EXACT is bit-identical to the oracle, so every comparison here is exact."""
import numpy as np
import pytest

from prographmsa_b200 import alphabet, model, synth, progressive
from prographmsa_b200.graph import sequence_graph, CleanedGraph
from prographmsa_b200.options import Options
from common import golden, aa_codes

pytestmark = pytest.mark.gpu
AA = alphabet.AA


@pytest.fixture(scope="module")
def ga():
    from prographmsa_b200 import graphalign
    return graphalign


@pytest.fixture(scope="module")
def po():
    from oracle import pyoracle
    return pyoracle


def _interior(S):
    return S[1:-1, 1:-1]


def _same_alignment(a, b):
    assert a["score"] == b["score"], (a["score"], b["score"])
    assert a["n_tr_indels"] == b["n_tr_indels"]
    assert np.array_equal(a["mapping1"], b["mapping1"]) and np.array_equal(a["mapping2"], b["mapping2"])


def test_emission_exact_bitwise(ga, po):
    opts = Options(mldist_flag=True)
    mf = model.getDefault(opts)
    a, b = synth.random_pair(97, 0.4, mf, seed=4)
    g1, g2 = sequence_graph(a, 20), sequence_graph(b, 20)
    m = mf.getModel(0.5, 0.5)
    S_ref = po.precompute_scores(g1, g2, m, -0.25)
    S = ga.emission_scores(g1, g2, m, -0.25)
    assert np.array_equal(_interior(S).view(np.uint32), _interior(S_ref).view(np.uint32))


def test_dynprog_scores_host(ga, po):
    opts = Options(mldist_flag=True)
    mf = model.getDefault(opts)
    a, b = synth.random_pair(60, 0.4, mf, seed=5)
    g1, g2 = sequence_graph(a, 20), sequence_graph(b, 20)
    for o in (opts, Options(mldist_flag=True, end_indel_prob=-1.0), Options(mldist_flag=True, repeat_rate=0.0)):
        m = mf.getModel(0.7, 0.9)
        assert np.array_equal(ga.dynprog_scores(g1, g2, m, o).view(np.uint32), po.dynprog_scores(g1, g2, m, o).view(np.uint32))


@pytest.mark.parametrize("L,t,seed", [(1, 0.3, 1), (2, 0.3, 2), (40, 0.1, 3), (150, 0.8, 4), (300, 1.5, 5), (700, 0.4, 6)])
def test_leaf_pair_alignment_identical(ga, po, L, t, seed):
    opts = Options(mldist_flag=True)
    mf = model.getDefault(opts)
    a, b = synth.random_pair(L, t, mf, seed=seed)
    g1, g2 = sequence_graph(a, 20), sequence_graph(b, 20)
    m = mf.getModel(t, t)
    _same_alignment(ga.alignGraphs(g1, g2, m, opts), po.align_graphs(g1, g2, m, opts))


@pytest.mark.parametrize("L", [1100, 1500, 2300, 3300, 4200])
def test_long_graphs_identical(ga, po, L):
    """graph 1 beyond 1024 rows (BASELINE config 3 has 1500-nt sequences, the root of a 5000-sequence tree > 6000 nodes): the
    rows are swept in bands (128 rows, or one warp) by CTAs that hand over through HBM; PGM_FILL_BAND_MIN_N1 raises the threshold,
    which brings back the single-CTA flavours with several rows per thread"""
    opts = Options(mldist_flag=True)
    mf = model.getDefault(opts)
    a, b = synth.random_pair(L, 0.4, mf, seed=L)
    g1, g2 = sequence_graph(a, 20), sequence_graph(b, 20)
    m = mf.getModel(0.4, 0.4)
    _same_alignment(ga.alignGraphs(g1, g2, m, opts), po.align_graphs(g1, g2, m, opts))


def test_too_long_graph_is_an_error_not_a_crash(ga):
    from prographmsa_b200._lib import PgmError
    opts = Options(mldist_flag=True)
    mf = model.getDefault(opts)
    a, b = synth.random_pair(61000, 0.01, mf, seed=1)
    with pytest.raises(PgmError):
        ga.alignGraphs(sequence_graph(a, 20), sequence_graph(b, 20), mf.getModel(0.4, 0.4), opts)


def test_empty_sequences(ga, po):
    opts = Options(mldist_flag=True)
    mf = model.getDefault(opts)
    e = sequence_graph(np.zeros(0, np.uint8), 20)
    s = sequence_graph(np.array([3, 4, 5], np.uint8), 20)
    m = mf.getModel(0.3, 0.3)
    for g1, g2 in ((e, e), (e, s), (s, e)):
        _same_alignment(ga.alignGraphs(g1, g2, m, opts), po.align_graphs(g1, g2, m, opts))


@pytest.mark.parametrize("idx", range(10))
def test_progressive_alignment_matches_reference_binary(ga, idx):
    """whole families: every internal node's alignGraphs runs on the GPU; the final text equals the
    reference binary's output"""
    from test_oracle_golden import _opts_from_extra
    fam = golden("alignments.json")["families"][idx]
    opts = _opts_from_extra(fam["extra"])
    mf = model.getDefault(opts)
    seqs = {k: aa_codes(v) for k, v in fam["seqs"].items()}
    tree = progressive.parse_newick(fam["newick"])
    res = progressive.progressive_alignment(seqs, tree, mf, opts, batch_align_fn=lambda jobs, o: ga.alignGraphsBatch(jobs, o))
    assert {k: AA.decode(v) for k, v in res.aligned.items()} == fam["aligned"]


def test_merged_graph_jobs_identical_to_oracle(ga, po):
    """every job of a 40-sequence progressive pass (skip edges, multi-predecessor nodes): score and
    mappings identical to the oracle, job by job"""
    opts = Options(mldist_flag=True)
    mf = model.getDefault(opts)
    seqs, root = synth.simulate_family(40, 120, mf, seed=33, indel_rate=0.05)
    tree = progressive.from_synth(root)
    seen = []

    def batch(jobs, o):
        out = ga.alignGraphsBatch(jobs, o)
        for (g1, g2, m), r in zip(jobs, out):
            ref = po.align_graphs(g1, g2, m, o)
            _same_alignment(r, ref)
            assert r["cell_updates"] == ref["cell_updates"]
            seen.append(g1.N * g2.N)
        return out
    progressive.progressive_alignment(seqs, tree, mf, opts, batch_align_fn=batch)
    assert len(seen) == 39


def test_repeat_edges_identical_to_oracle(ga, po):
    """tandem-repeat edges: repeat predecessors, n_tr_indels and markAlternativePath"""
    opts = Options(mldist_flag=True)
    mf = model.getDefault(opts)
    rng = np.random.default_rng(12)
    unit = rng.integers(0, 20, 7).astype(np.uint8)
    flank1, flank2 = rng.integers(0, 20, 15).astype(np.uint8), rng.integers(0, 20, 12).astype(np.uint8)
    a = np.concatenate([flank1, np.tile(unit, 5), flank2])
    b = np.concatenate([flank1, np.tile(unit, 3), flank2])
    reps = {}
    for name, s, nu in (("a", a, 5), ("b", b, 3)):
        reps[name] = [dict(start=len(flank1), tr_hom=list(range(7)) * nu)]
    results = {}
    for name, s in (("a", a), ("b", b)):
        results[name] = progressive._leaf(name, {name: s}, reps, 20)
    ga_, gb_ = results["a"].graph, results["b"].graph
    assert ga_.r_col.size > 0
    m = mf.getModel(0.2, 0.2)
    cg1, cg2 = CleanedGraph(ga_, opts.repeatext_prob), CleanedGraph(gb_, opts.repeatext_prob)
    ref = po.align_graphs(cg1, cg2, m, opts)
    got = ga.alignGraphs(cg1, cg2, m, opts)
    _same_alignment(got, ref)
    assert ref["n_tr_indels"] > 0


# ---- tensor-core emission (tcgen05, split-TF32): tolerance-based ------------------------------------
S_TOL_BITS = 2e-5          # |S_tensor - S_exact| in log2 units (per cell); dropped lo*lo term + fp32 accumulate order
SCORE_RTOL = 1e-5          # north_star: DP scores within 1e-5 relative


@pytest.mark.parametrize("n,L", [(20, 97), (20, 300), (4, 333), (61, 150)])
def test_emission_tensor_close_to_exact(ga, n, L):
    opts = Options(mldist_flag=True)
    if n == 61:
        opts = opts.for_codon(); mf = model.getDefault(opts, alphabet.Codon)
    elif n == 4:
        mf = model.getDefault(opts, alphabet.DNA, custom_text="1 2 1 1 2 1 0.3 0.2 0.2 0.3")
    else:
        mf = model.getDefault(opts)
    a, b = synth.random_pair(L, 0.5, mf, seed=n + L)
    g1, g2 = sequence_graph(a, n), sequence_graph(b, n)
    m = mf.getModel(0.6, 0.6)
    # profile (non one-hot) columns too: push one graph through P
    g2.sites[1:-1] = g2.sites[1:-1] @ m.P.T
    S0 = ga.emission_scores(g1, g2, m, -0.3, ga.EMISSION_EXACT)
    S1 = ga.emission_scores(g1, g2, m, -0.3, ga.EMISSION_TENSOR)
    err = np.abs(_interior(S1) - _interior(S0)).max()
    assert err <= S_TOL_BITS, err


def test_tensor_alignments_within_tolerance(ga, po):
    """every job of a progressive pass in tensor mode: score within 1e-5 relative of the oracle; the
    traceback is identical unless the oracle's own score margin is inside the tolerance (ties reported)"""
    opts = Options(mldist_flag=True)
    mf = model.getDefault(opts)
    seqs, root = synth.simulate_family(32, 150, mf, seed=44, indel_rate=0.04)
    tree = progressive.from_synth(root)
    stats = dict(jobs=0, identical=0)

    def batch(jobs, o):
        out = ga.alignGraphsBatch(jobs, o, emission_mode=ga.EMISSION_TENSOR)
        ref = [po.align_graphs(g1, g2, m, o) for g1, g2, m in jobs]
        for r, e in zip(out, ref):
            assert abs(r["score"] - e["score"]) <= SCORE_RTOL * abs(e["score"])
            stats["jobs"] += 1
            stats["identical"] += int(np.array_equal(r["mapping1"], e["mapping1"]) and np.array_equal(r["mapping2"], e["mapping2"]))
        return ref                      # continue the pass on the reference path so every job sees reference inputs
    progressive.progressive_alignment(seqs, tree, mf, opts, batch_align_fn=batch)
    print("tensor-mode tracebacks identical: %d / %d jobs" % (stats["identical"], stats["jobs"]))
    assert stats["identical"] >= stats["jobs"] - 2          # near-tie flips are reported, not hidden


def test_large_batch_is_deterministic_and_correct(ga, po):
    """600 jobs in one batch (several CTAs per SM, multiple waves): identical results on two runs and
    agreement with the oracle on a sample -- guards the asynchronous S prefetch ring of the fill kernel"""
    opts = Options(mldist_flag=True)
    mf = model.getDefault(opts)
    m = mf.getModel(0.5, 0.5)
    jobs = []
    for k in range(600):
        a, b = synth.random_pair(180 + (k % 7) * 11, 0.5, mf, seed=1000 + k)
        jobs.append((sequence_graph(a, 20), sequence_graph(b, 20), m))
    r1 = ga.alignGraphsBatch(jobs, opts)
    r2 = ga.alignGraphsBatch(jobs, opts)
    assert [r["score"] for r in r1] == [r["score"] for r in r2]
    for k in range(0, 600, 37):
        _same_alignment(r1[k], po.align_graphs(jobs[k][0], jobs[k][1], m, opts))


# ---- packed traceback: every fill flavour / direction-word width must give the oracle's path -------------------
FLAVOURS = {
    "default": {},
    "block": {"PGM_FILL_LAYOUT": "block"},
    "diag": {"PGM_FILL_LAYOUT": "diag", "PGM_FILL_WAVE": "0"},
    "banded": {"PGM_FILL_LAYOUT": "diag", "PGM_FILL_BAND_MIN_N1": "66", "PGM_FILL_BAND_ROWS": "64"},
    "banded_one_warp": {"PGM_FILL_LAYOUT": "diag", "PGM_FILL_BAND_MIN_N1": "66", "PGM_FILL_BAND_ROWS": "32"},
    "banded_auto_height": {"PGM_FILL_LAYOUT": "diag", "PGM_FILL_BAND_MIN_N1": "66"},
    "strip_cta_barrier": {"PGM_STRIP_DEBUG": "16", "PGM_FILL_STRIP": "3"},
    "nostrip": {"PGM_FILL_STRIP": "0"},
    "strip_chain_only": {"PGM_FILL_STRIP": "1"},
    "strip_c4": {"PGM_STRIP_C": "4"},
    "strip_c8": {"PGM_STRIP_C": "8"},
    "strip_c16": {"PGM_STRIP_C": "16"},
    "strip_c2": {"PGM_STRIP_C": "2"},
    "strip_small_rings": {"PGM_STRIP_HIST_KB": "6"},
    "strip_general_only": {"PGM_FILL_STRIP": "3"},
    "strip_full_s": {"PGM_SITE_CLASSES": "0"},
    "wave": {"PGM_FILL_LAYOUT": "diag", "PGM_FILL_WAVE": "1"},
    "wave_l1": {"PGM_FILL_LAYOUT": "diag", "PGM_FILL_WAVE": "1", "PGM_WAVE_L": "1"},
    "wave_l2_ring8": {"PGM_FILL_LAYOUT": "diag", "PGM_FILL_WAVE": "1", "PGM_WAVE_L": "2", "PGM_WAVE_RING": "8"},
    "wave_pipeline3": {"PGM_FILL_LAYOUT": "diag", "PGM_GRAPH_PIPELINE": "3"},
    "pipeline2": {"PGM_GRAPH_PIPELINE": "2"},
    "wave_ring16": {"PGM_FILL_LAYOUT": "diag", "PGM_FILL_WAVE": "1", "PGM_WAVE_RING": "16"},
}


@pytest.fixture(params=sorted(FLAVOURS))
def flavour(request, monkeypatch):
    for k, v in FLAVOURS[request.param].items():
        monkeypatch.setenv(k, v)
    return request.param


def test_every_fill_flavour_identical_to_oracle(ga, po, flavour):
    """a 40-sequence progressive pass with many indels (skip edges, multi-predecessor nodes, far predecessors that
    are read back from the value store) under every fill kernel and strip width"""
    opts = Options(mldist_flag=True)
    mf = model.getDefault(opts)
    seqs, root = synth.simulate_family(40, 150, mf, seed=51, indel_rate=0.08)
    tree = progressive.from_synth(root)
    n = [0]

    def batch(jobs, o):
        out = ga.alignGraphsBatch(jobs, o)
        for (g1, g2, m), r in zip(jobs, out):
            _same_alignment(r, po.align_graphs(g1, g2, m, o))
            n[0] += 1
        return out
    progressive.progressive_alignment(seqs, tree, mf, opts, batch_align_fn=batch)
    assert n[0] == 39


@pytest.mark.parametrize("E", [0.12, 0.01, 0.5, 0.99, -1.0])
def test_boundary_row_and_column_follow_the_reference_rule(ga, po, flavour, E):
    """paths that run along row 0 / column 0 (one sequence is a suffix, a prefix or an infix of the other): there the
    reference's traceback recomputes the candidates with gap_init although the fill used start_gap
    (GraphAlign.h:212-234 against :378-398), which the packed traceback reproduces from the kept boundary values"""
    opts = Options(mldist_flag=True, end_indel_prob=E)
    mf = model.getDefault(opts)
    rng = np.random.default_rng(7)
    a = rng.integers(0, 20, 90).astype(np.uint8)
    m = mf.getModel(0.2, 0.2)
    cases = [(a, a[35:]), (a[35:], a), (a, a[:40]), (a[:40], a), (a, a[20:70]), (a[20:70], a),
             (a, rng.integers(0, 20, 30).astype(np.uint8)), (a[:3], a[50:])]
    jobs = [(sequence_graph(p, 20), sequence_graph(q, 20), m) for p, q in cases]
    for (g1, g2, mm), r in zip(jobs, ga.alignGraphsBatch(jobs, opts)):
        _same_alignment(r, po.align_graphs(g1, g2, mm, opts))


def _branchy_graph(rng, N, extra_preds, n=20):
    """a chain of N nodes in which some nodes get `extra_preds` more predecessors at random distances (ascending costs)"""
    from prographmsa_b200.graph import Graph
    sites = np.zeros((N, n))
    sites[np.arange(1, N - 1), rng.integers(0, n, N - 2)] = 1.0
    edges = {(i, i - 1): 0.0 for i in range(1, N)}
    for to in rng.choice(np.arange(2, N), size=max(1, N // 6), replace=False):
        for frm in rng.choice(np.arange(0, to - 1), size=min(extra_preds, to - 1), replace=False):
            edges[(int(to), int(frm))] = float(rng.uniform(0.5, 6.0))
    return Graph(sites, edges)


@pytest.mark.parametrize("extra", [3, 7, 20])
def test_many_predecessors(ga, po, flavour, extra):
    """nodes with up to 21 predecessors at arbitrary distances: most predecessors are far (value store, list walks)"""
    opts = Options(mldist_flag=True)
    mf = model.getDefault(opts)
    rng = np.random.default_rng(100 + extra)
    m = mf.getModel(0.3, 0.3)
    jobs = [(_branchy_graph(rng, 70 + 13 * k, extra), _branchy_graph(rng, 90 - 7 * k, extra), m) for k in range(4)]
    for (g1, g2, mm), r in zip(jobs, ga.alignGraphsBatch(jobs, opts)):
        _same_alignment(r, po.align_graphs(g1, g2, mm, opts))


def test_more_predecessors_than_the_packed_traceback_holds(ga, po):
    """a node with 141 predecessors: beyond the index fields of the packed traceback, so the job takes the dense kernels"""
    opts = Options(mldist_flag=True)
    mf = model.getDefault(opts)
    rng = np.random.default_rng(5)
    g1, g2 = _branchy_graph(rng, 200, 140), _branchy_graph(rng, 50, 2)
    m = mf.getModel(0.3, 0.3)
    _same_alignment(ga.alignGraphs(g1, g2, m, opts), po.align_graphs(g1, g2, m, opts))


def test_failed_emission_fails_the_job(ga, monkeypatch):
    """a tensor-emission failure (bounded mbarrier wait timed out) must surface as a failed job, not as an
    alignment of a partly written S.  The failure is injected into the emission status word of job 1."""
    from prographmsa_b200._lib import PgmError
    opts = Options(mldist_flag=True)
    mf = model.getDefault(opts)
    m = mf.getModel(0.5, 0.5)
    jobs = []
    for k in range(3):
        a, b = synth.random_pair(60, 0.5, mf, seed=70 + k)
        jobs.append((sequence_graph(a, 20), sequence_graph(b, 20), m))
    for mode in (ga.EMISSION_EXACT, ga.EMISSION_TENSOR):
        assert len(ga.alignGraphsBatch(jobs, opts, emission_mode=mode)) == 3
        monkeypatch.setenv("PGM_INJECT_EMISSION_FAIL", "1")
        with pytest.raises(PgmError):
            ga.alignGraphsBatch(jobs, opts, emission_mode=mode)
        monkeypatch.delenv("PGM_INJECT_EMISSION_FAIL")


@pytest.mark.parametrize("L1,L2", [(3, 3), (1, 40), (40, 1), (13, 14), (130, 417), (417, 130), (400, 400), (760, 300), (300, 767), (2000, 700)])
def test_strip_kernel_leaf_pairs(ga, po, flavour, L1, L2):
    """chain x chain jobs go to the strip kernel (every strip width, partial last strips, graphs shorter than a strip);
    unknown residues make uniform profile columns (their own site class)"""
    opts = Options(mldist_flag=True)
    mf = model.getDefault(opts)
    rng = np.random.default_rng(L1 * 1000 + L2)
    a = rng.integers(0, 20, L1).astype(np.uint8)
    b = rng.integers(0, 20, L2).astype(np.uint8)
    k = min(L1, L2)
    b[:k // 2] = a[:k // 2]                                     # some homology
    if L1 > 5:
        a[rng.integers(0, L1, max(1, L1 // 20))] = 20           # unknown residues
    if L2 > 5:
        b[rng.integers(0, L2, max(1, L2 // 25))] = 21
    g1, g2 = sequence_graph(a, 20), sequence_graph(b, 20)
    m = mf.getModel(0.4, 0.4)
    _same_alignment(ga.alignGraphs(g1, g2, m, opts), po.align_graphs(g1, g2, m, opts))


def test_strip_kernel_profile_chain_graphs(ga, po, flavour):
    """chain graphs whose columns are real profiles (not class-coded): S is read as a full matrix by the strip kernel"""
    opts = Options(mldist_flag=True)
    mf = model.getDefault(opts)
    m = mf.getModel(0.5, 0.5)
    jobs = []
    for k in range(6):
        a, b = synth.random_pair(150 + 37 * k, 0.5, mf, seed=300 + k)
        g1, g2 = sequence_graph(a, 20), sequence_graph(b, 20)
        if k % 3 != 1:
            g1.sites[1:-1] = g1.sites[1:-1] @ m.P.T
        if k % 3 != 2:
            g2.sites[1:-1] = g2.sites[1:-1] @ m.P.T
        jobs.append((g1, g2, m))
    for (g1, g2, mm), r in zip(jobs, ga.alignGraphsBatch(jobs, opts)):
        _same_alignment(r, po.align_graphs(g1, g2, mm, opts))


def test_more_jobs_than_sms_multi_warp_strips(ga, po):
    """every job of a pass, replicated to more jobs than the GPU has SMs: CTAs of the general strip kernel then start on an SM
    whose shared memory still holds the previous CTA's data (regression: the warp hand-over slot of row -1 must be -inf,
    not whatever was left there)"""
    opts = Options(mldist_flag=True)
    mf = model.getDefault(opts)
    seqs, root = synth.simulate_family(48, 300, mf, seed=61, indel_rate=0.03)
    tree = progressive.from_synth(root)
    jobs, refs = [], []

    def batch(js, o):
        out = [po.align_graphs(g1, g2, m, o) for g1, g2, m in js]
        jobs.extend(js); refs.extend(out)
        return out
    progressive.progressive_alignment(seqs, tree, mf, opts, batch_align_fn=batch)
    rep = 8
    for _ in range(2):
        got = ga.alignGraphsBatch(jobs * rep, opts)
        for k, r in enumerate(got):
            _same_alignment(r, refs[k % len(jobs)])


def test_wave_kernel_large_and_mixed_shapes(ga, po):
    """the role-sorted anti-diagonal kernel near its limits: ~1000 rows (threads), long graph 2 (shared-memory staging), many
    non-chain columns (column threads), and a launch whose jobs fit one by one but not with the maxima of all of them"""
    opts = Options(mldist_flag=True)
    mf = model.getDefault(opts)
    rng = np.random.default_rng(77)
    m = mf.getModel(0.3, 0.3)
    shapes = [(1000, 120, 2, 2), (120, 1400, 2, 2), (700, 260, 3, 1), (260, 700, 1, 3), (1020, 90, 1, 1), (60, 2600, 1, 2)]
    jobs = [(_branchy_graph(rng, a, ea), _branchy_graph(rng, b, eb), m) for a, b, ea, eb in shapes]
    for (g1, g2, mm), r in zip(jobs, ga.alignGraphsBatch(jobs, opts)):
        _same_alignment(r, po.align_graphs(g1, g2, mm, opts))
