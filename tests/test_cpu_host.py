"""CPU-side tests (no GPU): the C-ABI library loads and exports every declared symbol, the host-only
entry points agree with the oracle, the eigen-basis reproduces exp(Qt), and the 2-rank sharding logic
reassembles the full matrix (gloo)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from prographmsa_b200 import alphabet, model, synth, sharding, _lib
from prographmsa_b200.graph import sequence_graph
from prographmsa_b200.options import Options
from oracle import pyoracle as po

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_library_exports_every_declared_symbol():
    lib = _lib.lib()
    header = open(os.path.join(ROOT, "include", "pgm_b200.h")).read()
    declared = set(re.findall(r"\b(pgm_[a-z0-9_]+)\s*\(", header))
    declared -= {"pgm_ctx", "pgm_graph_batch"}
    assert declared, "no declarations parsed"
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert set(_lib.SYMBOLS) <= declared
    assert lib.pgm_abi_version() == 1


def test_no_cpu_fallback_without_a_gpu():
    """the product fails loudly when no B200 is present"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.PgmError) as e:
        _lib.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "prographmsa_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pyoracle" not in txt and "liboracle" not in txt and "oracle/" not in txt.replace("tests use the oracle", ""), f


@pytest.mark.parametrize("kind", ["wag", "darwin", "ecm", "custom_dna", "plusf"])
def test_eigen_basis_reproduces_expm(kind):
    o = Options(mldist_flag=True)
    if kind == "wag": mf = model.ModelFactoryWag(o)
    elif kind == "darwin": mf = model.ModelFactoryDarwin(o)
    elif kind == "ecm": mf = model.ModelFactoryEcm(o.for_codon())
    elif kind == "custom_dna": mf = model.ModelFactoryCustom(o, alphabet.DNA, "1 2 1 1 2 1 0.3 0.2 0.2 0.3")
    else:
        mf = model.ModelFactoryPlusF(model.ModelFactoryWag(o))
        rng = np.random.default_rng(0)
        mf.estimateFreqs([rng.integers(0, 20, 300).astype(np.uint8) for _ in range(20)])
    Q = np.ascontiguousarray(mf.Q)
    n = Q.shape[0]
    V = np.zeros((n, n)); Vi = np.zeros((n, n)); s = np.zeros(n)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert _lib.lib().pgm_debug_eigen_basis(p(Q), n, p(V), p(s), p(Vi)) == 0, _lib.lib().pgm_last_error()
    for t in (0.05, 0.7, 2.2, 5.0):
        P = (V * np.exp(s * t)) @ Vi
        np.testing.assert_allclose(P, po.expm(Q, t), rtol=0, atol=5e-14)
        np.testing.assert_allclose(P, mf.P_of(t), rtol=0, atol=5e-14)


def test_dynprog_scores_host_bitwise():
    from prographmsa_b200 import graphalign
    opts = Options(mldist_flag=True)
    mf = model.getDefault(opts)
    a, b = synth.random_pair(80, 0.4, mf, seed=5)
    g1, g2 = sequence_graph(a, 20), sequence_graph(b, 20)
    for o in (opts, Options(mldist_flag=True, end_indel_prob=-1.0), Options(mldist_flag=True, repeat_rate=0.0),
              Options(mldist_flag=True, darwin_flag=True)):
        m = model.getDefault(o).getModel(0.7, 0.9)
        got = graphalign.dynprog_scores(g1, g2, m, o)
        ref = po.dynprog_scores(g1, g2, m, o)
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


def test_pair_shards_partition_the_triangle():
    for N in (2, 3, 17, 500):
        for world in (1, 2, 3, 8):
            cuts = [sharding.pair_shard(N, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == sharding.n_pairs(N)
            assert all(a[1] == b[0] for a, b in zip(cuts[:-1], cuts[1:]))
    assert sharding.pair_index(0, 1, 5) == 0 and sharding.pair_index(3, 4, 5) == 9


_GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from prographmsa_b200 import alphabet, model, synth, sharding, distance
from prographmsa_b200.options import Options
from oracle import pyoracle as po
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
opts = Options(mldist_flag=True, nwdist_flag=True)
mf = model.getDefault(opts)
seqs, _ = synth.simulate_family(14, 60, mf, seed=3, indel_rate=0.03)
names = sorted(seqs); N = len(names)
codes = [alphabet.AA.nw_codes(seqs[k]) for k in names]
tab, go, ge = distance.nw_table(alphabet.AA)
pb, pe = sharding.pair_shard(N, rank, world)
per = sharding.shard_size(N, world)
local = torch.zeros(2, per, dtype=torch.float64)
pairs = [(i, j) for i in range(N) for j in range(i + 1, N)]
for k, (i, j) in enumerate(pairs[pb:pe]):          # the rank's pair range; the oracle stands in for the GPU kernels
    c, g, _ = po.nw_align_pair(codes[i], codes[j], 20, tab, go, ge)
    d, v, _ = po.compute_distance(c, g, (len(codes[i]) + len(codes[j])) / 2.0, mf.Q, mf.freqs, opts)
    local[0, k], local[1, k] = d, v
gathered = sharding.all_gather_shards(dist, local, world)
D, V = sharding.scatter_to_matrices(gathered.numpy(), N)
D_ref, V_ref = po.nw_ml_all_pairs(codes, 20, tab, go, ge, mf.Q, mf.freqs, opts)
assert np.array_equal(D, D_ref) and np.array_equal(V, V_ref), rank
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_two_rank_sharding_reassembles_full_matrix_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611", str(script), ROOT],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert r.stdout.count("ok") == 2


# ---- native host glue (csrc/host_graph.cpp) vs its Python specification (graph.py) -----------------------------
def _graph_arrays(g):
    return [g.sites, g.e_rowptr, g.e_col, g.e_val, g.r_rowptr, g.r_col, g.r_units]


def _run_progressive(native, seqs, tree, mf, opts, repeats=None):
    from prographmsa_b200 import graph, progressive
    old = graph.NATIVE
    graph.NATIVE = native
    try:
        seen = []

        def align(a, b, m, o):
            seen.append((a, b))
            return po.align_graphs(a, b, m, o)
        res = progressive.progressive_alignment(seqs, tree, mf, opts, align_fn=align, repeats=repeats)
        return res, seen
    finally:
        graph.NATIVE = old


@pytest.mark.parametrize("seed,indel", [(3, 0.02), (5, 0.08)])
def test_native_host_glue_equals_python_specification(seed, indel):
    """CleanedGraph, merge plan and merged-graph edges: every array of every graph of a progressive pass is identical
    whether the C++ glue (pgm_host_*) or the Python restatement built it."""
    from prographmsa_b200 import progressive
    opts = Options(mldist_flag=True)
    mf = model.getDefault(opts)
    seqs, root = synth.simulate_family(12, 90, mf, seed=seed, indel_rate=indel)
    tree = progressive.from_synth(root)
    r_nat, seen_nat = _run_progressive(True, seqs, tree, mf, opts)
    r_py, seen_py = _run_progressive(False, seqs, tree, mf, opts)
    assert len(seen_nat) == len(seen_py) == 11
    for (a1, b1), (a2, b2) in zip(seen_nat, seen_py):
        for x, y in zip(_graph_arrays(a1) + _graph_arrays(b1) + [a1.outmapping, b1.outmapping],
                        _graph_arrays(a2) + _graph_arrays(b2) + [a2.outmapping, b2.outmapping]):
            assert x.dtype == y.dtype and np.array_equal(x, y)
    for x, y in zip(_graph_arrays(r_nat.graph), _graph_arrays(r_py.graph)):
        assert np.array_equal(x, y)
    assert all(np.array_equal(r_nat.aligned[k], r_py.aligned[k]) for k in r_py.aligned)


def test_native_host_glue_with_repeat_edges():
    """tandem-repeat edges through CleanedGraph (repeatext_prob = 0 drops multi-unit repeats) and mergeGraphs."""
    from prographmsa_b200 import progressive
    rng = np.random.default_rng(12)
    unit = rng.integers(0, 20, 7).astype(np.uint8)
    f1, f2 = rng.integers(0, 20, 15).astype(np.uint8), rng.integers(0, 20, 12).astype(np.uint8)
    seqs = {"a": np.concatenate([f1, np.tile(unit, 5), f2]), "b": np.concatenate([f1, np.tile(unit, 3), f2]),
            "c": np.concatenate([f1, np.tile(unit, 4), f2[:9]])}
    reps = {k: [dict(start=len(f1), tr_hom=list(range(7)) * nu)] for k, nu in (("a", 5), ("b", 3), ("c", 4))}
    tree = progressive.parse_newick("((a:0.1,b:0.12):0.05,c:0.2);")
    for opts in (Options(mldist_flag=True), Options(mldist_flag=True, repeatext_prob=0.0)):
        mf = model.getDefault(opts)
        r_nat, seen_nat = _run_progressive(True, seqs, tree, mf, opts, reps)
        r_py, seen_py = _run_progressive(False, seqs, tree, mf, opts, reps)
        assert seen_nat[0][0].r_col.size > 0 or opts.repeatext_prob == 0.0
        for (a1, b1), (a2, b2) in zip(seen_nat, seen_py):
            for x, y in zip(_graph_arrays(a1) + _graph_arrays(b1), _graph_arrays(a2) + _graph_arrays(b2)):
                assert np.array_equal(x, y)
        for x, y in zip(_graph_arrays(r_nat.graph), _graph_arrays(r_py.graph)):
            assert np.array_equal(x, y)
        assert r_nat.n_tr_indels == r_py.n_tr_indels


def test_first_maximum_rule_reproduces_the_reference_traceback():
    """groundwork for direction-word tracebacks: on the oracle's dumped matrices, following the first
    candidate that attains the maximum gives the same path as the reference's closest-|delta| search"""
    import importlib.util
    import numpy as np
    from prographmsa_b200 import model, synth, progressive
    from prographmsa_b200.options import Options
    from oracle import pyoracle as po
    spec = importlib.util.spec_from_file_location("check_first_max", os.path.join(os.path.dirname(__file__), "first_max_traceback.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    opts = Options(mldist_flag=True)
    mf = model.getDefault(opts)
    seqs, root = synth.simulate_family(14, 70, mf, seed=3, indel_rate=0.06)
    checked = []

    def batch(jobs, o):
        out = []
        for g1, g2, m in jobs:
            r = po.align_graphs(g1, g2, m, o, dump=True)
            with np.errstate(invalid="ignore", over="ignore"):
                p = mod.first_max_trace(g1, g2, r, po.dynprog_scores(g1, g2, m, o))
            assert p == list(zip(r["mapping1"].tolist()[::-1], r["mapping2"].tolist()[::-1]))
            checked.append(len(p))
            out.append(r)
        return out
    progressive.progressive_alignment(seqs, progressive.from_synth(root), mf, opts, batch_align_fn=batch)
    assert len(checked) == 13
