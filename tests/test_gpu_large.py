"""GPU parity at BASELINE.json configs[4]'s shape: a whole progressive pass (every alignGraphs + mergeGraphs node) over
1000 and 5000 synthetic 500-aa sequences under a fixed guide tree, compared with the reference CPU binary's alignment
through a SHA-256 (tests/golden/large.json, made by tools/make_golden_large.py from the same seed)."""
import hashlib
import os
import sys
import time

import pytest

from prographmsa_b200 import alphabet, model, synth, progressive
from prographmsa_b200.options import Options
from common import golden

pytestmark = pytest.mark.gpu
AA = alphabet.AA


def _digest(aligned):
    h = hashlib.sha256()
    for k in sorted(aligned):
        h.update(k.encode()); h.update(b"\n"); h.update(aligned[k].encode()); h.update(b"\n")
    return h.hexdigest()


@pytest.mark.parametrize("n", [1000, 5000])
def test_progressive_pass_matches_reference_binary_at_scale(n):
    from prographmsa_b200 import graphalign
    rec = [r for r in golden("large.json")["runs"] if r["n"] == n][0]
    opts = Options(mldist_flag=True)
    mf = model.getDefault(opts)
    seqs, root = synth.simulate_family(rec["n"], rec["L"], mf, seed=rec["seed"], forbid_first=10)
    fa = "".join(">%s\n%s\n" % (k, "".join(alphabet.AA_LETTERS[c] for c in v)) for k, v in seqs.items())
    assert hashlib.sha256(fa.encode()).hexdigest() == rec["fasta_sha256"], "the synthetic family is not the one the golden was made from"
    tree = progressive.parse_newick(root.newick(6) + ";")
    t0 = time.perf_counter()
    res = progressive.progressive_alignment(seqs, tree, mf, opts, batch_align_fn=lambda jobs, o: graphalign.alignGraphsBatch(jobs, o),
                                            batch_merge_fn=lambda jobs: graphalign.merge_profiles_batch(jobs))
    wall = time.perf_counter() - t0
    aligned = {k: AA.decode(v) for k, v in res.aligned.items()}
    cols = len(next(iter(aligned.values())))
    print("n=%d: %d columns, progressive pass %.1f s (reference binary: %.1f s)" % (n, cols, wall, rec["reference_wall_s"]), file=sys.stderr)
    assert cols == rec["columns"]
    assert _digest(aligned) == rec["sha256"]


@pytest.mark.parametrize("idx", range(len(golden("large.json")["whole_runs"])))
def test_whole_run_matches_reference_binary_at_scale(idx):
    """BASELINE.json configs[4]'s pipeline end to end on the GPU back-ends -- k-mer angle distances, the BioNJ join loop, every
    alignGraphs and mergeGraphs node, and for -i 2 the Prealigned ML distance matrices of both guide-tree iterations -- against
    whole runs of the reference binary (SHA-256 of the final alignment): 1000 sequences with -i 2 (the binary: 157 s) and 5000
    sequences with -i 0 (the binary: 47 minutes)."""
    from prographmsa_b200 import pipeline
    rec = golden("large.json")["whole_runs"][idx]
    iters = int(rec["flags"][rec["flags"].index("-i") + 1])
    opts = Options(mldist_flag=True, aafreqs_flag="-F" in rec["flags"], nwdist_flag="-a" in rec["flags"], darwin_flag="-w" in rec["flags"])
    mf0 = model.getDefault(Options(mldist_flag=True))
    seqs, _ = synth.simulate_family(rec["n"], rec["L"], mf0, seed=rec["seed"], forbid_first=10)
    strs = {k: "".join(alphabet.AA_LETTERS[c] for c in v) for k, v in seqs.items()}
    t0 = time.perf_counter()
    out = pipeline.do_align(strs, opts, lambda mf, o: pipeline.GpuBackend(mf, o), iters=iters)
    wall = time.perf_counter() - t0
    cols = len(next(iter(out.values())))
    h = hashlib.sha256("".join("%s\n%s\n" % (k, out[k]) for k in sorted(out)).encode()).hexdigest()
    print("whole run, %d x %d aa, %s: %d columns in %.1f s (reference binary: %.0f s)" % (rec["n"], rec["L"], " ".join(rec["flags"]), cols, wall, rec["reference_wall_s"]), file=sys.stderr)
    assert cols == rec["columns"]
    assert h == rec["sha256"]
