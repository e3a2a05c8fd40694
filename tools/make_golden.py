#!/usr/bin/env python3
"""Generate golden vectors with the reference's own prebuilt CPU binary.

The reference ships no tests or fixtures, so the parity pins are produced here,
in the build container, by running /root/reference/bin/ProGraphMSA_64 (staged as
oracle/_ref/ProGraphMSA_64) on seeded synthetic inputs.  Output: tests/golden/*.json (committed).

  pairs.json      2-sequence runs: a 2-leaf tree's branch length is distance/2, printed with 6
                  significant digits, for every distance method on the HP-2 path
                  (-a -m: NW+ML, -a: NW p-distance, -a -M: NW+ML with gap term, -a -m -w: Gonnet,
                   -m -i 1 --all_trees: Prealigned ML on the binary's own first alignment).
  alignments.json multi-sequence runs with a fixed guide tree (-t tree -m -i 0 -f): the final
                  alignment text pins alignGraphs + CleanedGraph + mergeGraphs end to end.
"""
import json, os, subprocess, sys, tempfile
import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from prographmsa_b200 import alphabet, model, synth
from prographmsa_b200.options import Options

BIN = os.path.join(ROOT, "oracle", "_ref", "ProGraphMSA_64")
AA = alphabet.AA_LETTERS


def to_str(codes):
    return "".join(AA[c] for c in codes)


def run(args, fasta_text, extra_files=None):
    with tempfile.TemporaryDirectory() as d:
        with open(os.path.join(d, "in.fa"), "w") as f:
            f.write(fasta_text)
        for name, text in (extra_files or {}).items():
            with open(os.path.join(d, name), "w") as f:
                f.write(text)
        r = subprocess.run([BIN] + args + ["in.fa"], cwd=d, capture_output=True, text=True, timeout=600)
        if r.returncode != 0:
            raise RuntimeError("binary failed: %s\n%s" % (args, r.stderr[-500:]))
        return r.stdout, r.stderr


def parse_fasta(text):
    out, name = {}, None
    for line in text.splitlines():
        if line.startswith(">"):
            name = line[1:].strip(); out[name] = ""
        elif name is not None:
            out[name] += line.strip()
    return out


def branch_of(tree_line):
    # "(s1:0.131988,s0:0.131988);"
    import re
    return [float(x) for x in re.findall(r":([-+0-9.eE]+)", tree_line)]


def main():
    wag = model.ModelFactoryWag(Options(mldist_flag=True))
    pairs = []
    cases = [(60, 0.05), (60, 0.3), (120, 0.02), (120, 0.6), (200, 0.15), (200, 1.0), (300, 0.4),
             (300, 1.8), (90, 0.9), (150, 2.5), (400, 0.25), (80, 0.005), (250, 0.7), (50, 1.3)]
    for k, (L, t) in enumerate(cases):
        a, b = synth.random_pair(L, t, wag, seed=1000 + k)
        if a[0] == 10: a[0] = 0
        if b[0] == 10: b[0] = 0
        sa, sb = to_str(a), to_str(b)
        fa = ">s0\n%s\n>s1\n%s\n" % (sa, sb)
        rec = dict(s0=sa, s1=sb, t=t)
        for mode, args in (("nw_ml", ["-a", "-m", "-i", "0", "-T"]),
                           ("nw_pdist", ["-a", "-i", "0", "-T"]),
                           ("nw_ml_gap", ["-a", "-M", "-i", "0", "-T"]),
                           ("nw_ml_darwin", ["-a", "-m", "-w", "-i", "0", "-T"])):
            out, _ = run(args, fa)
            rec[mode] = branch_of(out.strip().splitlines()[-1])
        out, _ = run(["-m", "-i", "1", "-T", "--all_trees"], fa)
        lines = out.strip().splitlines()
        rec["prealigned_ml"] = branch_of(lines[1])
        rec["angle_ml"] = branch_of(lines[0])
        out, _ = run(["-m", "-i", "0", "-f"], fa)
        rec["alignment_i0"] = parse_fasta(out)
        pairs.append(rec)
        print("pair", k, L, t, rec["nw_ml"], rec["prealigned_ml"], file=sys.stderr)
    with open(os.path.join(ROOT, "tests", "golden", "pairs.json"), "w") as f:
        json.dump(dict(generator="tools/make_golden.py", binary="ProGraphMSA_64 (2013-07-09 rev 5e7b708)", pairs=pairs), f, indent=1)

    alis = []
    fam_cases = [(4, 40, 11, []), (6, 60, 12, []), (8, 80, 13, ["-E", "-1"]), (12, 120, 14, []),
                 (16, 100, 15, ["-w"]), (20, 200, 16, []), (10, 150, 17, ["-g", "0.03"]), (32, 120, 18, []),
                 (24, 90, 19, ["-s", "0.1", "-l", "0.1"]), (50, 300, 20, [])]
    for n, L, seed, extra in fam_cases:
        opts = Options(mldist_flag=True, darwin_flag=("-w" in extra))
        mf = model.getDefault(opts)
        seqs, root = synth.simulate_family(n, L, mf, seed=seed, indel_rate=0.03, forbid_first=10)
        nwk = root.newick(6) + ";"
        fa = "".join(">%s\n%s\n" % (k, to_str(v)) for k, v in seqs.items())
        out, _ = run(["-t", "tree.nwk", "-m", "-i", "0", "-f"] + extra, fa, {"tree.nwk": nwk + "\n"})
        alis.append(dict(n=n, L=L, seed=seed, extra=extra, newick=nwk, seqs={k: to_str(v) for k, v in seqs.items()},
                         aligned=parse_fasta(out)))
        print("family", n, L, "cols", len(next(iter(alis[-1]["aligned"].values()))), file=sys.stderr)
    with open(os.path.join(ROOT, "tests", "golden", "alignments.json"), "w") as f:
        json.dump(dict(generator="tools/make_golden.py", binary="ProGraphMSA_64 (2013-07-09 rev 5e7b708)", families=alis), f, indent=1)


def pipelines():
    """whole runs (initial tree from k-mer angle or NW+ML distances, two guide-tree iterations, final
    alignment): `-i 2 -f` text + `-T --all_trees`.  (50, 300, -m) is BASELINE.json configs[0]."""
    out = []
    # the last case has sequences that start with M (every other one): main.cpp:333-353 strips that start residue before
    # aligning and re-inserts it as an extra first column (:455-463)
    for n, L, seed, flags, start_m in [(8, 60, 5, ["-m"], False), (20, 120, 6, ["-m"], False), (20, 120, 6, ["-a", "-m"], False),
                                       (50, 300, 7, ["-m"], False), (10, 80, 9, ["-m"], True)]:
        opts = Options(mldist_flag=True, nwdist_flag=("-a" in flags))
        mf = model.getDefault(opts)
        seqs, _ = synth.simulate_family(n, L, mf, seed=seed, indel_rate=0.03, forbid_first=10)
        if start_m:
            for i, k in enumerate(sorted(seqs)):
                if i % 2 == 0:
                    seqs[k][0] = 10                                  # 'M'
        strs = {k: to_str(v) for k, v in seqs.items()}
        fa = "".join(">%s\n%s\n" % (k, v) for k, v in strs.items())
        trees, _ = run(flags + ["-i", "2", "-T", "--all_trees"], fa)
        ali, _ = run(flags + ["-i", "2", "-f"], fa)
        parsed = parse_fasta(ali)
        out.append(dict(n=n, L=L, seed=seed, flags=flags, seqs=strs, trees=trees.strip().splitlines(),
                        aligned=parsed, order=list(parsed)))
        print("pipeline", n, L, flags, "trees", len(out[-1]["trees"]), file=sys.stderr)
    with open(os.path.join(ROOT, "tests", "golden", "pipelines.json"), "w") as f:
        json.dump(dict(generator="tools/make_golden.py", binary="ProGraphMSA_64 (2013-07-09 rev 5e7b708)", runs=out), f, indent=1)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "pipelines":
        pipelines()
        sys.exit(0)
    main()
    pipelines()
