#!/usr/bin/env python3
"""Golden for BASELINE.json configs[4]'s shape: a 5000-sequence x 500-aa synthetic protein family aligned by the
reference's own CPU binary under a FIXED guide tree (-t tree -m -i 0 -f), i.e. all 4,999 alignGraphs +
mergeGraphs nodes of one progressive pass.  Only the seed, the shape and a SHA-256 of the alignment are
committed (tests/golden/large.json); the family and the tree are regenerated from the seed.

  python tools/make_golden_large.py [n_seq] [length] [seed]
"""
import hashlib, json, os, subprocess, sys, tempfile, time

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from prographmsa_b200 import alphabet, model, synth
from prographmsa_b200.options import Options

BIN = os.path.join(ROOT, "oracle", "_ref", "ProGraphMSA_64")
AA = alphabet.AA_LETTERS


def family(n, L, seed):
    opts = Options(mldist_flag=True)
    mf = model.getDefault(opts)
    seqs, root = synth.simulate_family(n, L, mf, seed=seed, forbid_first=10)
    return opts, mf, seqs, root, root.newick(6) + ";"


def digest(aligned):
    """aligned: {name: gapped string}; canonical text = names in sorted order"""
    h = hashlib.sha256()
    for k in sorted(aligned):
        h.update(k.encode()); h.update(b"\n"); h.update(aligned[k].encode()); h.update(b"\n")
    return h.hexdigest()


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 500
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 50
    from make_golden import parse_fasta
    t0 = time.time()
    opts, mf, seqs, root, nwk = family(n, L, seed)
    print("family built in %.1f s" % (time.time() - t0), file=sys.stderr)
    fa = "".join(">%s\n%s\n" % (k, "".join(AA[c] for c in v)) for k, v in seqs.items())
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "in.fa"), "w").write(fa)
        open(os.path.join(d, "tree.nwk"), "w").write(nwk + "\n")
        t0 = time.time()
        r = subprocess.run([BIN, "-t", "tree.nwk", "-m", "-i", "0", "-f", "in.fa"], cwd=d, capture_output=True, text=True)
        wall = time.time() - t0
        if r.returncode != 0:
            raise RuntimeError(r.stderr[-500:])
    aligned = parse_fasta(r.stdout)
    rec = dict(n=n, L=L, seed=seed, flags=["-t", "tree.nwk", "-m", "-i", "0", "-f"], columns=len(next(iter(aligned.values()))),
               sha256=digest(aligned), reference_wall_s=round(wall, 1), fasta_sha256=hashlib.sha256(fa.encode()).hexdigest(),
               newick_sha256=hashlib.sha256(nwk.encode()).hexdigest())
    print(json.dumps(rec), file=sys.stderr)
    path = os.path.join(ROOT, "tests", "golden", "large.json")
    doc = dict(generator="tools/make_golden_large.py", binary="ProGraphMSA_64 (2013-07-09 rev 5e7b708)", runs=[])
    if os.path.exists(path):
        doc = json.load(open(path))
    doc["runs"] = [x for x in doc["runs"] if (x["n"], x["L"], x["seed"]) != (n, L, seed)] + [rec]
    json.dump(doc, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
