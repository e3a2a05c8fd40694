#!/usr/bin/env python3
"""Extract the NUMERIC TABLES (not code) the hot paths consume from the reference tree.

Run once in the build container (where /root/reference is mounted); the result,
prographmsa_b200/data/reference_tables.npz, is committed so nothing at test/bench time
needs /root/reference.  The tables are published scientific constants:

  * WAG rate matrix            (ModelFactoryWag.cpp:18-418, Whelan & Goldman 2001)
  * Gonnet/"Darwin" rate matrix(ModelFactoryDarwin.cpp:10-410)
  * ECM codon rate matrix      (ModelFactoryEcm.cpp:20-3741, Kosiol et al. 2007)
  * NW score tables + gap costs(DistanceFactoryAlign.cpp:8-34, 41-234, 243-247)
  * alphabet translation tables(Alphabet.cpp:6-76)

All matrices are stored exactly as the reference holds them in memory: Eigen column-major
flat arrays ("data[i + j*n] = A(i,j)").
"""
import re, sys, os
import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/src"
OUT = os.path.join(os.path.dirname(__file__), "..", "prographmsa_b200", "data", "reference_tables.npz")

def braces_after(text, marker, start=0):
    i = text.index(marker, start)
    a = text.index("{", i)
    b = text.index("}", a)
    return text[a + 1:b], b

def numbers(s, dtype):
    toks = [t for t in re.split(r"[\s,]+", s.strip()) if t]
    return np.array([dtype(t) for t in toks])

def read(name):
    with open(os.path.join(REF, name), encoding="latin-1") as f:
        return f.read()

out = {}
for key, fn, n in (("wag", "ModelFactoryWag.cpp", 20), ("darwin", "ModelFactoryDarwin.cpp", 20),
                   ("ecm", "ModelFactoryEcm.cpp", 61)):
    body, _ = braces_after(read(fn), "static double data[]")
    v = numbers(body, float)
    assert v.size == n * n, (key, v.size)
    out[key + "_q_raw"] = v.astype(np.float64)

txt = read("DistanceFactoryAlign.cpp")
body, end = braces_after(txt, "static int data[]")
v = numbers(body, int); assert v.size == 21 * 21
out["nw_aa"] = v.astype(np.int32)
body, end = braces_after(txt, "static int data[]", end)
v = numbers(body, int); assert v.size == 62 * 62
out["nw_codon"] = v.astype(np.int32)
m = re.search(r"scoring_matrix\s*<<([^;]*);", txt)
v = numbers(m.group(1), int); assert v.size == 25
# Eigen's comma initialiser fills row-major; store column-major like the others
out["nw_dna"] = v.reshape(5, 5).T.reshape(-1).astype(np.int32)
out["nw_gaps"] = np.array([[-10, -2], [-10, -2], [-5, -2]], dtype=np.int32)  # aa, codon, dna (open, extend)

atxt = read("Alphabet.cpp")
for key, marker, cnt in (("aa_translation", "aa_translation_table[256]", 256),
                         ("dna_translation", "dna_translation_table[256]", 256),
                         ("codon_from_product", "codon_from_product_table[64]", 64),
                         ("codon_translation", "codon_translation_table[256]", 256)):
    body, _ = braces_after(atxt, marker)
    v = numbers(body, int); assert v.size == cnt, (key, v.size)
    out[key] = v.astype(np.int8)

np.savez_compressed(OUT, **out)
print("wrote", os.path.abspath(OUT), {k: v.shape for k, v in out.items()})
