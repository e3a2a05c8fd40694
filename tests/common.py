"""Shared helpers for the parity tests."""
import json
import math
import os
import numpy as np

from prographmsa_b200 import alphabet, model
from prographmsa_b200.options import Options

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TABLES = np.load(os.path.join(os.path.dirname(__file__), "..", "prographmsa_b200", "data", "reference_tables.npz"))


def golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def six_digit_close(value, printed):
    """`printed` carries 6 significant digits (PhyTree::formatNewick); accept half an ulp + slack."""
    if printed == 0:
        return abs(value) < 5e-7
    ulp = 10.0 ** (math.floor(math.log10(abs(printed))) - 5)
    return abs(value - printed) <= 0.51 * ulp


def aa_codes(s):
    return alphabet.AA.encode(s)


def nw_table(alpha):
    key = {"aa": "nw_aa", "codon": "nw_codon", "dna": "nw_dna"}[alpha.name]
    idx = {"aa": 0, "codon": 1, "dna": 2}[alpha.name]
    go, ge = (int(v) for v in TABLES["nw_gaps"][idx])
    return TABLES[key].astype(np.int32), go, ge


def rows_to_codes(aligned, names):
    return np.stack([alphabet.AA.encode(aligned[n]) for n in names])
