"""uint8 state encodings consumed by the kernels (reference: Alphabet.h:37-117, Alphabet.cpp:6-214).

Codes: ``0..DIM-1`` valid states, ``DIM`` = unknown ('X' / 'XXX' and, for AA, every
non-residue letter the reference's table maps to 20), ``GAP`` (254) = alignment gap,
``INVALID`` (255) = ``value() < 0`` in the reference.

DNA deviation (SURVEY.md §8 note 1): the reference's ``DNA::value()`` indexes the amino-acid
table (Alphabet.cpp:204-206) so plain ACGT input crashes its binary.  We implement the
*intended* 4-state alphabet with the reference's ``dna_translation_table`` order T,C,A,G;
parity runs against the binary use DNA re-lettered T,C,A,G -> A,C,D,E.
"""
import os
import numpy as np

GAP = 254
INVALID = 255

_T = np.load(os.path.join(os.path.dirname(__file__), "data", "reference_tables.npz"))
_AA_TAB = _T["aa_translation"].astype(np.int16)
_DNA_TAB = _T["dna_translation"].astype(np.int16)
_CODON_FROM_PRODUCT = _T["codon_from_product"].astype(np.int16)

AA_LETTERS = "ACDEFGHIKLMNPQRSTVWYX"
DNA_LETTERS = "TCAGX"
_GAP_CHARS = b"_-. "


class Alphabet:
    name = ""
    DIM = 0
    chars_per_state = 1

    @classmethod
    def encode(cls, s: str) -> np.ndarray:
        raise NotImplementedError

    @classmethod
    def nw_codes(cls, codes: np.ndarray) -> np.ndarray:
        """State indices as alignPair sees them: ``value() < 0`` -> hard-coded 20
        (DistanceFactoryAlign.h:72,79), for every alphabet.  Clamped to the table size."""
        c = codes.astype(np.int32).copy()
        c[codes == INVALID] = 20
        return np.minimum(c, cls.DIM).astype(np.uint8)


class AA(Alphabet):
    name = "aa"
    DIM = 20

    @classmethod
    def encode(cls, s):
        b = np.frombuffer(s.encode("latin-1"), dtype=np.uint8)
        v = _AA_TAB[b]
        out = np.where(v < 0, INVALID, v).astype(np.uint8)
        out[np.isin(b, np.frombuffer(_GAP_CHARS, dtype=np.uint8))] = GAP
        return out

    @classmethod
    def decode(cls, codes):
        return "".join("-" if c == GAP else (AA_LETTERS[c] if c <= 20 else "?") for c in codes)


class DNA(Alphabet):
    name = "dna"
    DIM = 4

    @classmethod
    def encode(cls, s):
        b = np.frombuffer(s.encode("latin-1"), dtype=np.uint8)
        v = _DNA_TAB[b]
        out = np.where(v < 0, INVALID, v).astype(np.uint8)
        out[np.isin(b, np.frombuffer(_GAP_CHARS, dtype=np.uint8))] = GAP
        return out

    @classmethod
    def decode(cls, codes):
        return "".join("-" if c == GAP else (DNA_LETTERS[c] if c <= 4 else "?") for c in codes)


class Codon(Alphabet):
    name = "codon"
    DIM = 61
    chars_per_state = 3

    @classmethod
    def encode(cls, s):
        """Codon::Codon(c1,c2,c3) (Alphabet.cpp:121-156) + sequenceFromString<Codon> (:262-276)."""
        b = np.frombuffer(s.encode("latin-1"), dtype=np.uint8)
        n = len(b) // 3
        out = np.empty(n + (1 if len(b) % 3 else 0), dtype=np.uint8)
        gaps = np.frombuffer(_GAP_CHARS, dtype=np.uint8)
        for i in range(n):
            c1, c2, c3 = b[3 * i], b[3 * i + 1], b[3 * i + 2]
            if c1 in gaps or c2 in gaps or c3 in gaps:
                out[i] = GAP
                continue
            code = None
            c = int(_DNA_TAB[c3])
            if c < 0: code = INVALID
            elif c >= 4: code = cls.DIM
            if code is None:
                c += 4 * int(_DNA_TAB[c2])
                if c < 0: code = INVALID
                elif c >= 16: code = cls.DIM
            if code is None:
                c += 16 * int(_DNA_TAB[c1])
                if c < 0: code = INVALID
                elif c >= 64: code = cls.DIM
            if code is None:
                p = int(_CODON_FROM_PRODUCT[c])
                code = INVALID if p < 0 else p
            out[i] = code
        if len(b) % 3:
            out[n] = INVALID
        return out


def by_name(name: str):
    return {"aa": AA, "dna": DNA, "codon": Codon}[name]
