"""Host-side mirror of the reference's substitution/indel model layer.

Reference: Model.h:9-24, ModelFactory.h:18-127, ModelFactoryWag.cpp:420-440,
ModelFactoryDarwin.cpp:412-445, ModelFactoryEcm.cpp:3743-3764, ModelFactoryCustom.h:36-71,
ModelFactoryPlusF.h:36-106.  These run once per job on the host; the kernels only consume
``Q``, ``pi`` (and a re-derived eigen-basis) -- see include/pgm_b200.h.

Eigen's EigenSolver is replaced by the symmetric route: every rate matrix the reference can
build is reversible with respect to its own stationary distribution rho (WAG/Gonnet/ECM and
+F: rho = pi; custom qmat: Q is symmetric, rho uniform -- the reference never multiplies the
exchangeabilities by the frequencies, ModelFactoryCustom.h:37-48), so
``Q = D^-1/2 (U L U^T) D^1/2`` with ``D = diag(rho)``.  Agreement with a general solver is
~1e-13 on P(t), far inside the 1e-6 / 1e-5 tolerances.
"""
from dataclasses import dataclass
import math
import os
import numpy as np

from .options import Options
from . import alphabet as _alpha

_T = np.load(os.path.join(os.path.dirname(__file__), "data", "reference_tables.npz"))


@dataclass
class Model:
    """struct Model (Model.h:9-24)."""
    M: np.ndarray
    P: np.ndarray
    Q: np.ndarray
    pi: np.ndarray
    delta: float
    epsilon: float
    distance: float
    divergence: float


def stationary(Q: np.ndarray) -> np.ndarray:
    """Left null vector of Q, normalised to sum 1 (EigenSolver(Q^T) + maxCoeff in the reference)."""
    n = Q.shape[0]
    A = np.vstack([Q.T, np.ones((1, n))])
    b = np.zeros(n + 1); b[n] = 1.0
    rho, *_ = np.linalg.lstsq(A, b, rcond=None)
    return rho


def _normalize_rate(Q: np.ndarray, freqs: np.ndarray) -> np.ndarray:
    Q = Q.copy()
    np.fill_diagonal(Q, 0.0)
    np.fill_diagonal(Q, -Q.sum(axis=1))
    Q /= -(freqs @ np.diag(Q))
    return Q


def eigen_basis(Q: np.ndarray):
    """(V, sigma, Vi) with Q = V diag(sigma) Vi via rho-symmetrisation."""
    rho = stationary(Q)
    r = np.sqrt(rho)
    S = (r[:, None] * Q) / r[None, :]
    asym = np.abs(S - S.T).max()
    if asym > 1e-12 * max(1.0, np.abs(S).max()):
        # tables printed with few digits (ECM) are reversible only to ~1e-6: use the general solver,
        # as the reference does (EigenSolver, real parts)
        lam, V = np.linalg.eig(Q)
        order = np.argsort(lam.real)
        lam, V = lam.real[order], V.real[:, order]
        return V, lam, np.linalg.inv(V)
    S = 0.5 * (S + S.T)
    lam, U = np.linalg.eigh(S)
    V = U / r[:, None]
    Vi = U.T * r[None, :]
    return V, lam, Vi


class ModelFactory:
    """ModelFactory<A> (ModelFactory.h:10-34).  Subclasses fill Q and freqs."""
    alphabet = _alpha.AA

    def __init__(self, opts: Options):
        self.opts = opts
        self.Q = None
        self.freqs = None

    def _finish(self):
        self.V, self.sigma, self.Vi = eigen_basis(self.Q)

    # ModelFactory.h:92-101
    def getEpsilon(self, distance: float) -> float:
        return self.opts.gapext_prob

    def getDelta(self, distance: float) -> float:
        return (1.0 - math.exp(-distance * self.opts.indel_rate)) / 2.0

    # ModelFactory.h:103-127
    def parseDistance(self, distance: float):
        o = self.opts
        distance = distance if 0.0 < distance else 0.0      # std::max(0.0, NaN) == 0.0
        if o.mldist_flag or o.mldist_gap_flag:
            if math.isnan(distance):
                distance = 5.2
            d = distance
            ed = math.exp(d)
            div = -0.5 * (5.0 * ed - math.sqrt(45.0 * ed * ed - 20.0 * ed)) / ed
        else:
            if math.isnan(distance):
                distance = 1.0
            if distance > 0.85:
                d = 5.2
            else:
                d = -math.log(1.0 - distance - 0.2 * distance * distance)
            div = distance
        d = max(min(d, o.max_dist), o.min_dist)
        div = max(min(div, o.max_pdist), o.min_pdist)
        return d, div

    def P_of(self, d: float) -> np.ndarray:
        return (self.V * np.exp(self.sigma * d)[None, :]) @ self.Vi

    # ModelFactory.h:48-90
    def getModel(self, distance: float, gap_distance: float = None) -> Model:
        if gap_distance is None:
            d, div = self.parseDistance(distance)
            eps, delta = self.getEpsilon(d), self.getDelta(d)
        else:
            gd, _ = self.parseDistance(gap_distance)
            eps, delta = self.getEpsilon(gd), self.getDelta(gd)
            d, div = self.parseDistance(distance)
        P = self.P_of(d)
        M = self.freqs[:, None] * P
        if self.opts.mldist_flag or self.opts.mldist_gap_flag:
            div = 1.0 - float(np.trace(M))
        return Model(M=M, P=P, Q=self.Q, pi=self.freqs, delta=delta, epsilon=eps, distance=d, divergence=div)


class _TableFactory(ModelFactory):
    key = ""
    n = 20

    def __init__(self, opts: Options):
        super().__init__(opts)
        n = self.n
        Q = _T[self.key].reshape(n, n).T.copy()        # column-major data -> Q(i,j)
        self.freqs = stationary(Q)
        self.Q = _normalize_rate(Q, self.freqs)
        self._finish()


class ModelFactoryWag(_TableFactory):
    key = "wag_q_raw"


class ModelFactoryDarwin(_TableFactory):
    """Gonnet; overrides epsilon/delta (ModelFactoryDarwin.cpp:435-445)."""
    key = "darwin_q_raw"

    def getEpsilon(self, distance):
        return 0.72508680168437312

    def getDelta(self, distance):
        return min(0.5, 0.0052820194976999644 * math.pow(distance, 0.7434))


class ModelFactoryEcm(_TableFactory):
    key = "ecm_q_raw"
    n = 61
    alphabet = _alpha.Codon


class ModelFactoryCustom(ModelFactory):
    """qmat reader (ModelFactoryCustom.h:36-71): lower-triangular exchangeabilities then DIM freqs.
    Q is the *symmetric* exchangeability matrix, rate-normalised with freqs -- as in the reference."""

    def __init__(self, opts: Options, alphabet, text: str):
        super().__init__(opts)
        self.alphabet = alphabet
        n = alphabet.DIM
        vals = [float(t) for t in text.split()]
        need = n * (n - 1) // 2 + n
        if len(vals) < need:
            raise ValueError("error reading exchangeability matrix from file")
        Q = np.zeros((n, n))
        k = 0
        for i in range(1, n):
            for j in range(i):
                v = vals[k]; k += 1
                if not (v > 0 and v < math.inf):
                    raise ValueError("negative/infinity/zero value in exchangeability matrix")
                Q[i, j] = Q[j, i] = v
        f = np.array(vals[k:k + n])
        if not np.all((f > 0) & (f < math.inf)):
            raise ValueError("negative/infinity/zero value in amino acid frequencies")
        self.freqs = f / f.sum()
        self.Q = _normalize_rate(Q, self.freqs)
        self._finish()


class ModelFactoryPlusF(ModelFactory):
    """-F frequency re-estimation (ModelFactoryPlusF.h:71-106)."""

    def __init__(self, base: ModelFactory):
        super().__init__(base.opts)
        self.base = base
        self.alphabet = base.alphabet
        m = base.getModel(1.0)
        self.Q, self.freqs = m.Q.copy(), m.pi.copy()
        self._finish()

    def getEpsilon(self, d):
        return self.base.getEpsilon(d)

    def getDelta(self, d):
        return self.base.getDelta(d)

    def estimateFreqs(self, code_seqs):
        m = self.base.getModel(1.0)
        Q, old = m.Q.copy(), m.pi.copy()
        f = old * self.opts.pseudo_count
        n = len(f)
        for codes in code_seqs:
            c = np.asarray(codes)
            c = c[c < n]
            f += np.bincount(c, minlength=n)[:n]
        self.freqs = f / f.sum()
        Q = Q * (self.freqs / old)[None, :]
        self.Q = _normalize_rate(Q, self.freqs)
        self._finish()


def getDefault(opts: Options, alphabet=_alpha.AA, code_seqs=None, custom_text: str = None) -> ModelFactory:
    """ModelFactory<A>::getDefault (ModelFactory.cpp:10-100)."""
    if custom_text is not None:
        mf = ModelFactoryCustom(opts, alphabet, custom_text)
    elif alphabet is _alpha.AA:
        mf = ModelFactoryDarwin(opts) if opts.darwin_flag else ModelFactoryWag(opts)
    elif alphabet is _alpha.Codon:
        mf = ModelFactoryEcm(opts)
    else:
        raise ValueError("custom model file necessary for DNA alignments")
    if opts.aafreqs_flag:
        pf = ModelFactoryPlusF(mf)
        pf.estimateFreqs(code_seqs or [])
        mf = pf
    return mf
