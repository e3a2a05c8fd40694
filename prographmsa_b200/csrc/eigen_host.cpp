// eigen_host.cpp -- host-side eigen-basis of a reversible rate matrix.
//
// The reference's ModelFactory keeps V, Vi, sigma protected (ModelFactory.h:29-33); a shim that only
// sees Model{Q, pi} must re-derive them.  Every Q the reference can build is
// reversible w.r.t. its OWN stationary distribution rho (not necessarily Model::pi: the built-in WAG /
// Gonnet / ECM tables end up with a uniform pi and a row-renormalised Q, ModelFactoryWag.cpp:422-435,
// and custom models keep the symmetric exchangeabilities, ModelFactoryCustom.h:37-48), so
//   S = D^1/2 Q D^-1/2 (D = diag(rho)) is symmetric,  S = U L U^T (cyclic Jacobi),
//   V = D^-1/2 U,  Vi = U^T D^1/2,  sigma = diag(L).
// Tables printed with few digits (ECM: 6) are reversible only to ~1e-6, and symmetrising them would
// move P(t) by about that much -- too close to the 1e-6 distance tolerance.  So the symmetric
// solution is only the starting point: B = U^T S U = L + (small) is driven to diagonal form by Newton
// steps on the invariant-subspace equation, X = I + Z with Z_jk = B_jk / (B_kk - B_jj), B <- X^-1 B X,
// which converges quadratically and reproduces the exact (non-symmetric) Q to rounding error.
#include "internal.h"
#include <cmath>
#include <vector>

static bool stationary(const double* Q, int n, std::vector<double>& rho) {
    // solve [Q^T ; 1..1] rho = [0 ; 1] in the least-squares sense via normal equations with pivoting
    // (the system is consistent with a one-dimensional null space of Q^T)
    std::vector<double> A((size_t)n * n), b(n, 0.0);
    // replace the last equation of Q^T rho = 0 by sum(rho) = 1
    for (int r = 0; r < n; ++r)
        for (int c = 0; c < n; ++c) A[(size_t)r * n + c] = Q[(size_t)c * n + r];
    // choose the row to replace: the one whose removal keeps the system best conditioned is not
    // critical for these matrices; use the last
    for (int c = 0; c < n; ++c) A[(size_t)(n - 1) * n + c] = 1.0;
    b[n - 1] = 1.0;
    for (int col = 0; col < n; ++col) {
        int piv = col;
        for (int r = col + 1; r < n; ++r)
            if (std::fabs(A[(size_t)r * n + col]) > std::fabs(A[(size_t)piv * n + col])) piv = r;
        if (std::fabs(A[(size_t)piv * n + col]) < 1e-300) return false;
        if (piv != col) {
            for (int c = 0; c < n; ++c) std::swap(A[(size_t)piv * n + c], A[(size_t)col * n + c]);
            std::swap(b[piv], b[col]);
        }
        for (int r = col + 1; r < n; ++r) {
            const double f = A[(size_t)r * n + col] / A[(size_t)col * n + col];
            if (f == 0) continue;
            for (int c = col; c < n; ++c) A[(size_t)r * n + c] -= f * A[(size_t)col * n + c];
            b[r] -= f * b[col];
        }
    }
    rho.assign(n, 0.0);
    for (int r = n - 1; r >= 0; --r) {
        double s = b[r];
        for (int c = r + 1; c < n; ++c) s -= A[(size_t)r * n + c] * rho[c];
        rho[r] = s / A[(size_t)r * n + r];
    }
    for (int i = 0; i < n; ++i) if (!(rho[i] > 0)) return false;
    return true;
}

int pgm_eigen_basis(const double* Q, int n, double* V, double* sigma, double* Vi) {
    std::vector<double> rho;
    if (!stationary(Q, n, rho)) { pgm_set_error("rate matrix has no positive stationary distribution"); return PGM_ERR_MODEL; }
    std::vector<double> r(n), S((size_t)n * n), U((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i) r[i] = std::sqrt(rho[i]);
    double amax = 0, asym = 0;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) { S[(size_t)i * n + j] = r[i] * Q[(size_t)i * n + j] / r[j]; amax = std::fmax(amax, std::fabs(S[(size_t)i * n + j])); }
    const std::vector<double> Sfull = S;
    for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j) {
            asym = std::fmax(asym, std::fabs(S[(size_t)i * n + j] - S[(size_t)j * n + i]));
            const double m = 0.5 * (S[(size_t)i * n + j] + S[(size_t)j * n + i]);
            S[(size_t)i * n + j] = S[(size_t)j * n + i] = m;
        }
    if (asym > 1e-3 * std::fmax(1.0, amax)) {
        pgm_set_error("rate matrix is too far from reversible w.r.t. its stationary distribution (asymmetry %g)", asym);
        return PGM_ERR_MODEL;
    }
    for (int i = 0; i < n; ++i) U[(size_t)i * n + i] = 1.0;
    // cyclic Jacobi
    for (int sweep = 0; sweep < 100; ++sweep) {
        double off = 0;
        for (int i = 0; i < n; ++i)
            for (int j = i + 1; j < n; ++j) off += S[(size_t)i * n + j] * S[(size_t)i * n + j];
        if (off < 1e-30 * std::fmax(1.0, amax * amax)) break;
        for (int p = 0; p < n; ++p)
            for (int q = p + 1; q < n; ++q) {
                const double apq = S[(size_t)p * n + q];
                if (std::fabs(apq) < 1e-300) continue;
                const double app = S[(size_t)p * n + p], aqq = S[(size_t)q * n + q];
                const double theta = (aqq - app) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; ++k) {
                    const double akp = S[(size_t)k * n + p], akq = S[(size_t)k * n + q];
                    S[(size_t)k * n + p] = c * akp - s * akq;
                    S[(size_t)k * n + q] = s * akp + c * akq;
                }
                for (int k = 0; k < n; ++k) {
                    const double apk = S[(size_t)p * n + k], aqk = S[(size_t)q * n + k];
                    S[(size_t)p * n + k] = c * apk - s * aqk;
                    S[(size_t)q * n + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < n; ++k) {
                    const double ukp = U[(size_t)k * n + p], ukq = U[(size_t)k * n + q];
                    U[(size_t)k * n + p] = c * ukp - s * ukq;
                    U[(size_t)k * n + q] = s * ukp + c * ukq;
                }
            }
    }
    // T: current (right) eigenvector matrix of Sfull, Ti its inverse
    std::vector<double> T = U, Ti((size_t)n * n);
    for (int i = 0; i < n; ++i) for (int k = 0; k < n; ++k) Ti[(size_t)k * n + i] = U[(size_t)i * n + k];
    for (int k = 0; k < n; ++k) sigma[k] = S[(size_t)k * n + k];
    if (asym > 1e-14 * std::fmax(1.0, amax)) {
        std::vector<double> B((size_t)n * n), tmp((size_t)n * n), X((size_t)n * n);
        auto mm = [n](const std::vector<double>& A_, const std::vector<double>& B_, std::vector<double>& C_) {
            for (int i = 0; i < n; ++i)
                for (int j = 0; j < n; ++j) {
                    double s = 0;
                    for (int k = 0; k < n; ++k) s += A_[(size_t)i * n + k] * B_[(size_t)k * n + j];
                    C_[(size_t)i * n + j] = s;
                }
        };
        auto invert = [n](const std::vector<double>& A_, std::vector<double>& Inv) -> bool {     // Gauss-Jordan, partial pivoting
            std::vector<double> M(A_);
            Inv.assign((size_t)n * n, 0.0);
            for (int i = 0; i < n; ++i) Inv[(size_t)i * n + i] = 1.0;
            for (int col = 0; col < n; ++col) {
                int piv = col;
                for (int rr = col + 1; rr < n; ++rr) if (std::fabs(M[(size_t)rr * n + col]) > std::fabs(M[(size_t)piv * n + col])) piv = rr;
                if (std::fabs(M[(size_t)piv * n + col]) < 1e-300) return false;
                if (piv != col) for (int c2 = 0; c2 < n; ++c2) { std::swap(M[(size_t)piv * n + c2], M[(size_t)col * n + c2]); std::swap(Inv[(size_t)piv * n + c2], Inv[(size_t)col * n + c2]); }
                const double d = 1.0 / M[(size_t)col * n + col];
                for (int c2 = 0; c2 < n; ++c2) { M[(size_t)col * n + c2] *= d; Inv[(size_t)col * n + c2] *= d; }
                for (int rr = 0; rr < n; ++rr) {
                    if (rr == col) continue;
                    const double f = M[(size_t)rr * n + col];
                    if (f == 0) continue;
                    for (int c2 = 0; c2 < n; ++c2) { M[(size_t)rr * n + c2] -= f * M[(size_t)col * n + c2]; Inv[(size_t)rr * n + c2] -= f * Inv[(size_t)col * n + c2]; }
                }
            }
            return true;
        };
        bool converged = false;
        for (int it = 0; it < 30 && !converged; ++it) {
            mm(Ti, Sfull, tmp); mm(tmp, T, B);                     // B = T^-1 Sfull T
            double off = 0;
            for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) if (i != j) off = std::fmax(off, std::fabs(B[(size_t)i * n + j]));
            for (int k = 0; k < n; ++k) sigma[k] = B[(size_t)k * n + k];
            if (off < 4e-16 * std::fmax(1.0, amax) * n) { converged = true; break; }
            std::fill(X.begin(), X.end(), 0.0);
            for (int i = 0; i < n; ++i) X[(size_t)i * n + i] = 1.0;
            for (int j = 0; j < n; ++j)
                for (int k = 0; k < n; ++k) {
                    if (j == k) continue;
                    const double gap = B[(size_t)k * n + k] - B[(size_t)j * n + j];
                    if (std::fabs(gap) > 1e-9 * std::fmax(1.0, amax)) X[(size_t)j * n + k] = B[(size_t)j * n + k] / gap;
                }
            mm(T, X, tmp); T = tmp;
            if (!invert(T, Ti)) { pgm_set_error("eigenvector matrix became singular"); return PGM_ERR_MODEL; }
            if (it >= 6) {                                           // stagnation (degenerate + non-reversible): accept if small
                if (off < 1e-12 * std::fmax(1.0, amax)) converged = true;
            }
        }
        if (!converged) { pgm_set_error("eigen-basis refinement did not converge"); return PGM_ERR_MODEL; }
    }
    for (int i = 0; i < n; ++i)
        for (int k = 0; k < n; ++k) {
            V[(size_t)i * n + k] = T[(size_t)i * n + k] / r[i];
            Vi[(size_t)k * n + i] = Ti[(size_t)k * n + i] * r[i];
        }
    return 0;
}

// exported for the CPU-side unit test of the eigen-basis (not part of the hot path)
extern "C" int pgm_debug_eigen_basis(const double* Q, int n, double* V, double* sigma, double* Vi) {
    return pgm_eigen_basis(Q, n, V, sigma, Vi);
}
