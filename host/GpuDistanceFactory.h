/*
 * GpuDistanceFactory.h -- reference-side shims for HP-2.  Drop this header into the reference's src/
 * (next to DistanceFactoryAlign.h), link libpgm_b200.so, and the unchanged callers (TreeNJ<A>(), which
 * asks DistanceFactory<A>::getDefault for a factory; TreeNJ.h:41-43) get the GPU path.
 *
 * It needs the reference tree + Eigen to compile, so it is NOT built in this repository (Eigen and tclap
 * are absent here); tests exercise the same C ABI through ctypes instead (prographmsa_b200/distance.py
 * mirrors these classes method by method).
 *
 * Interfaces replaced (unchanged signatures):
 *   DistanceFactory<A>::computePwDistances(sequences, order)         DistanceFactory.h:44-51
 *   DistanceFactoryAlign<A> / DistanceFactoryPrealigned<A>           DistanceFactoryAlign.h:8-27, DistanceFactoryPrealigned.h:8-19
 *   DistanceFactoryAngle<A, K>                                        DistanceFactoryAngle.h:16-131
 */
#ifndef GPUDISTANCEFACTORY_H_
#define GPUDISTANCEFACTORY_H_

#include "DistanceFactory.h"
#include "DistanceFactoryAngle.h"
#include "ModelFactory.h"
#include "main.h"
#include "debug.h"
#include "pgm_b200.h"

#include <vector>
#include <map>
#include <string>

namespace pgm_shim {

inline pgm_ctx* context() {                       // one context per process (the reference is single-threaded)
    static pgm_ctx* ctx = NULL;
    if (!ctx && pgm_create(0, &ctx) != 0) error("libpgm_b200: %s", pgm_last_error());
    return ctx;
}

template <class ALPHABET> struct MlConsts;        // DistanceFactoryML.cpp:3-32
template <> struct MlConsts<AA> { static double dist_max() { return 2.2; } static double var_max() { return 1e3; } };
#ifdef WITH_DNA
template <> struct MlConsts<DNA> { static double dist_max() { return 2.2; } static double var_max() { return 1e3; } };
#endif
#ifdef WITH_CODON
template <> struct MlConsts<Codon> { static double dist_max() { return 5.2; } static double var_max() { return 5e3; } };
#endif

/* Model::Q / Model::pi are the only model data the C ABI needs; ModelFactory keeps V, Vi, sigma
 * protected (ModelFactory.h:29-33), so they are fetched like ModelFactoryPlusF does (:42-44). */
template <class ALPHABET>
struct MlParams {
    pgm_ml_params p;
    std::vector<double> Q, pi;
    explicit MlParams(const ModelFactory<ALPHABET>* mf) {
        const int n = ALPHABET::DIM;
        Q.resize(n * n); pi.resize(n);
        if (mf) {
            Model<ALPHABET> m = mf->getModel(1.0);
            for (int i = 0; i < n; ++i) { pi[i] = m.pi(i); for (int j = 0; j < n; ++j) Q[i * n + j] = m.Q(i, j); }
        }
        p.n_states = n; p.Q = &Q[0]; p.pi = &pi[0];
        p.mldist_flag = cmdlineopts.mldist_flag; p.mldist_gap_flag = cmdlineopts.mldist_gap_flag;
        p.indel_rate = cmdlineopts.indel_rate; p.min_dist = cmdlineopts.min_dist; p.max_dist = cmdlineopts.max_dist;
        p.cutoff_dist = cmdlineopts.cutoff_dist;
        p.DIST_MAX = MlConsts<ALPHABET>::dist_max(); p.VAR_MAX = MlConsts<ALPHABET>::var_max(); p.VAR_MIN = 1e-5;
    }
};

inline DistanceMatrix to_matrix(int N, const std::vector<double>& d, const std::vector<double>& v) {
    DistanceMatrix out(N);                         // symmetric, so row- vs column-major does not matter
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) { out.distances(i, j) = d[(size_t)i * N + j]; out.variances(i, j) = v[(size_t)i * N + j]; }
    return out;
}

} // namespace pgm_shim

template <class ALPHABET>
class GpuDistanceFactoryAlign : public DistanceFactoryAlign<ALPHABET> {
public:
    GpuDistanceFactoryAlign(const ModelFactory<ALPHABET>* mf) : DistanceFactoryAlign<ALPHABET>(mf) {}
    virtual DistanceMatrix computePwDistances(const std::map<std::string, sequence_t<ALPHABET> >& sequences,
                                              const std::vector<std::string>& order) {
        const int N = (int)order.size(), n = ALPHABET::DIM;
        std::vector<uint8_t> codes; std::vector<uint32_t> off(N + 1, 0);
        for (int k = 0; k < N; ++k) {
            typename std::map<std::string, sequence_t<ALPHABET> >::const_iterator it = sequences.find(order[k]);
            assert(it != sequences.end());
            for (size_t c = 0; c < it->second.size(); ++c) {
                int v = it->second[c].value();
                if (v < 0) v = 20;                 // DistanceFactoryAlign.h:72,79
                codes.push_back((uint8_t)(v > n ? n : v));
            }
            off[k + 1] = (uint32_t)codes.size();
        }
        std::vector<int32_t> tab((n + 1) * (n + 1));
        for (int c = 0; c <= n; ++c) for (int r = 0; r <= n; ++r) tab[r + (n + 1) * c] = this->scoring_matrix(r, c);
        pgm_shim::MlParams<ALPHABET> prm(this->model_factory);
        std::vector<double> d((size_t)N * N), v((size_t)N * N);
        if (pgm_nw_ml_all_pairs(pgm_shim::context(), &prm.p, codes.empty() ? NULL : &codes[0], &off[0], N, &tab[0],
                                this->gap_open, this->gap_extend, &d[0], &v[0]) != 0)
            error("%s", pgm_last_error());         // reference: error("error while backtracking") -> abort
        return pgm_shim::to_matrix(N, d, v);
    }
};

template <class ALPHABET>
class GpuDistanceFactoryPrealigned : public DistanceFactoryPrealigned<ALPHABET> {
public:
    GpuDistanceFactoryPrealigned(const ModelFactory<ALPHABET>* mf) : DistanceFactoryPrealigned<ALPHABET>(mf) {}
    virtual DistanceMatrix computePwDistances(const std::map<std::string, sequence_t<ALPHABET> >& sequences,
                                              const std::vector<std::string>& order) {
        const int N = (int)order.size(), n = ALPHABET::DIM;
        const int L = N ? (int)sequences.find(order[0])->second.length() : 0;
        std::vector<uint8_t> rows((size_t)N * L);
        for (int k = 0; k < N; ++k) {
            const sequence_t<ALPHABET>& s = sequences.find(order[k])->second;
            assert((int)s.length() == L);
            for (int c = 0; c < L; ++c) {
                if (s[c].isGap()) rows[(size_t)k * L + c] = 254;
                else { int v = s[c].value(); rows[(size_t)k * L + c] = (uint8_t)((v < 0 || v > n) ? n : v); }
            }
        }
        pgm_shim::MlParams<ALPHABET> prm(this->model_factory);
        std::vector<double> d((size_t)N * N), v((size_t)N * N);
        if (pgm_prealigned_ml_all_pairs(pgm_shim::context(), &prm.p, rows.empty() ? NULL : &rows[0], N, L, &d[0], &v[0]) != 0)
            error("%s", pgm_last_error());
        return pgm_shim::to_matrix(N, d, v);
    }
};

/* k-mer angle distances (the default initial-tree factory).  A k-mer containing an invalid state (value() < 0 or
 * >= DIM) is not counted (DistanceFactoryAngle.h:76-91): such states are passed as code 255 (any code >= n_states). */
template <class ALPHABET, int K>
class GpuDistanceFactoryAngle : public DistanceFactoryAngle<ALPHABET, K> {
public:
    GpuDistanceFactoryAngle() : DistanceFactoryAngle<ALPHABET, K>() {}
    virtual DistanceMatrix computePwDistances(const std::map<std::string, sequence_t<ALPHABET> >& sequences,
                                              const std::vector<std::string>& order) {
        const int N = (int)order.size(), n = ALPHABET::DIM;
        std::vector<uint8_t> codes; std::vector<uint32_t> off(N + 1, 0);
        for (int k = 0; k < N; ++k) {
            const sequence_t<ALPHABET>& s = sequences.find(order[k])->second;
            for (size_t c = 0; c < s.size(); ++c) { const int v = s[c].value(); codes.push_back((uint8_t)((v < 0 || v >= n) ? 255 : v)); }
            off[k + 1] = (uint32_t)codes.size();
        }
        std::vector<double> d((size_t)N * N), v((size_t)N * N);
        if (pgm_angle_all_pairs(pgm_shim::context(), n, K, codes.empty() ? NULL : &codes[0], &off[0], N,
                                cmdlineopts.mldist_flag || cmdlineopts.mldist_gap_flag, &d[0], &v[0]) != 0)
            error("%s", pgm_last_error());
        return pgm_shim::to_matrix(N, d, v);
    }
};

/* In DistanceFactory.cpp (getDefault, :8-55) replace
 *     new DistanceFactoryAngle<A, K>()                 by  new GpuDistanceFactoryAngle<A, K>()
 *     new DistanceFactoryAlign<A>(model_factory)       by  new GpuDistanceFactoryAlign<A>(model_factory)
 *     new DistanceFactoryPrealigned<A>(model_factory)  by  new GpuDistanceFactoryPrealigned<A>(model_factory)   */
#endif
