/*
 * GpuGraphAlign.h -- reference-side shim for HP-1: alignGraphs<A>(g1, g2, model) -> AlignmentResult<A>
 * (GraphAlign.h:6-23,201) served by libpgm_b200.so.  Needs the reference tree + Eigen; not built here.
 *
 * Graph<A> keeps sites / edges / repeats protected (Graph.h:27-33); the shim derives a reader class,
 * which is all that is needed to hand the three arrays to the C ABI unchanged (Eigen's RowMajor
 * SparseMatrix in compressed mode IS the CSR triple outerIndexPtr / innerIndexPtr / valuePtr).
 */
#ifndef GPUGRAPHALIGN_H_
#define GPUGRAPHALIGN_H_

#include "GraphAlign.h"
#include "pgm_b200.h"
#include <vector>

namespace pgm_shim {

pgm_ctx* context();                                 // GpuDistanceFactory.h

template <class ALPHABET>
struct GraphView : public Graph<ALPHABET> {         // derived only to read the protected members
    static pgm_graph of(const Graph<ALPHABET>& g) {
        const GraphView& v = static_cast<const GraphView&>(g);
        pgm_graph out;
        out.N = (int)v.sites.cols();
        out.sites = v.sites.data();                 // DIM x N, column-major (Model.h:13)
        out.e_rowptr = v.edges.outerIndexPtr(); out.e_col = v.edges.innerIndexPtr(); out.e_val = v.edges.valuePtr();
        const bool reps = v.repeats.nonZeros() > 0;
        out.r_rowptr = v.repeats.outerIndexPtr();
        out.r_col = reps ? v.repeats.innerIndexPtr() : NULL;
        out.r_units = reps ? (const uint32_t*)v.repeats.valuePtr() : NULL;
        return out;
    }
};

} // namespace pgm_shim

/* Batched form: all ready guide-tree nodes of one level (independent subtrees never interact before
 * their parent; ProgressiveAlignment.h:92-93 recurses serially and exploits none of this). */
template <class ALPHABET>
std::vector<AlignmentResult<ALPHABET> > alignGraphsGpuBatch(const std::vector<const Graph<ALPHABET>*>& g1,
                                                            const std::vector<const Graph<ALPHABET>*>& g2,
                                                            const std::vector<Model<ALPHABET> >& models,
                                                            int emission_mode = PGM_EMISSION_EXACT) {
    const int n = ALPHABET::DIM, jobs = (int)g1.size();
    std::vector<pgm_graph> a(jobs), b(jobs);
    std::vector<pgm_dp_model> m(jobs);
    std::vector<std::vector<double> > Mrow(jobs), pi(jobs);
    std::vector<std::vector<int32_t> > m1(jobs), m2(jobs);
    std::vector<pgm_alignment> out(jobs);
    for (int k = 0; k < jobs; ++k) {
        a[k] = pgm_shim::GraphView<ALPHABET>::of(*g1[k]);
        b[k] = pgm_shim::GraphView<ALPHABET>::of(*g2[k]);
        Mrow[k].resize(n * n); pi[k].resize(n);
        for (int i = 0; i < n; ++i) { pi[k][i] = models[k].pi(i); for (int j = 0; j < n; ++j) Mrow[k][i * n + j] = models[k].M(i, j); }
        m[k].M = &Mrow[k][0]; m[k].pi = &pi[k][0];
        m[k].distance = models[k].distance; m[k].epsilon = models[k].epsilon; m[k].delta = models[k].delta;
        const int cap = 2 * (a[k].N + b[k].N) + 8;
        m1[k].resize(cap); m2[k].resize(cap);
        out[k].mapping1 = &m1[k][0]; out[k].mapping2 = &m2[k][0]; out[k].capacity = cap;
    }
    pgm_dp_opts o;
    o.indel_rate = cmdlineopts.indel_rate; o.end_indel_prob = cmdlineopts.end_indel_prob;
    o.repeat_rate = cmdlineopts.repeat_rate; o.repeatext_prob = cmdlineopts.repeatext_prob;
    o.emission_mode = emission_mode;
    if (pgm_align_graphs_batch(pgm_shim::context(), n, jobs, jobs ? &a[0] : NULL, jobs ? &b[0] : NULL, jobs ? &m[0] : NULL, &o,
                               jobs ? &out[0] : NULL) != 0)
        error("%s", pgm_last_error());               // reference: error("backtracking failed") -> abort
    std::vector<AlignmentResult<ALPHABET> > res(jobs);
    for (int k = 0; k < jobs; ++k) {
        res[k].score = out[k].score; res[k].n_tr_indels = out[k].n_tr_indels;
        res[k].mapping1.assign(m1[k].begin(), m1[k].begin() + out[k].length);     // -1 -> (index_t)-1
        res[k].mapping2.assign(m2[k].begin(), m2[k].begin() + out[k].length);
    }
    return res;
}

/* Drop-in for the single call site style of the reference (ProgressiveAlignment.h:439, :170). */
template <class ALPHABET>
AlignmentResult<ALPHABET> alignGraphsGpu(const Graph<ALPHABET>& g1, const Graph<ALPHABET>& g2, const Model<ALPHABET>& model) {
    std::vector<const Graph<ALPHABET>*> a(1, &g1), b(1, &g2);
    std::vector<Model<ALPHABET> > m(1, model);
    return alignGraphsGpuBatch<ALPHABET>(a, b, m)[0];
}
#endif
