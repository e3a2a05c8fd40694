/*
 * pgm_b200.h -- C ABI of libpgm_b200.so: the B200 (sm_100a) implementation of ProGraphMSA's two
 * data-parallel hot paths.  Plain C, caller-allocated buffers, no exceptions; every function returns
 * 0 on success or a negative pgm_status, with a human-readable message in pgm_last_error().
 * One pgm_ctx per GPU; a ctx is not re-entrant (one call at a time), different ctxs are independent.
 *
 * The reference has no FFI layer: the boundary is its C++ interfaces, which stay unchanged
 * (SURVEY.md §8b).  Each entry point below names the reference interface it stands behind; the
 * host shims a maintainer adds on the reference side are in INTEGRATION.md and host/.
 * Every former dependency on the reference's global `cmdlineopts` (main.h:37-82) is an explicit
 * field here.
 *
 * Conventions
 *   states   uint8 codes 0..n-1 valid, n = unknown, 254 = alignment gap (Prealigned only)
 *   pairs    the i<j upper triangle of N sequences, linearised row-major:
 *            p(i,j) = i*N - i*(i+1)/2 + (j-i-1);  seq i is the reference's `seq1` (columns / count
 *            row), seq j its `seq2` (DistanceFactoryAlign.h:35-53)
 *   *_dev    same call with the large buffers already resident in device memory (what bench.py
 *            times as `value`); runs on the ctx stream and returns when the results are complete
 */
#ifndef PGM_B200_H_
#define PGM_B200_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct pgm_ctx pgm_ctx;

enum pgm_status {
    PGM_OK = 0,
    PGM_ERR_CUDA = -1,          /* CUDA runtime/driver failure (message has the CUDA string)     */
    PGM_ERR_ARG = -2,           /* bad argument                                                  */
    PGM_ERR_MODEL = -3,         /* rate matrix not reversible / eigen-solver failed              */
    PGM_ERR_BACKTRACK = -4,     /* reference: error("error while backtracking"/"backtracking failed") */
    PGM_ERR_CAPACITY = -5,      /* caller buffer too small                                       */
    PGM_ERR_NOMEM = -6
};

int pgm_create(int device, pgm_ctx** out);
void pgm_destroy(pgm_ctx* ctx);
int pgm_sync(pgm_ctx* ctx);
const char* pgm_last_error(void);
int pgm_abi_version(void);
/* scratch budget for packed traceback bits etc. (default 16 GiB); work is chunked to fit */
int pgm_set_scratch_bytes(pgm_ctx* ctx, uint64_t bytes);
/* cudaStream_t the ctx launches on (so callers can record events on it) */
void* pgm_stream(pgm_ctx* ctx);

/* per-kernel device times of the LAST call, measured with CUDA events on the ctx stream */
typedef struct {
    float nw_forward_ms, nw_walk_ms, ml_ms;          /* HP-2 */
    float emission_ms, graph_fill_ms, graph_trace_ms;/* HP-1 */
    uint32_t nw_forward_launches, nw_walk_launches, ml_launches;
    uint32_t emission_launches, graph_fill_launches, graph_trace_launches;
    uint64_t nw_cells;                               /* sum (L1+1)(L2+1) over pairs processed     */
    uint64_t ml_evaluations;                         /* Newton likelihood evaluations taken        */
    uint64_t graph_cell_updates;                     /* sum #preds(y)*#preds(x) over interior cells*/
    uint64_t graph_cells;                            /* sum N1*N2                                  */
    uint64_t tb_bytes;                               /* packed traceback bytes written to HBM      */
} pgm_timing;
int pgm_enable_timing(pgm_ctx* ctx, int on);
int pgm_get_timing(pgm_ctx* ctx, pgm_timing* out);

/* ------------------------------------------------------------------------------------------------
 * HP-2: all-pairs distances.  Stands behind
 *   DistanceFactory<A>::computePwDistances            (DistanceFactory.h:44-51)
 *   DistanceFactoryAlign<A>::computePwDistances/alignPair (DistanceFactoryAlign.h:29-127)
 *   DistanceFactoryPrealigned<A>::computePwDistances  (DistanceFactoryPrealigned.h:33-90)
 *   DistanceFactoryML<A>::computeDistance/computeMLDist (DistanceFactoryML.h:66-190)
 *   ModelFactory<A>::getModel(t) inside the Newton loop (ModelFactory.h:48-67,103-127)
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int n_states;               /* ALPHABET::DIM: 20 / 4 / 61                                     */
    const double* Q;            /* n*n row-major, Model::Q                                        */
    const double* pi;           /* n, Model::pi                                                   */
    int mldist_flag;            /* cmdlineopts.mldist_flag  (-m)                                  */
    int mldist_gap_flag;        /* cmdlineopts.mldist_gap_flag (-M)                               */
    double indel_rate;          /* cmdlineopts.indel_rate                                         */
    double min_dist, max_dist;  /* cmdlineopts.min_dist / max_dist (clamp inside getModel)        */
    double cutoff_dist;         /* cmdlineopts.cutoff_dist                                        */
    double DIST_MAX, VAR_MAX, VAR_MIN; /* DistanceFactoryML.cpp:3-32                              */
} pgm_ml_params;

/* NW + traceback counting + ML for pairs [pair_begin, pair_end).  Packed outputs, one entry per
 * pair in range.  score_tab: (n+1)*(n+1) column-major scoring_matrix(s2,s1) as the reference holds
 * it.  Optional outputs (may be NULL): counts (n*n int32 per pair, counts[s1*n+s2]), gaps,
 * nw_scores (W(L2,L1)).  Sequences are codes[offsets[k] .. offsets[k+1]).                        */
int pgm_nw_ml_pair_range(pgm_ctx* ctx, const pgm_ml_params* prm,
                         const uint8_t* codes, const uint32_t* offsets, int N,
                         const int32_t* score_tab, int gap_open, int gap_extend,
                         int64_t pair_begin, int64_t pair_end,
                         double* dist_pairs, double* var_pairs,
                         int32_t* counts, uint32_t* gaps, int32_t* nw_scores);
/* d_codes, d_dist_pairs, d_var_pairs (+ optional d_counts/d_gaps/d_nw_scores) are DEVICE pointers;
 * offsets stays a host pointer (it drives host-side scheduling).                                 */
int pgm_nw_ml_pair_range_dev(pgm_ctx* ctx, const pgm_ml_params* prm,
                             const uint8_t* d_codes, const uint32_t* offsets, int N,
                             const int32_t* score_tab, int gap_open, int gap_extend,
                             int64_t pair_begin, int64_t pair_end,
                             double* d_dist_pairs, double* d_var_pairs,
                             int32_t* d_counts, uint32_t* d_gaps, int32_t* d_nw_scores);
/* Full symmetric N*N matrices (row-major == column-major), diagonal 0: exactly what
 * DistanceFactoryAlign<A>::computePwDistances returns in DistanceMatrix.                          */
int pgm_nw_ml_all_pairs(pgm_ctx* ctx, const pgm_ml_params* prm,
                        const uint8_t* codes, const uint32_t* offsets, int N,
                        const int32_t* score_tab, int gap_open, int gap_extend,
                        double* dist, double* var);

/* Prealigned: rows is N*L_aln (row-major), 254 = gap.  The reference's hard-coded `< 20` count
 * filter (DistanceFactoryPrealigned.h:65) is applied as min(20, n).  Replaces
 * DistanceFactoryPrealigned<A>::computePwDistances (DistanceFactoryPrealigned.h:33-90): every row becomes the sorted
 * list of its residues and one thread merges the two lists of a pair (a column where both rows have a gap changes
 * nothing, :68-69); alignments of more than 65535 columns take the column scan inside the Newton kernel.
 * pair_range: packed results of the pairs [pair_begin, pair_end) in the order (0,1), (0,2), ..., (N-2,N-1);
 * all_pairs: the two symmetric N*N matrices (zero diagonal), scattered on the device.                */
int pgm_prealigned_ml_pair_range(pgm_ctx* ctx, const pgm_ml_params* prm,
                                 const uint8_t* rows, int N, int L_aln,
                                 int64_t pair_begin, int64_t pair_end,
                                 double* dist_pairs, double* var_pairs,
                                 int32_t* counts, uint32_t* gaps);
int pgm_prealigned_ml_pair_range_dev(pgm_ctx* ctx, const pgm_ml_params* prm,
                                     const uint8_t* d_rows, int N, int L_aln,
                                     int64_t pair_begin, int64_t pair_end,
                                     double* d_dist_pairs, double* d_var_pairs,
                                     int32_t* d_counts, uint32_t* d_gaps);
int pgm_prealigned_ml_all_pairs(pgm_ctx* ctx, const pgm_ml_params* prm,
                                const uint8_t* rows, int N, int L_aln, double* dist, double* var);

/* k-mer angle distances: DistanceFactoryAngle<A,K>::computePwDistances
 * (DistanceFactoryAngle.h:54-131).  K = 2 for AA / codon, 6 for DNA (DistanceFactory.cpp:14,32,50).
 * ml_flag = cmdlineopts.mldist_flag || mldist_gap_flag.  dist / var: full N*N matrices (host).            */
int pgm_angle_all_pairs(pgm_ctx* ctx, int n_states, int K, const uint8_t* codes, const uint32_t* offsets, int N,
                        int ml_flag, double* dist, double* var);

/* ---- N2: profile part of mergeGraphs (GraphAlign.h:569-620) ------------------------------
 * For every merged node: P * g of the child node(s), Hadamard product when both exist, L2 normalisation.
 * All jobs of one guide-tree level go in one call.  Edge construction stays on the host. */
typedef struct {
    int N1, N2;                 /* nodes of the two child graphs                                        */
    const double* sites1;       /* N1 * n_states, node-major (Graph::sites column = node)               */
    const double* sites2;       /* N2 * n_states                                                        */
    const double* P1;           /* n*n row-major, Model::P of the first  child's branch                 */
    const double* P2;           /* n*n row-major, Model::P of the second child's branch                 */
    int len;                    /* merged nodes                                                         */
    const int32_t* rm1;         /* per merged node: index into graph 1, or -1                           */
    const int32_t* rm2;         /* per merged node: index into graph 2, or -1                           */
    const uint8_t* matched;     /* 0: node skipped by the alignment path (uses P1, GraphAlign.h:591)    */
    double* out;                /* len * n_states                                                       */
} pgm_merge_job;
int pgm_merge_profiles_batch(pgm_ctx* ctx, int n_states, int n_jobs, const pgm_merge_job* jobs);

/* computeDistance on caller-supplied count matrices (n*n int32 each, counts[s1*n+s2]).            */
int pgm_ml_distances(pgm_ctx* ctx, const pgm_ml_params* prm, int64_t n_pairs,
                     const int32_t* counts, const uint32_t* gaps, const double* seqlen,
                     double* dist, double* var);

/* ------------------------------------------------------------------------------------------------
 * HP-1: sequence-graph DP.  Stands behind
 *   alignGraphs<A>(g1, g2, model) -> AlignmentResult   (GraphAlign.h:6-23,201-534)
 * including DynProgScores (:98-143), precomputeScores (:145-163) + ls_log_add (ls_log.h:22-59),
 * the fill (:205-260), the end-node match (:262-280) and the traceback (:283-521,165-198).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {                /* exactly Graph<A>'s members (Graph.h:31-33)                      */
    int N;                      /* nodes incl. start (0) and end (N-1)                             */
    const double* sites;        /* n*N column-major: sites[a + n*node]                             */
    const int32_t* e_rowptr;    /* N+1; row = to-node, entries ascending by from-node              */
    const int32_t* e_col;
    const float* e_val;         /* stored value min(cost,1e4)-1e4 (Graph.h:85); 0 => +inf          */
    const int32_t* r_rowptr;    /* N+1; repeat edges (may be NULL when there are none)             */
    const int32_t* r_col;
    const uint32_t* r_units;
} pgm_graph;

typedef struct {                /* the fields of Model<A> alignGraphs reads (Model.h:9-24)         */
    const double* M;            /* n*n row-major, M(i,j) = pi_i P_ij                               */
    const double* pi;           /* n                                                               */
    double distance, epsilon, delta;
} pgm_dp_model;

enum pgm_emission_mode {
    PGM_EMISSION_EXACT = 0,     /* fp32 CUDA cores, sequential-k mul/add: bit-identical to oracle  */
    PGM_EMISSION_TENSOR = 1     /* tcgen05 split-TF32 tensor-core GEMM, fp32 accumulate            */
};

typedef struct {
    double indel_rate;          /* cmdlineopts.indel_rate                                          */
    double end_indel_prob;      /* cmdlineopts.end_indel_prob                                      */
    double repeat_rate;         /* cmdlineopts.repeat_rate                                         */
    double repeatext_prob;      /* cmdlineopts.repeatext_prob                                      */
    int emission_mode;          /* pgm_emission_mode                                               */
} pgm_dp_opts;

typedef struct {                /* AlignmentResult<A> (GraphAlign.h:6-12)                          */
    float score;
    uint32_t n_tr_indels;
    int32_t* mapping1;          /* caller-allocated, capacity entries; -1 = gap                    */
    int32_t* mapping2;
    int32_t capacity;           /* >= 2*(N1+N2) is always enough                                   */
    int32_t length;             /* out                                                             */
    int32_t status;             /* out: 0 or pgm_status of this job                                */
    uint64_t cell_updates;      /* out: sum #preds(y)*#preds(x) over interior cells                */
} pgm_alignment;

/* n_jobs independent alignGraphs calls (all "ready" guide-tree nodes of one level).               */
int pgm_align_graphs_batch(pgm_ctx* ctx, int n_states, int n_jobs,
                           const pgm_graph* g1, const pgm_graph* g2, const pgm_dp_model* models,
                           const pgm_dp_opts* opts, pgm_alignment* out);
/* Stage a batch on the device once, then run it repeatedly without host<->device copies of the
 * graphs (what bench.py times as `value`).  Results are fetched with pgm_graph_batch_fetch.       */
typedef struct pgm_graph_batch pgm_graph_batch;
int pgm_graph_batch_create(pgm_ctx* ctx, int n_states, int n_jobs,
                           const pgm_graph* g1, const pgm_graph* g2, const pgm_dp_model* models,
                           const pgm_dp_opts* opts, pgm_graph_batch** out);
int pgm_graph_batch_run(pgm_ctx* ctx, pgm_graph_batch* b);            /* async on ctx stream       */
int pgm_graph_batch_fetch(pgm_ctx* ctx, pgm_graph_batch* b, pgm_alignment* out);
void pgm_graph_batch_destroy(pgm_ctx* ctx, pgm_graph_batch* b);

/* The ten transition scores (order: gap_init, gap_extend, match_init, end_match, end_gap, end_skip,
 * start_gap, start_init, repeat_init, repeat_ext) -- host arithmetic, exposed for parity tests.   */
int pgm_dynprog_scores(int n_states, const pgm_graph* g1, const pgm_graph* g2, const pgm_dp_model* model,
                       const pgm_dp_opts* opts, float out[10]);
/* precomputeScores alone: S is N1*N2 column-major (S[y + N1*x]) in host memory.                   */
int pgm_emission_scores(pgm_ctx* ctx, int n_states, const pgm_graph* g1, const pgm_graph* g2,
                        const pgm_dp_model* model, float match_init, int emission_mode, float* S);

/* ---- N3: BioNJ join loop on the GPU (csrc/nj_bionj.cu) ------------------------------------------------------------
 * Replaces the O(N^3) part of TreeNJ<A>::buildNJTree (reference TreeNJ.cpp:132-281) for N taxa: D, V are the N*N
 * row-major symmetric distance / variance matrices (diagonals are ignored).  Join k (k < N-3) merges the taxa at the
 * CURRENT indices join_i1[k] < join_i2[k] (index i2 is removed, later indices shift down, the new cluster takes i1)
 * with branch lengths join_d1[k], join_d2[k].  D_rest receives the remaining rest_dim x rest_dim (<= 3) distance
 * matrix, row-major, from which the caller places the last taxa as the reference does (TreeNJ.cpp:236-281).        */
int pgm_bionj_joins(pgm_ctx* ctx, int N, const double* D, const double* V,
                    int32_t* join_i1, int32_t* join_i2, double* join_d1, double* join_d2 /* N-3 each */,
                    double* D_rest /* min(N,3)^2 */, int* rest_dim);

/* Measurement helper (bench.py): FP64 FMA throughput of the device in TFLOP/s, measured in place -- the roofline
 * of the ML-distance kernel.                                                               */
int pgm_probe_fp64(pgm_ctx* ctx, double* tflops);

/* ---- native HOST glue around the graph DP (csrc/host_graph.cpp; no GPU work, callable without a device) -------
 * What the reference does in C++ between two alignGraphs calls, on the same CSR arrays as pgm_graph:
 * CleanedGraph (CleanedGraph.h:38-146), the node plan of mergeGraphs (GraphAlign.h:569-620) and its edge
 * construction (GraphAlign.h:622-727, Graph.h:81-100).  Outputs are caller-allocated; capacities in the comments. */
int pgm_host_clean_graph(const pgm_graph* g, double repeatext_prob,
                         int32_t* keep /* N */, int* new_N,
                         int32_t* e_rowptr /* N+1 */, int32_t* e_col /* E */, float* e_val /* E */,
                         int32_t* r_rowptr /* N+1 */, int32_t* r_col /* R */, uint32_t* r_units /* R */);
int pgm_host_merge_plan(const int64_t* mapping1, const int64_t* mapping2, int len,
                        int32_t* rm1 /* N1+N2 */, int32_t* rm2 /* N1+N2 */, uint8_t* matched /* N1+N2 */, int* out_len);
int pgm_host_merge_edges(const pgm_graph* g1, const pgm_graph* g2, int len,
                         const int32_t* rm1, const int32_t* rm2, const uint8_t* matched,
                         double support1, double support2, double altsplice_prob,
                         int32_t* e_rowptr /* len+1 */, int32_t* e_col /* E1+E2+4*len */, float* e_val, int* n_edges,
                         int32_t* r_rowptr /* len+1 */, int32_t* r_col /* R1+R2 */, uint32_t* r_units, int* n_repeats);

#ifdef __cplusplus
}
#endif
#endif /* PGM_B200_H_ */
