// internal.h -- shared between the C-ABI glue (api.cu) and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/pgm_b200.h"

void pgm_set_error(const char* fmt, ...);

#define PGM_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess) {                                                                \
            pgm_set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(e_), __FILE__, __LINE__, \
                          cudaGetErrorString(e_));                                              \
            return PGM_ERR_CUDA;                                                                \
        }                                                                                       \
    } while (0)

// ---- eigen_host.cpp ---------------------------------------------------------------------------
// Q (n*n row-major) -> V, sigma, Vi with Q = V diag(sigma) Vi (rho-symmetrised Jacobi). 0 or PGM_ERR_MODEL.
int pgm_eigen_basis(const double* Q, int n, double* V, double* sigma, double* Vi);

// ---- growable device buffer ---------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            pgm_set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
            cudaGetLastError();
            return PGM_ERR_NOMEM;
        }
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return (T*)p; }
};

struct HostPinned {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        if (cudaMallocHost(&p, want) != cudaSuccess) { cudaGetLastError(); pgm_set_error("cudaMallocHost(%zu) failed", want); return PGM_ERR_NOMEM; }
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return (T*)p; }
};

// ---- HP-2 device structs ------------------------------------------------------------------------
struct NwTask {             // 32 B: the pairs (i, j0 .. j1-1); one CTA builds seq i's profile once
    uint32_t i, j0, j1;
    uint32_t out_base;          // packed-output index of pair (i, j0)
    uint64_t tb_base;           // word offset of pair (i, j0)'s packed traceback in the chunk scratch
    uint32_t list_base;         // offset of pair (i, j0)'s matched-state list (uint16 entries)
    uint32_t C;                 // columns per lane of the forward kernel that wrote this task's traceback
};
#define NW_MODE_U16 0x100u      // NwTask::C flag: traceback written by the 16-bit two-pairs-per-warp flavour
#define NW_TWIN_A   0x200u      // ... and the NEXT task (same j1, row i') is aligned by this task's CTA, in the high halves
#define NW_TWIN_B   0x400u      // ... this task is the second of such a twin: the forward kernel skips its CTA
#define NW_TASK_PAIRS 32        // pairs per task (4 per warp of the forward kernel, 1 per lane of the walk)

// Everything else about a pair is derived on the device from the offsets array (a prefix sum of the
// lengths), so no per-pair descriptor is ever built or copied.  With span = off[j] - off[j0]:
//   32-bit flavour: tb_off(i,j) = tb_base + nb(L1) * 32 * (span + 31 * (j - j0)),   word (band, step t, lane) at
//                   ((band * (L2 + 31) + t) * 32 + lane
//   16-bit flavour: tb_off(i,j) = tb_base + 32 * roundup8(span + 46 * (j - j0)),    word (step t, lane) at
//                   ((t >> 3) * 32 + lane) * 8 + (t & 7): eight consecutive steps of a lane share one 32-byte
//                   sector, which is what the walk (mostly diagonal: same lane, t - 1) wants to read
//   list_off(i,j) = list_base + span                             (capacity L2 >= matched pairs)
#define NW_U16_STEP_SLACK 46u   // per-pair step budget L2 + 46 >= roundup8(L2 + 31) + 7 (alignment of the start)
struct NwPair { uint32_t s1_off, s2_off, L1, L2, out_idx, list_off; uint64_t tb_off; };
__host__ __device__ __forceinline__ uint64_t nw_tb_words(bool u16, uint32_t nb, uint64_t span, uint64_t cnt) {
    return u16 ? 32ull * (span + NW_U16_STEP_SLACK * cnt + 8u) : (uint64_t)nb * 32u * (span + 31u * cnt);
}
__device__ __forceinline__ NwPair nw_pair(const NwTask& t, uint32_t j, const uint32_t* __restrict__ off) {
    const uint32_t C = t.C & 0xffu;
    NwPair p;
    p.s1_off = off[t.i]; p.L1 = off[t.i + 1] - p.s1_off;
    p.s2_off = off[j]; p.L2 = off[j + 1] - p.s2_off;
    const uint32_t nb = (p.L1 + 32u * C - 1u) / (32u * C);
    const uint32_t span = (p.s2_off - off[t.j0]);
    if (t.C & NW_MODE_U16) p.tb_off = t.tb_base + 32ull * (((uint64_t)span + NW_U16_STEP_SLACK * (j - t.j0) + 7u) & ~7ull);
    else p.tb_off = t.tb_base + (uint64_t)nb * 32u * ((uint64_t)span + 31u * (j - t.j0));
    p.list_off = t.list_base + span;
    p.out_idx = t.out_base + (j - t.j0);
    return p;
}

struct MlDev {              // ML constants + eigen-basis in device memory
    int n;
    const double *V, *Vi, *sigma;   // n*n row-major, n*n row-major, n
    int mldist_flag, mldist_gap_flag;
    int symmetric;                  // V(i,k) Vi(k,j) = c_ij V(j,k) Vi(k,i) for all k (reversible Q): (i,j) and (j,i) counts can be merged
    double indel_rate, min_dist, max_dist, cutoff_dist, DIST_MAX, VAR_MAX, VAR_MIN;
};

// kernels (hp2_nw.cu / hp2_ml.cu); all launch on `st`
int launch_nw_forward(cudaStream_t st, int n_tasks, int n_states, int max_bands, const uint8_t* d_codes,
                      const uint32_t* d_off, const NwTask* d_tasks, const int8_t* d_tab, int go, int ge,
                      uint32_t* d_tb, int32_t* d_colbuf, uint32_t colbuf_stride, int32_t* d_scores, int u16_h, int u16_bias,
                      int u16_max_rows);
int launch_nw_walk(cudaStream_t st, int n_tasks, int n_states, const uint8_t* d_codes, const uint32_t* d_off,
                   const NwTask* d_tasks, const uint32_t* d_tb, uint16_t* d_list, uint32_t* d_list_len, uint32_t* d_gaps,
                   int* d_err, int lanes_per_warp);
// front-end LIST: histogram from matched-state lists; ROWS: scan two MSA rows
int launch_ml_from_lists(cudaStream_t st, const MlDev& ml, int n_tasks, const uint32_t* d_off, const NwTask* d_tasks,
                         const uint16_t* d_list, const uint32_t* d_list_len, const uint32_t* d_gaps, double* d_dist,
                         double* d_var, int32_t* d_counts, unsigned long long* d_evals, int task_pairs);
size_t prealigned_pitch(int L);
int launch_prealigned_repitch(cudaStream_t st, const uint8_t* d_rows, int N, int L, uint8_t* d_pitched);
int launch_ml_from_rows_split(cudaStream_t st, const MlDev& ml, int N, int L, int64_t pair_begin, int64_t n_pairs, const uint8_t* d_pitched,
                              uint16_t* d_counts16, uint32_t* d_gaps16, double* d_dist, double* d_var, int32_t* d_counts, uint32_t* d_gaps,
                              unsigned long long* d_evals);
int launch_ml_from_rows(cudaStream_t st, const MlDev& ml, int N, int L, int64_t pair_begin, int64_t n_pairs,
                        const uint8_t* d_rows, double* d_dist, double* d_var, int32_t* d_counts, uint32_t* d_gaps,
                        unsigned long long* d_evals);
int launch_ml_from_counts(cudaStream_t st, const MlDev& ml, int64_t n_pairs, const int32_t* d_counts, const uint32_t* d_gaps,
                          const double* d_seqlen, double* d_dist, double* d_var, unsigned long long* d_evals);

int launch_angle(cudaStream_t st, int n, int K, int dims, int N, const uint8_t* d_codes, const uint32_t* d_off,
                 uint16_t* d_counts, double* d_norm2, int ml_flag, double* d_dist, double* d_var);

// ---- HP-1 (hp1_*.cu) ------------------------------------------------------------------------------
struct NodeInfo { int ptr; int cnt; int idx0; float val0; };   // per node: pred list start, packed count word, first entry inline
// NodeInfo::cnt packs: bits 0..7 number of predecessors, bits 8..23 (index + 1) of the node's row / column in the value store
// (0 = its cells are not kept as values), bit 31 = the first predecessor is a repeat edge
#define NI_CNT(c) ((c) & 0xff)
#define NI_KEEP(c) ((int)(((unsigned)(c) >> 8) & 0xffffu) - 1)
#define NI_CLASS(c) ((int)(((unsigned)(c) >> 24) & 0x7fu))          // bits 24..30: site class of the node (class-coded graphs, see GraphJobDev::coded1)
#define NI_CHAIN(ni, node) ((((ni).cnt & 0x800000ff) == 1) && (ni).idx0 == (node) - 1)   // exactly one predecessor: the regular edge from node-1
// general strip kernel: everything a row needs in one 32-byte record, fetched two rows ahead (no dependent loads in a step)
struct RowInfo { int cnt; int idx0; float val0; int idx1; float val1; int kq01; int ncidx; int ptr; };   // cnt as NodeInfo::cnt; the first two
                             // predecessors inline; kq01 = (value-store row of idx0 + 1) | (... of idx1 + 1) << 16; ncidx = row of RowIdx or -1
// role-sorted anti-diagonal kernel (hp1_wave.cu): one record per row thread / per non-chain column, one per column for the row threads
struct WaveRow { int y; int nextra; float ey_adj; int yp0; float ey0; int yp1; float ey1; int rest; };   // y = -1: padding slot; ey_adj = +inf: no
                             // edge from y-1; extra predecessor k (every list entry but the adjacent one) inline for k < 2, else at wv_xi/wv_xv[rest + k]
struct WaveCol { int x; int cnt; int xp0; float ex0; int xp1; float ex1; int ptr; float ex_adj_col; };   // all predecessors of the column, k >= 2 at p2_idx/p2_val[ptr + k];
                             // ex_adj_col = value of the edge x-1 -> x, +inf if there is none
struct WaveCI { float ex_adj; int flags; };                  // flags bit 0: non-chain column (its cells belong to the column threads), bit 1: column x-1 is one
#define GF_MAX_PREDS 128     // predecessors per node the packed traceback's index fields can hold (graphs beyond it take the dense kernels)

#define GF_BAND_ROWS 128      // rows per band (= threads per CTA) of the banded anti-diagonal kernel (<= 512: its register budget)
#define GF_BAND_MIN_N1 700     // graph 1 from this many nodes on is swept in bands.  Measured on the upper levels (<= 12 jobs) of a
                              // 5000-sequence pass (tools/levels_knobs.py), fill time per pass: 1026 / 512 rows 261 ms, 1026 / 128 rows
                              // 237 ms, 600 / 128 rows 232 ms (graphs of ~600 rows with few non-chain rows lose 5-10 % in bands)
#define GF_DIAG_REGULAR_ROWS 1024   // anti-diagonal kernel: up to this many rows one row per thread and the 8-diagonal ring
#define GF_LAYOUT_BLOCK 0
#define GF_LAYOUT_DIAG 1
#define GF_LAYOUT_STRIP 2    // hp1_strip.cu: warp per job, a lane owns a strip of <= C columns in registers, 4 direction bits per cell
#define GF_STRIP_MAX_C 24    // most columns per strip (register arrays)
#define GF_K 8               // columns a fill-kernel thread sweeps per step (one block of a row)
#define GF_LOGK 3
#define GF_MAX_NB 512        // most column blocks per row the fill kernel's shared memory holds (graph 2 up to ~4000 nodes)

struct GraphJobDev {        // one alignGraphs job, all pointers into the batch's device arena
    int N1, N2;
    int bands;                  // > 1: graph 1 is swept in bands of band_rows rows (one CTA each) by the banded anti-diagonal kernel
    int band_rows;
    int* progress;              // bands + 1 ints (scratch, zeroed before every sweep): anti-diagonals completed per band, then the band ticket
    int layout;                 // cell layout of D = which fill kernel sweeps the job: GF_LAYOUT_BLOCK or GF_LAYOUT_DIAG
    int nb;                     // column blocks per row (columns 0 .. N2-2 cut into blocks of <= GF_K; a non-chain node starts a block)
    int pitch;                  // block-skewed layout (hp1_graphdp.cu): cell (y,x) at (((y + blk(x))*GF_K + off(x))*pitch + y%nb;
                                // diagonal layout (hp1_graphdp_diag.cu): (y+x)*pitch + y - max(0, y+x-(N2-1))
    const uint16_t* colinfo;    // N2 entries: (blk(x) << GF_LOGK) | off(x)
    const uint16_t* blkstart;   // nb + 1 entries: first column of each block, blkstart[nb] = N2 - 1
    // Emission inputs.  A graph whose every interior profile column is one-hot or uniform (a leaf: SequenceGraph.h:44-54)
    // is CLASS-CODED: the emission runs on its n + 1 distinct columns only (N1e / N2e = n + 3 pseudo-nodes) and the fill
    // reads S(class(y), class(x)) -- bit-identical, because S(y,x) depends on the two profile columns only.  Only jobs
    // swept by the strip kernel are coded (the other fill kernels stream full rows of S).
    int N1e, N2e;               // nodes the emission kernels see (= N1, N2 unless coded)
    int coded1, coded2;         // graph 1 / graph 2 class-coded: S row of node y = NI_CLASS(ni1[y].cnt), S column of node x = scol[x]
    const uint16_t* scol;       // N2 entries: S column of graph-2 node x (x itself unless coded2)
    const float* sites1;        // n * N1e (float cast of Graph::sites), column-major
    const float* sites2;        // n * N2e
    // strip layout (hp1_strip.cu): strip j = columns [sstart[j], sstart[j+1]) of graph 2, at most C of them; colstrip[x] = strip * 32 + slot
    int nstrips, C;
    const int* sstart; const uint32_t* colstrip;
    // general strip kernel (graphs with non-chain nodes): per strip the slots whose column is not a chain node (specmask) and
    // the slots whose column is the SOURCE of a non-adjacent column edge (srcmask; its (X,W) go to a shared-memory ring,
    // slot colsrc[x]); direction-word index fields of the cells in non-chain rows / columns (ncidx = row in RowIdx / ColIdx)
    // role-sorted anti-diagonal kernel (hp1_wave.cu; layout = GF_LAYOUT_DIAG)
    int wave;                   // 1: swept by graph_fill_wave_kernel
    int wv_tr, wv_ncols, wv_nx; // row-thread slots (chain rows, then non-chain rows, each padded to warps), non-chain columns, extra-predecessor entries
    int wv_chain_pad;           // slots of the chain rows (a multiple of 32): the non-chain rows' slots start here
    int wv_tcols;               // lane groups of the column threads: group g sweeps the non-chain columns g, g + wv_tcols, ... in turn
    const float* wv_adj1;       // N1 entries: value of the edge y-1 -> y, +inf if there is none
    const WaveRow* wv_rows; const WaveCol* wv_cols; const WaveCI* wv_ci; const uint16_t* wv_rowslot; const int* wv_xi; const float* wv_xv;
    int sgen;                   // 1: swept by graph_fill_strip_gen_kernel
    int dbg;                    // PGM_STRIP_DEBUG bits (experiments): 1 = never the fused fast path, 2 = fence after the step barrier
    int nsrc, ring_rows;        // source columns, rows per ring (power of two > nstrips + 1)
    int hist_rows;              // rows of the (Y,W) history ring (power of two >= 2)
    const uint32_t* specmask; const uint32_t* srcmask; const int16_t* colsrc;
    const int* ncidx1; const int* ncidx2;
    const RowInfo* rinfo;       // N1 records (general strip kernel)
    void* RowIdx; void* ColIdx; int ldRI, ldCI;
    const float* Mf;            // n*n row-major float cast of Model::M
    const float* pif;           // n
    // unified predecessor lists (regular edges ascending, then repeat edges ascending; Graph.h:196,241-247)
    const int* p1_ptr; const int* p1_idx; const float* p1_val; const uint8_t* p1_rep;
    const int* p2_ptr; const int* p2_idx; const float* p2_val; const uint8_t* p2_rep;
    // alternative-path preds (values with repeat_init = repeat_ext = +inf; GraphAlign.h:165-198)
    const float* p1_alt; const float* p2_alt;
    const NodeInfo* ni1; const NodeInfo* ni2;   // N1 / N2 entries
    int np1, np2;                               // total predecessor-list entries of graph 1 / graph 2
    float sc[10];               // DynProgScores
    float* S;                   // row-major by graph-1 node: S(y,x) at y*ldS + x (a DP thread streams its own row)
    int ldS;                    // N2 rounded up to a multiple of 4
    // Blocked / anti-diagonal kernels (latency-bound: the dense matrices cost them nothing, the first-maximum bookkeeping of
    // a packed traceback 25-70 % of their time, profiles/): {M,X,Y,W} of every cell, one float4, re-traced by the reference's
    // closest-|delta| search.
    float4* MXYW;               // block-skewed: (N1-1 + nb-1) * GF_K * pitch cells; diagonal: (N1+N2-1) * pitch
    // Strip kernels: PACKED traceback, 4 direction bits per cell (hp1_dir.cuh, hp1_strip.cu).  The DP values themselves live
    // only in registers / shared-memory rings, except for the rows of graph 1 and columns of graph 2 that a far predecessor
    // edge, the end-node match or the traceback of the boundary row / column reads back: those are kept as {M,X,Y,W} in the
    // value store.
    void* D;                    // strip j of row y at uint32 words ((y + j) * nstrips + j) * W32
    int dbits;                  // bits per predecessor index of the side arrays' words (7)
    int ldVr, ldVc;             // pitches of the value store
    float4* Vrow;               // kept row r (NI_KEEP of its graph-1 node), column x at r*ldVr + x
    float4* Vcol;               // kept column c (NI_KEEP of its graph-2 node), row y at c*ldVc + y
    float* Tbuf;                // n * N2 scratch: T = M^T g2
    float* Aimg; float* Bimg;   // tensor flavour: per-128-row-tile UMMA operand images (ceil(N/128) * 128 * KP floats)
    float* u; float* v;         // g1^T pi (N1), pi^T g2 (N2)
    int* map1; int* map2;       // output mappings (capacity cap), written reversed then flipped
    int cap;
    float* alt_score; int* alt_prev;   // markAlternativePath scratch, max(N1,N2) each
    int* result;                // [0]=status [1]=length [2]=n_tr_indels, score at ((float*)result)[3], [4]=emission status,
                                // [5]=end-node pick: state (0 M, 1 X, 2 Y, 3 end_skip) | ky << 8 | kx << 16 (set by the fill)
};

#ifdef __CUDACC__
// g^T pi in the order of Eigen's SSE2 row-major gemv kernel (what the reference binary does for the two normalisers of
// GraphAlign.h:154; pinned against the binary on a 1000-sequence family): four lane sums, (l0 + l2) + (l1 + l3), sequential tail
__device__ __forceinline__ float dot_gemv_order(const float* __restrict__ g, const float* __restrict__ pi, int n) {
    float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
    const int n4 = n & ~3;
    for (int k = 0; k < n4; k += 4) {
        l0 = __fadd_rn(l0, __fmul_rn(g[k], pi[k])); l1 = __fadd_rn(l1, __fmul_rn(g[k + 1], pi[k + 1]));
        l2 = __fadd_rn(l2, __fmul_rn(g[k + 2], pi[k + 2])); l3 = __fadd_rn(l3, __fmul_rn(g[k + 3], pi[k + 3]));
    }
    float s = __fadd_rn(__fadd_rn(l0, l2), __fadd_rn(l1, l3));
    for (int k = n4; k < n; ++k) s = __fadd_rn(s, __fmul_rn(g[k], pi[k]));
    return s;
}
#endif

int launch_emission_exact(cudaStream_t st, int n_states, int n_jobs, const GraphJobDev* d_jobs, const int* d_order, int max_N1, int max_N2);
int launch_emission_tensor(cudaStream_t st, int n_states, int n_jobs, const GraphJobDev* d_jobs, const int* d_order, int max_N1, int max_N2);
int launch_graph_trace(cudaStream_t st, int n_jobs, const GraphJobDev* d_jobs, const int* d_order);          // dense matrices (hp1_graphdp.cu)
int launch_graph_trace_packed(cudaStream_t st, int n_jobs, const GraphJobDev* d_jobs, const int* d_order);   // direction bits (hp1_trace.cu)
int launch_graph_fill_strip(cudaStream_t st, int C, int n_jobs, const GraphJobDev* d_jobs, const int* d_order);
int launch_graph_fill_strip_gen(cudaStream_t st, int C, int dbits, int warps, size_t smem, int n_jobs, const GraphJobDev* d_jobs, const int* d_order);
size_t wave_smem_bytes(int ring, int max_n1, int max_n2, int max_tr, int max_np1, int max_nx);
int launch_graph_fill_wave(cudaStream_t st, int L, int ring, int threads, int n_jobs, const GraphJobDev* d_jobs, const int* d_order, int max_n1,
                           int max_n2, int max_tr, int max_np1, int max_nx);
#define SG_MAX_WARPS 8       // warps per job of the general strip kernel (256 strips)
#define SG_SLACK 48          // steps a warp of a job may run ahead of the job's last warp (point-to-point coupling of the warps)
#define SG_HAND_ROWS 128     // rows of a warp's hand-over ring: >= SG_SLACK + 2 + the deepest history look-back (30 rows)
#define SG_MAX_SRC 256       // source columns of non-adjacent column edges per job (shared-memory rings)
__host__ __device__ __forceinline__ int strip_words(int C) { return C <= 8 ? 1 : (C <= 16 ? 2 : 4); }   // uint32 words of direction nibbles per strip and row
