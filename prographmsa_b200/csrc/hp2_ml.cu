// hp2_ml.cu -- K3 (+K2 fused): batched ML distance estimation, one warp per sequence pair.
//
// Replaces DistanceFactoryML<A>::computeDistance / computeMLDist (reference DistanceFactoryML.h:66-190)
// with ModelFactory<A>::getModel(t) inlined (ModelFactory.h:48-67,103-127), and -- for the ROWS
// front-end -- the column scan of DistanceFactoryPrealigned<A>::computePwDistances
// (DistanceFactoryPrealigned.h:57-84).
//
//   P(t) = V diag(exp(sigma t_c)) Vi at the CLAMPED t_c = clamp(max(0,t), min_dist, max_dist),
//   P' = Q P = V diag(sigma e) Vi,  P'' = Q P' = V diag(sigma^2 e) Vi
//   f = sum c_ij P'_ij / P_ij,  ff = sum c_ij (P''_ij P_ij - P'_ij^2) / P_ij^2   (+ the -M gap terms)
// Newton with the reference's bracket / midpoint / iteration-cap rules, iterating on the UNclamped t.
//
// Design: the n x n count matrix is a histogram in the warp's shared memory (built from the NW
// walk's matched-state list, from two MSA rows, or copied from caller counts); it is compacted in
// place to its non-zero entries, so every Newton evaluation costs nnz * 3n FP64 FMAs spread over
// the 32 lanes instead of the reference's three dense n^3 products.  All lanes carry the scalar
// Newton state redundantly (xor-shuffle reductions give every lane identical bits), so the loop has
// no divergence.  FP64 on CUDA cores; agreement with the reference is ~1e-13 relative on P(t).
#include "internal.h"
#include <mutex>

enum { FRONT_LIST = 0, FRONT_ROWS = 1, FRONT_COUNTS = 2, FRONT_COUNTS16 = 3 };

struct MlFront {
    // LIST
    const NwTask* tasks; const uint32_t* off; const uint16_t* list; const uint32_t* list_len; const uint32_t* gaps_in; int task_pairs;
    // ROWS
    const uint8_t* rows; int N; int L; long long pair_begin;
    // COUNTS
    const int32_t* counts_in; const double* seqlen_in;
    // COUNTS16 (the column-scan kernel's output: lim x lim uint16 per pair, lim = min(n, 20); gaps_in; every pair's length is L)
    const uint16_t* counts16_in; int lim;
    uint32_t* gaps_out; int32_t* counts_out;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void pair_from_index(long long p, int N, int& i, int& j) {
    // p = i*N - i(i+1)/2 + (j-i-1);  solve for the largest i with start(i) <= p
    const double b = 2.0 * N - 1.0;
    long long ii = (long long)floor((b - sqrt(b * b - 8.0 * (double)p)) * 0.5);
    if (ii < 0) ii = 0;
    while (ii > 0 && ii * N - ii * (ii + 1) / 2 > p) --ii;
    while ((ii + 1) * N - (ii + 1) * (ii + 2) / 2 <= p) ++ii;
    i = (int)ii;
    j = (int)(p - (ii * N - ii * (ii + 1) / 2) + ii + 1);
}

// ATAB (n <= 20): the products A[(i,j)][k] = V(i,k) Vi(k,j) live in shared memory (stride n+1 doubles,
// conflict-free for LDS.64), so one likelihood evaluation costs one load + three FMAs per (entry, k).
// Each lane register-blocks EB entries so the broadcast loads of the per-k weights are shared.
#define ML_EB 4

// One register block of the likelihood sums: lane handles entries q0 + lane + 32 r, r < EB.  The two quotients
// of the reference (P'/P and (P''P - P'^2)/P^2) share one reciprocal.
template <int EB, bool ATAB>
__device__ __forceinline__ void
ml_eval_block(const uint32_t* __restrict__ hist, uint32_t q0, uint32_t nnz, int lane, int n, int astride,
              const double* __restrict__ sA, const double* __restrict__ sV, const double* __restrict__ sVi,
              const double* __restrict__ wb, double& f, double& ff) {
    int ia[EB], ib[EB]; double cnt[EB], a0[EB], a1[EB], a2[EB];
#pragma unroll
    for (int r = 0; r < EB; ++r) {
        const uint32_t q = q0 + lane + 32 * r;
        const uint32_t h = (q < nnz) ? hist[q] : 0u;
        const int idx = (int)(h >> 20);
        cnt[r] = (double)(h & 0xfffffu);
        if (ATAB) { ia[r] = idx * astride; ib[r] = 0; }
        else { const int i = idx / n; ia[r] = i * n; ib[r] = idx - i * n; }
        a0[r] = a1[r] = a2[r] = 0;
    }
    for (int k = 0; k < n; ++k) {
        const double2 w01 = *reinterpret_cast<const double2*>(wb + 4 * k);
        const double w2 = wb[4 * k + 2];
#pragma unroll
        for (int r = 0; r < EB; ++r) {
            const double a = ATAB ? sA[ia[r] + k] : sV[ia[r] + k] * sVi[k * n + ib[r]];
            a0[r] = fma(a, w01.x, a0[r]); a1[r] = fma(a, w01.y, a1[r]); a2[r] = fma(a, w2, a2[r]);
        }
    }
#pragma unroll
    for (int r = 0; r < EB; ++r)
        if (cnt[r] != 0) {
            const double inv = __drcp_rn(a0[r]);
            const double g = a1[r] * inv;                       // P'/P
            f = fma(cnt[r], g, f);
            ff = fma(cnt[r], fma(a2[r], inv, -g * g), ff);      // P''/P - (P'/P)^2
        }
}

template <int FRONT, bool ATAB, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
ml_kernel(MlDev ml, MlFront fr, long long n_pairs, double* __restrict__ dist_out, double* __restrict__ var_out,
          unsigned long long* __restrict__ evals_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = ml.n, nn = n * n, astride = n + 1;
    double* sA = (double*)smem_raw;                                   // ATAB: nn*(n+1); else V (nn) then Vi (nn)
    double* sV = sA; double* sVi = sA + nn;
    double* sSig = sA + (ATAB ? nn * astride : 2 * nn);              // n
    double* sW = sSig + ((n + 1) & ~1);                               // WARPS * n * 4   {w0,w1,w2,pad}
    uint32_t* sHist = (uint32_t*)(sW + WARPS * n * 4);                // WARPS * nn

    if (ATAB) {
        for (int i = threadIdx.x; i < nn * n; i += WARPS * 32) {
            const int e = i / n, k = i - e * n, r = e / n, c = e - r * n;
            sA[e * astride + k] = ml.V[r * n + k] * ml.Vi[k * n + c];
        }
    } else {
        for (int i = threadIdx.x; i < nn; i += WARPS * 32) { sV[i] = ml.V[i]; sVi[i] = ml.Vi[i]; }
    }
    for (int i = threadIdx.x; i < n; i += WARPS * 32) sSig[i] = ml.sigma[i];
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* wb = sW + warp * n * 4;
    uint32_t* hist = sHist + warp * nn;
    unsigned long long my_evals = 0;
    const uint32_t div_n = ((1u << 20) + (uint32_t)n - 1u) / (uint32_t)n;

    for (long long p = (long long)blockIdx.x * WARPS + warp; p < n_pairs; p += (long long)gridDim.x * WARPS) {
        if (!(FRONT == FRONT_COUNTS16 && fr.lim == n))              // (that front-end overwrites every entry)
            for (int i = lane; i < nn; i += 32) hist[i] = 0;
        __syncwarp();
        uint32_t gaps = 0;
        double seqlen = 0;
        long long out_idx = p;

        if (FRONT == FRONT_LIST) {
            // p enumerates (task, slot); slots past the task's last pair are empty
            const long long ti = p / fr.task_pairs;
            const NwTask task = fr.tasks[ti];
            const uint32_t j = task.j0 + (uint32_t)(p - ti * fr.task_pairs);
            if (j >= task.j1) continue;
            const NwPair d = nw_pair(task, j, fr.off);
            out_idx = d.out_idx;
            const uint32_t len = fr.list_len[out_idx];
            const uint16_t* lp = fr.list + d.list_off;
            for (uint32_t k = lane; k < len; k += 32) atomicAdd(&hist[lp[k]], 1u);
            gaps = fr.gaps_in[out_idx];
            seqlen = ((double)d.L1 + (double)d.L2) / 2.0;
        } else if (FRONT == FRONT_ROWS) {
            int i, j;
            pair_from_index(fr.pair_begin + p, fr.N, i, j);
            const uint8_t* __restrict__ r1 = fr.rows + (size_t)i * fr.L;
            const uint8_t* __restrict__ r2 = fr.rows + (size_t)j * fr.L;
            const int lim = n < 20 ? n : 20;               // hard-coded 20, DistanceFactoryPrealigned.h:65
            int carry = 0;                                   // 0 = none/match, 1 = run in seq1 (A), 2 = run in seq2 (B)
            // The scan is a chain of L2 round trips (the rows are L2-resident, one 32-column group per trip: measured 84 ns per pair
            // at 6660 columns = ~1900 cycles per group): the residues of PF groups are loaded up front, then the groups are walked
            // in order as before
            constexpr int PF = 8;
            for (int base0 = 0; base0 < fr.L; base0 += 32 * PF) {
              int c1v[PF], c2v[PF];
#pragma unroll
              for (int u = 0; u < PF; ++u) {
                  const int k = base0 + 32 * u + lane;
                  c1v[u] = k < fr.L ? (int)r1[k] : 254; c2v[u] = k < fr.L ? (int)r2[k] : 254;
              }
#pragma unroll
              for (int u = 0; u < PF; ++u) {
                const int base = base0 + 32 * u;
                if (base >= fr.L) break;
                const int k = base + lane;
                int type = 3;                                // 3 = skip (both gaps / out of range)
                if (k < fr.L) {
                    const int c1 = c1v[u], c2 = c2v[u];
                    const bool g1 = (c1 == 254), g2 = (c2 == 254);
                    if (!g1 && !g2) { type = 0; if (c1 < lim && c2 < lim) atomicAdd(&hist[c1 * n + c2], 1u); }
                    else if (g1 && g2) type = 3;
                    else if (!g1) type = 1;
                    else type = 2;
                }
                const uint32_t mA = __ballot_sync(0xffffffffu, type == 1);
                const uint32_t mB = __ballot_sync(0xffffffffu, type == 2);
                const uint32_t mM = __ballot_sync(0xffffffffu, type == 0);
                const uint32_t ns = mA | mB | mM;
                const uint32_t below = ns & ((1u << lane) - 1u);
                int prev = carry;
                if (below) { const int q = 31 - __clz(below); prev = ((mA >> q) & 1u) ? 1 : (((mB >> q) & 1u) ? 2 : 0); }
                const bool ev = (type == 1 && prev != 1) || (type == 2 && prev != 2);
                gaps += __popc(__ballot_sync(0xffffffffu, ev));
                if (ns) { const int q = 31 - __clz(ns); carry = ((mA >> q) & 1u) ? 1 : (((mB >> q) & 1u) ? 2 : 0); }
              }
            }
            seqlen = ((double)fr.L + (double)fr.L) / 2.0;
        } else if (FRONT == FRONT_COUNTS16) {
            const int lim = fr.lim, ll = lim * lim;
            const uint16_t* __restrict__ cp = fr.counts16_in + (size_t)p * ll;
            if (lim == n) {                                         // the histogram is the count matrix: two counts per load
                const uint32_t* __restrict__ cw = reinterpret_cast<const uint32_t*>(cp);      // (lim x lim uint16 per pair, lim even or 1 pair: 4-byte aligned when ll is even)
                if ((ll & 1) == 0) { for (int q = lane; 2 * q < ll; q += 32) { const uint32_t v = cw[q]; hist[2 * q] = v & 0xffffu; hist[2 * q + 1] = v >> 16; } }
                else for (int q = lane; q < ll; q += 32) hist[q] = (uint32_t)cp[q];
            } else {
                for (int q = lane; q < ll; q += 32) { const int ci = q / lim, cj = q - ci * lim; hist[ci * n + cj] = (uint32_t)cp[q]; }
            }
            gaps = fr.gaps_in[p];
            seqlen = ((double)fr.L + (double)fr.L) / 2.0;
        } else {
            const int32_t* cp = fr.counts_in + (size_t)p * nn;
            for (int i = lane; i < nn; i += 32) hist[i] = (uint32_t)cp[i];
            gaps = fr.gaps_in[p];
            seqlen = fr.seqlen_in[p];
        }
        __syncwarp();
        if (fr.counts_out) {
            int32_t* co = fr.counts_out + (size_t)out_idx * nn;
            for (int i = lane; i < nn; i += 32) co[i] = (int32_t)hist[i];
        }
        if (fr.gaps_out && lane == 0) fr.gaps_out[out_idx] = gaps;

        // ident / total, then in-place compaction to (idx << 20 | count)
        double ident = 0, total = 0;
        uint32_t nnz = 0;
        for (int base = 0; base < nn; base += 32) {
            const int idx = base + lane;
            uint32_t c = (idx < nn) ? hist[idx] : 0u;
            const int ci = (int)(((uint32_t)idx * div_n) >> 20), cj = idx - ci * n;     // idx / n for idx < 4096, n <= 63 (div_n = ceil(2^20 / n))
            if (c) { total += c; if (ci == cj) ident += c; }
            // reversible model: P'/P and P''/P are the same for (i,j) and (j,i), so the two counts share one entry
            if (ml.symmetric && idx < nn) { if (ci > cj) c = 0u; else if (ci < cj) c += hist[cj * n + ci]; }
            const uint32_t m = __ballot_sync(0xffffffffu, c != 0);
            __syncwarp();
            if (c) hist[nnz + __popc(m & ((1u << lane) - 1u))] = ((uint32_t)idx << 20) | c;
            nnz += __popc(m);
            __syncwarp();
        }
        ident = warp_sum(ident); total = warp_sum(total);

        // computeDistance (DistanceFactoryML.h:138-190)
        double dist0 = 1.0 - ident / total;
        double dist, var;
        if (ml.mldist_flag || ml.mldist_gap_flag) {
            if (total == 0 || dist0 > 0.85) { dist = dist0 = ml.DIST_MAX; var = ml.VAR_MAX; }
            else { dist = dist0 = -log(1.0 - dist0 - 0.2 * dist0 * dist0); var = dist / total; }
            if (total > 0 && ident != total) {
                // computeMLDist (DistanceFactoryML.h:66-136)
                const double var0 = var;
                double dist_min = 0, dist_max = INFINITY, delta = 1;
                double last_tc = -1.0, last_f = 0, last_ff = 0;
                unsigned iteration = 0;
                while (fabs(delta) > 1e-5) {
                    if (iteration > 20) {
                        if (dist_max == INFINITY) { dist = ml.DIST_MAX; var = ml.VAR_MAX; }
                        else { dist = dist0; var = var0; }
                        break;
                    }
                    double tc = (0.0 < dist) ? dist : 0.0;                 // parseDistance (ModelFactory.h:104,125)
                    tc = fmax(fmin(tc, ml.max_dist), ml.min_dist);
                    double f, ff;
                    if (tc == last_tc) {
                        // the likelihood is evaluated at the CLAMPED distance: outside [min_dist, max_dist] the
                        // reference recomputes the same P, P', P'' every iteration -- reuse the identical sums
                        f = last_f; ff = last_ff;
                    } else {
                        __syncwarp();
                        for (int k = lane; k < n; k += 32) {
                            const double s = sSig[k], e = exp(s * tc);
                            wb[4 * k] = e; wb[4 * k + 1] = s * e; wb[4 * k + 2] = s * s * e;
                        }
                        __syncwarp();
                        f = 0; ff = 0;
                        uint32_t q0 = 0;
                        for (; q0 + 32 * ML_EB <= nnz; q0 += 32 * ML_EB)
                            ml_eval_block<ML_EB, ATAB>(hist, q0, nnz, lane, n, astride, sA, sV, sVi, wb, f, ff);
                        switch ((nnz - q0 + 31) >> 5) {         // the tail gets the smallest register block that covers it
                            case 1: ml_eval_block<1, ATAB>(hist, q0, nnz, lane, n, astride, sA, sV, sVi, wb, f, ff); break;
                            case 2: ml_eval_block<2, ATAB>(hist, q0, nnz, lane, n, astride, sA, sV, sVi, wb, f, ff); break;
                            case 3: ml_eval_block<3, ATAB>(hist, q0, nnz, lane, n, astride, sA, sV, sVi, wb, f, ff); break;
                            case 4: ml_eval_block<4, ATAB>(hist, q0, nnz, lane, n, astride, sA, sV, sVi, wb, f, ff); break;
                            default: break;
                        }
                        f = warp_sum(f); ff = warp_sum(ff);
                        ++my_evals;
                        last_tc = tc; last_f = f; last_ff = ff;
                    }
                    if (ml.mldist_gap_flag) {
                        const double grate = ml.indel_rate * seqlen * dist;
                        f += (-grate + (double)gaps) / dist;
                        // sic: the reference negates an unsigned (DistanceFactoryML.h:102)
                        ff += (double)(uint32_t)(0u - gaps) / (dist * dist);
                    }
                    var = -1.0 / ff;
                    if (f > 0) dist_min = fmax(dist_min, dist); else dist_max = fmin(dist_max, dist);
                    double nd = dist - f / ff;
                    if (!(nd < dist_max && nd > dist_min)) {
                        const double upper = (dist_max == INFINITY) ? dist * 3 : dist_max;
                        nd = (upper + dist_min) / 2.0;
                    }
                    delta = 1.0 - nd / dist;
                    dist = nd;
                    ++iteration;
                }
            }
        } else {
            if (total == 0) { dist = dist0 = 1.0; var = ml.VAR_MAX; }
            else { dist = dist0; var = dist0 / total; }
        }
        if (!(dist < ml.DIST_MAX)) { dist = ml.DIST_MAX; var = ml.VAR_MAX; }
        if (dist > ml.cutoff_dist) dist = ml.cutoff_dist;
        if (var < ml.VAR_MIN) var = ml.VAR_MIN;
        if (!(var < ml.VAR_MAX)) var = ml.VAR_MAX;
        if (lane == 0) { dist_out[out_idx] = dist; var_out[out_idx] = var; }
        __syncwarp();
    }
    if (evals_out && lane == 0 && my_evals) atomicAdd(evals_out, my_evals);
}

template <int FRONT, bool ATAB, int WARPS>
static int launch_ml_t(cudaStream_t st, const MlDev& ml, const MlFront& fr, long long n_pairs, double* d_dist, double* d_var,
                       unsigned long long* d_evals) {
    const int n = ml.n, nn = n * n;
    const size_t smem = sizeof(double) * ((ATAB ? (size_t)nn * (n + 1) : 2 * (size_t)nn) + ((n + 1) & ~1) + (size_t)WARPS * n * 4) +
                        sizeof(uint32_t) * (size_t)WARPS * nn;
    // attribute / occupancy queries are driver calls: do them once per (instantiation, shared-memory size), not per launch
    // (function attributes are per device: one cache slot per device, guarded for concurrent contexts)
    struct Cache { size_t smem = 0; int sms = 148, per_sm = 1; };
    static Cache cache[64]; static std::mutex cache_mu;
    int dev = 0;
    PGM_CUDA(cudaGetDevice(&dev));
    int sms, per_sm;
    {
        std::lock_guard<std::mutex> lk(cache_mu);
        Cache& cc = cache[dev & 63];
        if (cc.smem != smem) {
            PGM_CUDA(cudaFuncSetAttribute(ml_kernel<FRONT, ATAB, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            cudaDeviceGetAttribute(&cc.sms, cudaDevAttrMultiProcessorCount, dev);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cc.per_sm, ml_kernel<FRONT, ATAB, WARPS>, WARPS * 32, smem);
            if (cc.per_sm < 1) cc.per_sm = 1;
            cc.smem = smem;
        }
        sms = cc.sms; per_sm = cc.per_sm;
    }
    long long want = (n_pairs + WARPS - 1) / WARPS;
    long long grid = (long long)sms * per_sm;
    if (grid > want) grid = want;
    ml_kernel<FRONT, ATAB, WARPS><<<(unsigned)grid, WARPS * 32, smem, st>>>(ml, fr, n_pairs, d_dist, d_var, d_evals);
    PGM_CUDA(cudaGetLastError());
    return 0;
}

template <int FRONT>
static int launch_ml(cudaStream_t st, const MlDev& ml, const MlFront& fr, long long n_pairs, double* d_dist, double* d_var,
                     unsigned long long* d_evals) {
    if (n_pairs <= 0) return 0;
    if (ml.n * ml.n >= 4096) { pgm_set_error("ML kernel supports n_states <= 63"); return PGM_ERR_ARG; }
    if (ml.n <= 20) return launch_ml_t<FRONT, true, 8>(st, ml, fr, n_pairs, d_dist, d_var, d_evals);
    return launch_ml_t<FRONT, false, 4>(st, ml, fr, n_pairs, d_dist, d_var, d_evals);
}

int launch_ml_from_lists(cudaStream_t st, const MlDev& ml, int n_tasks, const uint32_t* d_off, const NwTask* d_tasks,
                         const uint16_t* d_list, const uint32_t* d_list_len, const uint32_t* d_gaps, double* d_dist,
                         double* d_var, int32_t* d_counts, unsigned long long* d_evals, int task_pairs) {
    MlFront fr{};
    fr.tasks = d_tasks; fr.off = d_off; fr.list = d_list; fr.list_len = d_list_len; fr.gaps_in = d_gaps;
    fr.counts_out = d_counts; fr.task_pairs = task_pairs;
    return launch_ml<FRONT_LIST>(st, ml, fr, (long long)n_tasks * task_pairs, d_dist, d_var, d_evals);
}

// ---- K2: the column scan of DistanceFactoryPrealigned<A>::computePwDistances (DistanceFactoryPrealigned.h:57-84) on its own ----
// The warp-per-pair scan above spends a shared-memory atomic, four ballots and two byte loads per 32 columns and runs at 16 warps per SM next
// to the 67 KB likelihood table: 69 ns per pair at 11,000 columns.  But a row of a large alignment is mostly gaps (500 residues in 11,000
// columns for 5000 sequences), and a column where both rows have a gap changes nothing (:68-69).  So:
//   * prealigned_lists_kernel turns every row into the sorted list of its residues, (column << 8 | state), closed by two sentinels;
//   * prealigned_count_kernel gives one THREAD a pair and MERGES the two lists by column: equal columns are a match (counted when both
//     states are < 20, :65), a column in one list only opens or continues a gap run in that row (:70-79) -- len1 + len2 - matches steps
//     instead of L_aln columns, every step branch-free (what counts nothing increments a dummy bin), the next entry of either list
//     already in a register;
//   * the thread's lim x lim histogram of 16-bit counters sits in shared memory as hist[bin][thread], laid out so that the 32 lanes of a
//     warp hit 32 different banks whatever their bins: no atomics, no conflicts, no ballots;
//   * the counts go to HBM as lim x lim uint16 per pair and the Newton kernel reads them back (ml_kernel<FRONT_COUNTS16>).
#define PC_THREADS 64
#define PC_GAP 254u
#define PC_SENT 0xffffffu
__global__ void __launch_bounds__(256)
prealigned_lists_kernel(const uint8_t* __restrict__ rows, int L, uint32_t* __restrict__ lists, size_t lp) {
    __shared__ int wsum[8];
    __shared__ int base_s;
    const uint8_t* __restrict__ r = rows + (size_t)blockIdx.x * L;
    uint32_t* __restrict__ out = lists + (size_t)blockIdx.x * lp;
    const int tid = (int)threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) base_s = 0;
    __syncthreads();
    for (int k0 = 0; k0 < L; k0 += 256) {
        const int k = k0 + tid;
        const uint32_t c = k < L ? (uint32_t)r[k] : PC_GAP;
        const bool res = c != PC_GAP;
        const uint32_t m = __ballot_sync(0xffffffffu, res);
        if (lane == 0) wsum[w] = __popc(m);
        __syncthreads();
        int off = base_s;
        for (int q = 0; q < w; ++q) off += wsum[q];
        if (res) out[off + __popc(m & ((1u << lane) - 1u))] = ((uint32_t)k << 8) | c;
        __syncthreads();
        if (tid == 0) { int t = 0; for (int q = 0; q < 8; ++q) t += wsum[q]; base_s += t; }
        __syncthreads();
    }
    if (tid < 2) out[base_s + tid] = PC_SENT << 8;
}

__global__ void __launch_bounds__(PC_THREADS)
prealigned_count_kernel(const uint32_t* __restrict__ lists, size_t lp, int N, long long pair_begin, long long n_pairs, int lim,
                        uint16_t* __restrict__ counts16, uint32_t* __restrict__ gaps_out) {
    extern __shared__ __align__(16) unsigned char pc_smem[];
    const int tid = (int)threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int ll = lim * lim;
    // halfword (bin, thread) at byte bin * 2 * PC_THREADS + my: warps 2k and 2k+1 share the words of one 128-byte row
    const uint32_t my = (uint32_t)__cvta_generic_to_shared(pc_smem) + (uint32_t)((lane + 32 * (w >> 1)) * 4 + (w & 1) * 2);
    constexpr uint32_t BIN = 2u * PC_THREADS;
    for (int b = 0; b <= ll; ++b) asm volatile("st.shared.u16 [%0], %1;" ::"r"(my + (uint32_t)b * BIN), "h"((unsigned short)0) : "memory");
    const long long p = (long long)blockIdx.x * PC_THREADS + tid;
    const bool on = p < n_pairs;
    int i = 0, j = 1;
    if (on) pair_from_index(pair_begin + p, N, i, j);
    const uint32_t* __restrict__ A = lists + (size_t)i * lp;
    const uint32_t* __restrict__ B = lists + (size_t)j * lp;
    uint32_t gaps = 0, carry = 0;                                // carry: 0 = none / match, 1 = run in row 1, 2 = run in row 2
    const uint32_t dummy = (uint32_t)ll;
    uint32_t ea = on ? __ldg(A) : (PC_SENT << 8), eb = on ? __ldg(B) : (PC_SENT << 8);
    uint32_t ean = __ldg(A + 1), ebn = __ldg(B + 1);
    int ia = 1, ib = 1;                                          // index of ean / ebn
    for (;;) {
        const uint32_t ca = ea >> 8, cb = eb >> 8;
        if ((ca & cb) == PC_SENT) break;                         // both lists are through
        const bool ta = ca <= cb, tb = cb <= ca;                 // the earlier column comes next; equal columns: both rows have a residue
        const uint32_t c1 = ea & 0xffu, c2 = eb & 0xffu;
        const bool cnt = ta && tb && c1 < (uint32_t)lim && c2 < (uint32_t)lim;          // hard-coded 20, DistanceFactoryPrealigned.h:65
        const uint32_t addr = my + (cnt ? c1 * (uint32_t)lim + c2 : dummy) * BIN;
        unsigned short v;
        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr) : "memory");
        v = (unsigned short)(v + 1);
        asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
        const uint32_t t = ta ? (tb ? 0u : 1u) : 2u;             // 0 match, 1 residue in row 1 only, 2 in row 2 only
        gaps += (uint32_t)(t != 0u && t != carry);               // a run opens (DistanceFactoryPrealigned.h:70-79)
        carry = t;
        ea = ta ? ean : ea; eb = tb ? ebn : eb;
        ia += ta ? 1 : 0; ib += tb ? 1 : 0;
        ean = __ldg(A + ia); ebn = __ldg(B + ib);                // (behind a sentinel sits the second sentinel: never past the list)
        // (measured and not kept: 8-byte loads of entry pairs with a position flag and predicated reloads -- half the list wavefronts,
        //  but 15.8 -> 19.8 ms per 4 Mi pairs)
    }
    if (on) {
        uint16_t* __restrict__ out = counts16 + (size_t)p * ll;
        if ((ll & 7) == 0) {
            for (int b0 = 0; b0 < ll; b0 += 8) {
                uint32_t h[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) { unsigned short v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(my + (uint32_t)(b0 + u) * BIN) : "memory"); h[u] = v; }
                *reinterpret_cast<uint4*>(out + b0) = make_uint4(h[0] | (h[1] << 16), h[2] | (h[3] << 16), h[4] | (h[5] << 16), h[6] | (h[7] << 16));
            }
        } else {
            for (int b0 = 0; b0 < ll; ++b0) { unsigned short v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(my + (uint32_t)b0 * BIN) : "memory"); out[b0] = v; }
        }
        gaps_out[p] = gaps;
    }
}

size_t prealigned_pitch(int L) { return (size_t)std::max(L, 0) + 2; }      // entries per row list: every column a residue + two sentinels

// rows (N x L, device) -> residue lists in d_lists (N x prealigned_pitch(L) uint32)
int launch_prealigned_repitch(cudaStream_t st, const uint8_t* d_rows, int N, int L, uint8_t* d_lists) {
    if (N <= 0) return 0;
    prealigned_lists_kernel<<<N, 256, 0, st>>>(d_rows, L, (uint32_t*)d_lists, prealigned_pitch(L));
    PGM_CUDA(cudaGetLastError());
    return 0;
}

// one chunk of pairs [pair_begin, pair_begin + n_pairs): list merge into d_counts16 / d_gaps16 (chunk-local indices), then the Newton
// kernel on those counts; d_dist / d_var / d_counts / d_gaps are the chunk's output positions
int launch_ml_from_rows_split(cudaStream_t st, const MlDev& ml, int N, int L, int64_t pair_begin, int64_t n_pairs, const uint8_t* d_lists,
                              uint16_t* d_counts16, uint32_t* d_gaps16, double* d_dist, double* d_var, int32_t* d_counts, uint32_t* d_gaps,
                              unsigned long long* d_evals) {
    if (n_pairs <= 0) return 0;
    const int lim = ml.n < 20 ? ml.n : 20;
    const size_t smem = (size_t)(lim * lim + 1) * 2 * PC_THREADS;
    static std::mutex mu; static size_t opted[64] = {0};
    int dev = 0; PGM_CUDA(cudaGetDevice(&dev));
    {
        std::lock_guard<std::mutex> lk(mu);
        if (opted[dev & 63] < smem) {
            PGM_CUDA(cudaFuncSetAttribute(prealigned_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            // the lists are read 4 bytes at a time, 64 different rows per CTA: with all 228 KB as shared memory (four CTAs per SM) the
            // L1 that is left thrashes and every step of the merge is an L2 round trip (430 cycles per step measured); three CTAs
            // leave ~90 KB of L1
            int carve = 64;
            if (const char* e = getenv("PGM_PC_CARVEOUT")) carve = atoi(e);
            PGM_CUDA(cudaFuncSetAttribute(prealigned_count_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
            opted[dev & 63] = smem;
        }
    }
    prealigned_count_kernel<<<(unsigned)((n_pairs + PC_THREADS - 1) / PC_THREADS), PC_THREADS, smem, st>>>(
        (const uint32_t*)d_lists, prealigned_pitch(L), N, (long long)pair_begin, (long long)n_pairs, lim, d_counts16, d_gaps16);
    PGM_CUDA(cudaGetLastError());
    MlFront fr{};
    fr.counts16_in = d_counts16; fr.lim = lim; fr.gaps_in = d_gaps16; fr.L = L; fr.counts_out = d_counts; fr.gaps_out = d_gaps;
    return launch_ml<FRONT_COUNTS16>(st, ml, fr, n_pairs, d_dist, d_var, d_evals);
}

int launch_ml_from_rows(cudaStream_t st, const MlDev& ml, int N, int L, int64_t pair_begin, int64_t n_pairs,
                        const uint8_t* d_rows, double* d_dist, double* d_var, int32_t* d_counts, uint32_t* d_gaps,
                        unsigned long long* d_evals) {
    MlFront fr{};
    fr.rows = d_rows; fr.N = N; fr.L = L; fr.pair_begin = pair_begin; fr.counts_out = d_counts; fr.gaps_out = d_gaps;
    return launch_ml<FRONT_ROWS>(st, ml, fr, n_pairs, d_dist, d_var, d_evals);
}

int launch_ml_from_counts(cudaStream_t st, const MlDev& ml, int64_t n_pairs, const int32_t* d_counts, const uint32_t* d_gaps,
                          const double* d_seqlen, double* d_dist, double* d_var, unsigned long long* d_evals) {
    MlFront fr{};
    fr.counts_in = d_counts; fr.gaps_in = d_gaps; fr.seqlen_in = d_seqlen;
    return launch_ml<FRONT_COUNTS>(st, ml, fr, n_pairs, d_dist, d_var, d_evals);
}
