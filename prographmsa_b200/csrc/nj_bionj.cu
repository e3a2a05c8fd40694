// nj_bionj.cu -- BioNJ join loop on the GPU.  Replaces the O(N^3) part of
// TreeNJ<A>::buildNJTree (reference TreeNJ.cpp:132-281): per join the clamped matrices, the column sums, the argmin of
// the Q criterion, the BioNJ lambda and the reduced distance / variance matrices.  The tree objects, names and the
// final <= 3 taxa stay on the host (tree.py / TreeNJ in the reference).
//
// The arithmetic follows the host specification (tree.py build_nj_tree) operation by operation in FP64 without
// contraction, so the join sequence and the branch lengths are bit-identical to it:
//   D <- max(D, 1e-4), V <- max(V, 1e-5), zero diagonals          (only the row/column written by the last join changes)
//   sums[j] = sum_i D[i][j], sequentially in i                    (tiles staged in shared memory, one adding thread per column)
//   Q[i][j] = 0.5 D[i][j] - (0.5 / (dim - 2)) (sums[j] + sums[i]); first minimum in column-major order
//   d1, d2, lambda (its row difference summed sequentially), new row, then row/column i2 removed
// The whole loop is enqueued without a host synchronisation: the join parameters live in device memory.
//
// Removing a row and a column by copying the matrices (what the reference and the first version here did) is 4 of the 6
// dim^2 passes of a join.  The matrices now stay where they are: `alive` maps the logical index of a taxon (the order the
// reference's compacted matrices would have) to its physical row / column, a join rewrites row and column alive[i1] in place
// and drops i2 from `alive`; every kernel walks logical indices in increasing order, so sums, ties and results are bit for bit
// those of the compacted matrices.  When a quarter of the physical rows are dead the matrices are compacted once.
#include "internal.h"
#include <math_constants.h>
#include <algorithm>

struct pgm_ctx;
cudaStream_t pgm_ctx_stream(pgm_ctx* c);
int pgm_ctx_device(pgm_ctx* c);
int pgm_ctx_take_dev(pgm_ctx* c, size_t bytes, DevBuf& out);
void pgm_ctx_give_dev(pgm_ctx* c, DevBuf& b);

namespace {

inline size_t align_up_sz(size_t v, size_t a) { return (v + a - 1) / a * a; }
constexpr double NJ_MIN_DIST = 1e-4, NJ_MIN_VAR = 1e-5;
constexpr int ARGMIN_BLOCKS = 592, ARGMIN_THREADS = 256;

struct NjParams { int i1, i2; double d1, d2, lam, v12; };
struct NjCand { double q; unsigned long long flat; };

// clamp: the first join touches everything, later joins only the row / column written by the previous join
__global__ void __launch_bounds__(256)
nj_clamp_kernel(double* __restrict__ D, double* __restrict__ V, int dim, int P, const int* __restrict__ alive, const NjParams* __restrict__ prm, int first) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    int i, j;                                                       // physical
    if (first) { if (e >= (size_t)dim * dim) return; i = (int)(e / dim); j = (int)(e - (size_t)i * dim); }     // (alive is the identity then)
    else {
        if (e >= 2 * (size_t)dim) return;
        const int t = alive[prm->i1], k = alive[(int)(e >> 1)];
        if (e & 1) { i = t; j = k; } else { i = k; j = t; }
    }
    const size_t o = (size_t)i * P + j;
    D[o] = (i == j) ? 0.0 : fmax(D[o], NJ_MIN_DIST);
    V[o] = (i == j) ? 0.0 : fmax(V[o], NJ_MIN_VAR);
}

// sums[j] = sum_i D[i][j], sequentially in i.  A CTA owns 32 columns: its eight warps stream the rows in stages of 64 through a
// six-stage cp.async ring in shared memory (five stages = 320 rows of loads in flight per CTA, which covers the memory latency),
// and one thread per column adds the values of a stage in row order -- the order of a scalar loop, at the pace of the
// dependent additions (64 per stage) instead of one memory round trip per tile (the first version: 81 us per launch on average,
// half of the loop's time).
#define CS_ROWS 64
#define CS_STAGES 6
__global__ void __launch_bounds__(256)
nj_colsum_kernel(const double* __restrict__ D, int dim, int P, const int* __restrict__ alive, double* __restrict__ sums) {
    extern __shared__ __align__(16) double cs_tile[];               // [CS_STAGES][CS_ROWS][32]
    const int j0 = blockIdx.x * 32, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int pj = j0 + lane < dim ? alive[j0 + lane] : alive[0];
    const int n_st = (dim + CS_ROWS - 1) / CS_ROWS;
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(cs_tile);
    auto issue = [&](int st) {                                      // rows st * CS_ROWS ... of the 32 columns -> ring slot st % CS_STAGES
        if (st < n_st) {
            const int slot = st % CS_STAGES;
            for (int r = w; r < CS_ROWS; r += 8) {
                const int i = st * CS_ROWS + r;
                const uint32_t dst = base + (uint32_t)(((slot * CS_ROWS + r) * 32 + lane) * 8);
                if (i < dim) {
                    const double* src = D + (size_t)alive[i] * P + pj;
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dst), "l"(src) : "memory");
                }
            }
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };
    for (int st = 0; st < CS_STAGES - 1; ++st) issue(st);
    double s = 0.0;
    for (int st = 0; st < n_st; ++st) {
        issue(st + CS_STAGES - 1);                                  // (overwrites the slot added in the previous iteration: the barrier below orders it)
        asm volatile("cp.async.wait_group %0;\n" ::"n"(CS_STAGES - 1) : "memory");
        __syncthreads();
        if (w == 0) {
            const int rows = min(CS_ROWS, dim - st * CS_ROWS);
            const double* t = cs_tile + (size_t)(st % CS_STAGES) * CS_ROWS * 32 + lane;
            int r = 0;
            for (; r + 16 <= rows; r += 16) {                       // the loads of 16 rows first (independent), then the additions in row order
                double v[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) v[q] = t[(r + q) * 32];
#pragma unroll
                for (int q = 0; q < 16; ++q) s = s + v[q];
            }
            for (; r < rows; ++r) s = s + t[r * 32];
        }
        __syncthreads();
    }
    if (w == 0 && j0 + lane < dim) sums[j0 + lane] = s;
}

__device__ __forceinline__ bool nj_better(double q, unsigned long long f, double bq, unsigned long long bf) {
    return q < bq || (q == bq && f < bf);
}

// Q criterion: per-block minimum, ties to the smallest column-major index (Eigen minCoeff / numpy argmin order)
__global__ void __launch_bounds__(ARGMIN_THREADS)
nj_argmin_kernel(const double* __restrict__ D, int dim, int P, const int* __restrict__ alive, const double* __restrict__ sums, NjCand* __restrict__ part) {
    const double c = 0.5 / ((double)dim - 2.0);
    double bq = CUDART_INF; unsigned long long bf = ~0ull;
    // the block's rows i = blockIdx.x, + gridDim.x, ...: their physical row and column sum once in shared memory; a thread keeps its
    // column's (alive, sum) in registers and reads four rows' entries at a time (eight: no faster) (independent loads).  The minimum is over a total
    // order (q, flat index), so the order of evaluation does not matter.
    constexpr int MAXR = 64;                                        // rows per block per pass
    __shared__ int r_phys[MAXR]; __shared__ double r_sum[MAXR];
    for (int rbase = blockIdx.x; rbase < dim; rbase += MAXR * gridDim.x) {
        const int nr = min(MAXR, (dim - rbase + (int)gridDim.x - 1) / (int)gridDim.x);
        __syncthreads();
        for (int r = threadIdx.x; r < nr; r += blockDim.x) { const int i = rbase + r * gridDim.x; r_phys[r] = alive[i]; r_sum[r] = sums[i]; }
        __syncthreads();
        for (int j = threadIdx.x; j < dim; j += blockDim.x) {
            const int aj = alive[j];
            const double sj = sums[j];
            for (int r0 = 0; r0 < nr; r0 += 4) {
                double dv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) dv[u] = r0 + u < nr ? D[(size_t)r_phys[r0 + u] * P + aj] : 0.0;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = rbase + (r0 + u) * gridDim.x;
                    if (r0 + u < nr && i != j) {
                        const double q = 0.5 * dv[u] - c * (sj + r_sum[r0 + u]);
                        const unsigned long long f = (unsigned long long)j * dim + i;
                        if (nj_better(q, f, bq, bf)) { bq = q; bf = f; }
                    }
                }
            }
        }
    }
    __shared__ double sq[ARGMIN_THREADS]; __shared__ unsigned long long sf[ARGMIN_THREADS];
    sq[threadIdx.x] = bq; sf[threadIdx.x] = bf;
    __syncthreads();
    for (int st = ARGMIN_THREADS / 2; st > 0; st >>= 1) {
        if ((int)threadIdx.x < st && nj_better(sq[threadIdx.x + st], sf[threadIdx.x + st], sq[threadIdx.x], sf[threadIdx.x])) {
            sq[threadIdx.x] = sq[threadIdx.x + st]; sf[threadIdx.x] = sf[threadIdx.x + st];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { part[blockIdx.x].q = sq[0]; part[blockIdx.x].flat = sf[0]; }
}

// one block: final argmin over the partials, then the join's scalars (TreeNJ.cpp:176-214 / tree.py)
#define JN_CHUNK 2048
__global__ void __launch_bounds__(ARGMIN_THREADS)
nj_join_kernel(const double* __restrict__ D, const double* __restrict__ V, int dim, int P, const int* __restrict__ alive, const double* __restrict__ sums,
               const NjCand* __restrict__ part, int n_part, NjParams* __restrict__ prm, int join,
               int32_t* __restrict__ j1, int32_t* __restrict__ j2, double* __restrict__ jd1, double* __restrict__ jd2) {
    __shared__ double sq[ARGMIN_THREADS]; __shared__ unsigned long long sf[ARGMIN_THREADS];
    __shared__ double diff[JN_CHUNK];
    __shared__ int s_i1, s_i2;
    double bq = CUDART_INF; unsigned long long bf = ~0ull;
    for (int k = threadIdx.x; k < n_part; k += blockDim.x) if (nj_better(part[k].q, part[k].flat, bq, bf)) { bq = part[k].q; bf = part[k].flat; }
    sq[threadIdx.x] = bq; sf[threadIdx.x] = bf;
    __syncthreads();
    for (int st = ARGMIN_THREADS / 2; st > 0; st >>= 1) {
        if ((int)threadIdx.x < st && nj_better(sq[threadIdx.x + st], sf[threadIdx.x + st], sq[threadIdx.x], sf[threadIdx.x])) {
            sq[threadIdx.x] = sq[threadIdx.x + st]; sf[threadIdx.x] = sf[threadIdx.x + st];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        int i2 = (int)(sf[0] % (unsigned long long)dim), i1 = (int)(sf[0] / (unsigned long long)dim);   // (row, col)
        if (i2 < i1) { const int t = i1; i1 = i2; i2 = t; }
        s_i1 = i1; s_i2 = i2;
    }
    __syncthreads();
    const int i1 = s_i1, i2 = s_i2;
    const int a1 = alive[i1], a2 = alive[i2];                       // physical rows of the two taxa
    // sum_k (V[i2][k] - V[i1][k]) sequentially in k: the differences are staged by the whole block, added by one thread
    double s = 0.0;
    for (int k0 = 0; k0 < dim; k0 += JN_CHUNK) {
        for (int k = threadIdx.x; k < JN_CHUNK && k0 + k < dim; k += blockDim.x) {
            const int pk = alive[k0 + k];
            diff[k] = V[(size_t)a2 * P + pk] - V[(size_t)a1 * P + pk];
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const int m = min(JN_CHUNK, dim - k0);
            int k = 0;
            for (; k + 16 <= m; k += 16) {                          // loads first, then the additions in order (see nj_colsum_kernel)
                double v[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) v[q] = diff[k + q];
#pragma unroll
                for (int q = 0; q < 16; ++q) s = s + v[q];
            }
            for (; k < m; ++k) s = s + diff[k];
        }
        __syncthreads();
    }
    if (threadIdx.x != 0) return;
    const double d12 = D[(size_t)a1 * P + a2];
    double d1 = (d12 + (sums[i1] - sums[i2]) / ((double)dim - 2.0)) / 2.0;
    d1 = fmin(fmax(d1, NJ_MIN_DIST), d12);
    const double d2 = fmax(D[(size_t)a2 * P + a1] - d1, NJ_MIN_DIST);
    const double v12 = V[(size_t)a1 * P + a2];
    double lam = (v12 != 0.0) ? 0.5 + s / ((double)(2 * (dim - 2)) * v12) : CUDART_NAN;
    lam = isnan(lam) ? 0.5 : fmin(fmax(0.0, lam), 1.0);
    prm->i1 = i1; prm->i2 = i2; prm->d1 = d1; prm->d2 = d2; prm->lam = lam; prm->v12 = v12;
    j1[join] = i1; j2[join] = i2; jd1[join] = d1; jd2[join] = d2;
}

// the join's new taxon takes row and column alive[i1], in place: the BioNJ combination of the two joined rows
__global__ void __launch_bounds__(256)
nj_update_kernel(double* __restrict__ D, double* __restrict__ V, int dim, int P, const int* __restrict__ alive, const NjParams* __restrict__ prm, int clamp) {
    const int i1 = prm->i1, i2 = prm->i2;
    const double d1 = prm->d1, d2 = prm->d2, lam = prm->lam, v12 = prm->v12;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;            // the other taxon, logical
    if (k >= dim || k == i1 || k == i2) return;
    const int a1 = alive[i1], a2 = alive[i2], ks = alive[k];
    double d = lam * (D[(size_t)a1 * P + ks] - d1) + (1.0 - lam) * (D[(size_t)a2 * P + ks] - d2);
    double v = lam * V[(size_t)a1 * P + ks] + (1.0 - lam) * V[(size_t)a2 * P + ks] - lam * (1.0 - lam) * v12;
    // the next iteration's clamp (TreeNJ.cpp: distances >= 1e-4, variances >= 1e-5 at the top of the loop) touches only what this
    // join writes, so it is applied here instead of by a launch of its own (the diagonal entry stays 0); the last join's values are
    // what the caller gets back, unclamped
    if (clamp) { d = fmax(d, NJ_MIN_DIST); v = fmax(v, NJ_MIN_VAR); }
    D[(size_t)a1 * P + ks] = d; D[(size_t)ks * P + a1] = d;
    V[(size_t)a1 * P + ks] = v; V[(size_t)ks * P + a1] = v;
}

// taxon i2 leaves: alive[k] <- alive[k + 1] for k >= i2 (one block, ascending chunks)
__global__ void __launch_bounds__(1024)
nj_drop_kernel(int* __restrict__ alive, int dim, const NjParams* __restrict__ prm) {
    const int i2 = prm->i2;
    for (int base = i2; base < dim - 1; base += (int)blockDim.x) {
        const int k = base + (int)threadIdx.x;
        const int v = k < dim - 1 ? alive[k + 1] : 0;
        __syncthreads();
        if (k < dim - 1) alive[k] = v;
        __syncthreads();
    }
}

// physical compaction: the alive rows / columns copied densely into the other buffer pair, alive <- identity
__global__ void __launch_bounds__(256)
nj_compact_kernel(const double* __restrict__ Ds, const double* __restrict__ Vs, double* __restrict__ Dd, double* __restrict__ Vd,
                  int dim, int P, const int* __restrict__ alive) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (c >= dim) return;
    const size_t src = (size_t)alive[r] * P + alive[c];
    Dd[(size_t)r * P + c] = Ds[src]; Vd[(size_t)r * P + c] = Vs[src];
}
__global__ void nj_iota_kernel(int* __restrict__ alive, int n) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) alive[k] = k;
}
// the remaining <= 3 x 3 distances, dense
__global__ void nj_rest_kernel(const double* __restrict__ D, int dim, int P, const int* __restrict__ alive, double* __restrict__ out) {
    const int k = threadIdx.x;
    if (k < dim * dim) out[k] = D[(size_t)alive[k / dim] * P + alive[k % dim]];
}

} // namespace

extern "C" int pgm_bionj_joins(pgm_ctx* c, int N, const double* D, const double* V, int32_t* join_i1, int32_t* join_i2,
                               double* join_d1, double* join_d2, double* D_rest, int* rest_dim) {
    if (!c || !D || !V || !rest_dim || !D_rest || N < 1) { pgm_set_error("bad arguments"); return PGM_ERR_ARG; }
    const int n_joins = N > 3 ? N - 3 : 0;
    if (n_joins && (!join_i1 || !join_i2 || !join_d1 || !join_d2)) { pgm_set_error("NULL join arrays"); return PGM_ERR_ARG; }
    *rest_dim = N - n_joins;
    if (n_joins == 0) { for (int i = 0; i < N * N; ++i) D_rest[i] = D[i]; return 0; }
    PGM_CUDA(cudaSetDevice(pgm_ctx_device(c)));
    cudaStream_t st = pgm_ctx_stream(c);
    const size_t mat = (size_t)N * N * sizeof(double);
    DevBuf buf;
    const size_t small = align_up_sz((size_t)N * 8, 256) + align_up_sz(sizeof(NjCand) * ARGMIN_BLOCKS, 256) + 256 +
                         2 * align_up_sz((size_t)n_joins * 4, 256) + 2 * align_up_sz((size_t)n_joins * 8, 256) +
                         align_up_sz((size_t)N * 4, 256) + 256;
    int rc = pgm_ctx_take_dev(c, 4 * mat + small, buf);
    if (rc) return rc;
    char* p = buf.as<char>();
    double* Dm[2] = {(double*)p, (double*)(p + mat)}; double* Vm[2] = {(double*)(p + 2 * mat), (double*)(p + 3 * mat)};
    p += 4 * mat;
    double* sums = (double*)p; p += align_up_sz((size_t)N * 8, 256);
    NjCand* part = (NjCand*)p; p += align_up_sz(sizeof(NjCand) * ARGMIN_BLOCKS, 256);
    NjParams* prm = (NjParams*)p; p += 256;
    int32_t* dj1 = (int32_t*)p; p += align_up_sz((size_t)n_joins * 4, 256);
    int32_t* dj2 = (int32_t*)p; p += align_up_sz((size_t)n_joins * 4, 256);
    double* dd1 = (double*)p; p += align_up_sz((size_t)n_joins * 8, 256);
    double* dd2 = (double*)p; p += align_up_sz((size_t)n_joins * 8, 256);
    int* alive = (int*)p; p += align_up_sz((size_t)N * 4, 256);
    double* rest = (double*)p;
    cudaError_t e = cudaMemcpyAsync(Dm[0], D, mat, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(Vm[0], V, mat, cudaMemcpyHostToDevice, st);
    nj_iota_kernel<<<(N + 255) / 256, 256, 0, st>>>(alive, N);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(nj_colsum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(CS_STAGES * CS_ROWS * 32 * sizeof(double)));
    int cur = 0, dim = N, phys = N;                                 // phys: physical rows in use (alive + dead)
    for (int join = 0; join < n_joins && e == cudaSuccess; ++join, --dim) {
        if (dim >= 64 && 4 * dim <= 3 * phys) {                     // a quarter of the physical rows are dead: compact once
            nj_compact_kernel<<<dim3((dim + 255) / 256, dim), 256, 0, st>>>(Dm[cur], Vm[cur], Dm[cur ^ 1], Vm[cur ^ 1], dim, N, alive);
            nj_iota_kernel<<<(dim + 255) / 256, 256, 0, st>>>(alive, dim);
            cur ^= 1; phys = dim;
        }
        if (join == 0) nj_clamp_kernel<<<(unsigned)(((size_t)dim * dim + 255) / 256), 256, 0, st>>>(Dm[cur], Vm[cur], dim, N, alive, prm, 1);   // later joins: nj_update_kernel clamps what it writes
        nj_colsum_kernel<<<(dim + 31) / 32, 256, CS_STAGES * CS_ROWS * 32 * sizeof(double), st>>>(Dm[cur], dim, N, alive, sums);
        const int nblk = std::min(ARGMIN_BLOCKS, dim);
        nj_argmin_kernel<<<nblk, ARGMIN_THREADS, 0, st>>>(Dm[cur], dim, N, alive, sums, part);
        nj_join_kernel<<<1, ARGMIN_THREADS, 0, st>>>(Dm[cur], Vm[cur], dim, N, alive, sums, part, nblk, prm, join, dj1, dj2, dd1, dd2);
        nj_update_kernel<<<(dim + 255) / 256, 256, 0, st>>>(Dm[cur], Vm[cur], dim, N, alive, prm, join + 1 < n_joins);
        nj_drop_kernel<<<1, 1024, 0, st>>>(alive, dim, prm);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) nj_rest_kernel<<<1, 32, 0, st>>>(Dm[cur], dim, N, alive, rest);
    if (e == cudaSuccess) e = cudaMemcpyAsync(join_i1, dj1, (size_t)n_joins * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(join_i2, dj2, (size_t)n_joins * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(join_d1, dd1, (size_t)n_joins * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(join_d2, dd2, (size_t)n_joins * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(D_rest, rest, (size_t)dim * dim * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    pgm_ctx_give_dev(c, buf);
    if (e != cudaSuccess) { pgm_set_error("CUDA error in BioNJ: %s", cudaGetErrorString(e)); cudaGetLastError(); return PGM_ERR_CUDA; }
    return 0;
}
