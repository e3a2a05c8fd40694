// hp1_emission.cu -- K4 (exact flavour): emission-score matrix on fp32 CUDA cores.
//
// Replaces precomputeScores (reference GraphAlign.h:145-163) + ls_log_add (ls_log.h:22-59):
//   T = M^T g2 (n x N2),  u = g1^T pi,  v = pi^T g2,
//   S(y,x) = [sum_a g1(a,y) T(a,x)] / (u(y) v(x)),  then  S <- ls_log2(S) + match_init
// all in fp32 with sequential-k accumulation and NO fused multiply-add (the reference is SSE2
// mul/add; this file is compiled with --fmad=false and uses explicit _rn intrinsics), so this
// flavour is bit-identical to the oracle and is what the traceback-identity tests run on.
//
// S is written row-major by graph-1 node (S(y,x) at y*ldS + x): the wavefront kernel maps one thread
// to one row y, which then streams its own row of S 16 bytes at a time.  Tiles are 32 rows x 32 columns;
// a thread computes 4 consecutive rows of one column, so per k it needs one conflict-free load of T and
// one broadcast 16-byte load of g1: 2 shared-memory loads per 8 flops.
#include "internal.h"

__device__ __forceinline__ float ls_log2_add(float x, float add) {
    // ls_log.h:30-56 (vector body): r = (b/(d-c) + add) + (float(e) + a)
    const float a = 2.44247459618085927548717403238913328776812604856113966238617812902399112761292613763080658235564f;
    const float b = -4.2040783745848554315883301529007786406310628696382695994938550046831869207082846248658671f;
    const float c = -0.72123729809042963774358701619456664388406302428056983119308906451199556380646306f;
    const uint32_t bits = __float_as_uint(x);
    const float e = (float)((int)(bits >> 23) - 126);
    const float d = __uint_as_float(((bits << 9) >> 9) ^ 0x3f000000u);
    float vd = __fdiv_rn(b, __fsub_rn(d, c));
    const float ve = __fadd_rn(e, a);
    vd = __fadd_rn(vd, add);
    return __fadd_rn(vd, ve);
}

// T = M^T g2, u, v.  grid = (ceil(maxN/128), n_jobs)
__global__ void __launch_bounds__(128)
emission_prep_kernel(int n, const GraphJobDev* __restrict__ jobs, const int* __restrict__ order) {
    const GraphJobDev& J = jobs[order ? order[blockIdx.y] : blockIdx.y];
    extern __shared__ float sm[];            // Mf (n*n) + pif (n)
    float* sM = sm; float* sPi = sm + n * n;
    for (int i = threadIdx.x; i < n * n; i += blockDim.x) sM[i] = J.Mf[i];
    for (int i = threadIdx.x; i < n; i += blockDim.x) sPi[i] = J.pif[i];
    __syncthreads();
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < J.N2e) {
        const float* g = J.sites2 + (size_t)n * k;
        for (int a = 0; a < n; ++a) {                       // T(a,x) = sum_b M(b,a) g2(b,x)
            float s = 0.f;
            for (int b = 0; b < n; ++b) s = __fadd_rn(s, __fmul_rn(sM[b * n + a], g[b]));
            J.Tbuf[a + (size_t)n * k] = s;
        }
        J.v[k] = dot_gemv_order(g, sPi, n);
    }
    if (k < J.N1e) {
        const float* g = J.sites1 + (size_t)n * k;
        J.u[k] = dot_gemv_order(g, sPi, n);
    }
}

// grid = (tiles_x, tiles_y, n_jobs), block = (32, 8).  NT: the number of states as a compile-time constant (20, 4, 61; 0 = run time):
// the tile staging divides by it, and with a run-time divisor that division was half of the kernel's instructions
// (ncu: 4.0 warp instructions per cell against 2 * 20 multiply / add).  The T tile is padded to 33 columns so that the transposing
// stores are conflict-free; the g1 tile is stored as one float4 per (row quad, state), read as a broadcast.
template <int NT>
__global__ void __launch_bounds__(256)
emission_exact_kernel(int n_rt, const GraphJobDev* __restrict__ jobs, const int* __restrict__ order) {
    const int n = NT ? NT : n_rt;
    const GraphJobDev& J = jobs[order ? order[blockIdx.z] : blockIdx.z];
    const int N1 = J.N1e, N2 = J.N2e;
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
    if (x0 >= N2 || y0 >= N1) return;
    extern __shared__ __align__(16) float sm[];
    float* g1q = sm;                       // [8 row quads][n][4]
    float* Tt = sm + n * 32;               // [n][33]
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const float* __restrict__ src1 = J.sites1 + (size_t)n * y0;      // the tile's 32 nodes x n states are contiguous
    const float* __restrict__ src2 = J.Tbuf + (size_t)n * x0;
    const int lim1 = n * min(32, N1 - y0), lim2 = n * min(32, N2 - x0);
    for (int i = tid; i < n * 32; i += 256) {
        const int yy = i / n, a = i - yy * n;
        g1q[((yy >> 2) * n + a) * 4 + (yy & 3)] = i < lim1 ? src1[i] : 0.f;
    }
    for (int i = tid; i < n * 32; i += 256) {
        const int xx = i / n, a = i - xx * n;
        Tt[a * 33 + xx] = i < lim2 ? src2[i] : 0.f;
    }
    __syncthreads();
    const int x = x0 + threadIdx.x;
    const int yl = threadIdx.y * 4;                      // 4 consecutive rows
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    const float4* __restrict__ gq = reinterpret_cast<const float4*>(g1q) + threadIdx.y * n;
#pragma unroll 4
    for (int a = 0; a < n; ++a) {                        // sequential-k, mul then add (no FMA)
        const float t = Tt[a * 33 + threadIdx.x];
        const float4 g = gq[a];
        s0 = __fadd_rn(s0, __fmul_rn(g.x, t)); s1 = __fadd_rn(s1, __fmul_rn(g.y, t));
        s2 = __fadd_rn(s2, __fmul_rn(g.z, t)); s3 = __fadd_rn(s3, __fmul_rn(g.w, t));
    }
    if (x >= N2) return;
    const float vx = J.v[x], add = J.sc[2];              // match_init
    const float ss[4] = {s0, s1, s2, s3};
    float* __restrict__ S = J.S;
    const int ldS = J.ldS;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int y = y0 + yl + r;
        if (y < N1) S[(size_t)y * ldS + x] = ls_log2_add(__fdiv_rn(ss[r], __fmul_rn(J.u[y], vx)), add);
    }
}

int launch_emission_prep(cudaStream_t st, int n, int n_jobs, const GraphJobDev* d_jobs, const int* d_order, int maxN) {
    dim3 grid((maxN + 127) / 128, n_jobs);
    emission_prep_kernel<<<grid, 128, sizeof(float) * (n * n + n), st>>>(n, d_jobs, d_order);
    PGM_CUDA(cudaGetLastError());
    return 0;
}

int launch_emission_exact(cudaStream_t st, int n, int n_jobs, const GraphJobDev* d_jobs, const int* d_order, int max_N1, int max_N2) {
    for (int j0 = 0; j0 < n_jobs; j0 += 65535) {                   // gridDim.y / .z hold at most 65535 jobs
        const int nj = n_jobs - j0 < 65535 ? n_jobs - j0 : 65535;
        const GraphJobDev* dj = d_order ? d_jobs : d_jobs + j0;
        const int* ord = d_order ? d_order + j0 : nullptr;
        int rc = launch_emission_prep(st, n, nj, dj, ord, max_N1 > max_N2 ? max_N1 : max_N2);
        if (rc) return rc;
        dim3 grid((max_N2 + 31) / 32, (max_N1 + 31) / 32, nj);
        const size_t smem = sizeof(float) * n * 65;
        if (n == 20) emission_exact_kernel<20><<<grid, dim3(32, 8), smem, st>>>(n, dj, ord);
        else if (n == 4) emission_exact_kernel<4><<<grid, dim3(32, 8), smem, st>>>(n, dj, ord);
        else if (n == 61) emission_exact_kernel<61><<<grid, dim3(32, 8), smem, st>>>(n, dj, ord);
        else emission_exact_kernel<0><<<grid, dim3(32, 8), smem, st>>>(n, dj, ord);
        PGM_CUDA(cudaGetLastError());
    }
    return 0;
}
