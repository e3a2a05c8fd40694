// hp1_tensor.cu -- K4 (tensor flavour): the emission-score matrix as a tcgen05 tensor-core GEMM.
//
// Replaces precomputeScores (reference GraphAlign.h:145-163) + ls_log_add (ls_log.h:22-59), like
// hp1_emission.cu, but the dense contraction  num(y,x) = sum_a g1(a,y) T(a,x),  T = M^T g2,  runs on
// the 5th-generation tensor cores:
//
//   * one CTA = one 128 x 128 tile of S for one job; accumulator in TMEM (128 lanes x 128 fp32 columns)
//   * operands are fp32; the tensor core reads them as TF32 (10-bit mantissa), so each operand is split
//     hi + lo (hi = top 19 bits, lo = remainder) and the three significant products are folded into ONE
//     GEMM by concatenating along K:  A' = [A_hi | A_hi | A_lo],  B' = [B_hi ; B_lo ; B_hi]
//     (K' = 3n padded to a multiple of 8: 16 / 64 / 184 for DNA / protein / codon), fp32 accumulation.
//     Dropped term lo*lo ~ 2^-22 relative: S agrees with the exact flavour to ~1e-6 in log2 units
//     (tolerance stated and tested in tests/test_gpu_hp1.py).
//   * operands are staged in shared memory in the canonical K-major no-swizzle UMMA layout
//     (8-row x 16-byte core matrices; LBO = stride between K chunks, SBO = stride between 8-row
//     groups), written with plain stores + fence.proxy.async -- K is tiny, so TMA would buy nothing
//   * one elected thread issues K'/8 tcgen05.mma (kind::tf32, M=128, N=128, K=8) and commits to an
//     mbarrier; the 4 warps then pull their 32 TMEM lanes with tcgen05.ld (32x32b.x32), apply the
//     epilogue  S = ls_log2( num / (u(y) v(x)) ) + match_init  with the reference's rational log2 and
//     store their row's 32 columns as eight 16-byte stores (S is row-major by graph-1 node, which is
//     exactly the TMEM lane = row ownership, so no transpose is needed).
//
// The GEMM has arithmetic intensity n/2 flop per output byte, so it is bound by the S write, not by the
// tensor pipe; the point of the tensor cores here is that the 2*n flop/cell leave the
// CUDA cores, which the wavefront kernel needs.
#include "internal.h"

#define ET_M 128
#define ET_N 128
#define ET_THREADS 128

__device__ __forceinline__ float ls_log2_add_t(float x, float add) {
    // ls_log.h:30-56 (vector body): r = (b/(d-c) + add) + (float(e) + a)
    const float a = 2.44247459618085927548717403238913328776812604856113966238617812902399112761292613763080658235564f;
    const float b = -4.2040783745848554315883301529007786406310628696382695994938550046831869207082846248658671f;
    const float c = -0.72123729809042963774358701619456664388406302428056983119308906451199556380646306f;
    const uint32_t bits = __float_as_uint(x);
    const float e = (float)((int)(bits >> 23) - 126);
    const float d = __uint_as_float(((bits << 9) >> 9) ^ 0x3f000000u);
    float vd = __fdividef(b, __fsub_rn(d, c));                // tolerance flavour: fast divide
    const float ve = __fadd_rn(e, a);
    vd = __fadd_rn(vd, add);
    return __fadd_rn(vd, ve);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle: element (row r, k) of an R-row operand, 16-byte chunk c = k/4:
//   byte = c * (R/8 * 128) + (r/8) * 128 + (r%8) * 16 + (k%4) * 4
__device__ __forceinline__ uint32_t op_offset(int r, int k, int R) {
    return (uint32_t)((k >> 2) * (R * 16) + (r >> 3) * 128 + (r & 7) * 16 + (k & 3) * 4);
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);               // start address, bits [0,14)
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;     // leading byte offset, bits [16,30)
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;     // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                               // descriptor version (Blackwell)
    // base_offset = 0, lbo_mode = 0, layout_type = 0 (SWIZZLE_NONE / interleave)
    return d;
}

// grid = (tiles_x * tiles_y of the largest job, n_jobs); block = 128
template <int KP>
__global__ void __launch_bounds__(ET_THREADS)
emission_tensor_kernel(int n, int tiles_x, const GraphJobDev* __restrict__ jobs) {
    const GraphJobDev& J = jobs[blockIdx.y];
    const int N1 = J.N1, N2 = J.N2;
    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
    const int y0 = ty * ET_M, x0 = tx * ET_N;
    if (y0 >= N1 || x0 >= N2) return;

    extern __shared__ __align__(128) unsigned char smem[];
    float* sA = (float*)smem;                                   // ET_M x KP, canonical layout
    float* sB = (float*)(smem + (size_t)ET_M * KP * 4);         // ET_N x KP
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base_slot;
    __shared__ float sRV[ET_N];                                  // 1 / v(x)

    const int tid = threadIdx.x, warp = tid >> 5;

    // ---- stage operands (hi/lo split, K-concatenated); node vectors are n contiguous floats: coalesced ----
    {
        const int rowsA = min(ET_M, N1 - y0), rowsB = min(ET_N, N2 - x0);
        const float* gA = J.sites1 + (size_t)n * y0;
        const float* gB = J.Tbuf + (size_t)n * x0;
        for (int i = tid; i < ET_M * n; i += ET_THREADS) {
            const int r = i / n, a = i - r * n;
            const float v = (r < rowsA) ? gA[i] : 0.f;
            const float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
            const float lo = __fsub_rn(v, hi);
            *(float*)((char*)sA + op_offset(r, a, ET_M)) = hi;
            *(float*)((char*)sA + op_offset(r, n + a, ET_M)) = hi;
            *(float*)((char*)sA + op_offset(r, 2 * n + a, ET_M)) = lo;
        }
        for (int i = tid; i < ET_N * n; i += ET_THREADS) {
            const int r = i / n, a = i - r * n;
            const float v = (r < rowsB) ? gB[i] : 0.f;
            const float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
            const float lo = __fsub_rn(v, hi);
            *(float*)((char*)sB + op_offset(r, a, ET_N)) = hi;
            *(float*)((char*)sB + op_offset(r, n + a, ET_N)) = lo;
            *(float*)((char*)sB + op_offset(r, 2 * n + a, ET_N)) = hi;
        }
        for (int i = tid; i < ET_M * (KP - 3 * n); i += ET_THREADS) {      // zero the K padding
            const int r = i / (KP - 3 * n), k = 3 * n + (i - r * (KP - 3 * n));
            *(float*)((char*)sA + op_offset(r, k, ET_M)) = 0.f;
            *(float*)((char*)sB + op_offset(r, k, ET_N)) = 0.f;
        }
        const int x = x0 + tid;
        sRV[tid] = (x < N2) ? __frcp_rn(J.v[x]) : 1.f;
    }
    // ---- TMEM allocation (warp 0) + mbarrier init ----
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_slot)), "n"(ET_N));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;\n");
    }
    // generic-proxy smem writes -> visible to the async proxy (tensor core reads)
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem_base = tmem_base_slot;

    // ---- MMA issue (one thread) ----
    if (tid == 0) {
        // instruction descriptor: D = F32 (1<<4), A = B = TF32 (2<<7, 2<<10), K-major both, N>>3 at [17,23), M>>4 at [24,29)
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(ET_N >> 3) << 17) | ((uint32_t)(ET_M >> 4) << 24);
        const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
        const uint32_t lboA = ET_M * 16, lboB = ET_N * 16, sbo = 128;
#pragma unroll 1
        for (int ks = 0; ks < KP / 8; ++ks) {
            const uint64_t adesc = make_desc(a_base + (uint32_t)(2 * ks) * lboA, lboA, sbo);
            const uint64_t bdesc = make_desc(b_base + (uint32_t)(2 * ks) * lboB, lboB, sbo);
            const uint32_t accum = ks > 0 ? 1u : 0u;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_base), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&mbar)) : "memory");
    }
    // ---- wait for the accumulator ----
    {
        uint32_t done = 0, spins = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0u) : "memory");
            if (++spins > (1u << 22)) { J.result[0] = PGM_ERR_CUDA; break; }   // never hang the GPU on a lost commit
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");

    // ---- epilogue: TMEM -> registers -> log-odds -> 16-byte row-major stores ----
    const int y = y0 + tid;
    const float ruy = (y < N1) ? __frcp_rn(J.u[y]) : 1.f;
    const float add = J.sc[2];
    const int ldS = J.ldS;
    const bool warp_has_rows = (y0 + warp * 32) < N1;            // warp-uniform: skip fully out-of-range row groups
#pragma unroll 1
    for (int cb = 0; cb < ET_N / 32; ++cb) {
        if (!warp_has_rows || x0 + cb * 32 >= N2) break;
        uint32_t r[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(cb * 32);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
            "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
              "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
              "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
              "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        if (y < N1) {
            float* dst = J.S + (size_t)y * ldS + x0 + cb * 32;
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
                if (x0 + cb * 32 + j4 * 4 < ldS) {                 // ldS is a multiple of 4: whole float4 in range
                    float4 o;
                    o.x = ls_log2_add_t(__uint_as_float(r[j4 * 4 + 0]) * ruy * sRV[cb * 32 + j4 * 4 + 0], add);
                    o.y = ls_log2_add_t(__uint_as_float(r[j4 * 4 + 1]) * ruy * sRV[cb * 32 + j4 * 4 + 1], add);
                    o.z = ls_log2_add_t(__uint_as_float(r[j4 * 4 + 2]) * ruy * sRV[cb * 32 + j4 * 4 + 2], add);
                    o.w = ls_log2_add_t(__uint_as_float(r[j4 * 4 + 3]) * ruy * sRV[cb * 32 + j4 * 4 + 3], add);
                    *reinterpret_cast<float4*>(dst + j4 * 4) = o;
                }
            }
        }
    }
    // ---- teardown ----
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "n"(ET_N));
}

// prep kernel of the exact flavour (T = M^T g2, u, v) lives in hp1_emission.cu
int launch_emission_prep(cudaStream_t st, int n, int n_jobs, const GraphJobDev* d_jobs, int maxN);

template <int KP>
static int launch_t(cudaStream_t st, int n, int n_jobs, const GraphJobDev* d_jobs, int tiles_x, int tiles_y) {
    const size_t smem = (size_t)(ET_M + ET_N) * KP * 4;
    const size_t need = smem;
    PGM_CUDA(cudaFuncSetAttribute(emission_tensor_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
    dim3 grid(tiles_x * tiles_y, n_jobs);
    emission_tensor_kernel<KP><<<grid, ET_THREADS, need, st>>>(n, tiles_x, d_jobs);
    PGM_CUDA(cudaGetLastError());
    return 0;
}

int launch_emission_tensor(cudaStream_t st, int n, int n_jobs, const GraphJobDev* d_jobs, const GraphJobDev* h_jobs) {
    if (n_jobs == 0) return 0;
    int mx1 = 0, mx2 = 0;
    for (int k = 0; k < n_jobs; ++k) { mx1 = mx1 > h_jobs[k].N1 ? mx1 : h_jobs[k].N1; mx2 = mx2 > h_jobs[k].N2 ? mx2 : h_jobs[k].N2; }
    int rc = launch_emission_prep(st, n, n_jobs, d_jobs, mx1 > mx2 ? mx1 : mx2);
    if (rc) return rc;
    const int tiles_y = (mx1 + ET_M - 1) / ET_M, tiles_x = (mx2 + ET_N - 1) / ET_N;
    const int kp = (3 * n + 7) & ~7;
    if (kp <= 16) return launch_t<16>(st, n, n_jobs, d_jobs, tiles_x, tiles_y);
    if (kp <= 64) return launch_t<64>(st, n, n_jobs, d_jobs, tiles_x, tiles_y);
    if (kp <= 184) return launch_t<184>(st, n, n_jobs, d_jobs, tiles_x, tiles_y);
    pgm_set_error("tensor emission supports n_states <= 61");
    return PGM_ERR_ARG;
}
