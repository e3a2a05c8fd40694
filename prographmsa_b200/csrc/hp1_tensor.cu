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

// ---- mbarrier / bulk-copy helpers ---------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
// bounded wait: a lost arrival must never hang the GPU; returns false on timeout
__device__ __forceinline__ bool mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t done = 0, spins = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
        if (++spins > (1u << 24)) return false;
    }
    return true;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

// Operand images: the prep kernel writes, per 128-row tile, the exact shared-memory image of the UMMA
// operand (hi/lo split, K-concatenated, canonical K-major no-swizzle layout), so the GEMM kernel stages a
// tile with ONE bulk async copy (cp.async.bulk + mbarrier complete_tx) and no register traffic.
//   block (tile, job): tiles [0, tiles_y) are A tiles (graph 1 rows), the rest B tiles (graph 2 columns)
template <int KP>
__global__ void __launch_bounds__(ET_M)
emission_images_kernel(int n, int tiles_y, int tiles_x, const GraphJobDev* __restrict__ jobs, const int* __restrict__ order) {
    const GraphJobDev& J = jobs[order ? order[blockIdx.y] : blockIdx.y];
    extern __shared__ float sm[];                     // Mf (n*n) + pif (n)
    float* sM = sm; float* sPi = sm + n * n;
    for (int i = threadIdx.x; i < n * n; i += blockDim.x) sM[i] = J.Mf[i];
    for (int i = threadIdx.x; i < n; i += blockDim.x) sPi[i] = J.pif[i];
    __syncthreads();
    const int t = blockIdx.x, r = threadIdx.x;
    const bool isA = t < tiles_y;
    const int tile = isA ? t : t - tiles_y;
    const int N = isA ? J.N1e : J.N2e;
    if (tile * ET_M >= N) return;
    const int node = tile * ET_M + r;
    char* img = (char*)(isA ? J.Aimg : J.Bimg) + (size_t)tile * ET_M * KP * 4;
    // The node's profile column sits in registers (NMAX = KP / 3 >= n is a compile-time bound, loops are unrolled and
    // guarded by n), so T = M^T g runs as n independent sequential-k chains the scheduler can interleave.
    constexpr int NMAX = KP / 3;
    float g[NMAX];
    const float* gp = (isA ? J.sites1 : J.sites2) + (size_t)n * node;
#pragma unroll
    for (int a = 0; a < NMAX; ++a) g[a] = (node < N && a < n) ? gp[a] : 0.f;
    if (node < N) {                                   // g^T pi in Eigen's gemv order (see dot_gemv_order in internal.h)
        float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
        const int n4 = n & ~3;
#pragma unroll
        for (int k = 0; k + 3 < NMAX; k += 4)
            if (k < n4) {
                l0 = __fadd_rn(l0, __fmul_rn(g[k], sPi[k])); l1 = __fadd_rn(l1, __fmul_rn(g[k + 1], sPi[k + 1]));
                l2 = __fadd_rn(l2, __fmul_rn(g[k + 2], sPi[k + 2])); l3 = __fadd_rn(l3, __fmul_rn(g[k + 3], sPi[k + 3]));
            }
        float sd = __fadd_rn(__fadd_rn(l0, l2), __fadd_rn(l1, l3));
#pragma unroll
        for (int k = 0; k < NMAX; ++k) if (k >= n4 && k < n) sd = __fadd_rn(sd, __fmul_rn(g[k], sPi[k]));
        if (isA) J.u[node] = sd; else J.v[node] = sd;
    }
#pragma unroll
    for (int a = 0; a < NMAX; ++a) {
        if (a < n) {
            float v = g[a];
            if (!isA) {                               // T(a,x) = sum_b M(b,a) g2(b,x), sequentially in b
                float sacc = 0.f;
#pragma unroll
                for (int bb = 0; bb < NMAX; ++bb) if (bb < n) sacc = __fadd_rn(sacc, __fmul_rn(sM[bb * n + a], g[bb]));
                v = sacc;
            }
            const float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
            const float lo = __fsub_rn(v, hi);
            // A' = [hi | hi | lo],  B' = [hi ; lo ; hi]
            *(float*)(img + op_offset(r, a, ET_M)) = hi;
            *(float*)(img + op_offset(r, n + a, ET_M)) = isA ? hi : lo;
            *(float*)(img + op_offset(r, 2 * n + a, ET_M)) = isA ? lo : hi;
        }
    }
    for (int k = 3 * n; k < KP; ++k) *(float*)(img + op_offset(r, k, ET_M)) = 0.f;
}

// Persistent GEMM + epilogue: one CTA per (job, 128-row tile of graph 1), looping over graph 2's 128-column
// tiles.  Warp roles: 0 = bulk-copy producer, 1 = TMEM owner + single-thread MMA issuer, 2..5 = epilogue
// (TMEM lanes 32*(warp%4)).  B operand tiles and TMEM accumulators are double-buffered so the copy of tile
// i+1, the MMA of tile i+1 and the epilogue of tile i overlap.
#define ETP_THREADS 192
template <int KP, int BST>
__global__ void __launch_bounds__(ETP_THREADS)
emission_tensor_kernel(int tiles_y, const GraphJobDev* __restrict__ jobs, const int* __restrict__ order) {
    const GraphJobDev& J = jobs[order ? order[blockIdx.y] : blockIdx.y];
    const int N1 = J.N1e, N2 = J.N2e;
    const int y0 = blockIdx.x * ET_M;
    if (y0 >= N1) return;
    const int ntx = (N2 + ET_N - 1) / ET_N;
    constexpr uint32_t IMG = ET_M * KP * 4;

    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* sA = smem;
    unsigned char* sB = smem + IMG;                               // BST stages
    __shared__ __align__(8) uint64_t a_full, b_full[2], b_empty[2], t_full[2], t_empty[2];
    __shared__ uint32_t tmem_base_slot;
    __shared__ float sRV[2][ET_N];
    __shared__ int s_fail;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        mbar_init(&a_full, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); mbar_init(&t_full[s], 1); mbar_init(&t_empty[s], 4); }
        s_fail = 0;
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_slot)), "n"(2 * ET_N));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem_base = tmem_base_slot;
    bool ok = true;

    if (warp == 0) {
        if (lane == 0) {                                          // ---- producer ----
            mbar_expect_tx(&a_full, IMG);
            bulk_g2s(sA, (const char*)J.Aimg + (size_t)blockIdx.x * IMG, IMG, &a_full);
            for (int i = 0; i < ntx && ok; ++i) {
                const int s = i % BST;
                if (i >= BST) ok = mbar_wait(&b_empty[s], ((i / BST) - 1) & 1);
                mbar_expect_tx(&b_full[s], IMG);
                bulk_g2s(sB + (size_t)s * IMG, (const char*)J.Bimg + (size_t)i * IMG, IMG, &b_full[s]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                          // ---- MMA issuer ----
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(ET_N >> 3) << 17) | ((uint32_t)(ET_M >> 4) << 24);
            const uint32_t lbo = ET_M * 16, sbo = 128;
            ok = mbar_wait(&a_full, 0);
            for (int i = 0; i < ntx && ok; ++i) {
                const int s = i % BST, ts = i & 1;
                ok = mbar_wait(&b_full[s], (i / BST) & 1);
                if (ok && i >= 2) ok = mbar_wait(&t_empty[ts], ((i >> 1) - 1) & 1);
                if (!ok) break;
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB + (size_t)s * IMG);
                const uint32_t d_tmem = tmem_base + (uint32_t)(ts * ET_N);
#pragma unroll 1
                for (int ks = 0; ks < KP / 8; ++ks) {
                    const uint64_t adesc = make_desc(a_base + (uint32_t)(2 * ks) * lbo, lbo, sbo);
                    const uint64_t bdesc = make_desc(b_base + (uint32_t)(2 * ks) * lbo, lbo, sbo);
                    const uint32_t accum = ks > 0 ? 1u : 0u;
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
                }
                umma_commit(&b_empty[s]);                         // shared-memory stage free once these MMAs retire
                umma_commit(&t_full[ts]);                         // accumulator ready
            }
        }
    } else {                                                      // ---- epilogue warps 2..5 ----
        const int q = warp & 3;                                   // TMEM lane quarter this warp may touch
        const int row = q * 32 + lane, y = y0 + row;
        const float ruy = (y < N1) ? __frcp_rn(J.u[y]) : 1.f;
        const float add = J.sc[2];
        const int ldS = J.ldS;
        const bool warp_has_rows = (y0 + q * 32) < N1;
        const int et = tid - 64;                                  // 0..127
        for (int i = 0; i < ntx && ok; ++i) {
            const int ts = i & 1, x0 = i * ET_N;
            { const int x = x0 + et; sRV[ts][et] = (x < N2) ? __frcp_rn(J.v[x]) : 1.f; }
            asm volatile("bar.sync 1, 128;\n" ::: "memory");
            ok = mbar_wait(&t_full[ts], (i >> 1) & 1);
            if (!ok) break;
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll 1
            for (int cb = 0; cb < ET_N / 32; ++cb) {
                if (!warp_has_rows || x0 + cb * 32 >= N2) break;
                uint32_t r[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ts * ET_N + cb * 32);
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
                        if (x0 + cb * 32 + j4 * 4 < ldS) {         // ldS is a multiple of 4: whole float4 in range
                            float4 o;
                            o.x = ls_log2_add_t(__uint_as_float(r[j4 * 4 + 0]) * ruy * sRV[ts][cb * 32 + j4 * 4 + 0], add);
                            o.y = ls_log2_add_t(__uint_as_float(r[j4 * 4 + 1]) * ruy * sRV[ts][cb * 32 + j4 * 4 + 1], add);
                            o.z = ls_log2_add_t(__uint_as_float(r[j4 * 4 + 2]) * ruy * sRV[ts][cb * 32 + j4 * 4 + 2], add);
                            o.w = ls_log2_add_t(__uint_as_float(r[j4 * 4 + 3]) * ruy * sRV[ts][cb * 32 + j4 * 4 + 3], add);
                            *reinterpret_cast<float4*>(dst + j4 * 4) = o;
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_empty[ts]);             // 4 arrivals (one per epilogue warp) free the accumulator
        }
    }
    if (!ok) s_fail = 1;
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "n"(2 * ET_N));
    if (tid == 0 && s_fail) J.result[4] = PGM_ERR_CUDA;           // emission status: the traceback kernel turns it into the job's status (a failed job, never a hang)
}

template <int KP, int BST>
static int launch_t(cudaStream_t st, int n, int n_jobs, const GraphJobDev* d_jobs, const int* d_order, int tiles_x, int tiles_y) {
    const size_t smem = (size_t)(1 + BST) * ET_M * KP * 4;
    PGM_CUDA(cudaFuncSetAttribute(emission_tensor_kernel<KP, BST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int j0 = 0; j0 < n_jobs; j0 += 65535) {                   // gridDim.y holds at most 65535 jobs
        const int nj = n_jobs - j0 < 65535 ? n_jobs - j0 : 65535;
        const GraphJobDev* dj = d_order ? d_jobs : d_jobs + j0;
        const int* ord = d_order ? d_order + j0 : nullptr;
        {
            dim3 grid(tiles_y + tiles_x, nj);
            emission_images_kernel<KP><<<grid, ET_M, sizeof(float) * (n * n + n), st>>>(n, tiles_y, tiles_x, dj, ord);
            PGM_CUDA(cudaGetLastError());
        }
        dim3 grid(tiles_y, nj);
        emission_tensor_kernel<KP, BST><<<grid, ETP_THREADS, smem, st>>>(tiles_y, dj, ord);
        PGM_CUDA(cudaGetLastError());
    }
    return 0;
}

// d_order: optional list of the n_jobs job indices (relative to d_jobs) this launch covers; mx1 / mx2: their largest graphs
int launch_emission_tensor(cudaStream_t st, int n, int n_jobs, const GraphJobDev* d_jobs, const int* d_order, int mx1, int mx2) {
    if (n_jobs == 0) return 0;
    const int tiles_y = (mx1 + ET_M - 1) / ET_M, tiles_x = (mx2 + ET_N - 1) / ET_N;
    const int kp = (3 * n + 7) & ~7;
    if (kp <= 16) return launch_t<16, 2>(st, n, n_jobs, d_jobs, d_order, tiles_x, tiles_y);
    if (kp <= 64) return launch_t<64, 2>(st, n, n_jobs, d_jobs, d_order, tiles_x, tiles_y);
    if (kp <= 184) return launch_t<184, 1>(st, n, n_jobs, d_jobs, d_order, tiles_x, tiles_y);
    pgm_set_error("tensor emission supports n_states <= 61");
    return PGM_ERR_ARG;
}
