// hp1_wave.cu -- K5, anti-diagonal wavefront fill with ROLE-SORTED threads (the flavour for graphs with many non-chain
// nodes: the upper levels of a guide tree, whose sweeps end a mixed batch).
//
// Same recurrences as hp1_graphdp_diag.cu (reference GraphAlign.h:205-280), same {M,X,Y,W} cell layout (anti-diagonal-major,
// so the dense traceback of hp1_graphdp.cu re-traces the job unchanged), but the work of one anti-diagonal is dealt so that
// no warp ever serialises a fast path with a slow one (ncu on the thread-per-row kernel, profiles/r2_*: half of all
// instructions ran with 1.7 of 32 lanes active and 60 % of the warp samples sat at the barrier waiting for those warps):
//
//   * chain rows (one predecessor, the regular edge from y-1) own a thread each, packed into their own warps;
//   * rows with other / more predecessors own a thread each in SEPARATE warps: the adjacent-edge candidates as above
//     (an absent adjacent edge is the value +inf, which turns its candidates into -inf without a branch), then a loop
//     over the row's extra predecessors;
//   * every NON-CHAIN COLUMN owns L threads (1, 2 or 4) that compute the column's cell of the current anti-diagonal in
//     full -- the predecessor pairs (yp, xp) are dealt over the L lanes and folded with shuffles.  max is exact,
//     associative and idempotent, so the fold order does not change a bit.  Row threads skip these cells.
//
// All candidates of a cell lie on strictly earlier anti-diagonals, so one barrier per anti-diagonal orders everything.
// Cells of the last RING (8 or 16) anti-diagonals are read from a shared-memory ring indexed by row; older ones (edges
// spanning more positions: < 1 % on the bench families with 16) from the {M,X,Y,W} array in L2.
#include "internal.h"
#include <math_constants.h>
#include <algorithm>
#include <cstdlib>

template <int RING, int L>
__global__ void __launch_bounds__(1024)
graph_fill_wave_kernel(const GraphJobDev* __restrict__ jobs, const int* __restrict__ order, int ring_pitch, int max_tr,
                       int max_n1, int max_n2, int max_np1, int max_nx) {
    const GraphJobDev& J = jobs[order[blockIdx.x]];
    const int N1 = J.N1, N2 = J.N2, pitch = J.pitch;
    const float gap_init = J.sc[0], gap_extend = J.sc[1], start_gap = J.sc[6], start_init = J.sc[7];
    const float NINF = -CUDART_INF_F;
    float4* __restrict__ C = J.MXYW;
    const int tid = (int)threadIdx.x, T = (int)blockDim.x;
    const int TR = J.wv_tr, NSC = J.wv_ncols;

    extern __shared__ __align__(16) unsigned char wsm[];
    float4* ring = (float4*)wsm;                                        // [RING][ring_pitch], row y at index y + 1 (index 0: -inf guard)
    float4* sring = ring + (size_t)RING * ring_pitch;                   // [4][max_tr]: S(y, 4q .. 4q+3) of row slot t at [(q & 3)][t]
    NodeInfo* n1 = (NodeInfo*)(sring + (size_t)4 * max_tr);             // [max_n1] graph 1's node infos (column threads)
    WaveCI* ci = (WaveCI*)(n1 + max_n1);                                // [max_n2]
    int* p1i = (int*)(ci + max_n2); float* p1v = (float*)(p1i + max_np1);   // graph 1's predecessor lists
    int* sxi = (int*)(p1v + max_np1); float* sxv = (float*)(sxi + max_nx);  // extra predecessors of the non-chain rows
    uint16_t* rowslot = (uint16_t*)(sxv + max_nx);                      // [max_n1] row -> thread slot
    {
        const float4 ninf4 = make_float4(NINF, NINF, NINF, NINF);
        for (int i = tid; i < RING * ring_pitch; i += T) ring[i] = ninf4;
        for (int i = tid; i < N1; i += T) { n1[i] = J.ni1[i]; rowslot[i] = J.wv_rowslot[i]; }
        for (int i = tid; i < N2; i += T) ci[i] = J.wv_ci[i];
        for (int i = tid; i < J.np1; i += T) { p1i[i] = J.p1_idx[i]; p1v[i] = J.p1_val[i]; }
        for (int i = tid; i < J.wv_nx; i += T) { sxi[i] = J.wv_xi[i]; sxv[i] = J.wv_xv[i]; }
    }
    // ---- roles -------------------------------------------------------------------------------------------------------------
    const bool is_row = tid < TR;
    const int cq = tid - TR;                                            // column threads: L consecutive lanes per column
    const bool is_col = cq >= 0 && cq < NSC * L;
    WaveRow wr = {-1, 0, 0.f, 0, 0.f, 0, 0.f, 0};
    WaveCol wc = {-1, 0, 0, 0.f, 0, 0.f, 0, 0};
    if (is_row) wr = J.wv_rows[tid];
    if (is_col) wc = J.wv_cols[cq / L];
    const int y = wr.y;                                                 // -1: padding slot
    const int xc = wc.x;
    const int lq = cq & (L - 1);
    const int kyq = L == 4 ? (lq >> 1) : 0, kxq = L >= 2 ? (lq & 1) : 0;
    constexpr int SY = L == 4 ? 2 : 1, SX = L >= 2 ? 2 : 1;
    const unsigned gmask = L == 1 ? 0u : (((1u << L) - 1u) << ((tid & 31) & ~(L - 1)));
    const float gix = y == 0 ? start_gap : gap_init;                    // row 0: GraphAlign.h:224-234 opens with start_gap
    float4 left = make_float4(NINF, NINF, NINF, NINF);
    const float* __restrict__ S = J.S;
    const int ldS = J.ldS;
    auto fetch_s4 = [&](int xq) {                                       // S(y, xq .. xq+3) -> this slot's ring, asynchronously
        const float* src = S + (size_t)y * ldS + xq;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sring + ((xq >> 2) & 3) * max_tr + tid);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src) : "memory");
    };
    if (y >= 1) { fetch_s4(0); if (4 < ldS) fetch_s4(4); }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    __syncthreads();

    auto cell = [&](int yy, int xx, int d) -> float4 {                  // any earlier cell: ring when recent, else L2
        const int dd = yy + xx;
        if (d - dd < RING) return ring[(dd & (RING - 1)) * ring_pitch + yy + 1];
        return C[(size_t)dd * pitch + (yy - max(0, dd - (N2 - 1)))];
    };
    const int dmax = (N1 - 2) + (N2 - 2);
    for (int d = 0; d <= dmax; ++d) {
        const int ylo = max(0, d - (N2 - 1));
        float4* rcur = ring + (d & (RING - 1)) * ring_pitch;
        const float4* row1 = ring + ((d - 1) & (RING - 1)) * ring_pitch;
        const float4* row2 = ring + ((d - 2) & (RING - 1)) * ring_pitch;
        if (is_row) {
            const int x = d - y;
            if (y >= 1 && x >= 0 && (x & 3) == 0 && x + 8 < ldS) fetch_s4(x + 8);
            asm volatile("cp.async.commit_group;\n" ::: "memory");
            asm volatile("cp.async.wait_group 7;\n" ::: "memory");     // everything issued >= 8 anti-diagonals ago has landed
            if (y >= 0 && (unsigned)x <= (unsigned)(N2 - 2)) {
                const WaveCI c = ci[x];
                if (!(c.flags & 1)) {                                   // a non-chain column's cell belongs to its column threads
                    const float s = reinterpret_cast<const float*>(sring + ((x >> 2) & 3) * max_tr + tid)[x & 3];
                    const float4 cu = row1[y];                          // (y-1, x)
                    const float4 cd = row2[y];                          // (y-1, x-1)
                    float4 cl = left;                                   // (y, x-1): this thread, last anti-diagonal ...
                    if (c.flags & 2) cl = row1[y + 1];                  // ... unless that column's threads made it
                    const float ex = c.ex_adj;
                    const float giy = x == 0 ? start_gap : gap_init;    // column 0: GraphAlign.h:212-222
                    float ys = __fsub_rn(fmaxf(__fadd_rn(cu.z, gap_extend), __fadd_rn(cu.w, giy)), wr.ey_adj);
                    float m = __fsub_rn(__fsub_rn(__fadd_rn(cd.w, s), wr.ey_adj), ex);
                    float xs = __fsub_rn(fmaxf(__fadd_rn(cl.y, gap_extend), __fadd_rn(cl.w, gix)), ex);
                    for (int k = 0; k < wr.nextra; ++k) {               // non-chain rows (their own warps)
                        const int yp = k == 0 ? wr.yp0 : (k == 1 ? wr.yp1 : sxi[wr.rest + k]);
                        const float ey = k == 0 ? wr.ey0 : (k == 1 ? wr.ey1 : sxv[wr.rest + k]);
                        const float4 cu2 = cell(yp, x, d);
                        ys = fmaxf(ys, __fsub_rn(fmaxf(__fadd_rn(cu2.z, gap_extend), __fadd_rn(cu2.w, giy)), ey));
                        if (x > 0) {
                            const float4 cd2 = cell(yp, x - 1, d);
                            m = fmaxf(m, __fsub_rn(__fsub_rn(__fadd_rn(cd2.w, s), ey), ex));
                        }
                    }
                    if (x == 0) { m = NINF; xs = NINF; }
                    if (y == 0) { m = NINF; ys = NINF; }
                    float w = fmaxf(m, fmaxf(xs, ys));
                    if (d == 0) w = start_init;                         // cell (0, 0)
                    const float4 out = make_float4(m, xs, ys, w);
                    left = out;
                    rcur[y + 1] = out;
                    C[(size_t)d * pitch + (y - ylo)] = out;
                }
            }
        } else {
            asm volatile("cp.async.commit_group;\n" ::: "memory");
            if (is_col) {
                const int yy = d - xc;
                if (xc >= 0 && (unsigned)yy <= (unsigned)(N1 - 2)) {   // the same for the L lanes of a column
                    float m = NINF, xs = NINF, ys = NINF;
                    const int c2 = wc.cnt;
                    if (yy == 0) {                                      // first row: X over the column's predecessors, start_gap
                        if (kyq == 0)
                            for (int kx = kxq; kx < c2; kx += SX) {
                                const int xp = kx == 0 ? wc.xp0 : (kx == 1 ? wc.xp1 : J.p2_idx[wc.ptr + kx]);
                                const float ex = kx == 0 ? wc.ex0 : (kx == 1 ? wc.ex1 : J.p2_val[wc.ptr + kx]);
                                const float4 cl = cell(0, xp, d);
                                xs = fmaxf(xs, __fsub_rn(fmaxf(__fadd_rn(cl.y, gap_extend), __fadd_rn(cl.w, start_gap)), ex));
                            }
                    } else {
                        const NodeInfo ny = n1[yy];
                        const int c1 = NI_CNT(ny.cnt);
                        const float s = reinterpret_cast<const float*>(sring + ((xc >> 2) & 3) * max_tr + rowslot[yy])[xc & 3];
                        if (kyq == 0)
                            for (int kx = kxq; kx < c2; kx += SX) {
                                const int xp = kx == 0 ? wc.xp0 : (kx == 1 ? wc.xp1 : J.p2_idx[wc.ptr + kx]);
                                const float ex = kx == 0 ? wc.ex0 : (kx == 1 ? wc.ex1 : J.p2_val[wc.ptr + kx]);
                                const float4 cl = cell(yy, xp, d);
                                xs = fmaxf(xs, __fsub_rn(fmaxf(__fadd_rn(cl.y, gap_extend), __fadd_rn(cl.w, gap_init)), ex));
                            }
                        for (int ky = kyq; ky < c1; ky += SY) {
                            const int yp = ky ? p1i[ny.ptr + ky] : ny.idx0;
                            const float ey = ky ? p1v[ny.ptr + ky] : ny.val0;
                            if (kxq == 0) {
                                const float4 cu = cell(yp, xc, d);
                                ys = fmaxf(ys, __fsub_rn(fmaxf(__fadd_rn(cu.z, gap_extend), __fadd_rn(cu.w, gap_init)), ey));
                            }
                            for (int kx = kxq; kx < c2; kx += SX) {
                                const int xp = kx == 0 ? wc.xp0 : (kx == 1 ? wc.xp1 : J.p2_idx[wc.ptr + kx]);
                                const float ex = kx == 0 ? wc.ex0 : (kx == 1 ? wc.ex1 : J.p2_val[wc.ptr + kx]);
                                const float4 cd = cell(yp, xp, d);
                                m = fmaxf(m, __fsub_rn(__fsub_rn(__fadd_rn(cd.w, s), ey), ex));
                            }
                        }
                    }
                    if (L > 1) {
#pragma unroll
                        for (int o = 1; o < L; o <<= 1) {
                            m = fmaxf(m, __shfl_xor_sync(gmask, m, o));
                            xs = fmaxf(xs, __shfl_xor_sync(gmask, xs, o));
                            ys = fmaxf(ys, __shfl_xor_sync(gmask, ys, o));
                        }
                    }
                    if (lq == 0) {
                        const float4 out = make_float4(m, xs, ys, fmaxf(m, fmaxf(xs, ys)));
                        rcur[yy + 1] = out;
                        C[(size_t)d * pitch + (yy - ylo)] = out;
                    }
                }
            }
        }
        __syncthreads();
    }
    // end-node match (GraphAlign.h:262-280)
    if (tid == 0) {
        const float end_match = J.sc[3], end_gap = J.sc[4], end_skip = J.sc[5];
        float Wend = NINF;
        for (int ky = J.p1_ptr[N1 - 1]; ky < J.p1_ptr[N1]; ++ky)
            for (int kx = J.p2_ptr[N2 - 1]; kx < J.p2_ptr[N2]; ++kx) {
                const int yp = J.p1_idx[ky], xp = J.p2_idx[kx];
                const float ey = J.p1_val[ky], ex = J.p2_val[kx];
                if (xp == 0 && yp == 0) {
                    Wend = fmaxf(__fsub_rn(__fsub_rn(end_skip, ey), ex), Wend);
                } else {
                    const int dd = yp + xp;
                    const float4 c = __ldcg(C + (size_t)dd * pitch + (yp - max(0, dd - (N2 - 1))));
                    Wend = fmaxf(__fsub_rn(__fsub_rn(__fadd_rn(c.y, end_gap), ey), ex), Wend);
                    Wend = fmaxf(__fsub_rn(__fsub_rn(__fadd_rn(c.z, end_gap), ey), ex), Wend);
                    Wend = fmaxf(__fsub_rn(__fsub_rn(__fadd_rn(c.x, end_match), ey), ex), Wend);
                }
            }
        ((float*)J.result)[3] = Wend;
    }
}

size_t wave_smem_bytes(int ring, int max_n1, int max_n2, int max_tr, int max_np1, int max_nx) {
    const size_t ring_pitch = (size_t)((max_n1 + 1 + 3) & ~3);
    return (size_t)ring * ring_pitch * 16 + (size_t)4 * max_tr * 16 + (size_t)max_n1 * sizeof(NodeInfo) + (size_t)max_n2 * sizeof(WaveCI) +
           (size_t)max_np1 * 8 + (size_t)max_nx * 8 + (size_t)max_n1 * 2 + 64;
}

template <int RING, int L>
static int launch_wave_t(cudaStream_t st, int n_jobs, const GraphJobDev* d_jobs, const int* d_order, int threads, int max_n1, int max_n2,
                         int max_tr, int max_np1, int max_nx) {
    const size_t smem = wave_smem_bytes(RING, max_n1, max_n2, max_tr, max_np1, max_nx);
    const int ring_pitch = (max_n1 + 1 + 3) & ~3;
    PGM_CUDA(cudaFuncSetAttribute(graph_fill_wave_kernel<RING, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    graph_fill_wave_kernel<RING, L><<<n_jobs, threads, smem, st>>>(d_jobs, d_order, ring_pitch, max_tr, max_n1, max_n2, max_np1, max_nx);
    PGM_CUDA(cudaGetLastError());
    return 0;
}

// threads: the launch's largest wv_tr + 32 * ceil(wv_ncols * L / 32); ring: 8 or 16 anti-diagonals (the caller checked the
// shared memory with wave_smem_bytes)
int launch_graph_fill_wave(cudaStream_t st, int L, int ring, int threads, int n_jobs, const GraphJobDev* d_jobs, const int* d_order, int max_n1,
                           int max_n2, int max_tr, int max_np1, int max_nx) {
    if (n_jobs == 0) return 0;
    if (threads > 1024 || wave_smem_bytes(ring, max_n1, max_n2, max_tr, max_np1, max_nx) > 227 * 1024) {
        pgm_set_error("wave fill: launch does not fit (threads %d)", threads);
        return PGM_ERR_ARG;
    }
#define GO(R_, L_) return launch_wave_t<R_, L_>(st, n_jobs, d_jobs, d_order, threads, max_n1, max_n2, max_tr, max_np1, max_nx)
    if (ring == 16) { if (L == 4) GO(16, 4); if (L == 2) GO(16, 2); GO(16, 1); }
    if (L == 4) GO(8, 4); if (L == 2) GO(8, 2); GO(8, 1);
#undef GO
}
