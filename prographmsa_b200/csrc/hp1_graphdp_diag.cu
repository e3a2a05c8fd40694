// hp1_graphdp_diag.cu -- K5, latency-optimised flavour: the anti-diagonal wavefront fill (one thread per graph-1 row,
// one cell per thread and barrier).  It does ~3x the instructions per cell of the blocked kernel in hp1_graphdp.cu, but a
// job's sweep is N1 + N2 short steps with hundreds of threads at work, and multi-predecessor cells cost it almost nothing,
// so it is the faster one for graphs with many non-chain nodes (the upper levels of a guide tree) and for small batches.
// hp1_host.cu picks the flavour per job (GraphJobDev::layout); the traceback reads either cell layout.
//
// Recurrences, evaluation order and the reference lines are those of hp1_graphdp.cu (GraphAlign.h:205-280).
#include "internal.h"
#include <math_constants.h>
#include <algorithm>
#include <type_traits>
#include <cstdlib>

__device__ __forceinline__ size_t skew(int y, int x, int N2, int pitch) {
    const int d = y + x;
    const int ylo = max(0, d - (N2 - 1));
    return (size_t)d * pitch + (y - ylo);
}

#define GF_RING 8            // anti-diagonals of {M,X,Y,W} kept in shared memory (power of two): covers skip edges spanning <= 5 nodes
#define GF_RING_BAND 8       // banded flavour (16 or 32 diagonals: -5 % where graph 2's node infos still fit in shared memory beside the ring, +10 % where they no longer do)
#define GF_RING_LONG 4       // ... for graphs too long for 8 (>= 4: rows d, d-1, d-2 are live at once); S is then read directly
// S is row-major by y; each (thread,row) streams its own row 16 bytes (4 columns) at a time with cp.async,
// 8 columns (= 8 anti-diagonals) ahead, into a private ring of four float4 slots

// Thread <-> ROW mapping: thread t owns rows y = t + r*blockDim (r < RPT).  Along the sweep the same
// thread meets (y,0), (y,1), ... on consecutive anti-diagonals, so the LEFT neighbour (y, x-1) and the
// row's own predecessor info stay in registers; the UP / DIAGONAL neighbours (y-1, .) come from the
// previous two ring rows at index y-1 (no skew arithmetic).  STAGE: graph 2's node infos in shared memory.
//
// BANDED (graph 1 too long for one CTA's ring): the rows are cut into bands of blockDim rows, one CTA per band, all
// co-resident.  Every dependency of a cell lies on a strictly earlier global anti-diagonal, so band b may sweep
// diagonal d once every earlier band has finished its cells on diagonals < d: each band publishes the number of
// diagonals it has completed (fence + flag, every 4 diagonals) and waits on the earlier bands' flags (every 8).
// Cells of other bands are read from HBM/L2 (the {M,X,Y,W} array itself is the exchange buffer); the row just above the
// band, which the band's first row needs on every diagonal, is staged 8 diagonals at a time into shared memory right after
// the flag check (it is complete by then), so that row runs the chain fast path like any other.
template <bool STAGE, int RPT, int RING, bool SPF, bool BANDED, bool DEEP = false>
__global__ void __launch_bounds__(BANDED ? 512 : 1024)
graph_fill_diag_kernel(const GraphJobDev* __restrict__ jobs, const int* __restrict__ order, int ring_pitch, int max_bands) {
    const GraphJobDev& J = jobs[order[BANDED ? blockIdx.x / max_bands : blockIdx.x]];
    // A CTA takes its band index from a per-job atomic ticket, so a band only ever waits on bands whose CTAs have already
    // started: no assumption on the order in which the hardware dispatches CTAs, no co-residency requirement
    __shared__ int s_band;
    if (BANDED) {
        if (threadIdx.x == 0) s_band = atomicAdd(J.progress + J.bands, 1);
        __syncthreads();
    }
    const int band = BANDED ? s_band : 0;
    if (BANDED && band >= J.bands) return;
    const int N1 = J.N1, N2 = J.N2, pitch = J.pitch;
    const int y0 = BANDED ? band * (int)blockDim.x * RPT : 0;                      // first row of this band
    const float gap_init = J.sc[0], gap_extend = J.sc[1], start_gap = J.sc[6], start_init = J.sc[7];
    const float NINF = -CUDART_INF_F;
    float4* __restrict__ C = J.MXYW;
    const float* __restrict__ S = J.S;
    const int T = (int)blockDim.x, tid = (int)threadIdx.x;

    extern __shared__ __align__(16) unsigned char gsm[];
    __shared__ float4 prevrow[16];                                                  // BANDED: cell (y0-1, dd-y0+1) at [dd & 15]
    float4* ring = (float4*)gsm;                                                    // [RING][ring_pitch], indexed by y
    float4* sring = (float4*)(gsm + (size_t)RING * ring_pitch * sizeof(float4));    // SPF: [4][RPT*T], private slots per (thread,row)
    NodeInfo* s2 = (NodeInfo*)(sring + (SPF ? 4 * RPT * T : 0));
    // STAGE: graph 2's node infos and both predecessor lists live in shared memory, so the rare multi-predecessor
    // cell (met by SOME row on almost every anti-diagonal) does not put global-memory latency on the sweep's critical path
    int* sp1i = (int*)(s2 + N2); float* sp1v = (float*)(sp1i + J.np1);
    int* sp2i = (int*)(sp1v + J.np1); float* sp2v = (float*)(sp2i + J.np2);
    if (STAGE) {
        for (int i = tid; i < N2; i += T) s2[i] = J.ni2[i];
        for (int i = tid; i < J.np1; i += T) { sp1i[i] = J.p1_idx[i]; sp1v[i] = J.p1_val[i]; }
        for (int i = tid; i < J.np2; i += T) { sp2i[i] = J.p2_idx[i]; sp2v[i] = J.p2_val[i]; }
        __syncthreads();
    }
    const NodeInfo* __restrict__ I2 = STAGE ? s2 : J.ni2;
    const int* __restrict__ P1I = STAGE ? sp1i : J.p1_idx; const float* __restrict__ P1V = STAGE ? sp1v : J.p1_val;
    const int* __restrict__ P2I = STAGE ? sp2i : J.p2_idx; const float* __restrict__ P2V = STAGE ? sp2v : J.p2_val;

    auto cell = [&](int yy, int xx, int d) -> float4 {          // any earlier cell: ring when recent (and of this band), else HBM/L2
        const int dd = yy + xx;
        if (d - dd < RING && (!BANDED || yy >= y0)) return ring[(dd & (RING - 1)) * ring_pitch + (yy - y0)];
        const float4* p = C + (size_t)dd * pitch + (yy - max(0, dd - (N2 - 1)));
        return BANDED ? __ldcg(p) : *p;
    };

    // BANDED: the loads of a multi-predecessor cell are issued together (up to 4 x 4 predecessor pairs) instead of one L2 round
    // trip per loop iteration: the step of an anti-diagonal takes as long as its slowest cell, and on the deep profiles near
    // the root of a guide tree (a third of the nodes non-chain) some row meets a 3 x 3 cell on every diagonal.  One generic
    // load per value: the ring (shared) or the {M,X,Y,W} array (L2: other bands write it, so never through L1).
    auto cell_addr = [&](int yy, int xx, int d) -> const float4* {
        const int dd = yy + xx;
        const float4* g = C + (size_t)dd * pitch + (yy - max(0, dd - (N2 - 1)));
        const float4* q = ring + (dd & (RING - 1)) * ring_pitch + (yy - y0);
        return (d - dd < RING && yy >= y0) ? q : g;
    };
    auto ld_w = [](const float4* p) -> float {
        if (!BANDED) return reinterpret_cast<const float*>(p)[3];           // one CTA per job: its own cells, through L1 as before
        float v; asm volatile("ld.relaxed.gpu.f32 %0, [%1+12];" : "=f"(v) : "l"(p)); return v;
    };
    auto ld_zw = [](const float4* p) -> float2 {
        if (!BANDED) return reinterpret_cast<const float2*>(p)[1];
        float2 v; asm volatile("ld.relaxed.gpu.v2.f32 {%0,%1}, [%2+8];" : "=f"(v.x), "=f"(v.y) : "l"(p)); return v;
    };
    auto ld_yw = [](const float4* p) -> float2 {
        if (!BANDED) { const float* f = reinterpret_cast<const float*>(p); return make_float2(f[1], f[3]); }
        float4 v; asm volatile("ld.relaxed.gpu.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p)); return make_float2(v.y, v.w);
    };
    constexpr int GB = 4;                                        // predecessors per side the batched path holds (measured: 2 -> the top levels 178 -> 325 ms, cells with 3-4 predecessors are common; extra 2 x 8 / 8 x 2 shapes -> 195 ms, the variants diverge)
    constexpr bool COOP = BANDED;                                // cells with many predecessor pairs are folded by the whole warp.  (In the one-CTA
                                                                 // flavour the ballot per step cost 10 % on moderately deep jobs, 7.9 -> 8.7 ms at
                                                                 // 629 x 1743, and gained 3 % on the deepest, 53.1 -> 51.2 ms at 566 x 10982.)

    NodeInfo ny[RPT]; bool ysimple[RPT]; float4 left[RPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        const int y = y0 + tid + r * T;
        ny[r] = (y <= N1 - 2) ? J.ni1[y] : NodeInfo{0, 0, 0, 0.f};
        ysimple[r] = (ny[r].cnt == 1 && ny[r].idx0 == y - 1);
        left[r] = make_float4(NINF, NINF, NINF, NINF);
    }
    const int band_rows = T * RPT;
    const int ylast = BANDED ? min(N1 - 2, y0 + band_rows - 1) : N1 - 2;           // last row of this band
    const int dmin = y0, dmax = ylast + (N2 - 2);
    const int ldS = J.ldS;
    auto fetch_s4 = [&](int r, int y, int xq) {                  // S(y, xq..xq+3) -> sring[(xq/4)&3][r*T+tid], asynchronously
        const float* src = S + (size_t)y * ldS + xq;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sring + ((xq >> 2) & 3) * (RPT * T) + r * T + tid);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src) : "memory");
    };
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        const int y = y0 + tid + r * T;
        if (SPF && y >= 1 && y <= N1 - 2) { fetch_s4(r, y, 0); if (4 < ldS) fetch_s4(r, y, 4); }
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");      // the first rows need these within < 8 diagonals

    for (int d = dmin; d <= dmax; ++d) {
        if (BANDED && band > 0 && ((d - dmin) & 7) == 0) {      // the earlier bands must be past the next 8 diagonals
            if (tid == 0) {
                // only the previous band's flag: that band published diagonal d' after it had itself seen ITS predecessor past
                // d' + 8, so by induction every earlier band is past d' too (and the fences make their cells visible in turn);
                // a job of 78 bands otherwise spends 78 serial L2 reads here every 8 diagonals
                const int bq = band - 1;
                const int last_q = min(N1 - 2, (bq + 1) * band_rows - 1) + (N2 - 2);
                const int need = min(d + 8, last_q + 1);
                while (*(volatile int*)(J.progress + bq) < need) __nanosleep(64);
                __threadfence();
            }
            __syncthreads();
            // row y0-1 on the global diagonals [d, d+7] (the first time also d-2, d-1): what the band's first row reads as its
            // upper and diagonal neighbours during the next 8 diagonals
            const int lo = (d == dmin) ? d - 2 : d;
            if (tid < 10 && lo + tid <= d + 7) {
                const int dd = lo + tid, xx = dd - (y0 - 1);
                if (xx >= 0 && xx <= N2 - 2) prevrow[dd & 15] = __ldcg(C + (size_t)dd * pitch + ((y0 - 1) - max(0, dd - (N2 - 1))));
            }
            __syncthreads();
        }
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const int y = y0 + tid + r * T, x = d - y;
            if (SPF && y >= 1 && y <= N1 - 2 && x >= 0 && (x & 3) == 0 && x + 8 < ldS) fetch_s4(r, y, x + 8);
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
        asm volatile("cp.async.wait_group 7;\n" ::: "memory");             // everything issued >= 8 diagonals ago has landed
        const int ylo = max(0, d - (N2 - 1));
        float4* rcur = ring + (d & (RING - 1)) * ring_pitch;
        const float4* row1 = ring + ((d - 1) & (RING - 1)) * ring_pitch;
        const float4* row2 = ring + ((d - 2) & (RING - 1)) * ring_pitch;
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const int y = y0 + tid + r * T, x = d - y;
            const bool valid = !(y > N1 - 2 || x < 0 || x > N2 - 2);
            if (!COOP && !valid) continue;
            float m = NINF, xs = NINF, ys = NINF, w = NINF;
            bool heavy = false;                                  // BANDED: a cell with more than GB predecessors on a side: the whole warp folds it (below)
            NodeInfo hx = NodeInfo{0, 0, 0, 0.f}; float hs = 0.f;
            if (valid) {
            if (x == 0) {
                if (y == 0) w = start_init;
                else {
                    const int cnt = ny[r].cnt & 0x7fffffff;
                    for (int ky = 0; ky < cnt; ++ky) {
                        const int yp = ky ? P1I[ny[r].ptr + ky] : ny[r].idx0;
                        const float ey = ky ? P1V[ny[r].ptr + ky] : ny[r].val0;
                        const float4 c = cell(yp, 0, d);
                        ys = fmaxf(ys, __fsub_rn(fmaxf(__fadd_rn(c.z, gap_extend), __fadd_rn(c.w, start_gap)), ey));
                    }
                    w = ys;
                }
            } else {
                const NodeInfo nx = I2[x];
                if (y == 0) {
                    const int cnt = nx.cnt & 0x7fffffff;
                    for (int kx = 0; kx < cnt; ++kx) {
                        const int xp = kx ? P2I[nx.ptr + kx] : nx.idx0;
                        const float ex = kx ? P2V[nx.ptr + kx] : nx.val0;
                        const float4 c = cell(0, xp, d);
                        xs = fmaxf(xs, __fsub_rn(fmaxf(__fadd_rn(c.y, gap_extend), __fadd_rn(c.w, start_gap)), ex));
                    }
                    w = xs;
                } else {
                    const float s = SPF ? reinterpret_cast<const float*>(sring + ((x >> 2) & 3) * (RPT * T) + r * T + tid)[x & 3]
                                        : __ldg(S + (size_t)y * ldS + x);
                    if (ysimple[r] && nx.cnt == 1 && nx.idx0 == x - 1) {
                        // fast path (chain nodes, the overwhelming majority): one adjacent regular predecessor each
                        const float ey = ny[r].val0, ex = nx.val0;
                        const bool edge = BANDED && band > 0 && y == y0;             // row y-1 belongs to the previous band
                        const float4 cu = edge ? prevrow[(d - 1) & 15] : row1[y - y0 - 1];   // (y-1, x)
                        const float4 cd = edge ? prevrow[(d - 2) & 15] : row2[y - y0 - 1];   // (y-1, x-1)
                        const float4 cl = left[r];                                   // (y, x-1): this thread, last diagonal
                        ys = __fsub_rn(fmaxf(__fadd_rn(cu.z, gap_extend), __fadd_rn(cu.w, gap_init)), ey);
                        m = __fsub_rn(__fsub_rn(__fadd_rn(cd.w, s), ey), ex);
                        xs = __fsub_rn(fmaxf(__fadd_rn(cl.y, gap_extend), __fadd_rn(cl.w, gap_init)), ex);
                        w = fmaxf(m, fmaxf(xs, ys));
                    } else if (BANDED && ((ny[r].cnt & 0x7fffffff) > GB || (nx.cnt & 0x7fffffff) > GB || (ny[r].cnt & 0x7fffffff) < 1 || (nx.cnt & 0x7fffffff) < 1)) {
                        heavy = true; hx = nx; hs = s;
                    } else if (BANDED) {
                        // every candidate of the cell at once (same candidates as the loops below; a maximum does not depend on
                        // the order): absent predecessors get -inf cells, which never win
                        const int c1 = ny[r].cnt & 0x7fffffff, c2 = nx.cnt & 0x7fffffff;
                        int ypv[GB], xpv[GB]; float eyv[GB], exv[GB];
#pragma unroll
                        for (int j = 0; j < GB; ++j) {
                            ypv[j] = ny[r].idx0; eyv[j] = ny[r].val0; xpv[j] = nx.idx0; exv[j] = nx.val0;
                            if (j > 0 && j < c1) { ypv[j] = P1I[ny[r].ptr + j]; eyv[j] = P1V[ny[r].ptr + j]; }
                            if (j > 0 && j < c2) { xpv[j] = P2I[nx.ptr + j]; exv[j] = P2V[nx.ptr + j]; }
                        }
                        float2 cu[GB], cl[GB]; float cdw[GB][GB];
#pragma unroll
                        for (int j = 0; j < GB; ++j) {
                            cu[j] = make_float2(NINF, NINF); cl[j] = make_float2(NINF, NINF);
                            if (j < c1) cu[j] = ld_zw(cell_addr(ypv[j], x, d));                  // (yp, x): Y, W
                            if (j < c2) cl[j] = ld_yw(cell_addr(y, xpv[j], d));                  // (y, xp): X, W
#pragma unroll
                            for (int k = 0; k < GB; ++k) {
                                cdw[j][k] = NINF;
                                if (j < c1 && k < c2) cdw[j][k] = ld_w(cell_addr(ypv[j], xpv[k], d));   // (yp, xp): W
                            }
                        }
#pragma unroll
                        for (int k = 0; k < GB; ++k)
                            xs = fmaxf(xs, __fsub_rn(fmaxf(__fadd_rn(cl[k].x, gap_extend), __fadd_rn(cl[k].y, gap_init)), exv[k]));
#pragma unroll
                        for (int j = 0; j < GB; ++j) {
                            ys = fmaxf(ys, __fsub_rn(fmaxf(__fadd_rn(cu[j].x, gap_extend), __fadd_rn(cu[j].y, gap_init)), eyv[j]));
#pragma unroll
                            for (int k = 0; k < GB; ++k) m = fmaxf(m, __fsub_rn(__fsub_rn(__fadd_rn(cdw[j][k], s), eyv[j]), exv[k]));
                        }
                        w = fmaxf(m, fmaxf(xs, ys));
                    } else if (nx.cnt == 1 && nx.idx0 == x - 1 && (ny[r].cnt & 0x7fffffff) >= 1) {
                        // row with several (or a non-adjacent) predecessors at a chain column: X through the left neighbour
                        // once, Y and M folded over the row's predecessors (same candidates as the general loop below)
                        const float ex = nx.val0;
                        const float4 cl = left[r];
                        xs = __fsub_rn(fmaxf(__fadd_rn(cl.y, gap_extend), __fadd_rn(cl.w, gap_init)), ex);
                        const int c1 = ny[r].cnt & 0x7fffffff;
                        for (int ky = 0; ky < c1; ++ky) {
                            const int yp = ky ? P1I[ny[r].ptr + ky] : ny[r].idx0;
                            const float ey = ky ? P1V[ny[r].ptr + ky] : ny[r].val0;
                            const int dyy = y - yp;
                            float4 cu, cd;
                            if (dyy + 1 < RING && yp >= y0) {
                                cu = ring[((d - dyy) & (RING - 1)) * ring_pitch + (yp - y0)];          // (yp, x)
                                cd = ring[((d - dyy - 1) & (RING - 1)) * ring_pitch + (yp - y0)];      // (yp, x-1)
                            } else { cu = cell(yp, x, d); cd = cell(yp, x - 1, d); }
                            ys = fmaxf(ys, __fsub_rn(fmaxf(__fadd_rn(cu.z, gap_extend), __fadd_rn(cu.w, gap_init)), ey));
                            m = fmaxf(m, __fsub_rn(__fsub_rn(__fadd_rn(cd.w, s), ey), ex));
                        }
                        w = fmaxf(m, fmaxf(xs, ys));
                    } else {
                        const int c1 = ny[r].cnt & 0x7fffffff, c2 = nx.cnt & 0x7fffffff;
                        for (int ky = 0; ky < c1; ++ky) {
                            const int yp = ky ? P1I[ny[r].ptr + ky] : ny[r].idx0;
                            const float ey = ky ? P1V[ny[r].ptr + ky] : ny[r].val0;
                            const float4 cu = cell(yp, x, d);                        // (yp, x): Y, W
                            const float Sy = __fsub_rn(fmaxf(__fadd_rn(cu.z, gap_extend), __fadd_rn(cu.w, gap_init)), ey);
                            if (c2 > 0) ys = fmaxf(ys, Sy);                      // (a column without predecessors has no candidates at all)
                            // the column's predecessors four at a time, their cells loaded together: a deep profile column has a dozen
                            // predecessors, and one L2 round trip per predecessor made the thread that met it the whole diagonal's pace
                            // (a 632 x 6238 job took 6.8 us per anti-diagonal).  Absent slots repeat the first predecessor: a maximum
                            // does not mind duplicates.
                            auto fold_cols = [&](auto nb_tag) {
                                constexpr int NB = decltype(nb_tag)::value;
                                for (int kx0 = 0; kx0 < c2; kx0 += NB) {
                                    int xpv[NB]; float exv[NB], cdw[NB]; float2 clv[NB];
#pragma unroll
                                    for (int q = 0; q < NB; ++q) {
                                        const int kx = kx0 + q;
                                        const bool on = kx > 0 && kx < c2;
                                        xpv[q] = on ? P2I[nx.ptr + kx] : nx.idx0;
                                        exv[q] = on ? P2V[nx.ptr + kx] : nx.val0;
                                    }
#pragma unroll
                                    for (int q = 0; q < NB; ++q) {
                                        cdw[q] = ld_w(cell_addr(yp, xpv[q], d));             // (yp, xp): W
                                        clv[q] = ld_yw(cell_addr(y, xpv[q], d));             // (y, xp): X, W
                                    }
#pragma unroll
                                    for (int q = 0; q < NB; ++q) {
                                        m = fmaxf(m, __fsub_rn(__fsub_rn(__fadd_rn(cdw[q], s), ey), exv[q]));
                                        xs = fmaxf(xs, __fsub_rn(fmaxf(__fadd_rn(clv[q].x, gap_extend), __fadd_rn(clv[q].y, gap_init)), exv[q]));
                                    }
                                }
                            };
                            // (chosen per launch: per cell it would diverge, per job at run time both loops in one kernel ran slower.  Measured: 629 x 1743 with 571 non-chain columns 7.9 ms one at
                            // a time, 8.8 ms four at a time; 566 x 10982 74 against 53 ms)
                            fold_cols(std::integral_constant<int, DEEP ? 4 : 1>{});   // DEEP: the launch has a graph 2 of >= 4000 nodes (a profile of ~1000 sequences or more)
                        }
                        w = fmaxf(m, fmaxf(xs, ys));
                    }
                }
            }
            }
            if (COOP) {
                // Cells with many predecessors (deep profile columns reach dozens): one thread folding c1 x c2 pairs with an L2 round
                // trip per pair held the whole band for tens of microseconds per anti-diagonal (ncu: a fifth of the kernel's
                // instructions ran with 1.6 of 32 lanes active).  The warp folds such a cell together: lane l takes the pairs l, l + 32, ...,
                // three warp maxima, the owner keeps the result.  A maximum does not depend on the order, so the values are unchanged.
                unsigned hb = __ballot_sync(0xffffffffu, heavy);
                const int lane = tid & 31;
                // two cells per round: the second cell's loads are in flight while the first is finished (a warp of a deep profile meets
                // two or three such cells on most anti-diagonals, and each costs two dependent L2 round trips)
                struct Fold { int src, by, bx, c1, c2, p1, p2, yi0, xi0; float yv0, xv0, bs, eyP, exP, exX, eyY, cdw; float2 cl, cu; };
                auto fold_issue = [&](const int src, Fold& F) {
                    F.src = src;
                    F.by = __shfl_sync(0xffffffffu, y, src); F.bx = __shfl_sync(0xffffffffu, x, src);
                    F.c1 = __shfl_sync(0xffffffffu, ny[r].cnt, src) & 0x7fffffff; F.c2 = __shfl_sync(0xffffffffu, hx.cnt, src) & 0x7fffffff;
                    F.p1 = __shfl_sync(0xffffffffu, ny[r].ptr, src); F.p2 = __shfl_sync(0xffffffffu, hx.ptr, src);
                    F.yi0 = __shfl_sync(0xffffffffu, ny[r].idx0, src); F.xi0 = __shfl_sync(0xffffffffu, hx.idx0, src);
                    F.yv0 = __shfl_sync(0xffffffffu, ny[r].val0, src); F.xv0 = __shfl_sync(0xffffffffu, hx.val0, src);
                    F.bs = __shfl_sync(0xffffffffu, hs, src);
                    // first round: the lane's pair, X candidate and Y candidate -- index loads together, then the three cell loads together
                    const int n12 = F.c1 * F.c2;
                    const bool op = lane < n12, ox = lane < F.c2, oy = lane < F.c1;
                    const int j = op ? lane / F.c2 : 0, k = op ? lane - j * F.c2 : 0;
                    const int ypP = j ? P1I[F.p1 + j] : F.yi0, xpP = k ? P2I[F.p2 + k] : F.xi0;
                    F.eyP = j ? P1V[F.p1 + j] : F.yv0; F.exP = k ? P2V[F.p2 + k] : F.xv0;
                    const int xpX = (ox && lane) ? P2I[F.p2 + lane] : F.xi0; F.exX = (ox && lane) ? P2V[F.p2 + lane] : F.xv0;
                    const int ypY = (oy && lane) ? P1I[F.p1 + lane] : F.yi0; F.eyY = (oy && lane) ? P1V[F.p1 + lane] : F.yv0;
                    F.cdw = NINF; F.cl = make_float2(NINF, NINF); F.cu = make_float2(NINF, NINF);
                    if (op) F.cdw = ld_w(cell_addr(ypP, xpP, d));
                    if (ox) F.cl = ld_yw(cell_addr(F.by, xpX, d));
                    if (oy) F.cu = ld_zw(cell_addr(ypY, F.bx, d));
                };
                auto fold_finish = [&](const Fold& F) {
                    float pm = __fsub_rn(__fsub_rn(__fadd_rn(F.cdw, F.bs), F.eyP), F.exP);
                    float pxs = __fsub_rn(fmaxf(__fadd_rn(F.cl.x, gap_extend), __fadd_rn(F.cl.y, gap_init)), F.exX);
                    float pys = __fsub_rn(fmaxf(__fadd_rn(F.cu.x, gap_extend), __fadd_rn(F.cu.y, gap_init)), F.eyY);
                    for (int q = lane + 32; q < F.c1 * F.c2; q += 32) {
                        const int j = q / F.c2, k = q - j * F.c2;
                        const int yp = j ? P1I[F.p1 + j] : F.yi0, xp = k ? P2I[F.p2 + k] : F.xi0;
                        const float ey = j ? P1V[F.p1 + j] : F.yv0, ex = k ? P2V[F.p2 + k] : F.xv0;
                        pm = fmaxf(pm, __fsub_rn(__fsub_rn(__fadd_rn(ld_w(cell_addr(yp, xp, d)), F.bs), ey), ex));
                    }
                    for (int k = lane + 32; k < F.c2; k += 32) {
                        const int xp = P2I[F.p2 + k];
                        const float ex = P2V[F.p2 + k];
                        const float2 cl = ld_yw(cell_addr(F.by, xp, d));
                        pxs = fmaxf(pxs, __fsub_rn(fmaxf(__fadd_rn(cl.x, gap_extend), __fadd_rn(cl.y, gap_init)), ex));
                    }
                    for (int j = lane + 32; j < F.c1; j += 32) {
                        const int yp = P1I[F.p1 + j];
                        const float ey = P1V[F.p1 + j];
                        const float2 cu = ld_zw(cell_addr(yp, F.bx, d));
                        pys = fmaxf(pys, __fsub_rn(fmaxf(__fadd_rn(cu.x, gap_extend), __fadd_rn(cu.y, gap_init)), ey));
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        pm = fmaxf(pm, __shfl_xor_sync(0xffffffffu, pm, o));
                        pxs = fmaxf(pxs, __shfl_xor_sync(0xffffffffu, pxs, o));
                        pys = fmaxf(pys, __shfl_xor_sync(0xffffffffu, pys, o));
                    }
                    if (lane == F.src) { m = pm; xs = pxs; ys = pys; w = fmaxf(pm, fmaxf(pxs, pys)); }
                };
                while (hb) {
                    const int src0 = __ffs(hb) - 1;
                    hb &= hb - 1;
                    const bool two = hb != 0u;
                    const int src1 = two ? __ffs(hb) - 1 : src0;
                    hb &= hb - 1;                                    // (0 stays 0)
                    Fold A, B;
                    fold_issue(src0, A);
                    if (two) fold_issue(src1, B);
                    fold_finish(A);
                    if (two) fold_finish(B);
                }
            }
            if (valid) {
                const float4 out = make_float4(m, xs, ys, w);
                left[r] = out;
                rcur[y - y0] = out;
                C[(size_t)d * pitch + (y - ylo)] = out;
            }
        }
        __syncthreads();
        if (BANDED && tid == 0 && (((d - dmin) & 3) == 3 || d == dmax)) {
            __threadfence();                                     // the CTA's stores of diagonals <= d, ordered by the barrier, before the flag
            *(volatile int*)(J.progress + band) = d + 1;
        }
    }
    // end-node match (GraphAlign.h:262-280); banded: the last band has waited for all the others
    if (tid == 0 && (!BANDED || band == J.bands - 1)) {
        const float end_match = J.sc[3], end_gap = J.sc[4], end_skip = J.sc[5];
        float Wend = NINF;
        for (int ky = J.p1_ptr[N1 - 1]; ky < J.p1_ptr[N1]; ++ky)
            for (int kx = J.p2_ptr[N2 - 1]; kx < J.p2_ptr[N2]; ++kx) {
                const int yp = J.p1_idx[ky], xp = J.p2_idx[kx];
                const float ey = J.p1_val[ky], ex = J.p2_val[kx];
                if (xp == 0 && yp == 0) {
                    Wend = fmaxf(__fsub_rn(__fsub_rn(end_skip, ey), ex), Wend);
                } else {
                    const float4 c = __ldcg(C + skew(yp, xp, N2, pitch));
                    Wend = fmaxf(__fsub_rn(__fsub_rn(__fadd_rn(c.y, end_gap), ey), ex), Wend);
                    Wend = fmaxf(__fsub_rn(__fsub_rn(__fadd_rn(c.z, end_gap), ey), ex), Wend);
                    Wend = fmaxf(__fsub_rn(__fsub_rn(__fadd_rn(c.x, end_match), ey), ex), Wend);
                }
            }
        ((float*)J.result)[3] = Wend;
    }
}

template <bool STAGE, int RPT, int RING, bool SPF, bool DEEP>
static int launch_fill_diag_t(cudaStream_t st, int n_jobs, const GraphJobDev* d_jobs, const int* d_order, int threads, int ring_pitch, size_t smem) {
    PGM_CUDA(cudaFuncSetAttribute(graph_fill_diag_kernel<STAGE, RPT, RING, SPF, false, DEEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    graph_fill_diag_kernel<STAGE, RPT, RING, SPF, false, DEEP><<<n_jobs, threads, smem, st>>>(d_jobs, d_order, ring_pitch, 1);
    PGM_CUDA(cudaGetLastError());
    return 0;
}

// Long graphs: bands of band_rows rows, one CTA each (band index = ticket); the jobs' progress flags and tickets are zeroed
// by the caller.
template <bool STAGE>
static int launch_banded_t(cudaStream_t st, int n_jobs, const GraphJobDev* d_jobs, const int* d_order, int max_bands, int n_sms, int band_rows, size_t smem) {
    (void)n_sms;
    PGM_CUDA(cudaFuncSetAttribute(graph_fill_diag_kernel<STAGE, 1, GF_RING_BAND, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    graph_fill_diag_kernel<STAGE, 1, GF_RING_BAND, true, true><<<n_jobs * max_bands, band_rows, smem, st>>>(d_jobs, d_order, band_rows, max_bands);
    PGM_CUDA(cudaGetLastError());
    return 0;
}

// Long graphs: bands of band_rows rows, one CTA each.  All CTAs of a launch must be resident at once (they wait on each
// other's flags), so a launch holds at most as many CTAs as fit the GPU; the jobs' progress flags are zeroed by the caller.
// Graph 2's node infos and both predecessor lists are staged in shared memory when they fit next to the ring.
int launch_graph_fill_banded(cudaStream_t st, int n_jobs, const GraphJobDev* d_jobs, const int* d_order, int max_bands, int n_sms, int band_rows,
                             int max_n2, int max_np) {
    if (n_jobs == 0) return 0;
    const size_t ring = (size_t)GF_RING_BAND * band_rows * sizeof(float4) + (size_t)4 * band_rows * sizeof(float4);
    const size_t info = (size_t)max_n2 * sizeof(NodeInfo) + (size_t)max_np * 8 + 64;
    if (ring + info <= 200 * 1024) return launch_banded_t<true>(st, n_jobs, d_jobs, d_order, max_bands, n_sms, band_rows, ring + info);
    return launch_banded_t<false>(st, n_jobs, d_jobs, d_order, max_bands, n_sms, band_rows, ring);
}

// max_n1 / max_n2: largest graph-1 / graph-2 node counts in this launch
// d_order: the n_jobs job indices this launch sweeps (the host groups jobs of similar graph-1 length, because block size
// and shared memory -- hence CTAs per SM -- are set by the launch's largest job)
int launch_graph_fill_diag(cudaStream_t st, int n_jobs, const GraphJobDev* d_jobs, const int* d_order, int max_n1, int max_n2, int max_np) {
    if (n_jobs == 0) return 0;
    const int rows = max_n1 > 1 ? max_n1 - 1 : 1;                   // rows 0 .. N1-2
    // rows per thread: on an anti-diagonal sweep only ~half the rows are active at a time, so giving a thread
    // several rows (y, y+T, ..) halves the number of idle warps that still pay the per-diagonal loop overhead
    int rpt = rows <= 1024 ? 1 : (rows <= 2048 ? 2 : 4);       // measured: 1 row per thread is best while it fits
    if (const char* e = getenv("PGM_FILL_RPT")) { const int v = atoi(e); if (v == 1 || v == 2 || v == 4) rpt = v; }
    while ((rows + rpt - 1) / rpt > 1024) rpt *= 2;
    if (rows > 4096) { pgm_set_error("graph 1 has %d nodes; the wavefront kernel supports up to 4097", max_n1); return PGM_ERR_ARG; }
    const int threads = std::min(1024, std::max(64, ((rows + rpt - 1) / rpt + 31) & ~31));
    const int ring_pitch = (max_n1 + 3) & ~3;
    const size_t info = (size_t)max_n2 * sizeof(NodeInfo) + (size_t)max_np * 8 + 64;
    // regular flavour: 8-diagonal ring + cp.async S prefetch slots; long graphs: 4-diagonal ring, S read directly
    const size_t ring8 = (size_t)GF_RING * ring_pitch * sizeof(float4) + (size_t)4 * rpt * threads * sizeof(float4);
    const size_t ring4 = (size_t)GF_RING_LONG * ring_pitch * sizeof(float4);
    const bool regular = ring8 <= 227 * 1024;
    const size_t ring = regular ? ring8 : ring4;
    if (ring > 227 * 1024) { pgm_set_error("graph too large for the shared-memory diagonal ring (N1 = %d, limit ~3500)", max_n1); return PGM_ERR_ARG; }
    const bool stage = ring + info <= 200 * 1024;
    const size_t smem = ring + (stage ? info : 0);
    const bool deep = max_n2 >= 4000 && rpt == 1;
#define GO(S_, R_) do { if (regular && deep && R_ == 1) return launch_fill_diag_t<S_, 1, GF_RING, true, true>(st, n_jobs, d_jobs, d_order, threads, ring_pitch, smem); \
                        if (regular) return launch_fill_diag_t<S_, R_, GF_RING, true, false>(st, n_jobs, d_jobs, d_order, threads, ring_pitch, smem); \
                        return launch_fill_diag_t<S_, R_, GF_RING_LONG, false, false>(st, n_jobs, d_jobs, d_order, threads, ring_pitch, smem); } while (0)
    if (stage) { if (rpt == 1) GO(true, 1); if (rpt == 2) GO(true, 2); GO(true, 4); }
    if (rpt == 1) GO(false, 1); if (rpt == 2) GO(false, 2); GO(false, 4);
#undef GO
}
