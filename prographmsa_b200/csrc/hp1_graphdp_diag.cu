// hp1_graphdp_diag.cu -- K5, latency-optimised flavour: the anti-diagonal wavefront fill (one thread per graph-1 row,
// one cell per thread and barrier).  It does ~3x the instructions per cell of the blocked kernel in hp1_graphdp.cu, but a
// job's sweep is N1 + N2 short steps with hundreds of threads at work, and multi-predecessor cells cost it almost nothing,
// so it is the faster one for graphs with many non-chain nodes (the upper levels of a guide tree) and for small batches.
// hp1_host.cu picks the flavour per job (GraphJobDev::layout); the traceback reads either cell layout.
//
// Recurrences, evaluation order and the reference lines are those of hp1_graphdp.cu (GraphAlign.h:205-280).  The DP
// values live in a shared-memory ring of the last RING anti-diagonals; what goes to HBM is one packed direction word per
// cell (hp1_dir.cuh) plus the {M,X,Y,W} of the few kept rows / columns (value store) that a far predecessor edge, another
// band, the end-node match or the traceback of the boundary row / column reads back.
#include "hp1_dir.cuh"
#include <algorithm>
#include <cstdlib>

#define GF_RING 8            // anti-diagonals of {M,X,Y,W} kept in shared memory (power of two): covers skip edges spanning <= 3 nodes in both graphs
#define GF_RING_LONG 4       // ... for graphs too long for 8 (>= 4: rows d, d-1, d-2 are live at once); S is then read directly
// S is row-major by y; each (thread,row) streams its own row 16 bytes (4 columns) at a time with cp.async,
// 8 columns (= 8 anti-diagonals) ahead, into a private ring of four float4 slots

// Thread <-> ROW mapping: thread t owns rows y = t + r*blockDim (r < RPT).  Along the sweep the same
// thread meets (y,0), (y,1), ... on consecutive anti-diagonals, so the LEFT neighbour (y, x-1) and the
// row's own predecessor info stay in registers; the UP / DIAGONAL neighbours (y-1, .) come from the
// previous two ring rows at index y-1 (no skew arithmetic).  STAGE: graph 2's node infos in shared memory.
//
// BANDED (graph 1 too long for one CTA's ring): the rows are cut into bands of blockDim rows, one CTA per band.  Every
// dependency of a cell lies on a strictly earlier global anti-diagonal, so band b may sweep diagonal d once every earlier
// band has finished its cells on diagonals < d: each band publishes the number of diagonals it has completed (fence + flag,
// every 4 diagonals) and waits on the earlier bands' flags (every 8).  A CTA takes its band index from a per-job atomic
// ticket, so a band only ever waits on bands whose CTAs have already started (no assumption on the order in which the
// hardware dispatches CTAs, no co-residency requirement).  Cells of other bands are read from the value store (the host
// keeps every row an edge into a later band starts from); the row just above the band, which the band's first row needs on
// every diagonal, is staged 8 diagonals at a time into shared memory right after the flag check (it is complete by then),
// so that row runs the chain fast path like any other.
template <int B, bool STAGE, int RPT, int RING, bool SPF, bool BANDED, int MAXT>
__global__ void __launch_bounds__(MAXT)
graph_fill_diag_kernel(const GraphJobDev* __restrict__ jobs, const int* __restrict__ order, int ring_pitch, int max_bands) {
    typedef typename DirWord<B>::T DW;
    const GraphJobDev& J = jobs[order[BANDED ? blockIdx.x / max_bands : blockIdx.x]];
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
    DW* __restrict__ D = (DW*)J.D;
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

    auto cell = [&](int yy, int xx, int d) -> float4 {          // any earlier cell: ring when recent (and of this band), else the value store
        const int dd = yy + xx;
        if (d - dd < RING && (!BANDED || yy >= y0)) return ring[(dd & (RING - 1)) * ring_pitch + (yy - y0)];
        return dir_far_cell(J, NI_KEEP(__ldg(&J.ni1[yy].cnt)), NI_KEEP(I2[xx].cnt), yy, xx);
    };

    NodeInfo ny[RPT]; bool ysimple[RPT]; float4 left[RPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        const int y = y0 + tid + r * T;
        ny[r] = (y <= N1 - 2) ? J.ni1[y] : NodeInfo{0, 0, 0, 0.f};
        ysimple[r] = NI_CHAIN(ny[r], y);
        left[r] = make_float4(NINF, NINF, NINF, NINF);
    }
    const int band_rows = T * RPT;
    const int ylast = BANDED ? min(N1 - 2, y0 + band_rows - 1) : N1 - 2;           // last row of this band
    const int dmin = y0, dmax = ylast + (N2 - 2);
    const int ldS = J.ldS;
    const int keep_prev = (BANDED && y0 > 0) ? NI_KEEP(__ldg(&J.ni1[y0 - 1].cnt)) : -1;   // the row just above the band (always kept)
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
                for (int bq = 0; bq < band; ++bq) {
                    const int last_q = min(N1 - 2, (bq + 1) * band_rows - 1) + (N2 - 2);
                    const int need = min(d + 8, last_q + 1);
                    while (*(volatile int*)(J.progress + bq) < need) __nanosleep(64);
                }
                __threadfence();
            }
            __syncthreads();
            // row y0-1 on the global diagonals [d, d+7] (the first time also d-2, d-1): what the band's first row reads as its
            // upper and diagonal neighbours during the next 8 diagonals
            const int lo = (d == dmin) ? d - 2 : d;
            if (tid < 10 && lo + tid <= d + 7) {
                const int dd = lo + tid, xx = dd - (y0 - 1);
                if (xx >= 0 && xx <= N2 - 2) prevrow[dd & 15] = dir_far_cell(J, keep_prev, -1, y0 - 1, xx);
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
            if (y > N1 - 2 || x < 0 || x > N2 - 2) continue;
            float m = NINF, xs = NINF, ys = NINF, w = NINF;
            uint32_t kym = 0, kxm = 0, kyy = 0, kxx = 0; bool yopen = false, xopen = false;
            const NodeInfo nx = I2[x];
            if (x == 0) {
                if (y == 0) w = start_init;
                else {
                    const int cnt = NI_CNT(ny[r].cnt);
                    for (int ky = 0; ky < cnt; ++ky) {
                        const int yp = ky ? P1I[ny[r].ptr + ky] : ny[r].idx0;
                        const float ey = ky ? P1V[ny[r].ptr + ky] : ny[r].val0;
                        const float4 c = cell(yp, 0, d);
                        ys = fmaxf(ys, __fsub_rn(fmaxf(__fadd_rn(c.z, gap_extend), __fadd_rn(c.w, start_gap)), ey));
                    }
                    w = ys;
                }
            } else if (y == 0) {
                const int cnt = NI_CNT(nx.cnt);
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
                const bool xsimple = NI_CHAIN(nx, x);
                if (ysimple[r] && xsimple) {
                    // fast path (chain nodes, the overwhelming majority): one adjacent regular predecessor each
                    const float ey = ny[r].val0, ex = nx.val0;
                    const bool edge = BANDED && band > 0 && y == y0;             // row y-1 belongs to the previous band
                    const float4 cu = edge ? prevrow[(d - 1) & 15] : row1[y - y0 - 1];   // (y-1, x)
                    const float4 cd = edge ? prevrow[(d - 2) & 15] : row2[y - y0 - 1];   // (y-1, x-1)
                    const float4 cl = left[r];                                   // (y, x-1): this thread, last diagonal
                    const float yE = __fsub_rn(__fadd_rn(cu.z, gap_extend), ey), yO = __fsub_rn(__fadd_rn(cu.w, gap_init), ey);
                    const float xE = __fsub_rn(__fadd_rn(cl.y, gap_extend), ex), xO = __fsub_rn(__fadd_rn(cl.w, gap_init), ex);
                    ys = fmaxf(yE, yO); yopen = yO > yE;
                    xs = fmaxf(xE, xO); xopen = xO > xE;
                    m = __fsub_rn(__fsub_rn(__fadd_rn(cd.w, s), ey), ex);
                    w = fmaxf(m, fmaxf(xs, ys));
                } else {
                    // general cell: X over the column's predecessors, Y over the row's, M over the pairs in the reference's
                    // nested order; strict '>' keeps the FIRST maximum, which is what the reference's traceback would pick.
                    // (The reference's pair loop runs only when both lists are non-empty; so do these.)
                    const int c1 = NI_CNT(ny[r].cnt), c2 = NI_CNT(nx.cnt);
                    if (c1 > 0 && c2 > 0) {
                        if (xsimple) {                                             // left neighbour in registers
                            const float ex = nx.val0;
                            const float4 cl = left[r];
                            const float xE = __fsub_rn(__fadd_rn(cl.y, gap_extend), ex), xO = __fsub_rn(__fadd_rn(cl.w, gap_init), ex);
                            xs = fmaxf(xE, xO); xopen = xO > xE;
                        } else {
                            for (int kx = 0; kx < c2; ++kx) {
                                const int xp = kx ? P2I[nx.ptr + kx] : nx.idx0;
                                const float ex = kx ? P2V[nx.ptr + kx] : nx.val0;
                                const float4 cl = cell(y, xp, d);                    // (y, xp): X, W
                                const float xE = __fsub_rn(__fadd_rn(cl.y, gap_extend), ex), xO = __fsub_rn(__fadd_rn(cl.w, gap_init), ex);
                                if (xE > xs) { xs = xE; kxx = kx; xopen = false; }
                                if (xO > xs) { xs = xO; kxx = kx; xopen = true; }
                            }
                        }
                        for (int ky = 0; ky < c1; ++ky) {
                            const int yp = ky ? P1I[ny[r].ptr + ky] : ny[r].idx0;
                            const float ey = ky ? P1V[ny[r].ptr + ky] : ny[r].val0;
                            const float4 cu = cell(yp, x, d);                        // (yp, x): Y, W
                            const float yE = __fsub_rn(__fadd_rn(cu.z, gap_extend), ey), yO = __fsub_rn(__fadd_rn(cu.w, gap_init), ey);
                            if (yE > ys) { ys = yE; kyy = ky; yopen = false; }
                            if (yO > ys) { ys = yO; kyy = ky; yopen = true; }
                            for (int kx = 0; kx < c2; ++kx) {
                                const int xp = kx ? P2I[nx.ptr + kx] : nx.idx0;
                                const float ex = kx ? P2V[nx.ptr + kx] : nx.val0;
                                const float cw = cell(yp, xp, d).w;                  // (yp, xp): W
                                const float Sm = __fsub_rn(__fsub_rn(__fadd_rn(cw, s), ey), ex);
                                if (Sm > m) { m = Sm; kym = ky; kxm = kx; }
                            }
                        }
                        w = fmaxf(m, fmaxf(xs, ys));
                    }
                }
            }
            const float4 out = make_float4(m, xs, ys, w);
            left[r] = out;
            rcur[y - y0] = out;
            D[(size_t)d * pitch + (y - ylo)] = (DW)dir_pack<B>(dir_wstate(m, xs, ys, w), yopen, xopen, kym, kxm, kyy, kxx);
            const int kr = NI_KEEP(ny[r].cnt), kc = NI_KEEP(nx.cnt);
            if (kr >= 0) J.Vrow[(size_t)kr * J.ldVr + x] = out;
            if (kc >= 0) J.Vcol[(size_t)kc * J.ldVc + y] = out;
        }
        __syncthreads();
        if (BANDED && tid == 0 && (((d - dmin) & 3) == 3 || d == dmax)) {
            __threadfence();                                     // the CTA's stores of diagonals <= d, ordered by the barrier, before the flag
            *(volatile int*)(J.progress + band) = d + 1;
        }
    }
    // end-node match (GraphAlign.h:262-280) and the traceback's first pick (:283-311); banded: the last band has waited
    // for all the others.  The predecessors of graph 1's end node are kept rows.
    if (tid == 0 && (!BANDED || band == J.bands - 1)) dir_end_match(J);
}

template <int B, bool STAGE, int RPT, int RING, bool SPF>
static int launch_fill_diag_t(cudaStream_t st, int n_jobs, const GraphJobDev* d_jobs, const int* d_order, int threads, int ring_pitch, size_t smem) {
    // up to 512 threads the kernel may use 128 registers (the first-maximum bookkeeping of the packed traceback needs them)
    if (threads <= 512) {
        PGM_CUDA(cudaFuncSetAttribute(graph_fill_diag_kernel<B, STAGE, RPT, RING, SPF, false, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        graph_fill_diag_kernel<B, STAGE, RPT, RING, SPF, false, 512><<<n_jobs, threads, smem, st>>>(d_jobs, d_order, ring_pitch, 1);
    } else {
        PGM_CUDA(cudaFuncSetAttribute(graph_fill_diag_kernel<B, STAGE, RPT, RING, SPF, false, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        graph_fill_diag_kernel<B, STAGE, RPT, RING, SPF, false, 1024><<<n_jobs, threads, smem, st>>>(d_jobs, d_order, ring_pitch, 1);
    }
    PGM_CUDA(cudaGetLastError());
    return 0;
}

// Long graphs: bands of band_rows rows, one CTA each (band index = ticket, so no co-residency requirement); the jobs'
// progress flags and tickets are zeroed by the caller.
template <int B, bool STAGE>
static int launch_banded_t(cudaStream_t st, int n_jobs, const GraphJobDev* d_jobs, const int* d_order, int max_bands, int band_rows, size_t smem) {
    if (band_rows <= 512) {
        PGM_CUDA(cudaFuncSetAttribute(graph_fill_diag_kernel<B, STAGE, 1, GF_RING, true, true, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        graph_fill_diag_kernel<B, STAGE, 1, GF_RING, true, true, 512><<<n_jobs * max_bands, band_rows, smem, st>>>(d_jobs, d_order, band_rows, max_bands);
    } else {
        PGM_CUDA(cudaFuncSetAttribute(graph_fill_diag_kernel<B, STAGE, 1, GF_RING, true, true, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        graph_fill_diag_kernel<B, STAGE, 1, GF_RING, true, true, 1024><<<n_jobs * max_bands, band_rows, smem, st>>>(d_jobs, d_order, band_rows, max_bands);
    }
    PGM_CUDA(cudaGetLastError());
    return 0;
}

// Graph 2's node infos and both predecessor lists are staged in shared memory when they fit next to the ring.
int launch_graph_fill_banded(cudaStream_t st, int dbits, int n_jobs, const GraphJobDev* d_jobs, const int* d_order, int max_bands, int band_rows,
                             int max_n2, int max_np) {
    if (n_jobs == 0) return 0;
    const size_t ring = (size_t)GF_RING * band_rows * sizeof(float4) + (size_t)4 * band_rows * sizeof(float4);
    const size_t info = (size_t)max_n2 * sizeof(NodeInfo) + (size_t)max_np * 8 + 64;
    const bool stage = ring + info <= 200 * 1024;
    const size_t smem = ring + (stage ? info : 0);
    if (dbits == 3) return stage ? launch_banded_t<3, true>(st, n_jobs, d_jobs, d_order, max_bands, band_rows, smem)
                                 : launch_banded_t<3, false>(st, n_jobs, d_jobs, d_order, max_bands, band_rows, smem);
    return stage ? launch_banded_t<7, true>(st, n_jobs, d_jobs, d_order, max_bands, band_rows, smem)
                 : launch_banded_t<7, false>(st, n_jobs, d_jobs, d_order, max_bands, band_rows, smem);
}

// max_n1 / max_n2: largest graph-1 / graph-2 node counts in this launch
// d_order: the n_jobs job indices this launch sweeps (the host groups jobs of similar graph-1 length, because block size
// and shared memory -- hence CTAs per SM -- are set by the launch's largest job)
template <int B>
static int launch_graph_fill_diag_b(cudaStream_t st, int n_jobs, const GraphJobDev* d_jobs, const int* d_order, int max_n1, int max_n2, int max_np) {
    const int rows = max_n1 > 1 ? max_n1 - 1 : 1;                   // rows 0 .. N1-2
    if (rows > 4096) { pgm_set_error("graph 1 has %d nodes; the wavefront kernel supports up to 4097", max_n1); return PGM_ERR_ARG; }
    const int ring_pitch = (max_n1 + 3) & ~3;
    const size_t info = (size_t)max_n2 * sizeof(NodeInfo) + (size_t)max_np * 8 + 64;
    if (rows <= GF_DIAG_REGULAR_ROWS) {
        // regular flavour: one row per thread, 8-diagonal ring + cp.async S prefetch slots
        const int threads = std::min(1024, std::max(64, (rows + 31) & ~31));
        const size_t ring = (size_t)GF_RING * ring_pitch * sizeof(float4) + (size_t)4 * threads * sizeof(float4);
        const bool stage = ring + info <= 200 * 1024;
        const size_t smem = ring + (stage ? info : 0);
        return stage ? launch_fill_diag_t<B, true, 1, GF_RING, true>(st, n_jobs, d_jobs, d_order, threads, ring_pitch, smem)
                     : launch_fill_diag_t<B, false, 1, GF_RING, true>(st, n_jobs, d_jobs, d_order, threads, ring_pitch, smem);
    }
    // long graphs that are not swept in bands: 2 or 4 rows per thread (on an anti-diagonal sweep only ~half the rows are
    // active at a time), 4-diagonal ring, S read directly
    const int rpt = rows <= 2048 ? 2 : 4;
    const int threads = std::min(1024, std::max(64, ((rows + rpt - 1) / rpt + 31) & ~31));
    const size_t ring = (size_t)GF_RING_LONG * ring_pitch * sizeof(float4);
    const bool stage = ring + info <= 200 * 1024;
    const size_t smem = ring + (stage ? info : 0);
    if (rpt == 2) return stage ? launch_fill_diag_t<B, true, 2, GF_RING_LONG, false>(st, n_jobs, d_jobs, d_order, threads, ring_pitch, smem)
                               : launch_fill_diag_t<B, false, 2, GF_RING_LONG, false>(st, n_jobs, d_jobs, d_order, threads, ring_pitch, smem);
    return stage ? launch_fill_diag_t<B, true, 4, GF_RING_LONG, false>(st, n_jobs, d_jobs, d_order, threads, ring_pitch, smem)
                 : launch_fill_diag_t<B, false, 4, GF_RING_LONG, false>(st, n_jobs, d_jobs, d_order, threads, ring_pitch, smem);
}

int launch_graph_fill_diag(cudaStream_t st, int dbits, int n_jobs, const GraphJobDev* d_jobs, const int* d_order, int max_n1, int max_n2, int max_np) {
    if (n_jobs == 0) return 0;
    return dbits == 3 ? launch_graph_fill_diag_b<3>(st, n_jobs, d_jobs, d_order, max_n1, max_n2, max_np)
                      : launch_graph_fill_diag_b<7>(st, n_jobs, d_jobs, d_order, max_n1, max_n2, max_np);
}
