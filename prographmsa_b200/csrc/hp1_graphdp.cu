// hp1_graphdp.cu -- K5 (anti-diagonal wavefront fill + end-node match) and K6 (traceback walk).
//
// Replaces alignGraphs (reference GraphAlign.h:201-534):
//   boundary column/row   :212-234      fill over predecessor pairs :238-260
//   end-node match        :262-280      traceback                   :283-521 (+ :165-198)
// Recurrences (fp32, the reference's exact evaluation order; compiled with --fmad=false):
//   Sm = ((W(yp,xp) + S(y,x)) - e_y) - e_x
//   Sx = max(X(y,xp) + gap_extend, W(y,xp) + gap_init) - e_x
//   Sy = max(Y(yp,x) + gap_extend, W(yp,x) + gap_init) - e_y
//   M,X,Y,W <- running max over all predecessor pairs (yp,xp) of (y,x)
//
// Design: every dependency of a cell lies on a strictly earlier anti-diagonal (predecessors have
// smaller node indices in both graphs), so one CTA sweeps the anti-diagonals of one job with a
// barrier between them; a batch of independent jobs (all ready guide-tree nodes) fills the GPU.
// {M,X,Y,W} of a cell is one float4 in a skewed layout (diagonal-major), so a diagonal is one
// coalesced 16 B-per-thread store and the common predecessor (y-1, x-1 / y, x-1 / y-1, x) reads
// are coalesced too.  Arbitrary predecessors (skip edges after indels, tandem-repeat edges) are read
// from the same arrays, so no graph shape needs a fallback.
//
// The traceback re-runs the reference's "closest |delta|" search verbatim on the stored matrices
// (one thread per job): first-wins ties in predecessor order, M > Y > X when landing on W,
// alternative-path marking for repeat edges.  With identical S it is bit-identical to the oracle.
#include "internal.h"
#include <math_constants.h>
#include <algorithm>
#include <cstdlib>

__device__ __forceinline__ size_t skew(int y, int x, int N2, int pitch) {
    const int d = y + x;
    const int ylo = max(0, d - (N2 - 1));
    return (size_t)d * pitch + (y - ylo);
}

#define GF_RING 8            // anti-diagonals of {M,X,Y,W} kept in shared memory (power of two): covers skip edges spanning <= 5 nodes
#define GF_RING_LONG 4       // ... for graphs too long for 8 (>= 4: rows d, d-1, d-2 are live at once); S is then read directly
// S is row-major by y; each (thread,row) streams its own row 16 bytes (4 columns) at a time with cp.async,
// 8 columns (= 8 anti-diagonals) ahead, into a private ring of four float4 slots

// Thread <-> ROW mapping: thread t owns rows y = t + r*blockDim (r < RPT).  Along the sweep the same
// thread meets (y,0), (y,1), ... on consecutive anti-diagonals, so the LEFT neighbour (y, x-1) and the
// row's own predecessor info stay in registers; the UP / DIAGONAL neighbours (y-1, .) come from the
// previous two ring rows at index y-1 (no skew arithmetic).  STAGE: graph 2's node infos in shared memory.
template <bool STAGE, int RPT, int RING, bool SPF>
__global__ void __launch_bounds__(1024)
graph_fill_kernel(const GraphJobDev* __restrict__ jobs, const int* __restrict__ order, int ring_pitch) {
    const GraphJobDev& J = jobs[order[blockIdx.x]];
    const int N1 = J.N1, N2 = J.N2, pitch = J.pitch;
    const float gap_init = J.sc[0], gap_extend = J.sc[1], start_gap = J.sc[6], start_init = J.sc[7];
    const float NINF = -CUDART_INF_F;
    float4* __restrict__ C = J.MXYW;
    const float* __restrict__ S = J.S;
    const int T = (int)blockDim.x, tid = (int)threadIdx.x;

    extern __shared__ __align__(16) unsigned char gsm[];
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

    auto cell = [&](int yy, int xx, int d) -> float4 {          // any earlier cell: ring when recent, else HBM/L2
        const int dd = yy + xx;
        if (d - dd < RING) return ring[(dd & (RING - 1)) * ring_pitch + yy];
        return C[(size_t)dd * pitch + (yy - max(0, dd - (N2 - 1)))];
    };

    NodeInfo ny[RPT]; bool ysimple[RPT]; float4 left[RPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        const int y = tid + r * T;
        ny[r] = (y <= N1 - 2) ? J.ni1[y] : NodeInfo{0, 0, 0, 0.f};
        ysimple[r] = (ny[r].cnt == 1 && ny[r].idx0 == y - 1);
        left[r] = make_float4(NINF, NINF, NINF, NINF);
    }
    const int dmax = (N1 - 2) + (N2 - 2);
    const int ldS = J.ldS;
    auto fetch_s4 = [&](int r, int y, int xq) {                  // S(y, xq..xq+3) -> sring[(xq/4)&3][r*T+tid], asynchronously
        const float* src = S + (size_t)y * ldS + xq;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sring + ((xq >> 2) & 3) * (RPT * T) + r * T + tid);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src) : "memory");
    };
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        const int y = tid + r * T;
        if (SPF && y >= 1 && y <= N1 - 2) { fetch_s4(r, y, 0); if (4 < ldS) fetch_s4(r, y, 4); }
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");      // the first rows need these within < 8 diagonals

    for (int d = 0; d <= dmax; ++d) {
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const int y = tid + r * T, x = d - y;
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
            const int y = tid + r * T, x = d - y;
            if (y > N1 - 2 || x < 0 || x > N2 - 2) continue;
            float m = NINF, xs = NINF, ys = NINF, w = NINF;
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
                        const float4 cu = row1[y - 1];                               // (y-1, x)
                        const float4 cd = row2[y - 1];                               // (y-1, x-1)
                        const float4 cl = left[r];                                   // (y, x-1): this thread, last diagonal
                        ys = __fsub_rn(fmaxf(__fadd_rn(cu.z, gap_extend), __fadd_rn(cu.w, gap_init)), ey);
                        m = __fsub_rn(__fsub_rn(__fadd_rn(cd.w, s), ey), ex);
                        xs = __fsub_rn(fmaxf(__fadd_rn(cl.y, gap_extend), __fadd_rn(cl.w, gap_init)), ex);
                        w = fmaxf(m, fmaxf(xs, ys));
                    } else {
                        const int c1 = ny[r].cnt & 0x7fffffff, c2 = nx.cnt & 0x7fffffff;
                        for (int ky = 0; ky < c1; ++ky) {
                            const int yp = ky ? P1I[ny[r].ptr + ky] : ny[r].idx0;
                            const float ey = ky ? P1V[ny[r].ptr + ky] : ny[r].val0;
                            const float4 cu = cell(yp, x, d);                        // (yp, x): Y, W
                            const float Sy = __fsub_rn(fmaxf(__fadd_rn(cu.z, gap_extend), __fadd_rn(cu.w, gap_init)), ey);
                            for (int kx = 0; kx < c2; ++kx) {
                                const int xp = kx ? P2I[nx.ptr + kx] : nx.idx0;
                                const float ex = kx ? P2V[nx.ptr + kx] : nx.val0;
                                const float4 cd = cell(yp, xp, d);                   // (yp, xp): W
                                const float4 cl = cell(y, xp, d);                    // (y, xp): X, W
                                const float Sm = __fsub_rn(__fsub_rn(__fadd_rn(cd.w, s), ey), ex);
                                const float Sx = __fsub_rn(fmaxf(__fadd_rn(cl.y, gap_extend), __fadd_rn(cl.w, gap_init)), ex);
                                const float Sw = fmaxf(Sm, fmaxf(Sx, Sy));
                                m = fmaxf(m, Sm); xs = fmaxf(xs, Sx); ys = fmaxf(ys, Sy); w = fmaxf(w, Sw);
                            }
                        }
                    }
                }
            }
            const float4 out = make_float4(m, xs, ys, w);
            left[r] = out;
            rcur[y] = out;
            C[(size_t)d * pitch + (y - ylo)] = out;
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
                    const float4 c = C[skew(yp, xp, N2, pitch)];
                    Wend = fmaxf(__fsub_rn(__fsub_rn(__fadd_rn(c.y, end_gap), ey), ex), Wend);
                    Wend = fmaxf(__fsub_rn(__fsub_rn(__fadd_rn(c.z, end_gap), ey), ex), Wend);
                    Wend = fmaxf(__fsub_rn(__fsub_rn(__fadd_rn(c.x, end_match), ey), ex), Wend);
                }
            }
        ((float*)J.result)[3] = Wend;
    }
}

// ---- traceback ------------------------------------------------------------------------------------------
struct Tracer {
    const GraphJobDev& J;
    int len;
    __device__ Tracer(const GraphJobDev& j) : J(j), len(0) {}
    __device__ void push(int a, int b) {
        if (len < J.cap) { J.map1[len] = a; J.map2[len] = b; }
        ++len;
    }
    // markAlternativePath (GraphAlign.h:165-198); which = 0: graph 1 (self = map1), 1: graph 2
    __device__ void alt_path(int start, int end, int which) {
        const int* ptr = which ? J.p2_ptr : J.p1_ptr;
        const int* idx = which ? J.p2_idx : J.p1_idx;
        const float* alt = which ? J.p2_alt : J.p1_alt;
        const int n = end - start + 1;
        float* score = J.alt_score; int* prev = J.alt_prev;
        score[0] = 0.f; prev[0] = -1;
        for (int i = 1; i < n; ++i) {
            float sc = -CUDART_INF_F; int pv = -1;
            const int node = i + start;
            for (int k = ptr[node]; k < ptr[node + 1]; ++k) {
                const int f = idx[k];
                if (f >= start && f <= end) {
                    const float cand = __fsub_rn(score[f - start], alt[k]);
                    if (sc <= cand) { sc = cand; pv = f - start; }
                }
            }
            score[i] = sc; prev[i] = pv;
        }
        if (score[n - 1] > -CUDART_INF_F) {
            int i = prev[n - 1];
            while (i != 0) {
                if (which) push(-1, i + start); else push(i + start, -1);
                i = prev[i];
            }
        }
    }
};

// One WARP per job.  All 32 lanes run the walk redundantly (identical values, so the duplicate stores are
// benign and every load is a broadcast); the lanes earn their keep on DIAGONAL RUNS: in state M at chain
// nodes the next cell is forced to be (y-1, x-1), so lane k fetches cell (y-k, x-k) and its nodes, the warp
// finds with two ballots how far the run extends, and up to 32 path steps are emitted at once instead of
// one dependent load chain per step.
__global__ void __launch_bounds__(32)
graph_trace_kernel(int n_jobs, const GraphJobDev* __restrict__ jobs) {
    const int job = blockIdx.x;
    const int lane = threadIdx.x;
    if (job >= n_jobs) return;
    const GraphJobDev& J = jobs[job];
    const int N1 = J.N1, N2 = J.N2, pitch = J.pitch;
    const float gap_init = J.sc[0], gap_extend = J.sc[1], end_match = J.sc[3], end_gap = J.sc[4], end_skip = J.sc[5];
    const float4* __restrict__ C = J.MXYW;
    const float Wend = ((const float*)J.result)[3];
    const float FINF = CUDART_INF_F;
    Tracer T(J);
    int status = 0;
    unsigned n_tr = 0;
    int state = 0;                       // 0 = M, 1 = X, 2 = Y
    float cur = -FINF;
    int y = N1 - 1, x = N2 - 1;
    T.push(N1 - 1, N2 - 1);
    bool trx = false, try_ = false;
    float best = FINF;
    for (int ky = J.p1_ptr[N1 - 1]; ky < J.p1_ptr[N1]; ++ky)
        for (int kx = J.p2_ptr[N2 - 1]; kx < J.p2_ptr[N2]; ++kx) {
            const int yp = J.p1_idx[ky], xp = J.p2_idx[kx];
            const float ey = J.p1_val[ky], ex = J.p2_val[kx];
            const bool ry = J.p1_rep[ky] != 0, rx = J.p2_rep[kx] != 0;
            const float4 c = C[skew(yp, xp, N2, pitch)];
            float q = fabsf(__fsub_rn(Wend, __fsub_rn(__fsub_rn(__fadd_rn(c.x, end_match), ey), ex)));
            if (best > q) { best = q; trx = rx; try_ = ry; cur = c.x; state = 0; y = yp; x = xp; }
            q = fabsf(__fsub_rn(Wend, __fsub_rn(__fsub_rn(__fadd_rn(c.z, end_gap), ey), ex)));
            if (best > q) { best = q; trx = rx; try_ = ry; cur = c.z; state = 2; y = yp; x = xp; }
            q = fabsf(__fsub_rn(Wend, __fsub_rn(__fsub_rn(__fadd_rn(c.y, end_gap), ey), ex)));
            if (best > q) { best = q; trx = rx; try_ = ry; cur = c.y; state = 1; y = yp; x = xp; }
            if (xp == 0 && yp == 0) {
                q = fabsf(__fsub_rn(Wend, __fsub_rn(__fsub_rn(end_skip, ey), ex)));
                if (best > q) { best = q; trx = rx; try_ = ry; y = yp; x = xp; }
            }
        }
    n_tr += (trx ? 1u : 0u) + (try_ ? 1u : 0u);
    if (try_) T.alt_path(y, N1 - 1, 0);
    if (trx) T.alt_path(x, N2 - 1, 1);
    if (x != 0 || y != 0) {
        if (state == 0) T.push(y, x); else if (state == 1) T.push(-1, x); else T.push(y, -1);
    }
    float nscore = FINF; int nstate = 0, nx = -1, ny = -1;
    long guard = 0;
    const long guard_max = 4L * (N1 + N2) + 16;
    // predecessor k of a node: the first one is inline in its NodeInfo (one 16-byte load)
    auto pred1 = [&](const NodeInfo& ni, int k, int& idx, float& val, bool& rep) {
        if (k == 0) { idx = ni.idx0; val = ni.val0; rep = ni.cnt < 0; }
        else { idx = J.p1_idx[ni.ptr + k]; val = J.p1_val[ni.ptr + k]; rep = J.p1_rep[ni.ptr + k] != 0; }
    };
    auto pred2 = [&](const NodeInfo& ni, int k, int& idx, float& val, bool& rep) {
        if (k == 0) { idx = ni.idx0; val = ni.val0; rep = ni.cnt < 0; }
        else { idx = J.p2_idx[ni.ptr + k]; val = J.p2_val[ni.ptr + k]; rep = J.p2_rep[ni.ptr + k] != 0; }
    };
    const int PF = 8;                                   // prefetch distance (steps) along the diagonal
    while (x != 0 || y != 0) {
        if (++guard > guard_max) { status = PGM_ERR_BACKTRACK; break; }
        if (state == 0 && isfinite(cur)) {
            // ---- diagonal run: lane -> k = lane+1, position (y-k, x-k), reached from (y-k+1, x-k+1) ----
            const int k = lane + 1, yk = y - k, xk = x - k;
            bool tr = false, stM = false; int st = 3; float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
            if (yk >= 0 && xk >= 0) {
                const NodeInfo a1 = J.ni1[yk + 1], a2 = J.ni2[xk + 1];          // the nodes whose predecessor is followed
                const float sv = J.S[(size_t)(yk + 1) * J.ldS + (xk + 1)];
                c = C[skew(yk, xk, N2, pitch)];
                tr = a1.cnt == 1 && a1.idx0 == yk && a2.cnt == 1 && a2.idx0 == xk && isfinite(a1.val0) && isfinite(a2.val0) &&
                     isfinite(sv) && isfinite(c.w);
                if (yk == 0 && xk == 0) st = 4;                                 // start node: the walk ends here
                else if (c.w == c.x) st = 0; else if (c.w == c.z) st = 2; else if (c.w == c.y) st = 1;
                stM = (st == 0);
            }
            const unsigned mtr = __ballot_sync(0xffffffffu, tr), mM = __ballot_sync(0xffffffffu, stM);
            const unsigned cond = mtr & ((mM << 1) | 1u);                        // step k needs state M at k-1
            const int R = (cond == 0xffffffffu) ? 32 : (__ffs(~cond) - 1);
            if (R > 0) {
                const int stR = __shfl_sync(0xffffffffu, st, R - 1);
                if (stR == 3) { status = PGM_ERR_BACKTRACK; break; }            // reference: error("backtracking failed")
                if (lane < R && st != 4) {
                    const int pos = T.len + lane;
                    if (pos < J.cap) {
                        J.map1[pos] = (st == 1) ? -1 : yk;
                        J.map2[pos] = (st == 2) ? -1 : xk;
                    }
                }
                T.len += (stR == 4) ? R - 1 : R;
                const float cx = __shfl_sync(0xffffffffu, c.x, R - 1), cy = __shfl_sync(0xffffffffu, c.y, R - 1),
                            cz = __shfl_sync(0xffffffffu, c.z, R - 1);
                y -= R; x -= R;
                if (stR != 4) { state = stR; cur = (stR == 0) ? cx : (stR == 1 ? cy : cz); }
                trx = false; try_ = false; nx = x; ny = y; nstate = state; nscore = cur;
                guard += R;
                continue;
            }
        }
        {   // the path is mostly diagonal: pull the lines it will touch PF steps from now into L2
            const int py = y - PF, px = x - PF;
            if (py > 0 && px > 0) {
                const size_t o = skew(py, px, N2, pitch);
                asm volatile("prefetch.global.L2 [%0];" ::"l"(C + o));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(J.S + (size_t)py * J.ldS + px));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(J.ni1 + py));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(J.ni2 + px));
            }
        }
        best = FINF;
        bool landed_w = false;
        if (state == 2) {
            const NodeInfo ni = J.ni1[y];
            const int cnt = ni.cnt & 0x7fffffff;
            for (int ky = 0; ky < cnt; ++ky) {
                int yp; float ey; bool ry; pred1(ni, ky, yp, ey, ry);
                const float4 c = C[skew(yp, x, N2, pitch)];
                float q = fabsf(__fsub_rn(cur, __fsub_rn(__fadd_rn(c.z, gap_extend), ey)));
                if (best > q) { best = q; trx = false; try_ = ry; nx = x; ny = yp; nscore = c.z; nstate = 2; landed_w = false; }
                q = fabsf(__fsub_rn(cur, __fsub_rn(__fadd_rn(c.w, gap_init), ey)));
                if (best > q) { best = q; trx = false; try_ = ry; nx = x; ny = yp; landed_w = true; }
            }
        } else if (state == 1) {
            const NodeInfo ni = J.ni2[x];
            const int cnt = ni.cnt & 0x7fffffff;
            for (int kx = 0; kx < cnt; ++kx) {
                int xp; float ex; bool rx; pred2(ni, kx, xp, ex, rx);
                const float4 c = C[skew(y, xp, N2, pitch)];
                float q = fabsf(__fsub_rn(cur, __fsub_rn(__fadd_rn(c.y, gap_extend), ex)));
                if (best > q) { best = q; trx = rx; try_ = false; nx = xp; ny = y; nscore = c.y; nstate = 1; landed_w = false; }
                q = fabsf(__fsub_rn(cur, __fsub_rn(__fadd_rn(c.w, gap_init), ex)));
                if (best > q) { best = q; trx = rx; try_ = false; nx = xp; ny = y; landed_w = true; }
            }
        } else {
            const NodeInfo n1 = J.ni1[y], n2 = J.ni2[x];
            const float s = J.S[(size_t)y * J.ldS + x];
            const int c1 = n1.cnt & 0x7fffffff, c2 = n2.cnt & 0x7fffffff;
            for (int ky = 0; ky < c1; ++ky) {
                int yp; float ey; bool ry; pred1(n1, ky, yp, ey, ry);
                for (int kx = 0; kx < c2; ++kx) {
                    int xp; float ex; bool rx; pred2(n2, kx, xp, ex, rx);
                    const float4 c = C[skew(yp, xp, N2, pitch)];
                    const float q = fabsf(__fsub_rn(cur, __fsub_rn(__fsub_rn(__fadd_rn(c.w, s), ey), ex)));
                    if (best > q) { best = q; trx = rx; try_ = ry; ny = yp; nx = xp; landed_w = true; }
                }
            }
        }
        // Landing on W resolves the state by equality, M then Y then X (GraphAlign.h:399-411).  The
        // reference resolves at every improving candidate; only the last improvement survives, so
        // resolving once for the winner is equivalent -- except for its error(), which would fire on
        // any improving candidate whose W matches none of M/Y/X.  W is the max of exactly those, so
        // that cannot happen for a finite W; a non-finite W is reported as a backtracking failure.
        if (landed_w && (nx != 0 || ny != 0)) {
            const float4 c = C[skew(ny, nx, N2, pitch)];
            if (c.w == c.x) { nscore = c.x; nstate = 0; }
            else if (c.w == c.z) { nscore = c.z; nstate = 2; }
            else if (c.w == c.y) { nscore = c.y; nstate = 1; }
            else { status = PGM_ERR_BACKTRACK; break; }
        }
        n_tr += (trx ? 1u : 0u) + (try_ ? 1u : 0u);
        if (try_) T.alt_path(ny, y, 0);
        if (trx) T.alt_path(nx, x, 1);
        x = nx; y = ny; state = nstate; cur = nscore;
        if (x != 0 || y != 0) {
            if (state == 0) T.push(y, x); else if (state == 1) T.push(-1, x); else T.push(y, -1);
        }
    }
    T.push(0, 0);
    if (T.len > J.cap && status == 0) status = PGM_ERR_CAPACITY;
    J.result[0] = status; J.result[1] = T.len; J.result[2] = (int)n_tr;
}

template <bool STAGE, int RPT, int RING, bool SPF>
static int launch_fill_t(cudaStream_t st, int n_jobs, const GraphJobDev* d_jobs, const int* d_order, int threads, int ring_pitch, size_t smem) {
    PGM_CUDA(cudaFuncSetAttribute(graph_fill_kernel<STAGE, RPT, RING, SPF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    graph_fill_kernel<STAGE, RPT, RING, SPF><<<n_jobs, threads, smem, st>>>(d_jobs, d_order, ring_pitch);
    PGM_CUDA(cudaGetLastError());
    return 0;
}

// max_n1 / max_n2: largest graph-1 / graph-2 node counts in this launch
// d_order: the n_jobs job indices this launch sweeps (the host groups jobs of similar graph-1 length, because block size
// and shared memory -- hence CTAs per SM -- are set by the launch's largest job)
int launch_graph_fill(cudaStream_t st, int n_jobs, const GraphJobDev* d_jobs, const int* d_order, int max_n1, int max_n2, int max_np) {
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
#define GO(S_, R_) do { if (regular) return launch_fill_t<S_, R_, GF_RING, true>(st, n_jobs, d_jobs, d_order, threads, ring_pitch, smem); \
                        return launch_fill_t<S_, R_, GF_RING_LONG, false>(st, n_jobs, d_jobs, d_order, threads, ring_pitch, smem); } while (0)
    if (stage) { if (rpt == 1) GO(true, 1); if (rpt == 2) GO(true, 2); GO(true, 4); }
    if (rpt == 1) GO(false, 1); if (rpt == 2) GO(false, 2); GO(false, 4);
#undef GO
}

int launch_graph_trace(cudaStream_t st, int n_jobs, const GraphJobDev* d_jobs) {
    if (n_jobs == 0) return 0;
    graph_trace_kernel<<<n_jobs, 32, 0, st>>>(n_jobs, d_jobs);
    PGM_CUDA(cudaGetLastError());
    return 0;
}
