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

// Cell (y,x) of the {M,X,Y,W} matrices lives in a BLOCK-SKEWED layout.  The columns of graph 2 are cut (on the host,
// hp1_host.cu) into nb blocks: a block is a run of columns inside one aligned group of GF_K columns [8g, 8g+8), and every
// column whose node is not a chain node (several predecessors, or one that is not x-1) STARTS a block.  A column keeps its
// slot x % GF_K; colinfo[x] = (block << GF_LOGK) | slot.  Block b of row y is produced at sweep step st = y + b by thread
// y % nb, and the cell sits at ((st * GF_K + slot) * pitch + y % nb): a warp's store of one slot is a contiguous run.
__device__ __forceinline__ size_t cidx(int y, int x, const uint16_t* __restrict__ colinfo, int nb, int pitch) {
    const int ci = colinfo[x];
    return ((size_t)(y + (ci >> GF_LOGK)) * GF_K + (ci & (GF_K - 1))) * pitch + (y % nb);
}

// either layout, for the traceback
struct CellIndex {
    const uint16_t* __restrict__ colinfo; int nb, pitch, N2, layout;
    __device__ __forceinline__ size_t operator()(int y, int x) const {
        if (layout == GF_LAYOUT_DIAG) { const int d = y + x; return (size_t)d * pitch + (y - max(0, d - (N2 - 1))); }
        return cidx(y, x, colinfo, nb, pitch);
    }
};

#define GF_SPF 6             // S is pulled into L2 this many steps before the step that loads it into registers
#define GF_SMEM_PER_SLOT(R) ((R) * GF_K * 12 + GF_K * 4 + 2 + 1)   // shared-memory bytes per thread slot; R = ring depth in sweep steps (power of two >= 4)

__device__ __forceinline__ void st_global_f4(float4* p, const float4& v) {
    asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// K5: blocked wavefront.  The rows of a job are dealt round-robin to nb threads (nb = column blocks per row):
// thread t sweeps rows t, t+nb, t+2nb, ... one after the other, one block (<= GF_K columns) per step, so that at step
// s it is at block b = (s - t) % nb of its current row y, row y-1 (thread t-1) is exactly one step ahead, and a thread
// is never idle between its first and its last block.  One barrier per step; ~(N1 + N2/GF_K) steps per job instead of
// N1 + N2 anti-diagonals, and the per-step overhead is shared by the block's cells.
//   phase 1 (independent across the slots): for every predecessor yp of row y fetch the window (Y,W)(yp, block b) and
//            W(yp, x0-1) -- from the shared-memory ring when it was produced < R steps ago, else from HBM/L2 -- and
//            fold it into ys[k] and the diagonal term m[k];
//   phase 1b (only a block whose first column is not a chain node): that column's M and X terms over all predecessor
//            pairs; all its column predecessors lie in earlier blocks, i.e. in the ring or in HBM/L2;
//   phase 2 (sequential in x): the X recurrence through the left neighbour in registers.
// All eight slots of the group are always computed; the slots outside the block ([lo, hi)) are simply not stored and
// nothing valid ever reads them: a block starts inside a group only at a non-chain column, whose terms come from phase 1b
// and not from its left neighbour or the window's previous slot.
// max() is exact, so accumulating the candidates in this order gives the reference's running maxima bit for bit.
// The ring keeps {Y,W} and X of the last R steps (M is never read back by the sweep).  S (row-major) is loaded one
// step ahead into registers as the aligned 32-byte group (thread-per-row means every lane touches its own line, so the
// L1 tag stage is the scarce resource: two 16-byte requests per lane and step), and pulled into L2 a few steps earlier.
template <int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
graph_fill_kernel(const GraphJobDev* __restrict__ jobs, const int* __restrict__ order, int Tp, int R, int ci_cap, int np_cap) {
    const GraphJobDev* Jp = jobs + order[blockIdx.x];
    const int N1 = Jp->N1, N2 = Jp->N2, pitch = Jp->pitch, nb = Jp->nb, ldS = Jp->ldS;
    const float gap_init = Jp->sc[0], gap_extend = Jp->sc[1], start_gap = Jp->sc[6], start_init = Jp->sc[7];
    const float NINF = -CUDART_INF_F;
    float4* const C = Jp->MXYW;
    const float* const S = Jp->S;
    const NodeInfo* const ni1 = Jp->ni1;
    const NodeInfo* const ni2 = Jp->ni2;
    const int* const p1i = Jp->p1_idx; const float* const p1v = Jp->p1_val;
    const int t = (int)threadIdx.x;
    constexpr int K = GF_K;
    const int Rm = R - 1;                                          // ring depth R: sweep steps of {X,Y,W} blocks kept in shared memory

    extern __shared__ __align__(16) unsigned char gsm[];
    float2* ringYW = (float2*)gsm;                                 // [R][K][Tp]
    float* ringX = (float*)(ringYW + (size_t)R * K * Tp);          // [R][K][Tp]
    float* exs = ringX + (size_t)R * K * Tp;                       // [Tp * K]: first-predecessor edge value of graph-2 node x
    unsigned short* bst = (unsigned short*)(exs + (size_t)Tp * K); // [Tp + 2]: first column of block b (bst[nb] = N2 - 1)
    unsigned short* cis = bst + Tp + 2;                            // [ci_cap]: colinfo staged (launches whose jobs have many non-chain nodes)
    unsigned char* gfl = (unsigned char*)(cis + ci_cap);           // [Tp]: block's first column is not a chain node
    const uint16_t* colinfo = Jp->colinfo;
    const int* p2i = Jp->p2_idx; const float* p2v = Jp->p2_val;
    const NodeInfo* gni = ni2;                                     // node infos / predecessor lists read by phase 1b
    if (ci_cap > 0) {                                              // staged: phase 1b sits on the sweep's critical path in these jobs
        for (int x = t; x < N2; x += (int)blockDim.x) cis[x] = colinfo[x];
        colinfo = cis;
        int4* sni = (int4*)(gsm + (((size_t)GF_SMEM_PER_SLOT(R) * Tp + 2 * (size_t)ci_cap + 32 + 15) & ~(size_t)15));   // [N2] + lists [np2]
        const int np2 = Jp->np2;
        if (np_cap >= np2) {
            for (int x = t; x < N2; x += (int)blockDim.x) sni[x] = __ldg((const int4*)(ni2 + x));
            int* spi = (int*)(sni + ((ci_cap + 7) & ~7)); float* spv = (float*)(spi + np_cap);
            for (int e = t; e < np2; e += (int)blockDim.x) { spi[e] = p2i[e]; spv[e] = p2v[e]; }
            gni = (const NodeInfo*)sni; p2i = spi; p2v = spv;
        }
    }
    {
        const uint16_t* bs = Jp->blkstart;
        for (int b = t; b <= nb; b += (int)blockDim.x) bst[b] = bs[b];
        const int ncol = (N2 - 1 + K - 1) & ~(K - 1);
        for (int x = t; x < ncol; x += (int)blockDim.x) exs[x] = (x >= 1 && x <= N2 - 2) ? __ldg(&ni2[x].val0) : 0.f;
        for (int b = t; b < nb; b += (int)blockDim.x) {
            const int x = bs[b];
            bool g = false;
            if (x >= 1) { const int cnt = __ldg(&ni2[x].cnt), i0 = __ldg(&ni2[x].idx0); g = !(cnt == 1 && i0 == x - 1); }
            gfl[b] = g ? 1 : 0;
        }
    }
    __syncthreads();

    const bool lane_on = t < nb;
    int y = t, b = 0;
    // any earlier cell (yy, xx) as {X, Y, W}: ring when produced < R steps ago, else HBM/L2 (written by this CTA >= R barriers ago)
    auto cell_at = [&](int yy, int xx, int s) -> float3 {
        const int ci = colinfo[xx];
        const int st = yy + (ci >> GF_LOGK), kk = ci & (K - 1);
        int tp = t - (y - yy);                                     // yy % nb without the division when yy is a recent row
        if (tp < 0) { tp += nb; if (tp < 0) tp = yy % nb; }
        if (s - st < R) {
            const int o = ((st & Rm) * K + kk) * Tp + tp;
            const float2 yw = ringYW[o];
            return make_float3(ringX[o], yw.x, yw.y);
        }
        const float4 c = __ldcg(C + ((size_t)st * K + kk) * pitch + tp);
        return make_float3(c.y, c.z, c.w);
    };

    NodeInfo ny = NodeInfo{0, 0, 0, 0.f}, nyn = ny;
    if (lane_on && y <= N1 - 2) {
        const int4 v = __ldg((const int4*)(ni1 + y));
        ny.ptr = v.x; ny.cnt = v.y; ny.idx0 = v.z; ny.val0 = __int_as_float(v.w);
    }
    // ---- S stream: position p of this thread = (row t + (p / nb) * nb, block p % nb).  The aligned group of S for
    // position p+1 is loaded while position p is computed; reads past the row's end stay inside the S allocation (rows
    // are contiguous, 64 bytes of slack at the end) and feed only slots that are never stored ----
    float4 sna = make_float4(0.f, 0.f, 0.f, 0.f), snb = sna;
    if (lane_on && y >= 1 && y <= N1 - 2) {
        const float4* sp = (const float4*)(S + (size_t)y * ldS);
        sna = __ldg(sp); snb = __ldg(sp + 1);
    }
    int yF = t, bF = 0;                                            // L2 prefetch position, GF_SPF + 1 ahead
#pragma unroll 1
    for (int d = 0; d <= GF_SPF; ++d) if (++bF == nb) { bF = 0; yF += nb; }
    float leftX = NINF, leftW = NINF;
    const int n_steps = (N1 - 1) + nb - 1;
    float4* gcur = C + t;                                          // this thread's slot of sweep step s
    const size_t gstep = (size_t)K * pitch;
    for (int s = 0; s < n_steps; ++s, gcur += gstep) {
        if (lane_on && s >= t && y <= N1 - 2) {
            const int x0 = bst[b], x1 = bst[b + 1];                // this block: columns x0 .. x1-1, slots lo .. hi-1
            const int lo = x0 & (K - 1), hi = lo + (x1 - x0), g8 = x0 - lo;
            float sv[K];
            sv[0] = sna.x; sv[1] = sna.y; sv[2] = sna.z; sv[3] = sna.w; sv[4] = snb.x; sv[5] = snb.y; sv[6] = snb.z; sv[7] = snb.w;
            // ---- next position's S group (registers) and the L2 prefetch further ahead ----
            int bn = b + 1, yn = y;
            if (bn == nb) { bn = 0; yn = y + nb; }
            if (yn >= 1 && yn <= N1 - 2) {
                const float4* sp = (const float4*)(S + (size_t)yn * ldS + (bst[bn] & ~(K - 1)));
                sna = __ldg(sp); snb = __ldg(sp + 1);
            }
            {
                const int xf = bst[bF];
                if ((xf & 31) == 0 && yF >= 1 && yF <= N1 - 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(S + (size_t)yF * ldS + xf));
                if (++bF == nb) { bF = 0; yF += nb; }
            }
            if (b == 0 && y + nb <= N1 - 2) {                      // next row's node info, a whole row ahead
                const int4 v = __ldg((const int4*)(ni1 + y + nb));
                nyn.ptr = v.x; nyn.cnt = v.y; nyn.idx0 = v.z; nyn.val0 = __int_as_float(v.w);
            }
            float ex[K];
            {
                const float4 e0 = *(const float4*)(exs + g8), e1 = *(const float4*)(exs + g8 + 4);
                ex[0] = e0.x; ex[1] = e0.y; ex[2] = e0.z; ex[3] = e0.w; ex[4] = e1.x; ex[5] = e1.y; ex[6] = e1.z; ex[7] = e1.w;
            }
            const bool gen0 = gfl[b] != 0;
            float ys[K], m[K];
#pragma unroll
            for (int k = 0; k < K; ++k) { ys[k] = NINF; m[k] = NINF; }
            // ---- phase 1: row predecessors ----
            const int c1 = ny.cnt & 0x7fffffff;
            const float gi0 = (b == 0) ? start_gap : gap_init;
            auto fold_pred = [&](int yp, float ey) {
                const int dy = y - yp;
                float2 win[K]; float wprev = NINF;                 // wprev = W(yp, x0-1), used only when the block starts a group
                if (dy + 1 < R) {
                    int tp = t - dy;
                    if (tp < 0) { tp += nb; if (tp < 0) tp = yp % nb; }
                    const float2* r0 = ringYW + (((s - dy) & Rm) * K) * Tp + tp;
#pragma unroll
                    for (int k = 0; k < K; ++k) win[k] = r0[k * Tp];
                    if (b > 0) wprev = ringYW[(((s - dy - 1) & Rm) * K + (K - 1)) * Tp + tp].y;
                } else {
                    const int tp = yp % nb;
                    const float4* base = C + ((size_t)(yp + b) * K) * pitch + tp;
#pragma unroll
                    for (int k = 0; k < K; ++k) { const float4 c = __ldcg(base + (size_t)k * pitch); win[k] = make_float2(c.z, c.w); }
                    if (b > 0) wprev = __ldcg(C + ((size_t)(yp + b - 1) * K + (K - 1)) * pitch + tp).w;
                }
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const float gi = (k == 0) ? gi0 : gap_init;
                    ys[k] = fmaxf(ys[k], __fsub_rn(fmaxf(__fadd_rn(win[k].x, gap_extend), __fadd_rn(win[k].y, gi)), ey));
                    const float wd = (k == 0) ? wprev : win[k - 1].y;
                    m[k] = fmaxf(m[k], __fsub_rn(__fsub_rn(__fadd_rn(wd, sv[k]), ey), ex[k]));
                }
            };
            if (c1 > 0) fold_pred(ny.idx0, ny.val0);                // the first predecessor is inline in the node info
            for (int i = 1; i < c1; ++i) fold_pred(__ldg(p1i + ny.ptr + i), __ldg(p1v + ny.ptr + i));
            const float gix = (y == 0) ? start_gap : gap_init;
            // ---- phase 1b: a first column that is not a chain node (all its predecessors are in earlier blocks) ----
            float xsg = NINF, mg = NINF;
            if (gen0) {
                const int4 v = *(const int4*)(gni + x0);
                const int c2 = v.y & 0x7fffffff;
                const float s0 = __ldg(S + (size_t)y * ldS + x0);  // == sv[lo]; y >= 1 whenever it matters (c1 == 0 for row 0)
                for (int j = 0; j < c2; ++j) {
                    const int xp = j ? p2i[v.x + j] : v.z;
                    const float exj = j ? p2v[v.x + j] : __int_as_float(v.w);
                    const float3 cl = cell_at(y, xp, s);
                    xsg = fmaxf(xsg, __fsub_rn(fmaxf(__fadd_rn(cl.x, gap_extend), __fadd_rn(cl.z, gix)), exj));
                    for (int i = 0; i < c1; ++i) {
                        const int yp = i ? __ldg(p1i + ny.ptr + i) : ny.idx0;
                        const float ey = i ? __ldg(p1v + ny.ptr + i) : ny.val0;
                        const float cw = cell_at(yp, xp, s).z;
                        mg = fmaxf(mg, __fsub_rn(__fsub_rn(__fadd_rn(cw, s0), ey), exj));
                    }
                }
            }
            const int gslot = gen0 ? lo : -1;
            // ---- phase 2: along the row ----
            const int ro = ((s & Rm) * K) * Tp + t;
            float4* gk = gcur;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                float mk = m[k];
                float xs = __fsub_rn(fmaxf(__fadd_rn(leftX, gap_extend), __fadd_rn(leftW, gix)), ex[k]);
                if (k == gslot) { mk = mg; xs = xsg; }
                if (k == 0 && b == 0) { mk = NINF; xs = NINF; }     // column 0: Y only (GraphAlign.h:212-222)
                float wv = fmaxf(mk, fmaxf(xs, ys[k]));
                if (k == 0 && b == 0 && y == 0) wv = start_init;
                leftX = xs; leftW = wv;
                if (k >= lo && k < hi) {
                    ringYW[ro + k * Tp] = make_float2(ys[k], wv); ringX[ro + k * Tp] = xs;
                    st_global_f4(gk, make_float4(mk, xs, ys[k], wv));
                }
                gk += pitch;
            }
            // ---- advance ----
            b = bn;
            if (yn != y) { y = yn; ny = nyn; }
        }
        __syncthreads();
    }
    // end-node match (GraphAlign.h:262-280)
    if (t == 0) {
        const GraphJobDev& J = *Jp;
        const float end_match = J.sc[3], end_gap = J.sc[4], end_skip = J.sc[5];
        float Wend = NINF;
        for (int ky = J.p1_ptr[N1 - 1]; ky < J.p1_ptr[N1]; ++ky)
            for (int kx = J.p2_ptr[N2 - 1]; kx < J.p2_ptr[N2]; ++kx) {
                const int yp = J.p1_idx[ky], xp = J.p2_idx[kx];
                const float ey = J.p1_val[ky], ex = J.p2_val[kx];
                if (xp == 0 && yp == 0) {
                    Wend = fmaxf(__fsub_rn(__fsub_rn(end_skip, ey), ex), Wend);
                } else {
                    const float4 c = __ldcg(C + cidx(yp, xp, colinfo, nb, pitch));
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
graph_trace_kernel(int n_jobs, const GraphJobDev* __restrict__ jobs, const int* __restrict__ order) {
    const int lane = threadIdx.x;
    if ((int)blockIdx.x >= n_jobs) return;
    const GraphJobDev& J = jobs[order ? order[blockIdx.x] : blockIdx.x];
    if (J.layout == GF_LAYOUT_STRIP) return;             // the packed traceback (hp1_trace.cu) walks the strip kernels' jobs
    if (J.result[4] != 0) {                              // a failed emission (tensor flavour: bounded mbarrier wait timed out) fails the job
        if (lane == 0) { J.result[0] = J.result[4]; J.result[1] = 0; J.result[2] = 0; }
        return;
    }
    const int N1 = J.N1, N2 = J.N2, pitch = J.pitch, nb = J.nb;
    const float gap_init = J.sc[0], gap_extend = J.sc[1], end_match = J.sc[3], end_gap = J.sc[4], end_skip = J.sc[5];
    const float4* __restrict__ C = J.MXYW;
    const CellIndex cix{J.colinfo, nb, pitch, N2, J.layout};
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
            const float4 c = C[cix(yp, xp)];
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
                c = C[cix(yk, xk)];
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
                const size_t o = cix(py, px);
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
                const float4 c = C[cix(yp, x)];
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
                const float4 c = C[cix(y, xp)];
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
                    const float4 c = C[cix(yp, xp)];
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
            const float4 c = C[cix(ny, nx)];
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

// max_nb: most column blocks per row (= threads that work) among this launch's jobs.  d_order: the n_jobs job indices
// the launch sweeps (the host groups jobs of similar graph-2 length, because block size and shared memory -- hence CTAs
// per SM -- are set by the launch's largest job)
int launch_graph_fill(cudaStream_t st, int n_jobs, const GraphJobDev* d_jobs, const int* d_order, int max_nb, int max_n2, int max_np2) {
    if (n_jobs == 0) return 0;
    const int Tp = std::max(32, (max_nb + 31) & ~31);
    if (Tp > GF_MAX_NB) { pgm_set_error("graph 2 has more than %d nodes; the wavefront kernel's shared-memory ring does not hold it", GF_MAX_NB * GF_K); return PGM_ERR_ARG; }
    // Ring depth.  Launches of small blocks hold mostly chain graphs and want as many resident jobs per SM as possible:
    // 4 steps.  Larger blocks mean graphs with many non-chain nodes (every one starts a block), whose predecessors reach
    // further back: 8 steps (7 rows) and the column table in shared memory, as long as it fits.
    int R = 4, ci_cap = 0, np_cap = 0;
    size_t stage = 0;
    if (Tp > 64) {
        ci_cap = (max_n2 + 8) & ~7;
        np_cap = (max_np2 + 8) & ~7;
        stage = 16 + (size_t)((ci_cap + 7) & ~7) * 16 + (size_t)np_cap * 8;      // graph 2's node infos and predecessor lists
        if ((size_t)GF_SMEM_PER_SLOT(8) * Tp + 2 * ci_cap + 32 + stage <= 200 * 1024) R = 8;
        if ((size_t)GF_SMEM_PER_SLOT(R) * Tp + 2 * ci_cap + 32 + stage > 200 * 1024) { stage = 0; np_cap = 0; }
    }
    const size_t smem = (size_t)GF_SMEM_PER_SLOT(R) * Tp + 2 * (size_t)ci_cap + 32 + stage;
    if (smem > 227 * 1024) { pgm_set_error("graph 2 too large for the wavefront kernel's shared memory (%d blocks)", max_nb); return PGM_ERR_ARG; }
    // small blocks: cap the registers so that 8 CTAs of 64 threads (or 16 of 32) fit an SM; the sweep is latency-bound per
    // job, so resident jobs per SM is what fills the issue slots
    if (Tp <= 64) {
        graph_fill_kernel<64, 8><<<n_jobs, Tp, smem, st>>>(d_jobs, d_order, Tp, R, ci_cap, np_cap);
    } else if (Tp <= 128) {
        if (smem > 48 * 1024) PGM_CUDA(cudaFuncSetAttribute(graph_fill_kernel<128, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
        graph_fill_kernel<128, 4><<<n_jobs, Tp, smem, st>>>(d_jobs, d_order, Tp, R, ci_cap, np_cap);
    } else {
        PGM_CUDA(cudaFuncSetAttribute(graph_fill_kernel<GF_MAX_NB, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
        graph_fill_kernel<GF_MAX_NB, 1><<<n_jobs, Tp, smem, st>>>(d_jobs, d_order, Tp, R, ci_cap, np_cap);
    }
    PGM_CUDA(cudaGetLastError());
    return 0;
}

int launch_graph_trace(cudaStream_t st, int n_jobs, const GraphJobDev* d_jobs, const int* d_order) {
    if (n_jobs == 0) return 0;
    graph_trace_kernel<<<n_jobs, 32, 0, st>>>(n_jobs, d_jobs, d_order);
    PGM_CUDA(cudaGetLastError());
    return 0;
}
