// hp1_graphdp.cu -- K5 (blocked wavefront fill + end-node match) and K6 (traceback walk).
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
// smaller node indices in both graphs), so one CTA sweeps one job with a barrier between the steps; a
// batch of independent jobs (all ready guide-tree nodes) fills the GPU.  The DP values never leave the
// SM: they live in a shared-memory ring of the last few sweep steps.  What is written to HBM is the
// packed traceback, one direction word (2 bytes) per cell (hp1_dir.cuh), and the {M,X,Y,W} of the few
// rows / columns that are read back later (value store: far predecessor edges, end-node match, the
// boundary row and column for the traceback).  The traceback chases the direction words and is
// bit-identical to the reference's closest-|delta| search on the dense matrices.
#include "hp1_dir.cuh"
#include <algorithm>
#include <cstdlib>

// Direction words live in a BLOCK-SKEWED layout.  The columns of graph 2 are cut (on the host, hp1_host.cu) into nb
// blocks: a block is a run of columns inside one aligned group of GF_K columns [8g, 8g+8), and every column whose node is
// not a chain node (several predecessors, or one that is not x-1) STARTS a block.  A column keeps its slot x % GF_K;
// colinfo[x] = (block << GF_LOGK) | slot.  Block b of row y is produced at sweep step st = y + b by thread y % nb, which
// stores the group's 8 words as one 16-byte vector at ((st * pitch + y % nb) * GF_K): a warp's store is one contiguous run.
__device__ __forceinline__ size_t cidx(int y, int x, const uint16_t* __restrict__ colinfo, int nb, int pitch) {
    const int ci = colinfo[x];
    return ((size_t)(y + (ci >> GF_LOGK)) * pitch + (y % nb)) * GF_K + (ci & (GF_K - 1));
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

__device__ __forceinline__ void st_global_u4(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// K5: blocked wavefront.  The rows of a job are dealt round-robin to nb threads (nb = column blocks per row):
// thread t sweeps rows t, t+nb, t+2nb, ... one after the other, one block (<= GF_K columns) per step, so that at step
// s it is at block b = (s - t) % nb of its current row y, row y-1 (thread t-1) is exactly one step ahead, and a thread
// is never idle between its first and its last block.  One barrier per step; ~(N1 + N2/GF_K) steps per job instead of
// N1 + N2 anti-diagonals, and the per-step overhead is shared by the block's cells.
//   phase 1 (independent across the slots): for every predecessor yp of row y fetch the window (Y,W)(yp, block b) and
//            W(yp, x0-1) -- from the shared-memory ring when it was produced < R steps ago, else from the value store --
//            and fold it into ys[k] and the diagonal term m[k];
//   phase 1b (only a block whose first column is not a chain node): that column's M and X terms over all predecessor
//            pairs; all its column predecessors lie in earlier blocks, i.e. in the ring or in the value store;
//   phase 2 (sequential in x): the X recurrence through the left neighbour in registers, W, the direction word.
// All eight slots of the group are always computed and their words stored; the slots outside the block ([lo, hi)) are
// never read: a block starts inside a group only at a non-chain column, whose terms come from phase 1b and not from its
// left neighbour or the window's previous slot.
// max() is exact, so accumulating the candidates in this order gives the reference's running maxima bit for bit; strict
// '>' updates in the reference's candidate order record the FIRST maximum, which is what its traceback picks.
// The ring keeps {Y,W} and X of the last R steps (M is never read back by the sweep).  S (row-major) is loaded one
// step ahead into registers as the aligned 32-byte group (thread-per-row means every lane touches its own line, so the
// L1 tag stage is the scarce resource: two 16-byte requests per lane and step), and pulled into L2 a few steps earlier.
template <int B, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
graph_fill_kernel(const GraphJobDev* __restrict__ jobs, const int* __restrict__ order, int Tp, int R, int ci_cap, int np_cap) {
    typedef typename DirWord<B>::T DW;
    constexpr uint32_t YF = DIR_YOPEN | (dir_mask<B>() << (4 + 2 * B));        // the Y fields of a direction word
    constexpr uint32_t MF = (dir_mask<B>() << 4) | (dir_mask<B>() << (4 + B)); // the M fields
    const GraphJobDev* Jp = jobs + order[blockIdx.x];
    const int N1 = Jp->N1, N2 = Jp->N2, pitch = Jp->pitch, nb = Jp->nb, ldS = Jp->ldS;
    const float gap_init = Jp->sc[0], gap_extend = Jp->sc[1], start_gap = Jp->sc[6], start_init = Jp->sc[7];
    const float NINF = -CUDART_INF_F;
    DW* const D = (DW*)Jp->D;
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
    unsigned char* gfl = (unsigned char*)(cis + ci_cap);           // [Tp]: bit 0: block's first column is not a chain node, bit 1: block has a kept column
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
            const int x = bs[b], xe = bs[b + 1];
            unsigned char g = 0;
            if (x >= 1) { const int cnt = __ldg(&ni2[x].cnt), i0 = __ldg(&ni2[x].idx0); g = ((cnt & 0x800000ff) == 1 && i0 == x - 1) ? 0 : 1; }
            for (int xx = x; xx < xe; ++xx) if (NI_KEEP(__ldg(&ni2[xx].cnt)) >= 0) g |= 2;
            gfl[b] = g;
        }
    }
    __syncthreads();

    const bool lane_on = t < nb;
    int y = t, b = 0;
    // any earlier cell (yy, xx) as {X, Y, W}: ring when produced < R steps ago, else the value store (written by this CTA >= R barriers ago)
    auto cell_at = [&](int yy, int xx, int s) -> float3 {
        const int ci = colinfo[xx];
        const int st = yy + (ci >> GF_LOGK), kk = ci & (K - 1);
        if (s - st < R) {
            int tp = t - (y - yy);                                 // yy % nb without the division when yy is a recent row
            if (tp < 0) { tp += nb; if (tp < 0) tp = yy % nb; }
            const int o = ((st & Rm) * K + kk) * Tp + tp;
            const float2 yw = ringYW[o];
            return make_float3(ringX[o], yw.x, yw.y);
        }
        const float4 c = dir_far_cell(*Jp, NI_KEEP(__ldg(&ni1[yy].cnt)), NI_KEEP(gni[xx].cnt), yy, xx);
        return make_float3(c.y, c.z, c.w);
    };

    NodeInfo ny = NodeInfo{0, 0, 0, 0.f}, nyn = ny;
    if (lane_on && y <= N1 - 2) {
        const int4 v = __ldg((const int4*)(ni1 + y));
        ny.ptr = v.x; ny.cnt = v.y; ny.idx0 = v.z; ny.val0 = __int_as_float(v.w);
    }
    // ---- S stream: position p of this thread = (row t + (p / nb) * nb, block p % nb).  The aligned group of S for
    // position p+1 is loaded while position p is computed; reads past the row's end stay inside the S allocation (rows
    // are contiguous, 64 bytes of slack at the end) and feed only slots that are never read back ----
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
    DW* dcur = D + (size_t)t * K;                                  // this thread's 8 words of sweep step s
    const size_t dstep = (size_t)K * pitch;
    for (int s = 0; s < n_steps; ++s, dcur += dstep) {
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
            const unsigned gflag = gfl[b];
            const bool gen0 = (gflag & 1u) != 0;
            float ys[K], m[K]; uint32_t dr[K];                     // dr: the direction word's Y and M fields so far
#pragma unroll
            for (int k = 0; k < K; ++k) { ys[k] = NINF; m[k] = NINF; dr[k] = 0u; }
            // ---- phase 1: row predecessors ----
            const int c1 = NI_CNT(ny.cnt);
            const float gi0 = (b == 0) ? start_gap : gap_init;
            auto fold_pred = [&](const uint32_t ky, int yp, float ey, const bool first) {
                const int dy = y - yp;
                float2 win[K]; float wprev = NINF;                 // wprev = W(yp, x0-1), used only when the block starts a group
                if (dy + 1 < R) {
                    int tp = t - dy;
                    if (tp < 0) { tp += nb; if (tp < 0) tp = yp % nb; }
                    const float2* r0 = ringYW + (((s - dy) & Rm) * K) * Tp + tp;
#pragma unroll
                    for (int k = 0; k < K; ++k) win[k] = r0[k * Tp];
                    if (b > 0) wprev = ringYW[(((s - dy - 1) & Rm) * K + (K - 1)) * Tp + tp].y;
                } else {                                           // far row predecessor: a kept row of the value store
                    const int kr = NI_KEEP(__ldg(&ni1[yp].cnt));
                    const float4* base = Jp->Vrow + (size_t)(kr < 0 ? 0 : kr) * Jp->ldVr + g8;
#pragma unroll
                    for (int k = 0; k < K; ++k) { const float4 c = __ldcg(base + k); win[k] = kr < 0 ? make_float2(CUDART_NAN_F, CUDART_NAN_F) : make_float2(c.z, c.w); }
                    if (b > 0) wprev = kr < 0 ? CUDART_NAN_F : __ldcg(base - 1).w;
                }
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const float gi = (k == 0) ? gi0 : gap_init;
                    const float cE = __fsub_rn(__fadd_rn(win[k].x, gap_extend), ey), cO = __fsub_rn(__fadd_rn(win[k].y, gi), ey);
                    const float wd = (k == 0) ? wprev : win[k - 1].y;
                    const float v = __fsub_rn(__fsub_rn(__fadd_rn(wd, sv[k]), ey), ex[k]);
                    if (first) {
                        ys[k] = fmaxf(cE, cO); dr[k] = (cO > cE) ? DIR_YOPEN : 0u;
                        m[k] = v;
                    } else {
                        if (cE > ys[k]) { ys[k] = cE; dr[k] = (dr[k] & ~YF) | (ky << (4 + 2 * B)); }
                        if (cO > ys[k]) { ys[k] = cO; dr[k] = (dr[k] & ~YF) | (ky << (4 + 2 * B)) | DIR_YOPEN; }
                        if (v > m[k]) { m[k] = v; dr[k] = (dr[k] & ~MF) | (ky << 4); }
                    }
                }
            };
            if (c1 > 0) fold_pred(0u, ny.idx0, ny.val0, true);      // the first predecessor is inline in the node info
            for (int i = 1; i < c1; ++i) fold_pred((uint32_t)i, __ldg(p1i + ny.ptr + i), __ldg(p1v + ny.ptr + i), false);
            const float gix = (y == 0) ? start_gap : gap_init;
            // ---- phase 1b: a first column that is not a chain node (all its predecessors are in earlier blocks) ----
            float xsg = NINF, mg = NINF; uint32_t dg = 0u;          // dg: the X and M fields of that column's direction word
            if (gen0) {
                const int4 v = *(const int4*)(gni + x0);
                const int c2 = NI_CNT(v.y);
                const float s0 = __ldg(S + (size_t)y * ldS + x0);  // == sv[lo]; y >= 1 whenever it matters (c1 == 0 for row 0)
                if (c1 > 0 || y == 0)
                    for (int j = 0; j < c2; ++j) {
                        const int xp = j ? p2i[v.x + j] : v.z;
                        const float exj = j ? p2v[v.x + j] : __int_as_float(v.w);
                        const float3 cl = cell_at(y, xp, s);
                        const float cE = __fsub_rn(__fadd_rn(cl.x, gap_extend), exj), cO = __fsub_rn(__fadd_rn(cl.z, gix), exj);
                        if (cE > xsg) { xsg = cE; dg = (dg & MF) | ((uint32_t)j << (4 + 3 * B)); }
                        if (cO > xsg) { xsg = cO; dg = (dg & MF) | ((uint32_t)j << (4 + 3 * B)) | DIR_XOPEN; }
                    }
                for (int i = 0; i < c1; ++i) {                      // pairs in the reference's nested order: rows outer, columns inner
                    const int yp = i ? __ldg(p1i + ny.ptr + i) : ny.idx0;
                    const float ey = i ? __ldg(p1v + ny.ptr + i) : ny.val0;
                    for (int j = 0; j < c2; ++j) {
                        const int xp = j ? p2i[v.x + j] : v.z;
                        const float exj = j ? p2v[v.x + j] : __int_as_float(v.w);
                        const float cw = cell_at(yp, xp, s).z;
                        const float vm = __fsub_rn(__fsub_rn(__fadd_rn(cw, s0), ey), exj);
                        if (vm > mg) { mg = vm; dg = (dg & ~MF) | ((uint32_t)i << 4) | ((uint32_t)j << (4 + B)); }
                    }
                }
            }
            const int gslot = gen0 ? lo : -1;
            const int krow = NI_KEEP(ny.cnt);
            const bool keep_any = krow >= 0 || (gflag & 2u) != 0;
            // ---- phase 2: along the row ----
            const int ro = ((s & Rm) * K) * Tp + t;
            uint32_t dw[K];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                float mk = m[k];
                uint32_t dk = dr[k];
                const float cE = __fsub_rn(__fadd_rn(leftX, gap_extend), ex[k]), cO = __fsub_rn(__fadd_rn(leftW, gix), ex[k]);
                float xs = fmaxf(cE, cO);
                if (cO > cE) dk |= DIR_XOPEN;
                if (k == gslot) { mk = mg; xs = xsg; dk = (dr[k] & YF) | dg; }
                if (k == 0 && b == 0) { mk = NINF; xs = NINF; }     // column 0: Y only (GraphAlign.h:212-222)
                float wv = fmaxf(mk, fmaxf(xs, ys[k]));
                if (k == 0 && b == 0 && y == 0) wv = start_init;
                leftX = xs; leftW = wv;
                dw[k] = dk | dir_wstate(mk, xs, ys[k], wv);
                if (k >= lo && k < hi) {
                    ringYW[ro + k * Tp] = make_float2(ys[k], wv); ringX[ro + k * Tp] = xs;
                    if (keep_any) {                                 // value store: kept row and / or kept column
                        const float4 out = make_float4(mk, xs, ys[k], wv);
                        if (krow >= 0) Jp->Vrow[(size_t)krow * Jp->ldVr + g8 + k] = out;
                        if (gflag & 2u) { const int kc = NI_KEEP(gni[g8 + k].cnt); if (kc >= 0) Jp->Vcol[(size_t)kc * Jp->ldVc + y] = out; }
                    }
                }
            }
            if (B == 3) st_global_u4(dcur, dw[0] | (dw[1] << 16), dw[2] | (dw[3] << 16), dw[4] | (dw[5] << 16), dw[6] | (dw[7] << 16));
            else { st_global_u4(dcur, dw[0], dw[1], dw[2], dw[3]); st_global_u4(dcur + 4, dw[4], dw[5], dw[6], dw[7]); }
            // ---- advance ----
            b = bn;
            if (yn != y) { y = yn; ny = nyn; }
        }
        __syncthreads();
    }
    // end-node match (GraphAlign.h:262-280) and the traceback's first pick
    if (t == 0) dir_end_match(*Jp);
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

// One WARP per job chases the direction words from the end node to the start node.  All 32 lanes run the walk
// redundantly (identical values, so the duplicate stores are benign and every load is a broadcast); the lanes earn
// their keep on DIAGONAL RUNS: in state M at chain nodes the next cell is forced to be (y-1, x-1), so lane k fetches the
// word of (y-k, x-k) and its nodes, the warp finds with two ballots how far the run extends, and up to 32 path steps are
// emitted at once instead of one dependent load per step.
// The boundary row and column follow the reference's closest-|delta| search on the kept values (hp1_dir.cuh).
template <int B>
__device__ __forceinline__ void graph_trace_job(const GraphJobDev& J, const int lane) {
    typedef typename DirWord<B>::T DW;
    const int N1 = J.N1, N2 = J.N2;
    const float gap_init = J.sc[0], gap_extend = J.sc[1];
    const DW* __restrict__ D = (const DW*)J.D;
    const CellIndex cix{J.colinfo, J.nb, J.pitch, N2, J.layout};
    const bool strip = J.layout == GF_LAYOUT_STRIP;
    const int sC = J.C, sW = strip_words(J.C), sN = J.nstrips;
    // the direction word of cell (yy, xx) in the canonical format of hp1_dir.cuh, whatever layout the fill wrote
    auto dword = [&](int yy, int xx) -> uint32_t {
        if (strip) {                                     // 4-bit cells (hp1_strip.cu): mlose, xbeats, yopen, xopen
            const uint32_t cs = J.colstrip[xx];
            const int j = (int)(cs >> 5), k = (int)(cs & 31u), g = k >> 3, gsz = min(8, sC - 8 * g);
            const uint32_t word = ((const uint32_t*)J.D)[((size_t)(yy + j) * sN + j) * sW + g];
            const uint32_t nib = (word >> (4 * (gsz - 1 - (k & 7)))) & 15u;
            uint32_t idx = 0u;                           // predecessor-index fields: only cells in non-chain rows / columns have any
            if (J.sgen) {
                const int r1 = J.ncidx1[yy];
                if (r1 >= 0) idx = ((const uint32_t*)J.RowIdx)[(size_t)r1 * J.ldRI + xx];
                else { const int r2 = J.ncidx2[xx]; if (r2 >= 0) idx = ((const uint32_t*)J.ColIdx)[(size_t)r2 * J.ldCI + yy]; }
            }
            return ((nib & 1u) ? ((nib & 2u) ? 1u : 2u) : 0u) | (nib & 12u) | (idx & ~15u);
        }
        return D[cix(yy, xx)];
    };
    const float FINF = CUDART_INF_F;
    Tracer T(J);
    int status = J.result[4];            // a failed emission (tensor flavour: bounded mbarrier wait timed out) fails the job
    unsigned n_tr = 0;
    if (status != 0) { J.result[0] = status; J.result[1] = 0; J.result[2] = 0; return; }
    // predecessor k of a node: the first one is inline in its NodeInfo (one 16-byte load)
    auto pred1 = [&](const NodeInfo& ni, int k, int& idx, bool& rep) {
        if (k == 0) { idx = ni.idx0; rep = ni.cnt < 0; }
        else { idx = J.p1_idx[ni.ptr + k]; rep = J.p1_rep[ni.ptr + k] != 0; }
    };
    auto pred2 = [&](const NodeInfo& ni, int k, int& idx, bool& rep) {
        if (k == 0) { idx = ni.idx0; rep = ni.cnt < 0; }
        else { idx = J.p2_idx[ni.ptr + k]; rep = J.p2_rep[ni.ptr + k] != 0; }
    };
    int state = 0;                       // 0 = M, 1 = X, 2 = Y
    int y = N1 - 1, x = N2 - 1;
    T.push(N1 - 1, N2 - 1);
    bool trx = false, try_ = false;
    {   // the fill's end-node match recorded the reference's first pick
        const int pick = J.result[5];
        const int ky = (pick >> 8) & 0xff, kx = (pick >> 16) & 0xff, st = pick & 3;
        const int p1 = J.p1_ptr[N1 - 1], p2 = J.p2_ptr[N2 - 1];
        if (J.p1_ptr[N1] - p1 <= ky || J.p2_ptr[N2] - p2 <= kx || !(((const float*)J.result)[3] > -FINF)) status = PGM_ERR_BACKTRACK;
        else {
            y = J.p1_idx[p1 + ky]; x = J.p2_idx[p2 + kx];
            try_ = J.p1_rep[p1 + ky] != 0; trx = J.p2_rep[p2 + kx] != 0;
            state = (st == 3) ? 0 : st;
        }
    }
    if (status == 0) {
        n_tr += (trx ? 1u : 0u) + (try_ ? 1u : 0u);
        if (try_) T.alt_path(y, N1 - 1, 0);
        if (trx) T.alt_path(x, N2 - 1, 1);
        if (x != 0 || y != 0) {
            if (state == 0) T.push(y, x); else if (state == 1) T.push(-1, x); else T.push(y, -1);
        }
    }
    long guard = 0;
    const long guard_max = 4L * (N1 + N2) + 16;
    const int PF = 8;                                   // prefetch distance (steps) along the diagonal
    const int kr0 = NI_KEEP(J.ni1[0].cnt), kc0 = NI_KEEP(J.ni2[0].cnt);   // row 0 / column 0 of the value store
    while (status == 0 && (x != 0 || y != 0)) {
        if (++guard > guard_max) { status = PGM_ERR_BACKTRACK; break; }
        if (state == 0 && y > 0 && x > 0) {
            // ---- diagonal run: lane -> k = lane+1, position (y-k, x-k), reached from (y-k+1, x-k+1) ----
            const int k = lane + 1, yk = y - k, xk = x - k;
            bool tr = false, stM = false; int st = 3;
            if (yk >= 0 && xk >= 0) {
                const NodeInfo a1 = J.ni1[yk + 1], a2 = J.ni2[xk + 1];          // the nodes whose predecessor is followed
                tr = NI_CHAIN(a1, yk + 1) && NI_CHAIN(a2, xk + 1);
                if (yk == 0 && xk == 0) st = 4;                                 // start node: the walk ends here
                else st = (int)(dword(yk, xk) & 3u);
                stM = (st == 0) && yk > 0 && xk > 0;                            // a boundary cell ends the run (closest-|delta| rule there)
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
                y -= R; x -= R;
                if (stR != 4) state = stR;
                guard += R;
                continue;
            }
        }
        {   // the path is mostly diagonal: pull the lines it will touch PF steps from now into L2
            const int py = y - PF, px = x - PF;
            if (py > 0 && px > 0) {
                if (!strip) asm volatile("prefetch.global.L2 [%0];" ::"l"(D + cix(py, px)));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(J.ni1 + py));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(J.ni2 + px));
            }
        }
        int ny = -1, nx = -1, nstate = state;
        bool landed_w = false;
        trx = false; try_ = false;
        if (state == 2) {
            const NodeInfo ni = J.ni1[y];
            const int cnt = NI_CNT(ni.cnt);
            if (x == 0) {
                // boundary column: the reference's closest-|delta| search with gap_init on the kept values (GraphAlign.h:378-398)
                const float cur = dir_far_cell(J, -1, kc0, y, 0).z;
                float best = FINF;
                for (int ky = 0; ky < cnt; ++ky) {
                    int yp; bool ry; pred1(ni, ky, yp, ry);
                    const float ey = ky ? J.p1_val[ni.ptr + ky] : ni.val0;
                    const float4 c = dir_far_cell(J, -1, kc0, yp, 0);
                    float q = fabsf(__fsub_rn(cur, __fsub_rn(__fadd_rn(c.z, gap_extend), ey)));
                    if (best > q) { best = q; try_ = ry; nx = 0; ny = yp; nstate = 2; landed_w = false; }
                    q = fabsf(__fsub_rn(cur, __fsub_rn(__fadd_rn(c.w, gap_init), ey)));
                    if (best > q) { best = q; try_ = ry; nx = 0; ny = yp; landed_w = true; }
                }
            } else {
                const uint32_t w = dword(y, x);
                const int ky = (int)dir_kyy<B>(w);
                if (ky < cnt) { int yp; bool ry; pred1(ni, ky, yp, ry); ny = yp; nx = x; try_ = ry; landed_w = (w & DIR_YOPEN) != 0; nstate = 2; }
            }
        } else if (state == 1) {
            const NodeInfo ni = J.ni2[x];
            const int cnt = NI_CNT(ni.cnt);
            if (y == 0) {
                const float cur = dir_far_cell(J, kr0, -1, 0, x).y;
                float best = FINF;
                for (int kx = 0; kx < cnt; ++kx) {
                    int xp; bool rx; pred2(ni, kx, xp, rx);
                    const float ex = kx ? J.p2_val[ni.ptr + kx] : ni.val0;
                    const float4 c = dir_far_cell(J, kr0, -1, 0, xp);
                    float q = fabsf(__fsub_rn(cur, __fsub_rn(__fadd_rn(c.y, gap_extend), ex)));
                    if (best > q) { best = q; trx = rx; nx = xp; ny = 0; nstate = 1; landed_w = false; }
                    q = fabsf(__fsub_rn(cur, __fsub_rn(__fadd_rn(c.w, gap_init), ex)));
                    if (best > q) { best = q; trx = rx; nx = xp; ny = 0; landed_w = true; }
                }
            } else {
                const uint32_t w = dword(y, x);
                const int kx = (int)dir_kxx<B>(w);
                if (kx < cnt) { int xp; bool rx; pred2(ni, kx, xp, rx); nx = xp; ny = y; trx = rx; landed_w = (w & DIR_XOPEN) != 0; nstate = 1; }
            }
        } else if (y > 0 && x > 0) {
            const NodeInfo n1 = J.ni1[y], n2 = J.ni2[x];
            const uint32_t w = dword(y, x);
            const int ky = (int)dir_kym<B>(w), kx = (int)dir_kxm<B>(w);
            if (ky < NI_CNT(n1.cnt) && kx < NI_CNT(n2.cnt)) {
                int yp, xp; bool ry, rx; pred1(n1, ky, yp, ry); pred2(n2, kx, xp, rx);
                ny = yp; nx = xp; try_ = ry; trx = rx; landed_w = true;
            }
        }
        // state M on the boundary has no predecessor pair (the reference's search would find none either): failure
        if (ny < 0) { status = PGM_ERR_BACKTRACK; break; }
        // Landing on W resolves the state by equality, M then Y then X (GraphAlign.h:399-411): the cell's wstate.
        if (landed_w && (nx != 0 || ny != 0)) {
            nstate = (int)(dword(ny, nx) & 3u);
            if (nstate == 3) { status = PGM_ERR_BACKTRACK; break; }
        }
        n_tr += (trx ? 1u : 0u) + (try_ ? 1u : 0u);
        if (try_) T.alt_path(ny, y, 0);
        if (trx) T.alt_path(nx, x, 1);
        x = nx; y = ny; state = nstate;
        if (x != 0 || y != 0) {
            if (state == 0) T.push(y, x); else if (state == 1) T.push(-1, x); else T.push(y, -1);
        }
    }
    T.push(0, 0);
    if (T.len > J.cap && status == 0) status = PGM_ERR_CAPACITY;
    J.result[0] = status; J.result[1] = T.len; J.result[2] = (int)n_tr;
}

__global__ void __launch_bounds__(32)
graph_trace_kernel(int n_jobs, const GraphJobDev* __restrict__ jobs, const int* __restrict__ order) {
    if ((int)blockIdx.x >= n_jobs) return;
    const GraphJobDev& J = jobs[order ? order[blockIdx.x] : blockIdx.x];
    if (J.dbits == 3) graph_trace_job<3>(J, (int)threadIdx.x); else graph_trace_job<7>(J, (int)threadIdx.x);
}

// max_nb: most column blocks per row (= threads that work) among this launch's jobs.  d_order: the n_jobs job indices
// the launch sweeps (the host groups jobs of similar graph-2 length, because block size and shared memory -- hence CTAs
// per SM -- are set by the launch's largest job)
template <int B>
static int launch_graph_fill_b(cudaStream_t st, int n_jobs, const GraphJobDev* d_jobs, const int* d_order, int max_nb, int max_n2, int max_np2) {
    const int Tp = std::max(32, (max_nb + 31) & ~31);
    if (Tp > GF_MAX_NB) { pgm_set_error("graph 2 has more than %d nodes; the wavefront kernel's shared-memory ring does not hold it", GF_MAX_NB * GF_K); return PGM_ERR_ARG; }
    // Ring depth.  Launches of small blocks hold mostly chain graphs and want as many resident jobs per SM as possible:
    // 4 steps.  Larger blocks mean graphs with many non-chain nodes (every one starts a block), whose predecessors reach
    // further back: 8 steps (7 rows) and the column table in shared memory, as long as it fits.  (The host's keep sets
    // assume only the 4-step ring, so either depth is safe.)
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
        graph_fill_kernel<B, 64, 8><<<n_jobs, Tp, smem, st>>>(d_jobs, d_order, Tp, R, ci_cap, np_cap);
    } else if (Tp <= 128) {
        if (smem > 48 * 1024) PGM_CUDA(cudaFuncSetAttribute(graph_fill_kernel<B, 128, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
        graph_fill_kernel<B, 128, 4><<<n_jobs, Tp, smem, st>>>(d_jobs, d_order, Tp, R, ci_cap, np_cap);
    } else {
        PGM_CUDA(cudaFuncSetAttribute(graph_fill_kernel<B, GF_MAX_NB, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
        graph_fill_kernel<B, GF_MAX_NB, 1><<<n_jobs, Tp, smem, st>>>(d_jobs, d_order, Tp, R, ci_cap, np_cap);
    }
    PGM_CUDA(cudaGetLastError());
    return 0;
}

int launch_graph_fill(cudaStream_t st, int dbits, int n_jobs, const GraphJobDev* d_jobs, const int* d_order, int max_nb, int max_n2, int max_np2) {
    if (n_jobs == 0) return 0;
    return dbits == 3 ? launch_graph_fill_b<3>(st, n_jobs, d_jobs, d_order, max_nb, max_n2, max_np2)
                      : launch_graph_fill_b<7>(st, n_jobs, d_jobs, d_order, max_nb, max_n2, max_np2);
}

int launch_graph_trace(cudaStream_t st, int n_jobs, const GraphJobDev* d_jobs, const int* d_order) {
    if (n_jobs == 0) return 0;
    graph_trace_kernel<<<n_jobs, 32, 0, st>>>(n_jobs, d_jobs, d_order);
    PGM_CUDA(cudaGetLastError());
    return 0;
}
