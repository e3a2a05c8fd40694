// hp1_wave.cu -- K5, anti-diagonal wavefront fill with ROLE-SORTED warps and ONE-STEP LOOK-AHEAD (the flavour for graphs
// with many non-chain nodes: the upper levels of a guide tree, whose sweeps end a mixed batch).
//
// Same recurrences as hp1_graphdp_diag.cu (reference GraphAlign.h:205-280), same {M,X,Y,W} cell layout (anti-diagonal-major,
// so the dense traceback of hp1_graphdp.cu re-traces the job unchanged).  What changed is how one anti-diagonal's work is
// dealt.  ncu on the thread-per-row kernel (profiles/r2_diag_branchy_before.txt): half of all instructions ran with 1.7 of 32 lanes
// active (fast path, multi-predecessor row path and general path serialised in every warp) and 60 % of the warp samples
// sat at the barrier waiting for those warps.  Here
//
//   * chain rows (one predecessor, the regular edge from y-1) own a thread each, packed into their own warps;
//   * rows with other / more predecessors own a thread each in SEPARATE warps;
//   * a NON-CHAIN COLUMN is swept by a group of L lanes (1, 2 or 4) in further warps that compute the column's cell of each
//     anti-diagonal in full (row threads skip those cells); the predecessor pairs (yp, xp) are dealt over the L lanes and
//     folded with shuffles.  max is exact, associative and idempotent, so the fold order does not change a bit.  A column is
//     busy for N1 - 1 anti-diagonals; a lane group then moves on to the next column the host dealt it;
//   * every candidate of cell (y, x) except the two that come through the ADJACENT edges -- (y-1, x) and (y, x-1), made on
//     the previous anti-diagonal -- lies at least two anti-diagonals back.  So a thread folds all non-adjacent candidates of
//     its NEXT cell during the current step (predecessor-list walks, ring or L2 reads, shuffles: all off the critical
//     path), and the step that makes the cell only adds the adjacent candidates: two shared-memory reads and six
//     float operations for every role.  An absent adjacent edge is the edge value +inf, which turns its candidates into
//     -inf without a branch.
//
// One barrier per anti-diagonal orders everything.  Cells of the last RING (8 or 16) anti-diagonals are read from a
// shared-memory ring indexed by row ({X,W,Y,W} per cell: the halves a candidate needs are one 8-byte read each); older
// ones (edges spanning more positions) from the {M,X,Y,W} array in L2 -- during the look-ahead.  S is streamed per row
// thread with cp.async, four columns every fourth anti-diagonal (a warp-uniform event), 9 .. 12 columns ahead.
// Graphs in which a node has TWO list entries from its adjacent predecessor (a regular and a repeat edge) stay with the
// thread-per-row kernel (hp1_host.cu checks).
#include "internal.h"
#include <math_constants.h>
#include <algorithm>
#include <cstdlib>

#define WAVE_SYNC() asm volatile("bar.sync 0;" ::: "memory")
// Shared memory is addressed with 32-bit shared-window addresses and explicit ld.shared / st.shared: through generic
// pointers the compiler re-derived the window base (S2R SR_CgaCtaId + LEA) in front of most accesses of the loop
__device__ __forceinline__ float4 lds128(uint32_t a) {
    float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a)); return v;
}
__device__ __forceinline__ float2 lds64(uint32_t a) { float2 v; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ float lds32(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ int lds32i(uint32_t a) { int v; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ int lds16u(uint32_t a) { unsigned short v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a)); return (int)v; }
__device__ __forceinline__ void sts128(uint32_t a, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// a cell older than the ring, re-read from the {M,X,Y,W} array (L2).  Out of line on purpose: inlined, the compiler computed
// the 64-bit address in front of every ring read (12 % of the kernel's instructions, profiles/r2_wave_v3.txt) although far
// predecessors are rare
__device__ __noinline__ float4 wave_far_cell(const float4* __restrict__ C, int dd, int yy, int pitch, int N2) {
    return C[(size_t)dd * pitch + (yy - max(0, dd - (N2 - 1)))];
}

template <int RING, int L, int MAXT>
__global__ void __launch_bounds__(MAXT, MAXT <= 512 ? 2 : 1)
graph_fill_wave_kernel(const GraphJobDev* __restrict__ jobs, const int* __restrict__ order, int ring_pitch, int max_tr,
                       int max_n1, int max_n2, int max_np1, int max_nx) {
    const GraphJobDev& J = jobs[order[blockIdx.x]];
    const int N1 = J.N1, N2 = J.N2, pitch = J.pitch;
    const float gap_init = J.sc[0], gap_extend = J.sc[1], start_gap = J.sc[6], start_init = J.sc[7];
    const float NINF = -CUDART_INF_F;
    float4* __restrict__ C = J.MXYW;
    const int tid = (int)threadIdx.x, T = (int)blockDim.x;
    const int TR = J.wv_tr, NSC = J.wv_ncols, TRC = J.wv_chain_pad;

    extern __shared__ __align__(16) unsigned char wsm[];
    float4* ring = (float4*)wsm;                                        // [RING][ring_pitch], row y at index y + 1 (index 0: -inf guard)
    float4* sring = ring + (size_t)RING * ring_pitch;                   // [4][max_tr]: S(y, 4q .. 4q+3) of row slot t at [(q & 3)][t]
    NodeInfo* n1 = (NodeInfo*)(sring + (size_t)4 * max_tr);             // [max_n1] graph 1's node infos (column threads)
    WaveCI* ci = (WaveCI*)(n1 + max_n1);                                // [max_n2]
    int* p1i = (int*)(ci + max_n2); float* p1v = (float*)(p1i + max_np1);   // graph 1's predecessor lists
    int* sxi = (int*)(p1v + max_np1); float* sxv = (float*)(sxi + max_nx);  // extra predecessors of the non-chain rows
    float* adj1 = (float*)(sxv + max_nx);                               // [max_n1] value of the edge y-1 -> y, +inf if there is none
    uint16_t* rowslot = (uint16_t*)(adj1 + max_n1);                     // [max_n1] row -> thread slot
    {
        const float4 ninf4 = make_float4(NINF, NINF, NINF, NINF);
        for (int i = tid; i < RING * ring_pitch; i += T) ring[i] = ninf4;
        for (int i = tid; i < 4 * max_tr; i += T) sring[i] = make_float4(0.f, 0.f, 0.f, 0.f);      // row 0 reads S = 0 (its S row is never defined)
        for (int i = tid; i < N1; i += T) { n1[i] = J.ni1[i]; rowslot[i] = J.wv_rowslot[i]; adj1[i] = J.wv_adj1[i]; }
        for (int i = tid; i < N2; i += T) ci[i] = J.wv_ci[i];
        for (int i = tid; i < J.np1; i += T) { p1i[i] = J.p1_idx[i]; p1v[i] = J.p1_val[i]; }
        for (int i = tid; i < J.wv_nx; i += T) { sxi[i] = J.wv_xi[i]; sxv[i] = J.wv_xv[i]; }
    }
    __syncthreads();
    const int ringB = RING * ring_pitch * 16, pitchB = ring_pitch * 16;
    uint32_t ringa = (uint32_t)__cvta_generic_to_shared(wsm);
    asm volatile("mov.u32 %0, %0;" : "+r"(ringa));                     // opaque: kept in a register instead of being rematerialised (S2R) at every use
    const uint32_t sringa = ringa + (uint32_t)((unsigned char*)sring - wsm);
    const uint32_t n1a = ringa + (uint32_t)((unsigned char*)n1 - wsm), cia = ringa + (uint32_t)((unsigned char*)ci - wsm);
    const uint32_t p1ia = ringa + (uint32_t)((unsigned char*)p1i - wsm), p1va = ringa + (uint32_t)((unsigned char*)p1v - wsm);
    const uint32_t sxia = ringa + (uint32_t)((unsigned char*)sxi - wsm), sxva = ringa + (uint32_t)((unsigned char*)sxv - wsm);
    const uint32_t adj1a = ringa + (uint32_t)((unsigned char*)adj1 - wsm), rowslota = ringa + (uint32_t)((unsigned char*)rowslot - wsm);
    const int dmax = (N1 - 2) + (N2 - 2);
    const int xmax = N2 - 2;
    // any cell (yy, xx) on an anti-diagonal dd = yy + xx < d.  The ring read is unconditional (its address is always inside the
    // ring); cells older than the ring are re-read from the {M,X,Y,W} array in L2 (rare: lanes branch only then)
    // A ring cell is {X, W, Y, W}: the up candidate reads (Y, W) and the left candidate (X, W) as one 8-byte word each, the
    // diagonal candidate W alone; M is never read back.  (ncu on the first version, profiles/r2_wave_v3.txt: 62 % of the
    // shared-memory wavefronts were bank-conflict replays and the MIO queue was the top stall -- 16-byte reads of whole cells
    // at scattered rows, and the S reads 4-way conflicted.)
    auto cell_yw = [&](int yy, int xx, int d, bool usable) -> float2 {  // (Y, W) of cell (yy, xx)
        const int dd = yy + xx;
        float2 v = lds64(ringa + (uint32_t)((dd & (RING - 1)) * pitchB + (yy + 1) * 16 + 8));
        if (usable && d - dd >= RING) { const float4 c = wave_far_cell(C, dd, yy, pitch, N2); v = make_float2(c.z, c.w); }
        return v;
    };
    auto cell_xw = [&](int yy, int xx, int d, bool usable) -> float2 {  // (X, W)
        const int dd = yy + xx;
        float2 v = lds64(ringa + (uint32_t)((dd & (RING - 1)) * pitchB + (yy + 1) * 16));
        if (usable && d - dd >= RING) { const float4 c = wave_far_cell(C, dd, yy, pitch, N2); v = make_float2(c.y, c.w); }
        return v;
    };
    auto cell_w = [&](int yy, int xx, int d, bool usable) -> float {    // W
        const int dd = yy + xx;
        float v = lds32(ringa + (uint32_t)((dd & (RING - 1)) * pitchB + (yy + 1) * 16 + 4));
        if (usable && d - dd >= RING) v = wave_far_cell(C, dd, yy, pitch, N2).w;
        return v;
    };
    if (tid == 0) {                                                     // cell (0, 0); the sweep starts at anti-diagonal 1
        ring[1] = make_float4(NINF, start_init, NINF, start_init);
        C[0] = make_float4(NINF, NINF, NINF, start_init);
    }
    const float PINF = CUDART_INF_F;

    if (tid < TR) {
        // ================================================ row threads ========================================================
        // Branch-free step: every load has an always-valid (clamped) address and is issued up front, an absent edge is the value
        // +inf (its candidates come out as -inf), only the stores are predicated.
        const WaveRow wr = J.wv_rows[tid];
        const int y = wr.y;                                             // -1: padding slot
        const bool yvalid = y >= 0;
        const int yc = max(y, 0);
        const bool spec = tid >= TRC;                                   // warp-uniform: rows with extra predecessors
        const float ey = wr.ey_adj;
        const float gix = y == 0 ? start_gap : gap_init;                // row 0: GraphAlign.h:224-234 opens with start_gap
        float m_pre = NINF, ys_pre = NINF;
        // extra predecessors 0 and 1 in registers; an unused one is (row 0, +inf)
        const int yp0 = wr.nextra > 0 ? wr.yp0 : 0, yp1 = wr.nextra > 1 ? wr.yp1 : 0;
        const float ey0 = wr.nextra > 0 ? wr.ey0 : PINF, ey1 = wr.nextra > 1 ? wr.ey1 : PINF;
        const float* __restrict__ Srow = J.S + (size_t)yc * J.ldS;
        const int ldS = J.ldS;
        // S(y, x) of this thread's row sits at word (x & 15) * max_tr + tid of the S ring: the 32 lanes of a warp read 32
        // consecutive words whatever their columns are (one wavefront), at the price of four 4-byte cp.async per group
        const uint32_t smine = sringa + (uint32_t)tid * 4u;
        const int sstep = max_tr * 4;                                   // bytes between consecutive columns
        auto fetch_group = [&](int q) {                                 // columns 4q .. 4q+3 of the row, asynchronously
            const uint32_t dst = smine + (uint32_t)(((4 * q) & 15) * sstep);
            const float* src = Srow + 4 * q;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(dst), "l"(src) : "memory");
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(dst + (uint32_t)sstep), "l"(src + 1) : "memory");
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(dst + (uint32_t)(2 * sstep)), "l"(src + 2) : "memory");
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(dst + (uint32_t)(3 * sstep)), "l"(src + 3) : "memory");
        };
        if (y >= 1 && y <= 12) {                                        // the groups no event will fetch for the first rows (first event: d = 4)
            fetch_group(0);
            if (4 < ldS) fetch_group(1);
            if (8 < ldS) fetch_group(2);
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        WAVE_SYNC();
        const uint32_t a_up = ringa + (uint32_t)yc * 16u;               // + ring row: the cell of row y-1
        const uint32_t a_me = a_up + 16u;                               // ... of row y
        size_t cidx = (size_t)pitch + (size_t)yc;                       // cell index on anti-diagonal 1 (N2 >= 3: no row offset yet)
        int rb0 = pitchB, rb1 = 0;                                      // ring rows of anti-diagonals d, d-1 (byte offsets)
        float wdiag = NINF;                                             // W(y-1, x-1): the W this thread read as its up neighbour one step ago
        float2 left = make_float2(NINF, y == 0 ? start_init : NINF);    // (X, W)(y, x-1): this thread's last cell
        float w0p = NINF, w1p = NINF;                                   // W(yp_k, x): last step's look-ahead reads (the diagonal candidates' W now)
        for (int d = 1; d <= dmax; ++d) {
            const int x = d - y;
            if ((d & 3) == 0) {                                         // S prefetch event (warp-uniform): the group 9 .. 12 columns ahead
                const int q = (x >> 2) + 3;
                if (y >= 1 && q >= 0 && 4 * q < ldS) fetch_group(q);
                asm volatile("cp.async.commit_group;\n" ::: "memory");
                asm volatile("cp.async.wait_group 1;\n" ::: "memory");  // the group of the previous event has landed
            }
            const int xcl = min(max(x, 0), xmax);
            const float2 cc = lds64(cia + (uint32_t)xcl * 8u);          // ex of the column's adjacent edge, flags
            const float s = lds32(smine + (uint32_t)((xcl & 15) * sstep));
            const float2 cu = lds64(a_up + (uint32_t)rb1 + 8u);         // (Y, W)(y-1, x)
            const float ex = cc.x;
            const int cflags = __float_as_int(cc.y);
            float2 cl = left;                                           // (X, W)(y, x-1): this thread, last anti-diagonal ...
            if (cflags & 2) cl = lds64(a_me + (uint32_t)rb1);           // ... unless that column's threads made it
            const float giy = x == 0 ? start_gap : gap_init;            // column 0: GraphAlign.h:212-222
            float ys = __fsub_rn(fmaxf(__fadd_rn(cu.x, gap_extend), __fadd_rn(cu.y, giy)), ey);
            float m = __fsub_rn(__fsub_rn(__fadd_rn(wdiag, s), ey), ex);
            const float xs = __fsub_rn(fmaxf(__fadd_rn(cl.x, gap_extend), __fadd_rn(cl.y, gix)), ex);
            wdiag = cu.y;
            if (x == 0) m = NINF;                                       // S(y, 0) is not defined
            if (spec) { m = fmaxf(m, m_pre); ys = fmaxf(ys, ys_pre); }
            const float w = fmaxf(m, fmaxf(xs, ys));
            const bool active = yvalid && (unsigned)x <= (unsigned)xmax;
            if (active) left = make_float2(xs, w);
            if (active && !(cflags & 1)) {                              // a non-chain column's cell belongs to its column threads
                sts128(a_me + (uint32_t)rb0, make_float4(xs, w, ys, w));
                C[cidx] = make_float4(m, xs, ys, w);
            }
            if (spec) {                                                 // look-ahead: the extra predecessors' candidates of cell (y, x+1)
                const int xn = x + 1;
                const int xnc = min(max(xn, 0), xmax);
                const bool usable = yvalid && (unsigned)xn <= (unsigned)xmax;
                const float2 cn = lds64(cia + (uint32_t)xnc * 8u);
                const float sn = lds32(smine + (uint32_t)((xnc & 15) * sstep));
                const float gin = xn == 0 ? start_gap : gap_init;
                const bool use0 = usable && wr.nextra > 0, use1 = usable && wr.nextra > 1;      // (an unused slot must not fetch from L2)
                // (Y, W)(yp_k, x+1) is read now; W(yp_k, x), the diagonal candidate's, is what this thread read one step ago
                const float2 u0 = cell_yw(yp0, xnc, d, use0), u1 = cell_yw(yp1, xnc, d, use1);
                float yq = fmaxf(__fsub_rn(fmaxf(__fadd_rn(u0.x, gap_extend), __fadd_rn(u0.y, gin)), ey0),
                                 __fsub_rn(fmaxf(__fadd_rn(u1.x, gap_extend), __fadd_rn(u1.y, gin)), ey1));
                float mq = fmaxf(__fsub_rn(__fsub_rn(__fadd_rn(w0p, sn), ey0), cn.x), __fsub_rn(__fsub_rn(__fadd_rn(w1p, sn), ey1), cn.x));
                w0p = u0.y; w1p = u1.y;
                if (wr.nextra > 2)                                      // rare: the rest of the list
                    for (int k = 2; k < wr.nextra; ++k) {
                        const int yp = lds32i(sxia + (uint32_t)(wr.rest + k) * 4u);
                        const float eyk = lds32(sxva + (uint32_t)(wr.rest + k) * 4u);
                        const float2 u = cell_yw(yp, xnc, d, usable);
                        const float vw = cell_w(yp, xnc - 1, d, usable && xn > 0);
                        yq = fmaxf(yq, __fsub_rn(fmaxf(__fadd_rn(u.x, gap_extend), __fadd_rn(u.y, gin)), eyk));
                        mq = fmaxf(mq, __fsub_rn(__fsub_rn(__fadd_rn(vw, sn), eyk), cn.x));
                    }
                if (xn <= 0) mq = NINF;
                m_pre = mq; ys_pre = yq;
            }
            cidx += (size_t)pitch - (d >= N2 - 1 ? 1 : 0);
            rb1 = rb0; rb0 += pitchB; if (rb0 == ringB) rb0 = 0;
            WAVE_SYNC();
        }
    } else {
        // ================================================ column threads =====================================================
        // A group of L lanes sweeps the non-chain columns r, r + TCG, r + 2 TCG, ... one after the other: a column is busy for
        // N1 - 1 consecutive anti-diagonals, and the host chose TCG so that column r + TCG starts after column r has ended (graph 2
        // may be far longer than graph 1: at most ~N1 columns are in flight at a time, whatever N2 is)
        const int cq = tid - TR;
        const int TCG = J.wv_tcols;
        int cr = cq / L;                                                // the column's rank among the non-chain columns
        bool is_col = cq < TCG * L && cr < NSC;
        WaveCol wc = {0, 0, 0, 0.f, 0, 0.f, 0, 0.f};
        if (is_col) wc = J.wv_cols[cr];
        int xc = is_col ? wc.x : 1, c2 = wc.cnt;
        const int lq = cq & (L - 1);
        const int kyq = L == 4 ? (lq >> 1) : 0, kxq = L >= 2 ? (lq & 1) : 0;
        constexpr int SY = L == 4 ? 2 : 1, SX = L >= 2 ? 2 : 1;
        const unsigned gmask = L == 1 ? 0u : (((1u << L) - 1u) << ((tid & 31) & ~(L - 1)));
        float exl = is_col ? wc.ex_adj_col : PINF;                      // value of the edge x-1 -> x, +inf if there is none
        const int* __restrict__ P2I = J.p2_idx; const float* __restrict__ P2V = J.p2_val;
        // this lane's first column predecessor in registers; an absent one is (column 0, +inf)
        int xpq = kxq < c2 ? (kxq == 0 ? wc.xp0 : wc.xp1) : 0;
        float exq = kxq < c2 ? (kxq == 0 ? wc.ex0 : wc.ex1) : PINF;
        float m_pre = NINF, xs_pre = NINF, ys_pre = NINF, ey_up = PINF;
        WAVE_SYNC();
        size_t cidx0 = (size_t)pitch;                                   // index of row 0's cell on anti-diagonal d (as for the row threads)
        int rb0 = pitchB, rb1 = 0;
        uint32_t sbase = sringa + (uint32_t)((xc & 15) * max_tr * 4);   // S(., xc): word (xc & 15) * max_tr + row slot
        for (int d = 1; d <= dmax; ++d) {
            const int yy = d - xc;
            {   // fold the adjacent candidates, write the cell (lane 0 of the column's lanes)
                const int yyc = min(max(yy, 0), N1 - 2);
                const float2 cu = lds64(ringa + (uint32_t)(rb1 + yyc * 16 + 8));       // (Y, W)(yy-1, xc)
                const float2 cl = lds64(ringa + (uint32_t)(rb1 + yyc * 16 + 16));      // (X, W)(yy, xc-1)
                const float gi = yy == 0 ? start_gap : gap_init;
                const float ys = fmaxf(ys_pre, __fsub_rn(fmaxf(__fadd_rn(cu.x, gap_extend), __fadd_rn(cu.y, gap_init)), ey_up));
                const float xs = fmaxf(xs_pre, __fsub_rn(fmaxf(__fadd_rn(cl.x, gap_extend), __fadd_rn(cl.y, gi)), exl));
                const float w = fmaxf(m_pre, fmaxf(xs, ys));
                if (is_col && lq == 0 && (unsigned)yy <= (unsigned)(N1 - 2)) {
                    sts128(ringa + (uint32_t)(rb0 + yyc * 16 + 16), make_float4(xs, w, ys, w));
                    C[cidx0 + (size_t)yyc] = make_float4(m_pre, xs, ys, w);
                }
            }
            if (is_col && yy >= N1 - 2) {                               // the column's last cell is made: on to column r + TCG
                cr += TCG;
                is_col = cr < NSC;
                if (is_col) {
                    wc = J.wv_cols[cr];
                    xc = wc.x; c2 = wc.cnt; exl = wc.ex_adj_col;
                    xpq = kxq < c2 ? (kxq == 0 ? wc.xp0 : wc.xp1) : 0;
                    exq = kxq < c2 ? (kxq == 0 ? wc.ex0 : wc.ex1) : PINF;
                    sbase = sringa + (uint32_t)((xc & 15) * max_tr * 4);
                }
                m_pre = NINF; xs_pre = NINF; ys_pre = NINF; ey_up = PINF;
            }
            {   // look-ahead: every candidate of (yn, xc) that is at least two anti-diagonals back
                const int yn = d - xc + 1;                              // (the column may have just changed)
                const int ync = min(max(yn, 0), N1 - 2);
                const bool usable = is_col && (unsigned)yn <= (unsigned)(N1 - 2);
                const float4 nyv = lds128(n1a + (uint32_t)ync * 16u);          // NodeInfo: ptr, cnt, idx0, val0
                const int ny_ptr = __float_as_int(nyv.x), c1 = NI_CNT(__float_as_int(nyv.y));
                const float s = lds32(sbase + (uint32_t)lds16u(rowslota + (uint32_t)ync * 2u) * 4u);
                const float gi = yn == 0 ? start_gap : gap_init;
                float m = NINF, xs = NINF, ys = NINF;
                // one pair (row predecessor ky, column predecessor kx): its M candidate, and the Y candidate of ky / the X
                // candidate of kx (the same candidate may be folded by several lanes: max is idempotent); out-of-range ky or kx
                // is an absent edge (+inf)
                auto pair = [&](int ky, int kx, int xp, float exk) {
                    int yp = 0; float eyk = PINF;
                    if (ky < c1) {
                        yp = ky ? lds32i(p1ia + (uint32_t)(ny_ptr + ky) * 4u) : __float_as_int(nyv.z);
                        eyk = ky ? lds32(p1va + (uint32_t)(ny_ptr + ky) * 4u) : nyv.w;
                    }
                    const bool has_y = usable && ky < c1, has_x = usable && kx < c2;
                    const float cdw = cell_w(yp, xp, d, has_y && has_x);     // anti-diagonal d + 1 - dy - dx <= d - 1
                    const float2 cu = cell_yw(yp, xc, d, has_y && yp != yn - 1);
                    const float2 cl = cell_xw(ync, xp, d, has_x && xp != xc - 1);
                    const float mc = __fsub_rn(__fsub_rn(__fadd_rn(cdw, s), eyk), exk);
                    const float yv = __fsub_rn(fmaxf(__fadd_rn(cu.x, gap_extend), __fadd_rn(cu.y, gap_init)), eyk);
                    const float xv = __fsub_rn(fmaxf(__fadd_rn(cl.x, gap_extend), __fadd_rn(cl.y, gi)), exk);
                    m = fmaxf(m, mc);
                    ys = fmaxf(ys, yp == yn - 1 ? NINF : yv);           // the adjacent edges' candidates are folded when the cell is made
                    xs = fmaxf(xs, xp == xc - 1 ? NINF : xv);
                };
                pair(kyq, kxq, xpq, exq);
                if (c1 > kyq + SY || c2 > kxq + SX)                     // rare: more pairs than lanes
                    for (int ky = kyq; ky < max(c1, 1); ky += SY)
                        for (int kx = kxq; kx < max(c2, 1); kx += SX) {
                            if (ky == kyq && kx == kxq) continue;
                            const int xp = kx < c2 ? (kx == 0 ? wc.xp0 : (kx == 1 ? wc.xp1 : P2I[wc.ptr + kx])) : 0;
                            const float exk = kx < c2 ? (kx == 0 ? wc.ex0 : (kx == 1 ? wc.ex1 : P2V[wc.ptr + kx])) : PINF;
                            pair(ky, kx, xp, exk);
                        }
                if (yn == 0) m = NINF;                                  // first row: X only (start_gap, GraphAlign.h:224-234)
                if (L > 1) {
#pragma unroll
                    for (int o = 1; o < L; o <<= 1) {
                        m = fmaxf(m, __shfl_xor_sync(gmask, m, o));
                        xs = fmaxf(xs, __shfl_xor_sync(gmask, xs, o));
                        ys = fmaxf(ys, __shfl_xor_sync(gmask, ys, o));
                    }
                }
                m_pre = m; xs_pre = xs; ys_pre = ys;
                ey_up = lds32(adj1a + (uint32_t)ync * 4u);
            }
            cidx0 += (size_t)pitch - (d >= N2 - 1 ? 1 : 0);
            rb1 = rb0; rb0 += pitchB; if (rb0 == ringB) rb0 = 0;
            WAVE_SYNC();
        }
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
           (size_t)max_np1 * 8 + (size_t)max_nx * 8 + (size_t)max_n1 * 4 + (size_t)max_n1 * 2 + 64;
}

template <int RING, int L, int MAXT>
static int launch_wave_t(cudaStream_t st, int n_jobs, const GraphJobDev* d_jobs, const int* d_order, int threads, int max_n1, int max_n2,
                         int max_tr, int max_np1, int max_nx) {
    const size_t smem = wave_smem_bytes(RING, max_n1, max_n2, max_tr, max_np1, max_nx);
    const int ring_pitch = (max_n1 + 1 + 3) & ~3;
    PGM_CUDA(cudaFuncSetAttribute(graph_fill_wave_kernel<RING, L, MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    graph_fill_wave_kernel<RING, L, MAXT><<<n_jobs, threads, smem, st>>>(d_jobs, d_order, ring_pitch, max_tr, max_n1, max_n2, max_np1, max_nx);
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
#define GO(R_, L_) do { if (threads <= 512) return launch_wave_t<R_, L_, 512>(st, n_jobs, d_jobs, d_order, threads, max_n1, max_n2, max_tr, max_np1, max_nx); \
                        if (threads <= 768) return launch_wave_t<R_, L_, 768>(st, n_jobs, d_jobs, d_order, threads, max_n1, max_n2, max_tr, max_np1, max_nx); \
                        return launch_wave_t<R_, L_, 1024>(st, n_jobs, d_jobs, d_order, threads, max_n1, max_n2, max_tr, max_np1, max_nx); } while (0)
    if (ring == 16) { if (L == 4) GO(16, 4); if (L == 2) GO(16, 2); GO(16, 1); }
    if (L == 4) GO(8, 4); if (L == 2) GO(8, 2); GO(8, 1);
#undef GO
}
