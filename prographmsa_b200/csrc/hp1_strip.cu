// hp1_strip.cu -- K5, register-resident flavour: the strip kernel.
//
// Same recurrences, evaluation order and reference lines as hp1_graphdp.cu (GraphAlign.h:205-280), for graphs whose
// nodes are all chain nodes (one predecessor, the node before: every leaf x leaf job of a guide tree and the profile
// graphs that carry no indel yet).  The other two fill kernels spend most of their instructions on moving DP values
// through a shared-memory ring and on a CTA barrier per step; they are issue- and latency-bound long before HBM matters
// (A/B in profiles/).  Here nothing is exchanged through memory:
//
//   * one WARP per job, no barrier, no shared memory;
//   * lane j owns strip j = C consecutive columns of graph 2; the Y and W values of the previous row of those columns,
//     the columns' edge costs and S indices stay in registers;
//   * the rows stream through the warp with a skew of one row per lane: at step t lane j computes row y = t - j, so the
//     left neighbour (y, x0 - 1) is what lane j-1 produced one step earlier and the diagonal neighbour (y-1, x0 - 1)
//     what it produced two steps earlier -- two shuffles per step, C cells per lane and step;
//   * S(y, x) is read as S[srow(y)][scol(x)]: for class-coded (leaf) graphs srow / scol are the residue classes, so a
//     leaf x leaf job reads a 21 x 21 table instead of an N1 x N2 matrix and its emission launch is 400 times smaller;
//   * what goes to HBM is 4 direction bits per cell -- mlose, xbeats (which of M, Y, X a path landing on W continues in:
//     M unless mlose, then Y unless xbeats: the reference's W==M, W==Y, W==X chain), yopen, xopen -- taken as the SIGN BITS of
//     the differences of the candidates (a - b < 0 exactly when a < b: gradual underflow) and shifted into the strip's word
//     with one funnel shift each; plus the kept boundary row / column and end-node rows of the value store.
// Algorithmic HBM bytes per cell: 0.5 (direction bits) + 4 for S unless both graphs are class-coded (then ~0).
#include "hp1_dir.cuh"
#include <type_traits>

template <int C>
__global__ void __launch_bounds__(128)      // (register caps for 3 or 4 CTAs per SM were measured slower: spills)
graph_fill_strip_kernel(const GraphJobDev* __restrict__ jobs, const int* __restrict__ order, int n_jobs) {
    constexpr int W32 = C <= 8 ? 1 : (C <= 16 ? 2 : 4);
    const int lane = (int)threadIdx.x & 31;
    const int jq = (int)blockIdx.x * ((int)blockDim.x >> 5) + ((int)threadIdx.x >> 5);
    if (jq >= n_jobs) return;
    const GraphJobDev& J = jobs[order[jq]];
    const int N1 = J.N1, N2 = J.N2, ns = J.nstrips, ldS = J.ldS;
    const float gap_init = J.sc[0], gap_extend = J.sc[1], start_gap = J.sc[6], start_init = J.sc[7];
    const float NINF = -CUDART_INF_F;
    const float* __restrict__ S = J.S;
    const NodeInfo* __restrict__ ni1 = J.ni1;
    const bool coded1 = J.coded1 != 0;
    const int x0 = lane * C;
    const bool lane_on = lane < ns;
    float ex[C]; int sc[C];
#pragma unroll
    for (int k = 0; k < C; ++k) {
        const int x = x0 + k;
        const bool valid = lane_on && x <= N2 - 2;
        ex[k] = (valid && x >= 1) ? __ldg(&J.ni2[x].val0) : 0.f;
        sc[k] = valid ? (int)__ldg(&J.scol[x]) : 0;
    }
    float Yup[C], Wup[C];
#pragma unroll
    for (int k = 0; k < C; ++k) { Yup[k] = NINF; Wup[k] = NINF; }
    float outX = NINF, outW = NINF, diagW = NINF;
    const float gi0 = (lane == 0) ? start_gap : gap_init;          // column 0 opens with start_gap (GraphAlign.h:212-222)
    const int kc0 = NI_KEEP(__ldg(&J.ni2[0].cnt));
    uint32_t* __restrict__ D = (uint32_t*)J.D;
    float4* const Vrow = J.Vrow; float4* const Vcol0 = J.Vcol + (size_t)kc0 * J.ldVc;   // job fields in registers: the stores below could alias the job record
    const int ldVr = J.ldVr;
    auto row_ok = [&](int r) { return r >= 0 && r <= N1 - 2; };
    auto load_info = [&](int r) {
        const int4 v = __ldg((const int4*)(ni1 + r));
        NodeInfo o; o.ptr = v.x; o.cnt = v.y; o.idx0 = v.z; o.val0 = __int_as_float(v.w);
        return o;
    };
    const NodeInfo zero = NodeInfo{0, 0, 0, 0.f};
    int y = -lane;                                                  // the row of step t is t - lane
    // Software pipeline: while row y is computed, the S values of row y + 1 are loaded into the OTHER register set (the
    // step loop is unrolled by two so that the sets swap by name: a register move would wait for the loads inside the step
    // that issued them) and the node info of row y + 2 is fetched (mostly L1 hits: eight rows per line).  Nothing consumes a
    // load in the step that issues it.
    NodeInfo cur = zero, n1 = zero;
    if (lane_on) { if (row_ok(y)) cur = load_info(y); if (row_ok(y + 1)) n1 = load_info(y + 1); }
    float svA[C], svB[C];
#pragma unroll
    for (int k = 0; k < C; ++k) { svA[k] = 0.f; svB[k] = 0.f; }     // rows <= 0 read no S (row 0 is the boundary row)
    const int steps = (N1 - 1) + (ns - 1);
    auto step = [&](const int t, const float (&sv)[C], float (&sn)[C]) {
        float gX = __shfl_up_sync(0xffffffffu, outX, 1), gW = __shfl_up_sync(0xffffffffu, outW, 1);   // (y, x0 - 1) from lane - 1
        if (lane == 0) { gX = NINF; gW = NINF; }
        NodeInfo n2 = zero;
        if (lane_on) {
            if (row_ok(y + 2)) n2 = load_info(y + 2);
            const int r = y + 1;
            if (r >= 1 && r <= N1 - 2) {
                const float* sp = S + (size_t)(coded1 ? NI_CLASS(n1.cnt) : r) * ldS;
#pragma unroll
                for (int k = 0; k < C; ++k) sn[k] = __ldg(sp + sc[k]);
            }
        }
        if (lane_on && row_ok(y)) {
            const float ey = cur.val0;
            const float gix = (y == 0) ? start_gap : gap_init;      // row 0 opens with start_gap (GraphAlign.h:224-234)
            const int kr = NI_KEEP(cur.cnt);
            float wdiag = diagW, xl = gX, wl = gW;
            const float s0 = (lane == 0 || y == 0) ? 0.f : sv[0];   // column 0 and row 0 have no M term: keep S's garbage boundary (and stale registers) out
            uint32_t rr[W32];
#pragma unroll
            for (int g = 0; g < W32; ++g) rr[g] = 0u;
            float m0 = NINF, xs0 = NINF, ys0 = NINF, w0 = NINF;     // slot 0 (column 0 of the job lives there)
            // One row of the strip.  KEEP: also write {M,X,Y,W} of the row to the value store -- a second copy of the loop so
            // that the common rows carry no predicated-off store code (kept rows: the end node's predecessors; row 0 below)
            auto sweep = [&](auto keep_tag) {
                constexpr bool KEEP = decltype(keep_tag)::value;
                float4* vr = KEEP ? Vrow + (size_t)kr * ldVr + x0 : nullptr;
#pragma unroll
                for (int k = 0; k < C; ++k) {
                    const float yE = __fsub_rn(__fadd_rn(Yup[k], gap_extend), ey), yO = __fsub_rn(__fadd_rn(Wup[k], k == 0 ? gi0 : gap_init), ey);
                    const float ys = fmaxf(yE, yO);
                    const float m = __fsub_rn(__fsub_rn(__fadd_rn(wdiag, (k == 0) ? s0 : sv[k]), ey), ex[k]);
                    const float xE = __fsub_rn(__fadd_rn(xl, gap_extend), ex[k]), xO = __fsub_rn(__fadd_rn(wl, gix), ex[k]);
                    const float xs = fmaxf(xE, xO);
                    const float tt = fmaxf(xs, ys);
                    float w = fmaxf(m, tt);
                    if (k == 0) { if (lane == 0 && y == 0) w = start_init; m0 = m; xs0 = xs; ys0 = ys; w0 = w; }
                    // direction nibble, highest bit first: xopen, yopen, xbeats, mlose
                    uint32_t r = rr[k >> 3];
                    r = __funnelshift_l(__float_as_uint(__fsub_rn(xE, xO)), r, 1);
                    r = __funnelshift_l(__float_as_uint(__fsub_rn(yE, yO)), r, 1);
                    r = __funnelshift_l(__float_as_uint(__fsub_rn(ys, xs)), r, 1);
                    r = __funnelshift_l(__float_as_uint(__fsub_rn(m, tt)), r, 1);
                    rr[k >> 3] = r;
                    if (KEEP) { if (x0 + k <= N2 - 2) vr[k] = make_float4(m, xs, ys, w); }
                    wdiag = Wup[k]; Yup[k] = ys; Wup[k] = w; xl = xs; wl = w;
                }
            };
            if (kr >= 0 && y > 0) sweep(std::true_type{}); else sweep(std::false_type{});
            if (y == 0) {                                           // row 0 is always kept: M = Y = -inf and X = W there (W(0,0) = start_init, X(0,0) = -inf)
                float4* vr = Vrow + (size_t)kr * ldVr + x0;
#pragma unroll
                for (int k = 0; k < C; ++k)
                    if (x0 + k <= N2 - 2) vr[k] = make_float4(NINF, (k == 0 && lane == 0) ? NINF : Wup[k], NINF, Wup[k]);
            }
            if (lane == 0) Vcol0[y] = make_float4(m0, xs0, ys0, w0);   // column 0 is always kept
            outX = xl; outW = wl;
            uint32_t* dp = D + ((size_t)t * ns + lane) * W32;
            if (W32 == 1) dp[0] = rr[0];
            else if (W32 == 2) *reinterpret_cast<uint2*>(dp) = make_uint2(rr[0], rr[1]);
            else *reinterpret_cast<uint4*>(dp) = make_uint4(rr[0], rr[1], rr[W32 > 2 ? 2 : 0], rr[W32 > 3 ? 3 : 0]);
        }
        diagW = gW;
        cur = n1; n1 = n2;
        ++y;
    };
    for (int t = 0; t < steps; t += 2) {
        step(t, svA, svB);
        if (t + 1 < steps) step(t + 1, svB, svA);
    }
    // end-node match (GraphAlign.h:262-280) and the traceback's first pick: the rows it reads were written by other lanes
    __syncwarp();
    __threadfence_block();
    if (lane == 0) dir_end_match(J);
}

template <int C>
static int launch_strip_t(cudaStream_t st, int n_jobs, const GraphJobDev* d_jobs, const int* d_order) {
    graph_fill_strip_kernel<C><<<(n_jobs + 3) / 4, 128, 0, st>>>(d_jobs, d_order, n_jobs);
    PGM_CUDA(cudaGetLastError());
    return 0;
}

int launch_graph_fill_strip(cudaStream_t st, int C, int n_jobs, const GraphJobDev* d_jobs, const int* d_order) {
    if (n_jobs == 0) return 0;
    switch (C) {
        case 4: return launch_strip_t<4>(st, n_jobs, d_jobs, d_order);
        case 8: return launch_strip_t<8>(st, n_jobs, d_jobs, d_order);
        case 11: return launch_strip_t<11>(st, n_jobs, d_jobs, d_order);
        case 13: return launch_strip_t<13>(st, n_jobs, d_jobs, d_order);
        case 16: return launch_strip_t<16>(st, n_jobs, d_jobs, d_order);
        case 20: return launch_strip_t<20>(st, n_jobs, d_jobs, d_order);
        case 24: return launch_strip_t<24>(st, n_jobs, d_jobs, d_order);
    }
    pgm_set_error("strip kernel: unsupported strip width %d", C);
    return PGM_ERR_ARG;
}

// ======================================================================================================================
// General strip kernel: the same register-resident sweep for graphs with non-chain nodes, and with several warps per job.
//
//   * strip j = thread j of the CTA (up to SG_MAX_WARPS warps = 256 strips).  A strip holds at most C columns; the host
//     starts a new strip at every non-chain column, so the one place a column with several (or a non-adjacent) predecessor
//     can sit is SLOT 0.  The hand-over from the last lane of a warp to the first lane of the next goes through shared
//     memory, and one CTA barrier per step keeps the warps in step and publishes the shared-memory rings;
//   * phase 1 (row predecessors): a chain row folds its one predecessor from the registers (Yup / Wup / the diagonal
//     hand-over).  Any other row folds each predecessor yp in turn: the registers when yp = y - 1, the strip's (Y,W)
//     HISTORY RING in shared memory when yp is recent, else row yp of the value store in HBM/L2 (the host keeps every row a
//     non-adjacent edge starts from) -- recording the first maximum's predecessor index;
//   * phase 2 (along the strip): the X recurrence through the left neighbour in registers.  Slot 0 of a strip whose first
//     column is not a chain node takes its X and M terms from loops over that column's predecessors instead (their
//     descriptors sit in registers): the left neighbour from the hand-over, any other column xp from the SOURCE RING in shared
//     memory that the strip owning xp fills with (X, W)(y, xp) as it goes; pairs in the reference's nested order
//     (GraphAlign.h:238-260);
//   * the 4 direction bits of every cell go to the strip's word as in the chain kernel; the predecessor-index fields of the
//     cells in non-chain rows / columns go to two small side arrays (RowIdx / ColIdx) that the traceback consults.
// A float2 array in shared memory, addressed with a 32-bit shared-window address and explicit ld.shared / st.shared: through
// generic pointers the compiler re-derived the window base (S2R SR_CgaCtaId + LEA) in front of most of the 15 access sites of
// the step loop (39 S2R in the SASS of one instantiation)
struct SharedF2 {
    uint32_t a;
    __device__ __forceinline__ float2 ld(size_t i) const {
        float2 v; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a + (uint32_t)i * 8u)); return v;
    }
    __device__ __forceinline__ void st(size_t i, float2 v) const {
        asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(a + (uint32_t)i * 8u), "f"(v.x), "f"(v.y) : "memory");
    }
    __device__ __forceinline__ SharedF2 at(size_t i) const { return SharedF2{a + (uint32_t)i * 8u}; }
};
#define SG_MAXP 4            // column predecessors of a special column held in registers (more: read from the lists)
template <int C, int WB>     // WB: most warps per job of the launch (1, 2, 4 or SG_MAX_WARPS); the smaller ones leave room for several jobs per SM
__global__ void __launch_bounds__(32 * WB, WB >= 8 ? 1 : (C <= 4 ? 16 : 12) / WB)
graph_fill_strip_gen_kernel(const GraphJobDev* __restrict__ jobs, const int* __restrict__ order) {
    constexpr int B = 7;                                            // the side arrays hold 32-bit index words
    typedef uint32_t DW;
    constexpr int W32 = C <= 8 ? 1 : (C <= 16 ? 2 : 4);
    constexpr uint32_t YF = DIR_YOPEN | (dir_mask<B>() << (4 + 2 * B));        // the Y fields of a direction word
    constexpr uint32_t MF = (dir_mask<B>() << 4) | (dir_mask<B>() << (4 + B)); // the M fields
    constexpr uint32_t XF = DIR_XOPEN | (dir_mask<B>() << (4 + 3 * B));        // the X fields
    extern __shared__ __align__(16) unsigned char sg_smem[];
    __shared__ float2 hand[WB][SG_HAND_ROWS];                       // (X, W)(row, last column) of a warp's last strip: the next warp's left neighbour
    __shared__ int prog[WB];                                        // steps a warp has completed (release / acquire, see the step loop)
    const GraphJobDev& J = jobs[order[blockIdx.x]];
    const int lane = (int)threadIdx.x & 31, warp = (int)threadIdx.x >> 5, L = (int)threadIdx.x;
    const int N1 = J.N1, ns = J.nstrips, ldS = J.ldS, RRm = J.ring_rows - 1, RR = J.ring_rows, HR = J.hist_rows, HRm = J.hist_rows - 1;
    const int NSP = (ns + 31) & ~31;
    const int dbg = J.dbg;
    uint32_t smem_a = (uint32_t)__cvta_generic_to_shared(sg_smem), hand_a = (uint32_t)__cvta_generic_to_shared(&hand[0][0]);
    uint32_t prog_a = (uint32_t)__cvta_generic_to_shared(&prog[0]);
    asm volatile("mov.u32 %0, %0;" : "+r"(smem_a));                 // opaque: kept in registers instead of being rematerialised at every use
    asm volatile("mov.u32 %0, %0;" : "+r"(hand_a));
    asm volatile("mov.u32 %0, %0;" : "+r"(prog_a));
    constexpr int HDm = SG_HAND_ROWS - 1;
    constexpr int SLACK = WB > 1 ? SG_SLACK : 0;                    // steps a warp may be ahead of the job's last warp
    const SharedF2 ring{smem_a};                                    // [nsrc][RR]: (X, W)(y, source column) of the last RR rows
    const SharedF2 hist{smem_a + (uint32_t)((size_t)J.nsrc * RR * 8)};   // [HR][C + 1][NSP]: (Y, W)(row, slot) per strip; slot C = (X, W) of the strip's last column
    const SharedF2 handv{hand_a};                                   // hand[w][row & HDm] at w * SG_HAND_ROWS + (row & HDm)
    const float gap_init = J.sc[0], gap_extend = J.sc[1], start_gap = J.sc[6], start_init = J.sc[7];
    const float NINF = -CUDART_INF_F;
    const float* __restrict__ S = J.S;
    const NodeInfo* __restrict__ ni1 = J.ni1;
    const NodeInfo* __restrict__ ni2 = J.ni2;
    const bool coded1 = J.coded1 != 0;
    const bool lane_on = L < ns;
    const int x0 = lane_on ? __ldg(&J.sstart[L]) : 0;
    const int len = lane_on ? __ldg(&J.sstart[L + 1]) - x0 : 0;     // 1 .. C columns
    const bool spec0 = lane_on && (__ldg(&J.specmask[L]) & 1u);     // the strip's first column is not a chain node
    const uint32_t mysrc = lane_on ? __ldg(&J.srcmask[L]) : 0u;
    const bool wspec0 = __any_sync(0xffffffffu, spec0);
    const uint32_t wsrc = __reduce_or_sync(0xffffffffu, mysrc);
    // job fields the steps use, in registers (the stores below could alias the job record as far as the compiler knows)
    float4* const Vrow = J.Vrow; float4* const Vcol = J.Vcol;
    const int ldVr = J.ldVr, ldVc = J.ldVc, ldRI = J.ldRI, ldCI = J.ldCI;
    const int16_t* __restrict__ colsrc = J.colsrc;
    const int* __restrict__ p1_idx = J.p1_idx; const float* __restrict__ p1_val = J.p1_val;
    const RowInfo* __restrict__ rinfo = J.rinfo;
    uint32_t* __restrict__ D = (uint32_t*)J.D;
    DW* __restrict__ RowIdx = (DW*)J.RowIdx;
    DW* __restrict__ ColIdx = (DW*)J.ColIdx;
    auto load_info = [&](const NodeInfo* p) {
        const int4 v = __ldg((const int4*)p);
        NodeInfo o; o.ptr = v.x; o.cnt = v.y; o.idx0 = v.z; o.val0 = __int_as_float(v.w);
        return o;
    };
    const NodeInfo zero = NodeInfo{0, 0, 0, 0.f};
    // descriptors of the special column's predecessors (registers)
    NodeInfo nx0 = zero; int sp_c2 = 0, sp_cidx = -1;
    int sp_xp[SG_MAXP], sp_src[SG_MAXP]; float sp_ex[SG_MAXP];
#pragma unroll
    for (int j = 0; j < SG_MAXP; ++j) { sp_xp[j] = 0; sp_src[j] = -1; sp_ex[j] = CUDART_INF_F; }
    if (lane_on && !spec0 && x0 > 0) { sp_c2 = 1; sp_xp[0] = x0 - 1; sp_src[0] = -1; sp_ex[0] = __ldg(&ni2[x0].val0); }   // a chain column: its left neighbour
    if (spec0) {
        nx0 = load_info(ni2 + x0); sp_c2 = NI_CNT(nx0.cnt); sp_cidx = __ldg(&J.ncidx2[x0]);
#pragma unroll
        for (int j = 0; j < SG_MAXP; ++j)
            if (j < sp_c2) {
                sp_xp[j] = j ? __ldg(J.p2_idx + nx0.ptr + j) : nx0.idx0;
                sp_ex[j] = j ? __ldg(J.p2_val + nx0.ptr + j) : nx0.val0;
                sp_src[j] = (sp_xp[j] == x0 - 1) ? -1 : (int)__ldg(&colsrc[sp_xp[j]]);
            }
    }
    float ex[C]; int sc[C];
    int srck[C];                                                    // source slots: ring index | (value-store column + 1) << 16
#pragma unroll
    for (int k = 0; k < C; ++k) {
        const int x = x0 + k;
        const bool valid = k < len;
        ex[k] = (valid && x >= 1) ? __ldg(&ni2[x].val0) : 0.f;
        sc[k] = valid ? (int)__ldg(&J.scol[x]) : 0;
        srck[k] = 0;
        if ((mysrc >> k) & 1u) srck[k] = (int)__ldg(&colsrc[x]) | ((x == 0 ? 0 : NI_KEEP(__ldg(&ni2[x].cnt)) + 1) << 16);
    }
    float Yup[C], Wup[C];
#pragma unroll
    for (int k = 0; k < C; ++k) { Yup[k] = NINF; Wup[k] = NINF; }
    float outX = NINF, outW = NINF, diagW = NINF;
    const float gi0 = (L == 0) ? start_gap : gap_init;             // column 0 opens with start_gap (GraphAlign.h:212-222)
    const int kc0 = NI_KEEP(__ldg(&ni2[0].cnt));
    auto row_ok = [&](int r) { return r >= 0 && r <= N1 - 2; };
    auto load_row = [&](int r) {
        const int4 a = __ldg((const int4*)(rinfo + r)), b = __ldg((const int4*)(rinfo + r) + 1);
        RowInfo o; o.cnt = a.x; o.idx0 = a.y; o.val0 = __int_as_float(a.z); o.idx1 = a.w; o.val1 = __int_as_float(b.x); o.kq01 = b.y; o.ncidx = b.z; o.ptr = b.w;
        return o;
    };
    const RowInfo rzero = RowInfo{0, 0, 0.f, 0, 0.f, 0, -1, 0};
    // predecessor ky of the current row: index, edge value, value-store row (-1: not kept); the first two sit in the row record
    auto row_pred = [&](const RowInfo& ri, int ky, int& yp, float& ey, int& kq) {
        if (ky == 0) { yp = ri.idx0; ey = ri.val0; kq = (ri.kq01 & 0xffff) - 1; }
        else if (ky == 1) { yp = ri.idx1; ey = ri.val1; kq = (int)((unsigned)ri.kq01 >> 16) - 1; }
        else { yp = __ldg(p1_idx + ri.ptr + ky); ey = __ldg(p1_val + ri.ptr + ky); kq = NI_KEEP(__ldg(&ni1[yp].cnt)); }
    };
    // the rings start at -inf: a dummy predecessor (infinite edge) may read a row nobody has written yet, and -inf stays -inf
    for (int i = L; i < J.nsrc * RR + HR * (C + 1) * NSP; i += (int)blockDim.x) ring.st(i, make_float2(NINF, NINF));
    for (int i = L; i < WB * SG_HAND_ROWS; i += (int)blockDim.x) handv.st(i, make_float2(NINF, NINF));   // a warp's first strip reads the hand-over before the previous warp's last strip has started: row -1 is -inf
    if (L < WB) prog[L] = 0;
    __syncthreads();
    const int nwj = (ns + 31) >> 5;                                 // warps of this job that own strips (the launch may have more)
    int seen_lo = 0, seen_hi = 0;                                   // last progress read of the previous / of the job's last warp
    int y = -L;                                                     // the row of step t is t - L
    RowInfo cur = rzero, n1 = rzero;
    if (lane_on) { if (row_ok(y)) cur = load_row(y); if (row_ok(y + 1)) n1 = load_row(y + 1); }
    float sv[C];
#pragma unroll
    for (int k = 0; k < C; ++k) sv[k] = 0.f;
    const int steps = (WB == 1 || warp < nwj) ? (N1 - 1) + (ns - 1) : 0;   // (a warp without strips has nothing to hand over)
    for (int t = 0; t < steps; ++t) {
        float sn[C];
#pragma unroll
        for (int k = 0; k < C; ++k) sn[k] = 0.f;
        // Warps of a job are coupled point to point, not by a CTA barrier: warp w starts step t once warp w - 1 has completed
        // step t - 1 (everything a strip reads from other strips -- hand-over, source rings, value store -- was produced by a
        // strip to its left at an earlier step, and progress is monotone in w), and no warp runs more than SLACK steps ahead of
        // the job's last warp (the rings are that much deeper than their readers look back).  A warp that is slow for a while
        // (a row with several predecessors passing through its 32 strips) then delays its neighbours only once the slack is used
        // up, instead of all warps of the job at every step.
        __syncwarp();                                               // the warp's own ring / value-store writes of the previous step
        if (WB > 1 && (dbg & 16)) __syncthreads();                  // (A/B: the CTA barrier instead)
        else if (WB > 1) {
            // every lane polls (one broadcast read per warp, a warp-uniform loop): a loop run by lane 0 alone left the warp split in
            // {lane 0} and {lanes 1..31} for the rest of the step (ncu: 18 of 32 threads per instruction, twice the instructions)
            const uint32_t mine = prog_a + 4u * (uint32_t)warp;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, %2, 0;\n\t@p st.volatile.shared.b32 [%0], %1;\n\t}" ::"r"(mine), "r"(t), "r"(lane) : "memory");
            if (warp > 0)
                while (seen_lo < t) asm volatile("ld.volatile.shared.b32 %0, [%1];" : "=r"(seen_lo) : "r"(mine - 4u) : "memory");
            if (warp < nwj - 1)
                while (seen_hi < t - SLACK) asm volatile("ld.volatile.shared.b32 %0, [%1];" : "=r"(seen_hi) : "r"(prog_a + 4u * (uint32_t)(nwj - 1)) : "memory");
            __syncwarp();
        }
        if (dbg & 2) __threadfence();
        float gX = __shfl_up_sync(0xffffffffu, outX, 1), gW = __shfl_up_sync(0xffffffffu, outW, 1);   // (y, x0 - 1) from strip L - 1
        if (lane == 0) {
            if (WB == 1 || warp == 0) { gX = NINF; gW = NINF; }
            else { const float2 h = handv.ld((warp - 1) * SG_HAND_ROWS + (y & HDm)); gX = h.x; gW = h.y; }   // row y of strip L - 1 (produced at step t - 1)
        }
        RowInfo n2 = rzero;
        if (lane_on) {                                              // two rows ahead: row record; one row ahead: S (consumed next step)
            if (row_ok(y + 2)) n2 = load_row(y + 2);
            const int r = y + 1;
            if (r >= 1 && r <= N1 - 2) {
                const float* sp = S + (size_t)(coded1 ? NI_CLASS(n1.cnt) : r) * ldS;
#pragma unroll
                for (int k = 0; k < C; ++k) sn[k] = __ldg(sp + sc[k]);
            }
        }
        const bool active = lane_on && row_ok(y);
        const bool rowchain = ((cur.cnt & 0x800000ff) == 1) && cur.idx0 == y - 1;
        // one path per warp and step: the fused sweep when no strip of the warp needs anything else, else the uniform general sweep
        const bool fast_step = !(dbg & 1) && !wspec0 && !__any_sync(0xffffffffu, active && !(rowchain || y == 0));
        if (active) {
            const float gix = (y == 0) ? start_gap : gap_init;      // row 0 opens with start_gap (GraphAlign.h:224-234)
            const int kr = NI_KEEP(cur.cnt);
            const int c1 = NI_CNT(cur.cnt);
            const float s0 = (L == 0 || y == 0) ? 0.f : sv[0];      // column 0 and row 0 have no M term
            float xl = gX, wl = gW;
            uint32_t rr[W32];
#pragma unroll
            for (int g = 0; g < W32; ++g) rr[g] = 0u;
            float m0 = NINF, xs0 = NINF, ys0 = NINF, w0 = NINF;     // slot 0
            // the source of a non-adjacent column edge publishes (X, W) in its ring, and {M,X,Y,W} in the value store for the rows the ring has dropped
            auto publish = [&](const int k, float mk, float xs, float ysk, float w) {
                if ((wsrc >> k) & 1u) {
                    if ((mysrc >> k) & 1u) {
                        ring.st((size_t)(srck[k] & 0xffff) * RR + (y & RRm), make_float2(xs, w));
                        const int kc = (srck[k] >> 16) - 1;
                        if (kc >= 0) Vcol[(size_t)kc * ldVc + y] = make_float4(mk, xs, ysk, w);
                    }
                }
            };
            if (fast_step) {
                // ---- fast path (every strip of the warp is at a chain row or the boundary row, no special column in the warp): the chain kernel's fused sweep ----
                const float ey = cur.val0;
                float wdiag = diagW;
                auto sweep = [&](auto keep_tag) {
                    constexpr bool KEEP = decltype(keep_tag)::value;
                    float4* vr = KEEP ? Vrow + (size_t)kr * ldVr + x0 : nullptr;
#pragma unroll
                    for (int k = 0; k < C; ++k) {
                        const float yE = __fsub_rn(__fadd_rn(Yup[k], gap_extend), ey), yO = __fsub_rn(__fadd_rn(Wup[k], k == 0 ? gi0 : gap_init), ey);
                        const float ys = fmaxf(yE, yO);
                        const float m = __fsub_rn(__fsub_rn(__fadd_rn(wdiag, (k == 0) ? s0 : sv[k]), ey), ex[k]);
                        const float xE = __fsub_rn(__fadd_rn(xl, gap_extend), ex[k]), xO = __fsub_rn(__fadd_rn(wl, gix), ex[k]);
                        const float xs = fmaxf(xE, xO);
                        const float tt = fmaxf(xs, ys);
                        float w = fmaxf(m, tt);
                        if (k == 0) { if (L == 0 && y == 0) w = start_init; m0 = m; xs0 = xs; ys0 = ys; w0 = w; }
                        uint32_t r = rr[k >> 3];                    // direction nibble, highest bit first: xopen, yopen, xbeats, mlose
                        r = __funnelshift_l(__float_as_uint(__fsub_rn(xE, xO)), r, 1);
                        r = __funnelshift_l(__float_as_uint(__fsub_rn(yE, yO)), r, 1);
                        r = __funnelshift_l(__float_as_uint(__fsub_rn(ys, xs)), r, 1);
                        r = __funnelshift_l(__float_as_uint(__fsub_rn(m, tt)), r, 1);
                        rr[k >> 3] = r;
                        if (KEEP) { if (k < len) vr[k] = make_float4(m, xs, ys, w); }
                        publish(k, m, xs, ys, w);
                        wdiag = Wup[k]; Yup[k] = ys; Wup[k] = w;
                        if (k < len) { xl = xs; wl = w; }            // the strip's last valid column feeds the next strip
                    }
                };
                if (kr >= 0) sweep(std::true_type{}); else sweep(std::false_type{});
            } else {
                // W(yp, x0 - 1) of an earlier row: the neighbour strip's history, else the value store
                auto w_left_of_strip = [&](int yp, int kq) -> float {
                    if (x0 == 0) return NINF;
                    if (y - yp <= HR - 2) {
                        if (WB > 1 && lane == 0) return handv.ld((warp - 1) * SG_HAND_ROWS + (yp & HDm)).y;   // the neighbour is another warp's strip: its hand-over ring
                        return hist.ld(((size_t)(yp & HRm) * (C + 1) + C) * NSP + (L - 1)).y;
                    }
                    return kq < 0 ? CUDART_NAN_F : __ldcg(Vrow + (size_t)kq * ldVr + (x0 - 1)).w;
                };
                // ---- phase 1: row predecessors -> ys[k], m[k] (through the chain column edge x-1 -> x), index fields ----
                // Uniform across the warp: every row folds TWO predecessors whose windows come from the history ring (a chain row's
                // second predecessor is a dummy with an infinite edge, which yields -inf candidates that never win).  Rows with more
                // than two predecessors, or with one older than the ring, redo the fold from the lists / the value store afterwards.
                float ys[C], m[C]; uint32_t dr[C];
                auto fold = [&](const uint32_t ky, const float ey, const bool first, const float2 (&win)[C], const float wprev) {
#pragma unroll
                    for (int k = 0; k < C; ++k) {
                        const float cE = __fsub_rn(__fadd_rn(win[k].x, gap_extend), ey), cO = __fsub_rn(__fadd_rn(win[k].y, k == 0 ? gi0 : gap_init), ey);
                        const float wd = (k == 0) ? wprev : win[k - 1].y;
                        const float v = __fsub_rn(__fsub_rn(__fadd_rn(wd, k == 0 ? s0 : sv[k]), ey), ex[k]);
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
                const int yp0 = cur.idx0, yp1 = cur.idx1;
                const float ey0 = c1 > 0 ? cur.val0 : CUDART_INF_F, ey1 = c1 > 1 ? cur.val1 : CUDART_INF_F;
                const bool near0 = c1 < 1 || y - yp0 <= HR - 2, near1 = c1 < 2 || y - yp1 <= HR - 2;
                const bool slowrow = c1 > 2 || !near0 || !near1;
                {
                    const int q0 = (c1 > 0 && near0) ? yp0 : y, q1 = (c1 > 1 && near1) ? yp1 : q0;
                    const int r0 = q0 & HRm, r1 = q1 & HRm;
                    const SharedF2 h0 = hist.at(((size_t)r0 * (C + 1)) * NSP + L);
                    const SharedF2 h1 = hist.at(((size_t)r1 * (C + 1)) * NSP + L);
                    float2 win0[C], win1[C];
#pragma unroll
                    for (int k = 0; k < C; ++k) { win0[k] = h0.ld((size_t)k * NSP); win1[k] = h1.ld((size_t)k * NSP); }
                    float wp0, wp1;                                 // W(row, x0 - 1): slot C of the strip to the left
                    if (x0 == 0) { wp0 = NINF; wp1 = NINF; }
                    else if (WB > 1 && lane == 0) { wp0 = handv.ld((warp - 1) * SG_HAND_ROWS + (q0 & HDm)).y; wp1 = handv.ld((warp - 1) * SG_HAND_ROWS + (q1 & HDm)).y; }
                    else { wp0 = h0.ld((size_t)C * NSP - 1).y; wp1 = h1.ld((size_t)C * NSP - 1).y; }
                    fold(0u, ey0, true, win0, wp0);
                    fold(1u, ey1, false, win1, wp1);
                }
                if (slowrow) {                                      // rare: the lists, the value store
#pragma unroll
                    for (int k = 0; k < C; ++k) { ys[k] = NINF; m[k] = NINF; dr[k] = 0u; }
                    for (int ky = 0; ky < c1; ++ky) {
                        int yp, kq; float ey;
                        row_pred(cur, ky, yp, ey, kq);
                        float2 win[C]; float wprev;
                        if (y - yp <= HR - 2) {                     // recent row: the strip's history ring
                            const SharedF2 hb = hist.at(((size_t)(yp & HRm) * (C + 1)) * NSP + L);
#pragma unroll
                            for (int k = 0; k < C; ++k) win[k] = hb.ld((size_t)k * NSP);
                            wprev = w_left_of_strip(yp, -1);
                        } else {                                    // a kept row of the value store
                            const float4* base = Vrow + (size_t)(kq < 0 ? 0 : kq) * ldVr + x0;
#pragma unroll
                            for (int k = 0; k < C; ++k) { const float4 c = __ldcg(base + k); win[k] = kq < 0 ? make_float2(CUDART_NAN_F, CUDART_NAN_F) : make_float2(c.z, c.w); }
                            wprev = w_left_of_strip(yp, kq);
                        }
                        if (ky == 0) fold(0u, ey, true, win, wprev); else fold((uint32_t)ky, ey, false, win, wprev);
                    }
                }
                // ---- phase 2: along the strip ----
                float wdiag = diagW;
                float xsv[C];
                const int rix = (rowchain || y == 0) ? -1 : cur.ncidx;
                // slot 0 of EVERY strip is a generic cell: X over the column's predecessors, M over the pairs (rows outer, columns
                // inner: the reference's nested order).  A chain column has the one predecessor x0 - 1 (descriptor set up above),
                // column 0 has none.  Two by two pairs are computed uniformly; more go through the loops.
                float xg = NINF, mg = NINF; uint32_t dg = 0u;
                {
                    // W(yp, xp) for the M pairs: the source ring, the neighbour strip's history, the value store
                    auto w_cell = [&](int yp, int kq, int xp, int src) -> float {
                        if (src < 0) return w_left_of_strip(yp, kq);
                        if (yp == y - 1 || y - yp < RR - 1 - ns - SLACK) return ring.ld((size_t)src * RR + (yp & RRm)).y;   // (the owner may be SLACK steps ahead)
                        return kq < 0 ? CUDART_NAN_F : __ldcg(Vrow + (size_t)kq * ldVr + xp).w;
                    };
                    auto x_cand = [&](uint32_t kx, int src, float exj) {
                        float Xv = xl, Wv = wl;
                        if (src >= 0) { const float2 q = ring.ld((size_t)src * RR + (y & RRm)); Xv = q.x; Wv = q.y; }
                        const float cE = __fsub_rn(__fadd_rn(Xv, gap_extend), exj), cO = __fsub_rn(__fadd_rn(Wv, gix), exj);
                        if (cE > xg) { xg = cE; dg = (dg & ~XF) | (kx << (4 + 3 * B)); }
                        if (cO > xg) { xg = cO; dg = (dg & ~XF) | (kx << (4 + 3 * B)) | DIR_XOPEN; }
                    };
                    if (sp_c2 <= SG_MAXP) {
                        if (c1 > 0 || y == 0) {
#pragma unroll
                            for (int j = 0; j < SG_MAXP; ++j) if (j < sp_c2) x_cand((uint32_t)j, sp_src[j], sp_ex[j]);
                        }
                        for (int ky = 0; ky < c1; ++ky) {
                            int yp, kq; float ey;
                            row_pred(cur, ky, yp, ey, kq);
#pragma unroll
                            for (int j = 0; j < SG_MAXP; ++j)
                                if (j < sp_c2) {
                                    const float v = __fsub_rn(__fsub_rn(__fadd_rn(w_cell(yp, kq, sp_xp[j], sp_src[j]), s0), ey), sp_ex[j]);
                                    if (v > mg) { mg = v; dg = (dg & ~MF) | ((uint32_t)ky << 4) | ((uint32_t)j << (4 + B)); }
                                }
                        }
                    } else {                                        // many column predecessors: walk the lists
                        if (c1 > 0 || y == 0)
                            for (int kx = 0; kx < sp_c2; ++kx) {
                                const int xp = kx ? __ldg(J.p2_idx + nx0.ptr + kx) : nx0.idx0;
                                x_cand((uint32_t)kx, xp == x0 - 1 ? -1 : (int)__ldg(&colsrc[xp]), kx ? __ldg(J.p2_val + nx0.ptr + kx) : nx0.val0);
                            }
                        for (int ky = 0; ky < c1; ++ky) {
                            int yp, kq; float ey;
                            row_pred(cur, ky, yp, ey, kq);
                            for (int kx = 0; kx < sp_c2; ++kx) {
                                const int xp = kx ? __ldg(J.p2_idx + nx0.ptr + kx) : nx0.idx0;
                                const float exj = kx ? __ldg(J.p2_val + nx0.ptr + kx) : nx0.val0;
                                const float v = __fsub_rn(__fsub_rn(__fadd_rn(w_cell(yp, kq, xp, xp == x0 - 1 ? -1 : (int)__ldg(&colsrc[xp])), s0), ey), exj);
                                if (v > mg) { mg = v; dg = (dg & ~MF) | ((uint32_t)ky << 4) | ((uint32_t)kx << (4 + B)); }
                            }
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < C; ++k) {
                    const int x = x0 + k;
                    const float xE = __fsub_rn(__fadd_rn(xl, gap_extend), ex[k]), xO = __fsub_rn(__fadd_rn(wl, gix), ex[k]);
                    float xs = fmaxf(xE, xO);
                    float mk = m[k];
                    uint32_t dk = dr[k] | ((xO > xE) ? DIR_XOPEN : 0u);
                    if (k == 0) { mk = mg; xs = xg; dk = (dr[0] & YF) | dg; }
                    const float tt = fmaxf(xs, ys[k]);
                    float w = fmaxf(mk, tt);
                    if (k == 0) { if (L == 0 && y == 0) w = start_init; m0 = mk; xs0 = xs; ys0 = ys[k]; w0 = w; }
                    uint32_t r = rr[k >> 3];                        // direction nibble, highest bit first: xopen, yopen, xbeats, mlose
                    r = (r << 2) | ((dk >> 2) & 3u);                // DIR_XOPEN (8) and DIR_YOPEN (4) are adjacent: xopen, yopen
                    r = __funnelshift_l(__float_as_uint(__fsub_rn(ys[k], xs)), r, 1);
                    r = __funnelshift_l(__float_as_uint(__fsub_rn(mk, tt)), r, 1);
                    rr[k >> 3] = r;
                    if (rix >= 0) { if (k < len) RowIdx[(size_t)rix * ldRI + x] = dk; }   // predecessor-index fields for the traceback
                    if (k == 0) { if (spec0) ColIdx[(size_t)sp_cidx * ldCI + y] = dk; }
                    publish(k, mk, xs, ys[k], w);
                    m[k] = mk; xsv[k] = xs;
                    wdiag = Wup[k]; Yup[k] = ys[k]; Wup[k] = w;
                    if (k < len) { xl = xs; wl = w; }                // the strip's last valid column feeds the next strip
                }
                if (kr >= 0) {                                      // kept row: {M,X,Y,W} to the value store
                    float4* vr = Vrow + (size_t)kr * ldVr + x0;
#pragma unroll
                    for (int k = 0; k < C; ++k)
                        if (k < len) vr[k] = make_float4(m[k], xsv[k], Yup[k], Wup[k]);
                }
            }
            {   // history ring: this row's (Y,W) per slot, and (X,W) of the strip's last column
                const SharedF2 hb = hist.at(((size_t)(y & HRm) * (C + 1)) * NSP + L);
#pragma unroll
                for (int k = 0; k < C; ++k) hb.st((size_t)k * NSP, make_float2(Yup[k], Wup[k]));
                hb.st((size_t)C * NSP, make_float2(xl, wl));
            }
            if (L == 0) Vcol[(size_t)kc0 * ldVc + y] = make_float4(m0, xs0, ys0, w0);   // column 0 is always kept
            outX = xl; outW = wl;
            if (WB > 1 && lane == 31) handv.st(warp * SG_HAND_ROWS + (y & HDm), make_float2(xl, wl));
            uint32_t* dp = D + ((size_t)t * ns + L) * W32;
            if (W32 == 1) dp[0] = rr[0];
            else if (W32 == 2) *reinterpret_cast<uint2*>(dp) = make_uint2(rr[0], rr[1]);
            else *reinterpret_cast<uint4*>(dp) = make_uint4(rr[0], rr[1], rr[W32 > 2 ? 2 : 0], rr[W32 > 3 ? 3 : 0]);
        }
        // kept rows go to the value store in global memory, and a strip of another warp may read them back (far predecessors): the
        // stores are fenced before this warp's progress word moves on (the shared-memory rings need no fence: one in-order pipe)
        if (WB > 1 && __any_sync(0xffffffffu, active && NI_KEEP(cur.cnt) >= 0)) __threadfence_block();
        diagW = gW;
        cur = n1; n1 = n2;
#pragma unroll
        for (int k = 0; k < C; ++k) sv[k] = sn[k];                  // (the loads were issued at the top of this step)
        ++y;
    }
    // end-node match (GraphAlign.h:262-280) and the traceback's first pick
    __syncthreads();
    if (L == 0) dir_end_match(J);
}

template <int C, int WB>
static int launch_strip_gen_t(cudaStream_t st, int warps, size_t smem, int n_jobs, const GraphJobDev* d_jobs, const int* d_order) {
    // (static + dynamic shared memory beyond 48 KB needs the opt-in; the static part -- hand-over rings, progress words -- is up to 8.1 KB)
    if (smem > 32 * 1024) PGM_CUDA(cudaFuncSetAttribute(graph_fill_strip_gen_kernel<C, WB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    graph_fill_strip_gen_kernel<C, WB><<<n_jobs, 32 * warps, smem, st>>>(d_jobs, d_order);
    PGM_CUDA(cudaGetLastError());
    return 0;
}

int launch_graph_fill_strip_gen(cudaStream_t st, int C, int dbits, int warps, size_t smem, int n_jobs, const GraphJobDev* d_jobs, const int* d_order) {
    (void)dbits;
    if (n_jobs == 0) return 0;
    if (warps < 1 || warps > SG_MAX_WARPS) { pgm_set_error("general strip kernel: %d warps per job", warps); return PGM_ERR_ARG; }
#define SG_CASE(C_) case C_: return warps == 1 ? launch_strip_gen_t<C_, 1>(st, warps, smem, n_jobs, d_jobs, d_order) \
                                  : warps == 2 ? launch_strip_gen_t<C_, 2>(st, warps, smem, n_jobs, d_jobs, d_order) \
                                  : warps <= 4 ? launch_strip_gen_t<C_, 4>(st, warps, smem, n_jobs, d_jobs, d_order) \
                                               : launch_strip_gen_t<C_, 8>(st, warps, smem, n_jobs, d_jobs, d_order);
    switch (C) { SG_CASE(2) SG_CASE(4) SG_CASE(8) SG_CASE(13) SG_CASE(16) }
#undef SG_CASE
    pgm_set_error("general strip kernel: unsupported strip width %d", C);
    return PGM_ERR_ARG;
}
