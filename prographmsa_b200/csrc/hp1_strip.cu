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
__global__ void __launch_bounds__(128)
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
                float4* vr = KEEP ? J.Vrow + (size_t)kr * J.ldVr + x0 : nullptr;
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
                float4* vr = J.Vrow + (size_t)kr * J.ldVr + x0;
#pragma unroll
                for (int k = 0; k < C; ++k)
                    if (x0 + k <= N2 - 2) vr[k] = make_float4(NINF, (k == 0 && lane == 0) ? NINF : Wup[k], NINF, Wup[k]);
            }
            if (lane == 0) J.Vcol[(size_t)kc0 * J.ldVc + y] = make_float4(m0, xs0, ys0, w0);   // column 0 is always kept
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
