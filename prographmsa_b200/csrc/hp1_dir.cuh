// hp1_dir.cuh -- the packed traceback of the graph DP (K5 writes it, K6 chases it).
//
// The reference's traceback (GraphAlign.h:283-521) re-derives the path from the five dense float matrices: at every step
// it takes the candidate predecessor whose recomputed score is CLOSEST to the current cell's value, first candidate
// winning ties.  The fill computed that value as the maximum of the very same float expressions, so for every interior
// cell "closest" is "the first candidate that attains the maximum" (|delta| = 0), which the fill can record while it
// computes the maximum (checked against the reference rule on the CPU: tests/first_max_traceback.py).  One direction
// word per cell therefore replaces the 16 bytes of {M,X,Y,W}:
//
//   bits 0..1   wstate: the state a path LANDING on W(y,x) continues in -- 0 = M, 2 = Y, 1 = X by the reference's
//               equality chain W==M, W==Y, W==X (GraphAlign.h:399-411); 3 = none matches ("backtracking failed")
//   bit  2      yopen: in state Y the best candidate opens the gap (W(yp,x) + gap_init) instead of extending (Y(yp,x) + gap_extend)
//   bit  3      xopen: same for state X
//   B bits each ky_m, kx_m : state M came from predecessor pair (ky_m of row y, kx_m of column x), first maximum in the
//                            reference's nested order (row predecessors outer, column predecessors inner)
//               ky_y       : state Y came from row predecessor ky_y (candidates in order ext(0), open(0), ext(1), ...)
//               kx_x       : state X came from column predecessor kx_x
// B = 3 (uint16_t) when no node of either graph has more than 8 predecessors, else B = 7 (uint32_t, up to 128).
//
// The boundary row and column (y == 0 or x == 0) are the one place where the reference's rule is NOT a first maximum: the
// fill uses start_gap there while the traceback recomputes the candidates with gap_init (GraphAlign.h:212-234 against
// :378-398).  Row 0 and column 0 are therefore always kept as values (see GraphJobDev::Vrow / Vcol) and the traceback
// applies the reference's closest-|delta| search to them verbatim; their direction words carry only wstate.
#pragma once
#include "internal.h"
#include <math_constants.h>

template <int B> struct DirWord;
template <> struct DirWord<3> { typedef uint16_t T; };
template <> struct DirWord<7> { typedef uint32_t T; };

#define DIR_YOPEN 4u
#define DIR_XOPEN 8u
template <int B> __device__ __forceinline__ constexpr uint32_t dir_mask() { return (1u << B) - 1u; }
template <int B> __device__ __forceinline__ uint32_t dir_kym(uint32_t w) { return (w >> 4) & dir_mask<B>(); }
template <int B> __device__ __forceinline__ uint32_t dir_kxm(uint32_t w) { return (w >> (4 + B)) & dir_mask<B>(); }
template <int B> __device__ __forceinline__ uint32_t dir_kyy(uint32_t w) { return (w >> (4 + 2 * B)) & dir_mask<B>(); }
template <int B> __device__ __forceinline__ uint32_t dir_kxx(uint32_t w) { return (w >> (4 + 3 * B)) & dir_mask<B>(); }
template <int B> __device__ __forceinline__ uint32_t dir_pack(uint32_t wst, bool yopen, bool xopen, uint32_t kym, uint32_t kxm, uint32_t kyy, uint32_t kxx) {
    return wst | (yopen ? DIR_YOPEN : 0u) | (xopen ? DIR_XOPEN : 0u) | (kym << 4) | (kxm << (4 + B)) | (kyy << (4 + 2 * B)) | (kxx << (4 + 3 * B));
}
// the state a path landing on W continues in: M, then Y, then X (GraphAlign.h:399-411)
__device__ __forceinline__ uint32_t dir_wstate(float m, float xs, float ys, float w) {
    return (w == m) ? 0u : ((w == ys) ? 2u : ((w == xs) ? 1u : 3u));
}

// A cell that is no longer in a fill kernel's shared-memory ring: the host keeps every row / column such a read can hit
// (hp1_host.cu, keep sets).  A missing entry would be a host bug; it reads as NaN so that it surfaces as a failed job.
__device__ __forceinline__ float4 dir_far_cell(const GraphJobDev& J, int keep_row, int keep_col, int yy, int xx) {
    if (keep_row >= 0) return __ldcg(J.Vrow + (size_t)keep_row * J.ldVr + xx);
    if (keep_col >= 0) return __ldcg(J.Vcol + (size_t)keep_col * J.ldVc + yy);
    const float q = CUDART_NAN_F;
    return make_float4(q, q, q, q);
}

// End-node match (GraphAlign.h:262-280) and the traceback's first pick (:283-311), run by one thread after the sweep.
// The reference's pick is the candidate closest to Wend in the order (per predecessor pair, rows outer) M + end_match,
// Y + end_gap, X + end_gap, then end_skip for the pair (0,0); Wend is the maximum of those expressions, so the pick is
// the first maximum.  Every predecessor of graph 1's end node is a kept row.
__device__ __forceinline__ void dir_end_match(const GraphJobDev& J) {
    const int N1 = J.N1, N2 = J.N2;
    const float end_match = J.sc[3], end_gap = J.sc[4], end_skip = J.sc[5];
    float Wend = -CUDART_INF_F;
    int pick = 3;                                                // state | ky << 8 | kx << 16; default: nothing finite
    const int p1 = J.p1_ptr[N1 - 1], n1 = J.p1_ptr[N1] - p1, p2 = J.p2_ptr[N2 - 1], n2 = J.p2_ptr[N2] - p2;
    for (int ky = 0; ky < n1; ++ky) {
        const int yp = J.p1_idx[p1 + ky];
        const float ey = J.p1_val[p1 + ky];
        const int kr = NI_KEEP(J.ni1[yp].cnt);
        for (int kx = 0; kx < n2; ++kx) {
            const int xp = J.p2_idx[p2 + kx];
            const float ex = J.p2_val[p2 + kx];
            const float4 c = dir_far_cell(J, kr, -1, yp, xp);
            float v = __fsub_rn(__fsub_rn(__fadd_rn(c.x, end_match), ey), ex);
            if (v > Wend) { Wend = v; pick = 0 | (ky << 8) | (kx << 16); }
            v = __fsub_rn(__fsub_rn(__fadd_rn(c.z, end_gap), ey), ex);
            if (v > Wend) { Wend = v; pick = 2 | (ky << 8) | (kx << 16); }
            v = __fsub_rn(__fsub_rn(__fadd_rn(c.y, end_gap), ey), ex);
            if (v > Wend) { Wend = v; pick = 1 | (ky << 8) | (kx << 16); }
            if (xp == 0 && yp == 0) {
                v = __fsub_rn(__fsub_rn(end_skip, ey), ex);
                if (v > Wend) { Wend = v; pick = 3 | (ky << 8) | (kx << 16); }
            }
        }
    }
    ((float*)J.result)[3] = Wend;
    J.result[5] = pick;
}
