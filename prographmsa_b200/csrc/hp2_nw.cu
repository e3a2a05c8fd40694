// hp2_nw.cu -- K1: batched global (Gotoh) alignment + packed traceback, and the traceback walk.
//
// Replaces DistanceFactoryAlign<A>::alignPair (reference DistanceFactoryAlign.h:58-127):
//   W(0,0)=0; W(0,x)=X(0,x)=go+(x-1)ge; Y(0,x)=-10000; W(y,0)=Y(y,0)=go+(y-1)ge; X(y,0)=-10000
//   Wd=W(y-1,x-1)+T(s2(y),s1(x)); X=max(X(y,x-1)+ge, W(y,x-1)+go); Y=max(Y(y-1,x)+ge, W(y-1,x)+go)
//   W=max(max(X,Y),Wd)
// and its traceback, which looks only at the cell (not at a state): W==Wd -> diagonal (count the
// residue pair), else W==X -> left, else W==Y -> up; a gap is counted when a left-run or up-run opens.
// So the direction is a pure function of the cell: 2 bits per cell, computed in the forward pass.
//
// Design (B200): one warp per sequence pair, lanes own C consecutive columns in registers, the
// wavefront is skewed across lanes (lane l works on row t-l at step t) and the three values that
// cross a lane boundary (W, X of the lane's last column and the row's residue) move by warp
// shuffle.  A CTA handles pairs that share seq1 and builds its int8 query profile
// prof[band][residue][lane][16] once in shared memory: one conflict-free LDS.128 per lane per row
// yields all C substitution scores.  Direction bits: 2*C <= 32 bits per lane per step, one coalesced
// 128-byte store per warp-step into HBM scratch.  Sequences longer than 32*C columns are processed
// in column bands; the band's right boundary column (W, X per row) goes through a per-warp global
// buffer (L2-resident).  The walk is a separate kernel with one THREAD per pair: it is a dependent
// chain of ~L1+L2 loads, so it needs tens of thousands of chains in flight, not a warp each.
//
// Direction without compares: all DP values are carried scaled by 4 with a 2-bit priority tag in the
// low bits -- diagonal candidate 4v+3, X = 4v+2, Y = 4v+1, W = 4v -- so one 3-input max both selects
// the winner and, on ties, the reference's order (diagonal, then X, then Y); the tag of the maximum
// IS the traceback direction.  The int8 profile stores 4*T+3 directly.  Per cell this is two fused
// add-max (VIADDMNMX), one 3-input max, one mask and the adds -- no ISETP/SEL chain.
//
// 16-bit SIMD flavour (nw_forward_body_twin): when the scores fit, a warp aligns TWO pairs that share seq2
// at once, one per 16-bit half of every register (VIADDMNMX.U16x2 / VIMNMX3.U16x2).  To keep the plain
// adds on the FMA pipe as ordinary 32-bit integer multiply-adds, no carry may ever cross the halves, so
// all quantities are stored unsigned with a potential:  V^(y,x) = 4 V(y,x) + tag + H (x+y) + BIAS,
// with H >= -(4 go + 1), -4 ge so that every additive constant (gap open / extend, profile entry) is
// non-negative; the potential is the same for the three candidates of a cell, so maxima, ties and
// therefore directions are unchanged.  The host only selects this flavour when the largest possible value
// stays below 65536 (api.cu); otherwise the 32-bit flavour runs.
//
// Integer arithmetic only: bit-exact against the oracle.
#include "internal.h"
#include <mutex>

#define NW_WARPS 8
#define NW_THREADS (NW_WARPS * 32)
#define NW_MINFTY (-10000)

template <int C>
__device__ __forceinline__ void
nw_forward_body(const NwTask& task, const uint8_t* __restrict__ codes, const uint32_t* __restrict__ off,
                const int8_t* __restrict__ tab, int n1, int go, int ge,
                uint32_t* __restrict__ tb, int32_t* colbuf, uint32_t colbuf_stride, int32_t* __restrict__ scores_out,
                int k_one, int k_four, int k_mone) {
    // k_one == 1, k_four == 4, k_mone == -1 at run time: opaque multipliers that make the plain adds below integer
    // multiply-adds, i.e. FMA-pipe instructions, so the ALU pipe (which carries the max/logic ops and
    // is the kernel's limiter) and the FMA pipe are both busy
    extern __shared__ __align__(16) unsigned char smem[];
    int8_t* s_tab = (int8_t*)smem;                           // n1*n1, padded to 16
    const int tab_bytes = (n1 * n1 + 15) & ~15;
    int8_t* prof = (int8_t*)(smem + tab_bytes);              // [band][r][lane][16]

    const uint32_t s1_off = off[task.i];
    const int L1 = (int)(off[task.i + 1] - s1_off);
    const uint8_t* __restrict__ s1 = codes + s1_off;
    const int nbands = (L1 + 32 * C - 1) / (32 * C);

    for (int i = threadIdx.x; i < n1 * n1; i += NW_THREADS) s_tab[i] = tab[i];
    __syncthreads();
    // profile: prof[b][r][lane][k] = T(r, s1(x)), x = b*32C + lane*C + k (0-based); 0 beyond L1
    const int per_band = n1 * 512;
    for (int idx = threadIdx.x; idx < nbands * per_band; idx += NW_THREADS) {
        const int b = idx / per_band, rem = idx - b * per_band;
        const int r = rem >> 9, lk = rem & 511, lane = lk >> 4, k = lk & 15;
        const int x = b * 32 * C + lane * C + k;
        int8_t v = 0;
        if (k < C && x < L1) v = (int8_t)(4 * (int)s_tab[r + n1 * (int)s1[x]] + 3);   // 4*scoring_matrix(s2,s1)+3 (diag tag)
        prof[idx] = v;
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int32_t* colW = colbuf + (size_t)(blockIdx.x * NW_WARPS + warp) * 2 * colbuf_stride;
    int32_t* colX = colW + colbuf_stride;
    const int go4 = 4 * go, ge4 = 4 * ge;
    const int gox = go4 + 2, goy = go4 + 1;                   // gap-open into X (tag 2) / into Y (tag 1)
    const int XINF = 4 * NW_MINFTY + 2, YINF = 4 * NW_MINFTY + 1;

    for (uint32_t j = task.j0 + warp; j < task.j1; j += NW_WARPS) {
        const NwPair d = nw_pair(task, j, off);
        const int L2 = (int)d.L2;
        const uint8_t* __restrict__ s2 = codes + d.s2_off;
        const int steps = L2 + 31;                            // traceback layout: [band][step][lane]
        if (L2 > 0)
        for (int b = 0; b < nbands; ++b) {
            const int x0 = b * 32 * C + lane * C;            // columns x0+1 .. x0+C (1-based)
            const int cols_b = min(L1 - b * 32 * C, 32 * C);
            const int nsteps = L2 + (cols_b + C - 1) / C - 1; // lanes past the last column never start
            int Wup[C], Yup[C];                              // 4*W(y-1,x), 4*Y(y-1,x)+1
#pragma unroll
            for (int k = 0; k < C; ++k) { Wup[k] = go4 + (x0 + k) * ge4; Yup[k] = YINF; }
            int Wdiag = (x0 == 0) ? 0 : go4 + (x0 - 1) * ge4;  // 4*W(y-1, x0) for y-1 = 0
            int Wlast = 0, Xlast = 0, rcur = 0;
            // left-boundary inputs of lane 0 for the row it starts next (all lanes compute them: no branch)
            const bool first = (b == 0);
            int inW = first ? go4 : __ldcg(colW + 1);
            int inX = first ? XINF : __ldcg(colX + 1);
            int inR = s2[0];
            const int8_t* pb = prof + (size_t)b * per_band + lane * 16;
            uint32_t* __restrict__ tbl = tb + d.tb_off + (size_t)b * steps * 32 + lane;
            const bool has_next_band = (b + 1 < nbands);
            for (int t = 0; t < nsteps; ++t) {
                int Wl = __shfl_up_sync(0xffffffffu, Wlast, 1);
                int Xl = __shfl_up_sync(0xffffffffu, Xlast, 1);
                int r = __shfl_up_sync(0xffffffffu, rcur, 1);
                if (lane == 0) { Wl = inW; Xl = inX; r = inR; }
                {   // prefetch lane 0's inputs for row t+2 (clamped: rows past L2 are never used)
                    const int yn = min(t + 2, L2);
                    inR = s2[yn - 1];
                    if (first) { inW += ge4; } else { inW = __ldcg(colW + yn); inX = __ldcg(colX + yn); }
                }
                if ((unsigned)(t - lane) < (unsigned)L2) {      // row y = t - lane + 1 in [1, L2]
                    const uint4 pw = *reinterpret_cast<const uint4*>(pb + r * 512);
                    const uint32_t pws[4] = {pw.x, pw.y, pw.z, pw.w};
                    int Wleft = Wl, Xleft = Xl, dg = Wdiag;
                    uint32_t bits = 0;
#pragma unroll
                    for (int k = 0; k < C; ++k) {
                        const int s = (int)(int8_t)((pws[k >> 2] >> (8 * (k & 3))) & 0xff);   // 4*T+3
                        const int wd = dg + s;
                        const int xv = __viaddmax_s32(Xleft, ge4, Wleft * k_one + gox);
                        const int yv = __viaddmax_s32(Yup[k], ge4, Wup[k] * k_one + goy);
                        const int wt = __vimax3_s32(xv, yv, wd);                              // tagged maximum
                        const int wn = wt & ~3;
                        bits = bits * (uint32_t)k_four + (uint32_t)(wn * k_mone + wt);        // tag: 3 diag, 2 left, 1 up
                        dg = Wup[k];
                        Wup[k] = wn; Yup[k] = yv;
                        Wleft = wn; Xleft = xv;
                    }
                    Wdiag = Wl;
                    Wlast = Wleft; Xlast = Xleft; rcur = r;
                    *tbl = bits;
                    if (has_next_band && lane == 31) { const int y = t - 30; colW[y] = Wlast; colX[y] = Xlast; }
                }
                tbl += 32;
            }
            if (scores_out != nullptr) {
                const int kk = L1 - 1 - x0;                    // does this lane own column L1?
                if (kk >= 0 && kk < C) {
                    int v = 0;
#pragma unroll
                    for (int k = 0; k < C; ++k) if (k == kk) v = Wup[k];
                    scores_out[d.out_idx] = v >> 2;
                }
            }
            __syncwarp();
        }
        if (scores_out != nullptr && lane == 0 && (L1 == 0 || L2 == 0)) {
            // degenerate: W(L2,L1) is a boundary cell
            scores_out[d.out_idx] = (L1 == 0 && L2 == 0) ? 0 : go + (max(L1, L2) - 1) * ge;
        }
    }
}

// 16-bit twin flavour.  The CTA owns the rows (column sequences) of TWO tasks A = (i, j0..j1) and B = (i', j0'..j1)
// and every warp aligns the pairs (i, j) and (i', j) at once, A in the low and B in the high 16 bits of each
// register.  Because both halves share the row sequence j, the CTA's profile can hold both substitution scores
// of a cell in ONE word, prof[r][q][lane] (uint4 = 4 columns x (A | B << 16)), so a cell pair costs 9 instructions
// (5 multiply-adds on the FMA pipe, 4 max/logic on the ALU pipe) and no byte permutes.  The row residues are
// staged per warp in shared memory, pre-multiplied by the profile's row pitch.
#define NW_TWIN_ROW_BYTES 2048u                               // profile bytes per residue: 4 x 32 lanes x 16 B

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" :: "r"(addr), "r"(v) : "memory");
}
// write this lane's eight staged steps (one 32-byte sector) of an 8-step block
__device__ __forceinline__ void nw_flush_block(uint32_t tile_lane, uint32_t* __restrict__ dst) {
    uint32_t w[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w[q]) : "r"(tile_lane + 128u * q) : "memory");
    reinterpret_cast<uint4*>(dst)[0] = make_uint4(w[0], w[1], w[2], w[3]);
    reinterpret_cast<uint4*>(dst)[1] = make_uint4(w[4], w[5], w[6], w[7]);
}

template <int C>
__device__ __forceinline__ void
nw_forward_body_twin(const NwTask& ta, const NwTask* __restrict__ tasks, const uint8_t* __restrict__ codes,
                     const uint32_t* __restrict__ off, const int8_t* __restrict__ tab, int n1, int go, int ge,
                     uint32_t* __restrict__ tb, int32_t* __restrict__ scores_out, uint32_t rs_stride,
                     int k_one, int k_four, int k_mone, int H, int BIAS) {
    extern __shared__ __align__(16) unsigned char smem[];
    int8_t* s_tab = (int8_t*)smem;
    const int tab_bytes = (n1 * n1 + 15) & ~15;
    uint32_t* prof = (uint32_t*)(smem + tab_bytes);          // [r][q][lane][4] words
    uint32_t* tiles = prof + (size_t)n1 * (NW_TWIN_ROW_BYTES / 4);   // per warp [half][step & 7][lane] words
    uint32_t* rs_all = tiles + NW_WARPS * 512;

    const bool twin = (ta.C & NW_TWIN_A) != 0;
    NwTask tbk = ta;
    if (twin) tbk = tasks[blockIdx.x + 1];
    const uint32_t s1a_off = off[ta.i], s1b_off = off[tbk.i];
    const int L1a = (int)(off[ta.i + 1] - s1a_off), L1b = twin ? (int)(off[tbk.i + 1] - s1b_off) : 0;
    const uint8_t* __restrict__ s1a = codes + s1a_off;
    const uint8_t* __restrict__ s1b = codes + s1b_off;
    for (int i = threadIdx.x; i < n1 * n1; i += NW_THREADS) s_tab[i] = tab[i];
    __syncthreads();
    for (int idx = threadIdx.x; idx < n1 * 512; idx += NW_THREADS) {
        const int r = idx >> 9, q = (idx >> 7) & 3, ln = (idx >> 2) & 31, w = idx & 3;
        const int k = 4 * q + w, x = ln * C + k;
        uint32_t v = 0;
        if (k < C) {
            if (x < L1a) v = (uint32_t)(4 * (int)s_tab[r + n1 * (int)s1a[x]] + 3 + 2 * H);
            if (x < L1b) v |= (uint32_t)(4 * (int)s_tab[r + n1 * (int)s1b[x]] + 3 + 2 * H) << 16;
        }
        prof[idx] = v;
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* rs = rs_all + (size_t)warp * rs_stride;         // rs[32 + row] = residue(row) * row pitch
    const int go4 = 4 * go, ge4 = 4 * ge;
    auto pack2 = [](int v) -> uint32_t { return (uint32_t)v | ((uint32_t)v << 16); };
    uint32_t gex2 = pack2(ge4 + H), gox2 = pack2(go4 + 2 + H), goy2 = pack2(go4 + 1 + H);
    uint32_t u_one = (uint32_t)k_one, u_four = (uint32_t)k_four, u_mone = (uint32_t)k_mone;
    // pin the loop constants in registers (otherwise they are re-derived from the parameters every step)
    asm volatile("" : "+r"(gex2), "+r"(gox2), "+r"(goy2), "+r"(u_one), "+r"(u_four), "+r"(u_mone));
    const int L1m = max(L1a, L1b);
    const int nl = (L1m + C - 1) / C;
    const int x0 = lane * C;
    uint32_t prof_lane = (uint32_t)__cvta_generic_to_shared(prof) + (uint32_t)lane * 16u;
    uint32_t inWstep = pack2(ge4 + H);
    uint32_t tile_lane = (uint32_t)__cvta_generic_to_shared(tiles + warp * 512) + (uint32_t)lane * 4u;
    asm volatile("" : "+r"(prof_lane), "+r"(inWstep), "+r"(tile_lane));
    const uint32_t off_a0 = off[ta.j0], off_b0 = off[tbk.j0];

    for (uint32_t j = ta.j0 + warp; j < ta.j1; j += NW_WARPS) {
        const uint32_t s2_off = off[j];
        const int L2 = (int)(off[j + 1] - s2_off);
        const uint8_t* __restrict__ s2 = codes + s2_off;
        const bool hasB = twin && j >= tbk.j0;
        __syncwarp();
        for (int q = lane; q < L2; q += 32) rs[32 + q] = (uint32_t)s2[q] * NW_TWIN_ROW_BYTES;
        __syncwarp();
        // this lane's sector of the pair's first 8-step block (layout: internal.h)
        uint32_t* __restrict__ pa = tb + ta.tb_base + 32ull * (((uint64_t)(s2_off - off_a0) + NW_U16_STEP_SLACK * (j - ta.j0) + 7u) & ~7ull) + lane * 8;
        uint32_t* __restrict__ pb = pa;
        if (hasB) pb = tb + tbk.tb_base + 32ull * (((uint64_t)(s2_off - off_b0) + NW_U16_STEP_SLACK * (j - tbk.j0) + 7u) & ~7ull) + lane * 8;
        uint32_t toff = 0;                                    // (t & 7) * 128: byte offset of this step's tile row

        const int nsteps = L2 + nl - 1;
        uint32_t Wup[C], Yup[C];
#pragma unroll
        for (int k = 0; k < C; ++k) { Wup[k] = pack2(4 * (go + (x0 + k) * ge) + H * (x0 + k + 1) + BIAS); Yup[k] = 0u; }
        uint32_t Wdiag = (x0 == 0) ? pack2(BIAS) : pack2(4 * (go + (x0 - 1) * ge) + H * x0 + BIAS);
        uint32_t Wlast = 0, Xlast = 0;
        uint32_t inW = pack2(go4 + H + BIAS);                 // W^(1,0)
        const uint32_t* rp = rs + 32 - lane;                  // &rs[32 + row] for row = t - lane
        uint32_t rnext = rp[0];
        int row = -lane;
        for (int t = 0; t < nsteps; ++t) {
            uint32_t Wl = __shfl_up_sync(0xffffffffu, Wlast, 1);
            uint32_t Xl = __shfl_up_sync(0xffffffffu, Xlast, 1);
            if (lane == 0) { Wl = inW; Xl = 0u; }
            inW += inWstep;
            const uint32_t rcur = rnext;
            ++rp;
            rnext = *rp;                                      // next row's residue (slots past L2 are never used)
            if ((unsigned)row < (unsigned)L2) {
                const uint32_t pbase = prof_lane + rcur;
                uint32_t sw[16];
                {
                    const uint4 v = lds128(pbase);
                    sw[0] = v.x; sw[1] = v.y; sw[2] = v.z; sw[3] = v.w;
                }
                if (C > 4) { const uint4 v = lds128(pbase + 512u); sw[4] = v.x; sw[5] = v.y; sw[6] = v.z; sw[7] = v.w; }
                if (C > 8) { const uint4 v = lds128(pbase + 1024u); sw[8] = v.x; sw[9] = v.y; sw[10] = v.z; sw[11] = v.w; }
                if (C > 12) { const uint4 v = lds128(pbase + 1536u); sw[12] = v.x; sw[13] = v.y; sw[14] = v.z; sw[15] = v.w; }
                // phase 1: everything that only needs the previous row
                uint32_t wd[C];
                wd[0] = Wdiag * u_one + sw[0];
#pragma unroll
                for (int k = 1; k < C; ++k)                     // every third add stays on the ALU pipe: pipe balance
                    wd[k] = (k % 3 == 1) ? Wup[k - 1] + sw[k] : Wup[k - 1] * u_one + sw[k];
#pragma unroll
                for (int k = 0; k < C; ++k) Yup[k] = __viaddmax_u16x2(Yup[k], gex2, Wup[k] * u_one + goy2);
                // phase 2: the chain along the row
                uint32_t Wleft = Wl, Xleft = Xl, bits = 0, F = 0;
#pragma unroll
                for (int k = 0; k < C; ++k) {
                    Xleft = __viaddmax_u16x2(Xleft, gex2, Wleft * u_one + gox2);
                    const uint32_t wt = __vimax3_u16x2(Xleft, Yup[k], wd[k]);
                    Wleft = wt & 0xfffcfffcu;
                    bits = Wleft * u_mone + (bits * u_four + wt);   // per half: tag 3 diag, 2 left, 1 up
                    Wup[k] = Wleft;
                    if (k == 7) { F = bits; bits = 0; }
                }
                if (C < 8) { F = bits; bits = 0; }
                Wdiag = Wl;
                Wlast = Wleft; Xlast = Xleft;
                sts32(tile_lane + toff, __byte_perm(F, bits, 0x5410));
                sts32(tile_lane + toff + 1024u, __byte_perm(F, bits, 0x7632));
            }
            ++row;
            if (toff == 896u) {
                nw_flush_block(tile_lane, pa);
                if (hasB) nw_flush_block(tile_lane + 1024u, pb);
                pa += 256; pb += 256;
            }
            toff = (toff + 128u) & 1023u;
        }
        if (toff != 0u) {
            nw_flush_block(tile_lane, pa);
            if (hasB) nw_flush_block(tile_lane + 1024u, pb);
        }
        if (scores_out != nullptr) {
            // every lane stopped after its row L2, so Wup holds W^(L2, .)
            const int ka = L1a - 1 - x0, kb = L1b - 1 - x0;
            uint32_t va = 0, vb = 0;
#pragma unroll
            for (int k = 0; k < C; ++k) { if (k == ka) va = Wup[k] & 0xffffu; if (k == kb) vb = Wup[k] >> 16; }
            if (ka >= 0 && ka < C) scores_out[ta.out_base + (j - ta.j0)] = ((int)va - H * (L1a + L2) - BIAS) >> 2;
            if (hasB && kb >= 0 && kb < C) scores_out[tbk.out_base + (j - tbk.j0)] = ((int)vb - H * (L1b + L2) - BIAS) >> 2;
        }
    }
}

// One thread per pair (one warp per task): follow the 2-bit tags from (L2,L1) until x==0 or y==0
// (DistanceFactoryAlign.h:99-123).  Emits the matched state pairs (s1*n+s2, both < n) and the
// number of gap openings.
__global__ void __launch_bounds__(128)
nw_walk_kernel(int n_tasks, int n, const uint8_t* __restrict__ codes, const uint32_t* __restrict__ off,
               const NwTask* __restrict__ tasks, const uint32_t* __restrict__ tb, uint16_t* __restrict__ list,
               uint32_t* __restrict__ list_len, uint32_t* __restrict__ gaps_out, int* __restrict__ err, int lpw) {
    // lpw pairs per warp (32 or 8): 32 / lpw warps share a task
    const int wpt = 32 / lpw, gw = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int ti = gw / wpt;
    if (ti >= n_tasks || (int)(threadIdx.x & 31) >= lpw) return;
    const NwTask task = tasks[ti];
    const uint32_t j = task.j0 + (uint32_t)((gw - ti * wpt) * lpw) + (threadIdx.x & 31);
    if (j >= task.j1) return;
    const int C = (int)(task.C & 0xff);
    const bool u16 = (task.C & NW_MODE_U16) != 0;
    const NwPair d = nw_pair(task, j, off);
    const uint8_t* __restrict__ s1 = codes + d.s1_off;
    const uint8_t* __restrict__ s2 = codes + d.s2_off;
    const uint32_t* __restrict__ tbp = tb + d.tb_off;
    uint16_t* __restrict__ lp = list + d.list_off;
    const int steps = (int)d.L2 + 31;
    int x = (int)d.L1, y = (int)d.L2;
    uint32_t gaps = 0, nl = 0, prev = 3u;
    // (band, lane, k) of column x, kept incrementally: no divisions inside the chain
    int band = 0, lane = 0, k = 0;
    if (x > 0) { const int xz = x - 1, bandw = 32 * C; band = xz / bandw; const int rem = xz - band * bandw; lane = rem / C; k = rem - lane * C; }
    const int c8 = min(C, 8);
    bool pend = false;                                          // a diagonal step whose residues are still in flight
    int pc1 = 0, pc2 = 0;
    while (x != 0 && y != 0) {
        const int t = y - 1 + lane;
        const uint32_t* wp = u16 ? tbp + ((((uint32_t)t >> 3) * 32u + (uint32_t)lane) << 3) + (t & 7)
                                 : tbp + ((size_t)band * steps + t) * 32 + lane;
        const uint32_t w = __ldg(wp);
        if (pend && pc1 < n && pc2 < n) lp[nl++] = (uint16_t)(pc1 * n + pc2);   // previous step's pair: its loads overlap this one's
        // 32-bit flavour: cell k at bit 2(C-1-k).  16-bit flavour: cells 0..7 in the low half, 8.. in the high half
        const int sh = !u16 ? 2 * (C - 1 - k) : (k < 8 ? 2 * (c8 - 1 - k) : 16 + 2 * (C - 1 - k));
        const uint32_t tag = (w >> sh) & 3u;                       // 3 diagonal, 2 left (X), 1 up (Y)
        if (tag == 0u) { *err = 1; pend = false; break; }          // reference: error("error while backtracking")
        pend = (tag == 3u);
        if (pend) { pc1 = s1[x - 1]; pc2 = s2[y - 1]; }
        gaps += (tag != 3u && tag != prev) ? 1u : 0u;              // a left-run or an up-run opens
        prev = tag;
        const int dx = (int)(tag >> 1), dy = (int)(tag & 1u);
        x -= dx; y -= dy; k -= dx;
        if (k < 0) { k = C - 1; if (--lane < 0) { lane = 31; --band; } }
    }
    if (pend && pc1 < n && pc2 < n) lp[nl++] = (uint16_t)(pc1 * n + pc2);
    list_len[d.out_idx] = nl;
    gaps_out[d.out_idx] = gaps;
}

// One launch for every task: the CTA dispatches on its task's C (columns per lane), so rows of
// different lengths all keep their 32 lanes busy without paying one launch tail per length class.
__global__ void __launch_bounds__(NW_THREADS)
nw_forward_kernel(const uint8_t* __restrict__ codes, const uint32_t* __restrict__ off,
                  const NwTask* __restrict__ tasks, const int8_t* __restrict__ tab, int n1, int go, int ge,
                  uint32_t* __restrict__ tb, int32_t* colbuf, uint32_t colbuf_stride, int32_t* __restrict__ scores_out,
                  int k_one, int k_four, int k_mone, int u16_h, int u16_bias, uint32_t rs_stride) {
    const NwTask task = tasks[blockIdx.x];
    if (task.C & NW_TWIN_B) return;                            // aligned by the CTA of the preceding task
    if (task.C & NW_MODE_U16) {
#define CASE(c) case c: nw_forward_body_twin<c>(task, tasks, codes, off, tab, n1, go, ge, tb, scores_out, rs_stride, k_one, k_four, k_mone, u16_h, u16_bias); break;
        switch (task.C & 0xff) {
            CASE(4) CASE(6) CASE(8) CASE(10) CASE(12) CASE(13) CASE(14) CASE(16)
        }
#undef CASE
        return;
    }
#define CASE(c) case c: nw_forward_body<c>(task, codes, off, tab, n1, go, ge, tb, colbuf, colbuf_stride, scores_out, k_one, k_four, k_mone); break;
    switch (task.C & 0xff) {
        CASE(4) CASE(6) CASE(8) CASE(10) CASE(12) CASE(13) CASE(14) CASE(16)
    }
#undef CASE
}

int launch_nw_forward(cudaStream_t st, int n_tasks, int n_states, int max_bands, const uint8_t* d_codes,
                      const uint32_t* d_off, const NwTask* d_tasks, const int8_t* d_tab, int go, int ge,
                      uint32_t* d_tb, int32_t* d_colbuf, uint32_t colbuf_stride, int32_t* d_scores, int u16_h, int u16_bias,
                      int u16_max_rows) {
    if (n_tasks == 0) return 0;
    const int n1 = n_states + 1;
    const uint32_t rs_stride = (uint32_t)((u16_max_rows + 64 + 3) & ~3);   // 32 slots before row 0, >= 1 after the last row
    const size_t prof32 = (size_t)max_bands * n1 * 512;
    const size_t prof16 = u16_max_rows > 0 ? (size_t)n1 * NW_TWIN_ROW_BYTES + (size_t)NW_WARPS * (512 + rs_stride) * 4 : 0;
    const size_t smem = ((n1 * n1 + 15) & ~15) + (prof32 > prof16 ? prof32 : prof16);
    if (smem > 227 * 1024) { pgm_set_error("NW profile needs %zu B shared memory (> 227 KB)", smem); return PGM_ERR_ARG; }
    {   // the attribute call is a driver round trip: only when the size grows; attributes are per device
        static size_t attr_smem[64] = {0}; static std::mutex attr_mu;
        int dev = 0;
        PGM_CUDA(cudaGetDevice(&dev));
        std::lock_guard<std::mutex> lk(attr_mu);
        if (smem > attr_smem[dev & 63]) {
            PGM_CUDA(cudaFuncSetAttribute(nw_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_smem[dev & 63] = smem;
        }
    }
    nw_forward_kernel<<<n_tasks, NW_THREADS, smem, st>>>(d_codes, d_off, d_tasks, d_tab, n1, go, ge, d_tb,
                                                        d_colbuf, colbuf_stride, d_scores, 1, 4, -1, u16_h, u16_bias, rs_stride);
    PGM_CUDA(cudaGetLastError());
    return 0;
}

int launch_nw_walk(cudaStream_t st, int n_tasks, int n_states, const uint8_t* d_codes, const uint32_t* d_off,
                   const NwTask* d_tasks, const uint32_t* d_tb, uint16_t* d_list, uint32_t* d_list_len, uint32_t* d_gaps,
                   int* d_err, int lanes_per_warp) {
    if (n_tasks == 0) return 0;
    const int lpw = lanes_per_warp == 8 ? 8 : 32;
    const long long warps = (long long)n_tasks * (32 / lpw);
    nw_walk_kernel<<<(unsigned)((warps + 3) / 4), 128, 0, st>>>(n_tasks, n_states, d_codes, d_off, d_tasks, d_tb, d_list,
                                                                d_list_len, d_gaps, d_err, lpw);
    PGM_CUDA(cudaGetLastError());
    return 0;
}

int nw_pick_C(int max_L1) {
    static const int cs[] = {4, 6, 8, 10, 12, 13, 14, 16};
    for (int c : cs) if (32 * c >= max_L1) return c;
    return 16;
}
