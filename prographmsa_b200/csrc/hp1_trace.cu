// hp1_trace.cu -- K6 for the strip kernels: the traceback that chases the packed direction bits (hp1_dir.cuh).
//
// Replaces the reference's traceback (GraphAlign.h:283-521, markAlternativePath :165-198) for jobs swept by hp1_strip.cu.
#include "hp1_dir.cuh"

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
    const int N1 = J.N1, N2 = J.N2;
    const float gap_init = J.sc[0], gap_extend = J.sc[1];
    const int sC = J.C, sW = strip_words(J.C), sN = J.nstrips;
    // the direction word of cell (yy, xx) in the canonical format of hp1_dir.cuh, whatever layout the fill wrote
    auto dword = [&](int yy, int xx) -> uint32_t {
        {                                                // 4-bit cells (hp1_strip.cu): mlose, xbeats, yopen, xopen
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
graph_trace_packed_kernel(int n_jobs, const GraphJobDev* __restrict__ jobs, const int* __restrict__ order) {
    if ((int)blockIdx.x >= n_jobs) return;
    const GraphJobDev& J = jobs[order ? order[blockIdx.x] : blockIdx.x];
    if (J.layout != GF_LAYOUT_STRIP) return;             // the dense-matrix traceback (hp1_graphdp.cu) walks the other jobs
    graph_trace_job<7>(J, (int)threadIdx.x);
}

int launch_graph_trace_packed(cudaStream_t st, int n_jobs, const GraphJobDev* d_jobs, const int* d_order) {
    if (n_jobs == 0) return 0;
    graph_trace_packed_kernel<<<n_jobs, 32, 0, st>>>(n_jobs, d_jobs, d_order);
    PGM_CUDA(cudaGetLastError());
    return 0;
}

