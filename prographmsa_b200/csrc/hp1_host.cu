// hp1_host.cu -- host side of HP-1 behind pgm_align_graphs_batch / pgm_graph_batch_*:
// DynProgScores (reference GraphAlign.h:98-143, incl. averageAlignmentLength :56-96) in double on the
// host, predecessor lists in PredIterator order with their edge values (Graph.h:172-248), one packed
// host->device arena per batch, sub-batching by the scratch budget, result fetch.
#include "internal.h"
#include <atomic>
#include <thread>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <chrono>
#include <cstdio>

struct pgm_ctx;
cudaStream_t pgm_ctx_stream(pgm_ctx* c);
int pgm_ctx_device(pgm_ctx* c);
uint64_t pgm_ctx_scratch(pgm_ctx* c);
void pgm_ctx_time_begin(pgm_ctx* c, int kind, void** tok);
void pgm_ctx_time_begin_on(pgm_ctx* c, int kind, cudaStream_t st, void** tok);
void pgm_ctx_time_end(pgm_ctx* c, void* tok);
pgm_timing* pgm_ctx_timing(pgm_ctx* c);
void pgm_ctx_reset_timing(pgm_ctx* c);
int launch_graph_fill(cudaStream_t st, int n_jobs, const GraphJobDev* d_jobs, const int* d_order, int max_nb, int max_n2, int max_np2);
int launch_graph_fill_diag(cudaStream_t st, int n_jobs, const GraphJobDev* d_jobs, const int* d_order, int max_n1, int max_n2, int max_np);
int launch_graph_fill_banded(cudaStream_t st, int n_jobs, const GraphJobDev* d_jobs, const int* d_order, int max_bands, int n_sms, int band_rows,
                             int max_n2, int max_np);
int pgm_ctx_take_dev(pgm_ctx* c, size_t bytes, DevBuf& out);
static inline int ru32(int v) { return (v + 31) & ~31; }
int pgm_ctx_take_host(pgm_ctx* c, size_t bytes, HostPinned& out);
void pgm_ctx_give_dev(pgm_ctx* c, DevBuf& b);
void pgm_ctx_give_host(pgm_ctx* c, HostPinned& b);

namespace {

constexpr float FINF = std::numeric_limits<float>::infinity();
inline double LOG2(double x) { return std::log(x) / std::log(2); }     // GraphAlign.h:46-47 with USE_LS_LOG

double average_alignment_length(const pgm_graph& g) {                   // GraphAlign.h:56-96, bottom-up
    if (g.N == 0) return 0;
    std::vector<double> cache(g.N, -2.0);
    cache[0] = 0;
    for (int cur = 1; cur < g.N; ++cur) {
        double sum = 0; unsigned paths = 0;
        for (int k = g.e_rowptr[cur]; k < g.e_rowptr[cur + 1]; ++k) {
            const float c = g.e_val[k];
            const float v = c == 0 ? FINF : c + 10000.0f;
            if (v == 0.0f) {
                const double res = cache[g.e_col[k]];
                if (res >= 0.0) { sum += res + 1.0; ++paths; }
            }
        }
        cache[cur] = paths > 0 ? sum / paths : -2.0;
    }
    return cache[g.N - 1];
}

void dynprog_scores(const pgm_graph& g1, const pgm_graph& g2, const pgm_dp_model& m, const pgm_dp_opts& o, float* s) {
    const double l1 = average_alignment_length(g1), l2 = average_alignment_length(g2);
    const double eps = m.epsilon, delta = m.delta, E = o.end_indel_prob;
    const double exp_length = std::max(l1, l2) * std::exp(m.distance * o.indel_rate * (eps / (1.0 - eps) + 1.0));
    const double nu = 2.0 / (2.0 + l1 + l2);
    double tau = 1.0 / (1.0 + exp_length);
    if (eps + tau >= 1.0) tau = (1.0 - eps) / 2.0;
    s[0] = (float)LOG2(delta * (1.0 - eps - tau) / (1.0 - nu));                       // gap_init
    s[1] = (float)LOG2(eps / (1.0 - nu));                                             // gap_extend
    s[2] = (float)LOG2((1.0 - 2.0 * delta) * (1.0 - tau) / (1.0 - nu) / (1.0 - nu));  // match_init
    s[5] = (float)LOG2(tau);                                                          // end_skip
    if (E >= 0 && E <= 1) {
        s[3] = (float)LOG2(tau * (1.0 - E) / (1.0 - 2.0 * delta) / (1.0 - tau));      // end_match
        s[4] = (float)LOG2(tau * E / 2.0 / (1.0 - eps - tau) / delta);                // end_gap
        s[6] = (float)LOG2(E / 2.0 * (1.0 - eps - tau) / (1.0 - E) / (1.0 - nu));     // start_gap
        s[7] = (float)LOG2((1.0 - tau) * (1.0 - E));                                  // start_init
    } else {
        s[3] = (float)LOG2(tau / (1.0 - tau));
        s[4] = (float)LOG2(tau / (1.0 - eps - tau));
        s[6] = (float)LOG2(delta * (1.0 - eps - tau) / (1.0 - nu));
        s[7] = (float)LOG2(1.0 - tau);
    }
    const double repeat_prob = 1.0 - std::exp(-m.distance * o.repeat_rate);
    s[8] = (float)-LOG2(std::min<double>(1, repeat_prob / (1 - repeat_prob) * (1 - o.repeatext_prob)));   // repeat_init
    s[9] = (float)-LOG2(std::min<double>(1, std::max<double>(0, o.repeatext_prob)));                      // repeat_ext
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct JobLayout {                // byte offsets into the upload arena / scratch arena / output arena
    size_t sites1, sites2, Mf, pif, p1_ptr, p1_idx, p1_val, p1_rep, p1_alt, p2_ptr, p2_idx, p2_val, p2_rep, p2_alt, ni1, ni2, colinfo, blkstart;
    size_t S, MXYW, D, Vrow, Vcol, Tbuf, Aimg, Bimg, u, v, alt_score, alt_prev, progress;     // scratch
    size_t map1, map2, result;                           // output
    int np1, np2, cap, pitch, nb, layout, bands, band_rows;
    int dbits, ldVr, ldVc, nkr, nkc;
    int C, nstrips, N1e, N2e, coded1, coded2;
    int sgen, nsrc, ring_rows, hist_rows, swarps, nstrips_gen, nnc1, nnc2, ldRI, ldCI;
    int wave, wv_tr, wv_chain_pad, wv_ncols, wv_nx, wv_L, wv_tcols;       // role-sorted anti-diagonal kernel (hp1_wave.cu)
    size_t wv_rows, wv_cols, wv_ci, wv_rowslot, wv_xi, wv_xv, wv_adj1;
    size_t scol, sstart, colstrip, specmask, srcmask, colsrc_off, ncidx1_off, ncidx2_off, rinfo_off, RowIdx, ColIdx;
    std::vector<int16_t> colsrc; std::vector<int> ncidx1, ncidx2;
    double uniform1, uniform2;
    std::vector<int> keep1, keep2;                       // per node: index of its row / column in the value store, or -1
    std::vector<uint8_t> cls1, cls2;                     // class-coded graphs: site class per node
};

// Keep sets: which rows of graph 1 / columns of graph 2 the fill writes to the value store as {M,X,Y,W}.  A fill kernel
// finds a predecessor cell in its shared-memory ring only when it was produced recently: the blocked kernel's ring holds a
// cell (yp, xp) for the cell (y, x) while (y - yp) + (blk(x) - blk(xp)) < R (R >= 4 sweep steps), the anti-diagonal kernel
// while (y - yp) + (x - xp) < RING.  So a source row reached by an edge spanning >= t_row positions, or a source column
// reached by one spanning >= t_col, is kept, with t_row + t_col - 2 < ring depth: every ring miss then hits a kept row or
// column.  Also kept: row 0 and column 0 (the traceback's closest-|delta| rule on the boundary needs their values), the
// predecessors of graph 1's end node (end-node match), and every row an edge into a LATER band starts from.
void mark_far_sources(const pgm_graph& g, const int* pos, int thr, int band_rows, std::vector<int>& keep) {
    for (int to = 1; to <= g.N - 2; ++to) {
        const int pt = pos ? pos[to] : to;
        auto visit = [&](int from) {
            const int pf = pos ? pos[from] : from;
            if (pt - pf >= thr || (band_rows > 0 && from / band_rows != to / band_rows)) keep[from] = 1;
        };
        for (int e = g.e_rowptr[to]; e < g.e_rowptr[to + 1]; ++e) visit(g.e_col[e]);
        if (g.r_rowptr) for (int e = g.r_rowptr[to]; e < g.r_rowptr[to + 1]; ++e) visit(g.r_col[e]);
    }
}
int number_kept(std::vector<int>& keep) {                // flags -> indices
    int n = 0;
    for (int& k : keep) k = k ? n++ : -1;
    return n;
}
int max_preds(const pgm_graph& g) {
    int m = 0;
    for (int node = 0; node < g.N; ++node)
        m = std::max(m, g.e_rowptr[node + 1] - g.e_rowptr[node] + (g.r_rowptr ? g.r_rowptr[node + 1] - g.r_rowptr[node] : 0));
    return m;
}

// ---- job analysis (parallel over jobs): everything that decides sizes ------------------------------------------------------
struct BatchEnv {                 // environment switches, read once per create (tests, A-B runs, profiling)
    int force_layout = -1, diag_min_generic = 8, band_min_n1 = GF_BAND_MIN_N1, long_band_rows = GF_BAND_ROWS;
    bool band_rows_forced = false;
    int band_min_nowave = 550;
    int split_rows = 0, split_min_generic = 24, small_batch = 3 * 148, strip = 2, coding = 1, strip_c = 0, hist_kb = 0, dbg = 0, strip_max_generic = 16;
    int wave = 1, wave_L = 0, wave_ring = 0;
    BatchEnv() {
        if (const char* e = getenv("PGM_FILL_LAYOUT")) force_layout = !strcmp(e, "diag") ? GF_LAYOUT_DIAG : (!strcmp(e, "block") ? GF_LAYOUT_BLOCK : (!strcmp(e, "strip") ? GF_LAYOUT_STRIP : -1));
        if (const char* e = getenv("PGM_FILL_DIAG_MIN")) diag_min_generic = atoi(e);
        if (const char* e = getenv("PGM_FILL_BAND_MIN_N1")) { band_min_n1 = std::max(66, atoi(e)); band_min_nowave = band_min_n1; }
        if (const char* e = getenv("PGM_FILL_BAND_MIN_NOWAVE")) band_min_nowave = std::max(66, atoi(e));
        if (const char* e = getenv("PGM_FILL_BAND_ROWS")) { const int v = atoi(e) & ~31; if (v >= 32 && v <= 512) { long_band_rows = v; band_rows_forced = true; } }
        if (const char* e = getenv("PGM_FILL_SPLIT_ROWS")) split_rows = atoi(e) & ~31;
        if (const char* e = getenv("PGM_FILL_SPLIT_MIN")) split_min_generic = atoi(e);
        if (const char* e = getenv("PGM_FILL_SMALL_BATCH")) small_batch = atoi(e);      // fewer jobs than three per SM: latency-bound (measured crossover 300..440 jobs of 400 x 400)
        if (const char* e = getenv("PGM_FILL_STRIP")) strip = atoi(e);                  // 0: never a strip kernel, 1: chain graphs only, 2: also graphs with non-chain nodes
        if (const char* e = getenv("PGM_STRIP_HIST_KB")) hist_kb = atoi(e);             // shared-memory budget of the (Y,W) history ring
        if (const char* e = getenv("PGM_STRIP_DEBUG")) dbg = atoi(e);
        if (const char* e = getenv("PGM_STRIP_MAX_GENERIC")) strip_max_generic = atoi(e); // most non-chain nodes of a job the general strip kernel takes
        if (const char* e = getenv("PGM_STRIP_C")) strip_c = atoi(e);                   // force the general strip kernel's strip width
        if (const char* e = getenv("PGM_SITE_CLASSES")) coding = atoi(e);               // 0: never class-code leaf graphs
        if (const char* e = getenv("PGM_FILL_WAVE")) wave = atoi(e);                    // 0: never the role-sorted anti-diagonal kernel, 1: for every anti-diagonal job it fits
        if (const char* e = getenv("PGM_WAVE_L")) wave_L = atoi(e);                     // force the lanes per non-chain column (1, 2, 4)
        if (const char* e = getenv("PGM_WAVE_RING")) wave_ring = atoi(e);               // force the ring depth (8, 16)
    }
};

inline bool non_chain_node(const pgm_graph& g, int x) {
    const int nreg = g.e_rowptr[x + 1] - g.e_rowptr[x];
    const int nrep = g.r_rowptr ? g.r_rowptr[x + 1] - g.r_rowptr[x] : 0;
    return !(nreg == 1 && nrep == 0 && g.e_col[g.e_rowptr[x]] == x - 1);
}

// Site classes (see GraphJobDev::coded1): class a < n = one-hot column of state a (exactly 1.0), class n = uniform column
// (every entry equal to the graph's one uniform value, SequenceGraph.h:50-53).  -> false when some interior column is neither.
bool site_classes(const pgm_graph& g, int n, std::vector<uint8_t>& cls, double& uniform) {
    cls.assign(g.N, 0);
    uniform = 0.0;
    for (int node = 1; node <= g.N - 2; ++node) {
        const double* s = g.sites + (size_t)n * node;
        int hot = -1, nz = 0;
        for (int a = 0; a < n; ++a) if (s[a] != 0.0) { ++nz; hot = a; }
        if (nz == 1 && s[hot] == 1.0) { cls[node] = (uint8_t)hot; continue; }
        if (nz == n) {
            bool same = true;
            for (int a = 1; a < n; ++a) same &= (s[a] == s[0]);
            if (same && (uniform == 0.0 || uniform == s[0])) { uniform = s[0]; cls[node] = (uint8_t)n; continue; }
        }
        return false;
    }
    return true;
}

thread_local bool t_full_emission = false;   // pgm_emission_scores: the caller wants the whole S matrix, so no class coding

} // namespace

// side streams for concurrent fill launches: one set per device, created on first use and kept for the process
struct SideStreams { std::vector<cudaStream_t> streams; std::vector<cudaEvent_t> join, emit_done; cudaEvent_t fork = nullptr; };
static SideStreams& side_streams(int device) { static SideStreams s[64]; return s[device & 63]; }

struct pgm_graph_batch {
    int n_states = 0, n_jobs = 0, emission_mode = 0;
    std::vector<GraphJobDev> jobs;          // host copy with DEVICE pointers
    std::vector<JobLayout> lay;
    std::vector<uint64_t> cell_updates;
    DevBuf d_in, d_scratch, d_out, d_jobs;
    HostPinned h_in, h_out;
    size_t out_bytes = 0;
    std::vector<std::pair<int, int>> sub;   // [begin, end) job ranges that fit the scratch budget
    // Fill schedule.  The fill kernel's block size and shared memory -- hence CTAs per SM -- are set by the LARGEST job
    // of a launch, so each sub-batch is swept as a few launches of similar graph-2 length (largest first), on
    // separate streams so that a class with few jobs does not leave the GPU idle.
    struct FillLaunch { int sub, offset, count, layout, max_nb, max_n1, max_n2, max_np2, max_np, max_bands, band_rows, dbits, C, max_n1e, max_n2e, sgen, swarps; size_t ring_bytes;
                        int wave = 0, wL = 0, wring = 8, wthreads = 0, max_tr = 0, max_np1 = 0, max_nx = 0; };
    std::vector<int> order;                 // job indices relative to their sub-batch, grouped by FillLaunch
    std::vector<FillLaunch> fills;
    size_t scratch_bytes_used = 0;
    int max_N1 = 0, max_N2 = 0;
    uint64_t total_updates = 0, total_cells = 0;
};

namespace {
template <class F> void parallel_jobs(int n_jobs, F&& fn) {
    unsigned hw = std::thread::hardware_concurrency();
    const int nthreads = (int)std::max(1u, std::min<unsigned>(std::min<unsigned>(hw ? hw : 1u, 16u), (unsigned)(n_jobs / 8 + 1)));
    if (nthreads <= 1) { for (int k = 0; k < n_jobs; ++k) fn(k); return; }
    std::atomic<int> next(0);
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; ++t)
        pool.emplace_back([&] { for (int k = next.fetch_add(1); k < n_jobs; k = next.fetch_add(1)) fn(k); });
    for (auto& th : pool) th.join();
}
} // namespace

extern "C" {

int pgm_dynprog_scores(int n_states, const pgm_graph* g1, const pgm_graph* g2, const pgm_dp_model* model,
                       const pgm_dp_opts* opts, float out[10]) {
    (void)n_states;
    if (!g1 || !g2 || !model || !opts || !out) { pgm_set_error("NULL argument"); return PGM_ERR_ARG; }
    dynprog_scores(*g1, *g2, *model, *opts, out);
    return 0;
}

int pgm_graph_batch_create(pgm_ctx* c, int n, int n_jobs, const pgm_graph* g1, const pgm_graph* g2,
                           const pgm_dp_model* models, const pgm_dp_opts* opts, pgm_graph_batch** out) {
    if (!c || !out || !opts || n_jobs < 0 || (n_jobs > 0 && (!g1 || !g2 || !models))) { pgm_set_error("bad arguments"); return PGM_ERR_ARG; }
    if (n < 2 || n > 64) { pgm_set_error("n_states %d out of range", n); return PGM_ERR_ARG; }
    PGM_CUDA(cudaSetDevice(pgm_ctx_device(c)));
    cudaStream_t st = pgm_ctx_stream(c);
    for (int k = 0; k < n_jobs; ++k) {
        if (g1[k].N < 2 || g2[k].N < 2) { pgm_set_error("job %d: graphs need at least start and end nodes", k); return PGM_ERR_ARG; }
        if (g1[k].N > 60000 || g2[k].N > 60000) { pgm_set_error("job %d: graphs of more than 60000 nodes are not supported (%d x %d)", k, g1[k].N, g2[k].N); return PGM_ERR_ARG; }
    }
    const bool host_prof = getenv("PGM_HOST_PROFILE") != nullptr;    // stderr: where the host time of this call goes
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms_since = [&](std::chrono::steady_clock::time_point t0) { return std::chrono::duration<double, std::milli>(now() - t0).count(); };
    auto t_begin = now();
    pgm_graph_batch* b = new pgm_graph_batch();
    b->n_states = n; b->n_jobs = n_jobs; b->emission_mode = opts->emission_mode;
    b->lay.resize(n_jobs); b->jobs.resize(n_jobs); b->cell_updates.resize(n_jobs);
    const BatchEnv env;
    std::vector<size_t> scratch_need(n_jobs);
    // ---- pass A (parallel): per-job analysis -> layout decisions and byte sizes --------------------------------------------
    parallel_jobs(n_jobs, [&](int k) {
        const pgm_graph &a = g1[k], &bb = g2[k];
        JobLayout& L = b->lay[k];
        L.np1 = a.e_rowptr[a.N] + (a.r_rowptr ? a.r_rowptr[a.N] : 0);
        L.np2 = bb.e_rowptr[bb.N] + (bb.r_rowptr ? bb.r_rowptr[bb.N] : 0);
        L.cap = 2 * (a.N + bb.N) + 8;
        // column blocks of graph 2 (columns 0 .. N2-2) for the blocked fill kernel: a block never leaves its aligned group
        // of GF_K columns, and a column whose node is not a chain node (exactly one predecessor, the regular edge from
        // x-1) starts a block, so that inside a block every column's only predecessor is its left neighbour
        int nblk = 0, generic1 = 0, generic2 = 0;
        std::vector<int> blk(bb.N, 0);                              // column -> block (same rule as pack_job's colinfo)
        for (int x = 0; x <= bb.N - 2; ++x) {
            const bool nc = x >= 1 && non_chain_node(bb, x);
            generic2 += nc;
            if ((x & (GF_K - 1)) == 0 || nc) ++nblk;
            blk[x] = nblk - 1;
        }
        blk[bb.N - 1] = nblk;
        for (int y = 1; y <= a.N - 2; ++y) generic1 += non_chain_node(a, y);
        const int generic = generic1 + generic2;
        L.nb = nblk;
        const int mp = std::max(max_preds(a), max_preds(bb));      // the packed traceback's index fields hold GF_MAX_PREDS predecessors
        L.dbits = 7;
        // Which fill kernel.  Strip kernel (register-resident, warp per job): graphs whose nodes are all chain nodes (every
        // leaf x leaf job and the profile graphs that have no indel yet) with graph 2 short enough for 32 strips.  Otherwise: the
        // blocked sweep is ~2x faster on chain-dominated graphs; the anti-diagonal sweep pays almost nothing for
        // multi-predecessor nodes and has the shorter critical path on the few large, branchy graphs near the root.  It also
        // takes the graphs whose graph 2 is too long for the blocked kernel.  A batch too small to fill the GPU is latency-bound
        // per job, and there the anti-diagonal sweep is the shorter one even on chain graphs.
        const int cols = bb.N - 1;                                 // DP columns 0 .. N2-2
        int Cpick = 0;
        bool strip_ok = false;
        L.sgen = 0; L.nsrc = 0; L.ring_rows = 0; L.swarps = 1;
        if (env.strip && env.strip != 3 && generic == 0 && a.N >= 3 && bb.N >= 3) {  // chain x chain: warp per job (PGM_FILL_STRIP=3: the general kernel for everything)
            static const int strip_cs[] = {4, 8, 11, 13, 16, 20, 24};
            for (int q : strip_cs) if (q * 32 >= cols) { Cpick = q; break; }
            strip_ok = Cpick > 0;
        }
        // General strip kernel: measured at parity with the dense kernels on jobs with a few non-chain nodes (and its emission
        // is cheaper when one side is a leaf); on branchy graphs the anti-diagonal kernel's per-cell parallelism wins
        // (profiles/): up to strip_max_generic non-chain nodes by default.
        if (!strip_ok && env.strip >= 2 && a.N >= 3 && bb.N >= 3 && mp <= GF_MAX_PREDS && (generic <= env.strip_max_generic || env.strip == 3 || env.force_layout == GF_LAYOUT_STRIP)) {
            // general strip kernel: W warps per job.  Strip width: wide strips (one warp per job) for a batch that fills the GPU,
            // narrow strips (more lanes, shorter steps) when the batch is small and the job's own latency is what counts
            static const int gen_cs[] = {2, 4, 8, 13, 16};
            // strips of the general kernel: a strip starts at column 0, at every non-chain column (so that the special
            // handling sits in slot 0 only) and when the current strip is full
            auto count_strips = [&](int q) {
                int nsq = 0, len = 0;
                for (int x = 0; x < cols; ++x) {
                    if (x == 0 || len == q || non_chain_node(bb, x)) { ++nsq; len = 0; }
                    ++len;
                }
                return nsq;
            };
            double best = 0; int best_ns = 0;
            for (int q : gen_cs) {
                if (env.strip_c && q != env.strip_c) continue;
                if (q == 2 && !env.strip_c) continue;               // 8 warps per 400-column job: only on request
                const int ns_q = count_strips(q), w_q = (ns_q + 31) / 32;
                if (w_q > SG_MAX_WARPS) continue;
                if (q > 8 && Cpick && !env.strip_c) continue;       // wide strips (255 registers, one job per SM) only when the narrow ones do not fit
                // cost model (cycles): steps x (fixed + per-slot); a full GPU prefers the fewest warp-steps, a small batch the shortest sweep
                const double sweep = (double)(a.N - 1 + ns_q) * (500.0 + 70.0 * q);
                // (the switch to "fewest warp-steps" used to sit at 4 jobs per SM; on the 1020-job bench batch the narrow strips
                // are still the faster ones -- C = 4 with 4 warps 1.15 ms against 1.50 ms for C = 8 with 2 warps over its 452
                // jobs, although the latter executes 16 % fewer instructions -- so it now needs 16 jobs per SM)
                const double cost = n_jobs >= 16 * 148 ? sweep * w_q : sweep;
                if (!Cpick || cost < best) { Cpick = q; best = cost; best_ns = ns_q; }
            }
            if (Cpick) {
                L.colsrc.assign(bb.N, -1);
                int nsrc = 0;
                for (int x = 1; x <= bb.N - 2; ++x) {
                    auto visit = [&](int xp) { if (xp != x - 1 && L.colsrc[xp] < 0) L.colsrc[xp] = (int16_t)nsrc++; };
                    for (int e = bb.e_rowptr[x]; e < bb.e_rowptr[x + 1]; ++e) visit(bb.e_col[e]);
                    if (bb.r_rowptr) for (int e = bb.r_rowptr[x]; e < bb.r_rowptr[x + 1]; ++e) visit(bb.r_col[e]);
                }
                // source rings: a reader looks back at most rr - 2 - strips (- SG_SLACK: with several warps per job the owner of a
                // column may be that many steps ahead of the reader); at least the previous row must be there
                int rr = 64; while (rr < best_ns + 4 + (best_ns > 32 ? SG_SLACK : 0)) rr *= 2;
                // history ring: (Y,W) of the last hist_rows rows of every column, for the rows with non-adjacent predecessors
                int hr = 32;                                        // fewer jobs than SMs: the whole shared memory of an SM is the job's
                size_t hist_budget = (n_jobs < 148 ? 128 : 24) * 1024;
                if (env.hist_kb > 0) hist_budget = (size_t)env.hist_kb * 1024;      // PGM_STRIP_HIST_KB (tests)
                while (hr > 4 && (size_t)hr * (size_t)(Cpick + 1) * ((best_ns + 31) & ~31) * 8 > hist_budget) hr /= 2;
                const size_t smem = (size_t)nsrc * rr * 8 + (size_t)hr * (size_t)(Cpick + 1) * ((best_ns + 31) & ~31) * 8;
                if (nsrc <= SG_MAX_SRC && smem <= 200 * 1024) {
                    strip_ok = true; L.sgen = 1; L.nsrc = nsrc; L.ring_rows = rr; L.hist_rows = hr; L.swarps = (best_ns + 31) / 32; L.nstrips_gen = best_ns;
                }
            }
        }
        L.layout = (generic > env.diag_min_generic || nblk > GF_MAX_NB || n_jobs < env.small_batch) ? GF_LAYOUT_DIAG : GF_LAYOUT_BLOCK;
        if (strip_ok) L.layout = GF_LAYOUT_STRIP;
        if (env.force_layout >= 0 && !(env.force_layout == GF_LAYOUT_BLOCK && nblk > GF_MAX_NB) && !(env.force_layout == GF_LAYOUT_STRIP && !strip_ok)) L.layout = env.force_layout;
        if (L.layout != GF_LAYOUT_STRIP) L.sgen = 0;
        L.C = L.layout == GF_LAYOUT_STRIP ? Cpick : 0;
        L.nstrips = L.layout == GF_LAYOUT_STRIP ? (L.sgen ? L.nstrips_gen : (cols + Cpick - 1) / Cpick) : 0;
        // Bands: graphs too long for one CTA.  Short bands for the large branchy graphs (three SMs sharing one job's rows) are
        // wired and tested (PGM_FILL_SPLIT_ROWS) but off: measured 1.73 -> 2.8 ms on the 500-row jobs of the bench batch,
        // the first row of a band pays an L2 round trip per anti-diagonal
        L.band_rows = 0;
        if (L.layout == GF_LAYOUT_DIAG) {
            // band height: one warp per band (no CTA barrier per diagonal) for a short graph against a long profile -- 700-1100 rows against
            // 10,000 columns 55.9 -> 47.8 ms; on a tall graph the pipeline of 100+ bands costs more than the barrier saves (3455 x 1982:
            // 38.7 ms at 128 rows, 42.5 at 32), and squarish graphs lose as well (868 x 1120: 6.8 -> 7.5 ms).  PGM_FILL_BAND_ROWS forces one height.
            if (a.N >= env.band_min_n1) L.band_rows = env.band_rows_forced ? env.long_band_rows : (a.N <= 1100 && bb.N >= 4 * a.N ? 32 : env.long_band_rows);
            else if (env.split_rows > 0 && generic > env.split_min_generic && a.N - 1 >= 2 * env.split_rows) L.band_rows = env.split_rows;
        }
        L.bands = L.band_rows ? (a.N - 1 + L.band_rows - 1) / L.band_rows : 1;
        if (L.bands == 1) L.band_rows = 0;
        L.pitch = L.layout == GF_LAYOUT_DIAG ? (int)align_up((size_t)std::min(a.N, bb.N), 32) : (int)align_up((size_t)L.nb, 8);
        // Role-sorted anti-diagonal kernel (hp1_wave.cu): a thread per row (chain rows and non-chain rows in separate warps) plus
        // wv_L threads per non-chain column, all in one CTA
        L.wave = 0; L.wv_tr = L.wv_chain_pad = L.wv_ncols = L.wv_nx = L.wv_L = L.wv_tcols = 0;
        if (env.wave && L.layout == GF_LAYOUT_DIAG && L.bands == 1 && a.N >= 3 && bb.N >= 3) {
            const int nspec = generic1, nchain = (a.N - 1) - generic1;          // rows 0 .. N1-2; row 0 sweeps with the chain rows
            const int tr = ru32(nchain) + ru32(nspec);
            // column lane groups: a non-chain column is busy for N1 - 1 anti-diagonals, so a group can take column r + T once
            // column r is done -- T = the most non-chain columns in any window of N1 - 1 consecutive columns
            int tcols = 0;
            {
                std::vector<int> xs_;
                for (int x = 1; x <= bb.N - 2; ++x) if (non_chain_node(bb, x)) xs_.push_back(x);
                size_t lo = 0;
                for (size_t hi = 0; hi < xs_.size(); ++hi) {
                    while (xs_[hi] - xs_[lo] >= a.N - 1) ++lo;          // x_hi < x_lo + N1 - 1: still in flight together
                    tcols = std::max(tcols, (int)(hi - lo + 1));
                }
            }
            // Lanes per non-chain column.  Measured level by level on 1000- and 5000-sequence passes (tools/wave_vs_diag_levels.py): with
            // 4 lanes per column and up to ~130 columns in flight the wave kernel wins (0.92 against 1.20 ms ... 1.85 against 2.17 ms); on the
            // deep profiles further up (150-500 non-chain columns: only 1-2 lanes per column fit a CTA) the thread-per-row kernel is as fast
            // or faster (7.5 against 12.6 ms at 688 x 1389 with 134 + 486 non-chain nodes), so those keep it.  PGM_WAVE_L forces a width.
            int lw = 0;
            for (int q : {4, 2, 1}) if ((!env.wave_L || env.wave_L == q) && tr + ru32(tcols * q) <= 1024) { lw = q; break; }
            if (!env.wave_L && (lw != 4 || tcols > 128)) lw = 0;
            int nx = 0;
            bool dup = false;                                       // two list entries from the adjacent predecessor: not handled by the look-ahead split
            for (int w = 0; w < 2; ++w) {
                const pgm_graph& g = w ? bb : a;
                for (int y = 1; y <= g.N - 2; ++y)
                    if (non_chain_node(g, y)) {
                        bool adj = false;
                        auto visit = [&](int from) { if (from == y - 1) { if (adj) dup = true; adj = true; } else if (w == 0) ++nx; };
                        for (int e = g.e_rowptr[y]; e < g.e_rowptr[y + 1]; ++e) visit(g.e_col[e]);
                        if (g.r_rowptr) for (int e = g.r_rowptr[y]; e < g.r_rowptr[y + 1]; ++e) visit(g.r_col[e]);
                    }
            }
            if (lw && !dup && wave_smem_bytes(8, a.N, bb.N, tr, L.np1, nx) <= 227 * 1024) {
                L.wave = 1; L.wv_tr = tr; L.wv_chain_pad = ru32(nchain); L.wv_ncols = generic2; L.wv_nx = nx; L.wv_L = lw; L.wv_tcols = tcols;
            }
        }
        // Graphs of band_min_nowave .. band_min_n1 rows that the wave kernel does not take (too many non-chain columns in flight): bands too.
        // The one-CTA thread-per-row kernel folds a cell's predecessor pairs one at a time; the banded one does it with the warp
        // (629 x 1743 with 118 + 571 non-chain nodes 7.7 -> 6.9 ms, levels of 675-700-row graphs 3.1-3.6 -> 2.9-3.1 ms), while the
        // wave kernel's jobs lose in bands (levels of ~600-row graphs with ~50 + 50 non-chain nodes 1.5-1.9 -> 2.2-2.4 ms).
        if (!L.wave && L.layout == GF_LAYOUT_DIAG && L.band_rows == 0 && a.N >= env.band_min_nowave && a.N < env.band_min_n1 && !env.split_rows) {
            L.band_rows = env.band_rows_forced ? env.long_band_rows : (bb.N >= 4 * a.N ? 32 : env.long_band_rows);
            L.bands = (a.N - 1 + L.band_rows - 1) / L.band_rows;
            if (L.bands == 1) L.band_rows = 0;
        }
        {   // keep sets of the strip kernels (see mark_far_sources); the dense kernels keep everything
            L.keep1.assign(a.N, 0); L.keep2.assign(bb.N, 0);
            if (L.layout == GF_LAYOUT_STRIP) {
                L.keep1[0] = 1; L.keep2[0] = 1;
                for (int e = a.e_rowptr[a.N - 1]; e < a.e_rowptr[a.N]; ++e) L.keep1[a.e_col[e]] = 1;
                if (a.r_rowptr) for (int e = a.r_rowptr[a.N - 1]; e < a.r_rowptr[a.N]; ++e) L.keep1[a.r_col[e]] = 1;
                if (L.sgen) {                                       // general strip kernel: every source of a non-adjacent edge
                    mark_far_sources(a, nullptr, 2, 0, L.keep1);
                    mark_far_sources(bb, nullptr, 2, 0, L.keep2);
                }                                                   // chain strip kernel: nothing is far
            }
        }
        L.nkr = number_kept(L.keep1); L.nkc = number_kept(L.keep2);
        L.ldVr = (int)align_up((size_t)bb.N, 8) + 32; L.ldVc = (int)align_up((size_t)a.N, 4);
        L.nnc1 = L.nnc2 = 0;
        if (L.sgen) {
            L.ncidx1.assign(a.N, -1); L.ncidx2.assign(bb.N, -1);
            for (int y = 1; y <= a.N - 2; ++y) if (non_chain_node(a, y)) L.ncidx1[y] = L.nnc1++;
            for (int x = 1; x <= bb.N - 2; ++x) if (non_chain_node(bb, x)) L.ncidx2[x] = L.nnc2++;
        }
        L.ldRI = (int)align_up((size_t)bb.N, 8); L.ldCI = (int)align_up((size_t)a.N, 8);
        // class-coded emission inputs (strip jobs only: the other kernels stream full rows of S)
        L.coded1 = L.coded2 = 0;
        if (L.layout == GF_LAYOUT_STRIP && env.coding && !t_full_emission && n + 3 < 120) {
            double u1 = 0, u2 = 0;
            if (a.N > n + 3 && site_classes(a, n, L.cls1, u1)) { L.coded1 = 1; L.uniform1 = u1; }
            if (bb.N > n + 3 && site_classes(bb, n, L.cls2, u2)) { L.coded2 = 1; L.uniform2 = u2; }
        }
        L.N1e = L.coded1 ? n + 3 : a.N; L.N2e = L.coded2 ? n + 3 : bb.N;      // pseudo-nodes: 0 (start), 1 .. n+1 (classes 0 .. n), n+2 (end)
        const size_t mxyw_cells = L.layout == GF_LAYOUT_DIAG ? ((size_t)a.N + bb.N - 1) * L.pitch
                                : L.layout == GF_LAYOUT_BLOCK ? ((size_t)(a.N - 1) + L.nb - 1) * GF_K * L.pitch : 0;
        const size_t d_bytes = L.layout == GF_LAYOUT_STRIP ? ((size_t)(a.N - 1) + L.nstrips) * L.nstrips * strip_words(L.C) * 4 : 0;
        auto take = [&](size_t& off, size_t bytes) { size_t r = off; off = align_up(off + bytes, 128); return r; };
        size_t so = 0;
        L.S = take(so, (size_t)L.N1e * (((size_t)L.N2e + 3) & ~(size_t)3) * 4 + 128);
        L.MXYW = take(so, mxyw_cells * 16);
        L.D = take(so, d_bytes + 64);
        L.Vrow = take(so, (size_t)L.nkr * L.ldVr * 16 + 16); L.Vcol = take(so, (size_t)L.nkc * L.ldVc * 16 + 16);
        if (L.sgen) L.dbits = 7;                                    // the side arrays hold 32-bit index words
        L.RowIdx = take(so, (size_t)L.nnc1 * L.ldRI * 4 + 16); L.ColIdx = take(so, (size_t)L.nnc2 * L.ldCI * 4 + 16);
        L.Tbuf = take(so, sizeof(float) * n * L.N2e); L.u = take(so, 4 * L.N1e); L.v = take(so, 4 * L.N2e);
        if (opts->emission_mode == PGM_EMISSION_TENSOR) {
            const size_t kp = (size_t)((3 * n + 7) & ~7);
            L.Aimg = take(so, (size_t)((L.N1e + 127) / 128) * 128 * kp * 4);
            L.Bimg = take(so, (size_t)((L.N2e + 127) / 128) * 128 * kp * 4);
        } else { L.Aimg = L.Bimg = 0; }
        L.alt_score = take(so, 4 * std::max(a.N, bb.N)); L.alt_prev = take(so, 4 * std::max(a.N, bb.N));
        L.progress = take(so, 4 * ((size_t)L.bands + 1));
        scratch_need[k] = so;
    });
    const double ms_a = ms_since(t_begin); auto t_b = now();
    // ---- pass B (serial): arena offsets -------------------------------------------------------------------------------------
    size_t in_off = 0, out_off = 0;
    for (int k = 0; k < n_jobs; ++k) {
        const pgm_graph &a = g1[k], &bb = g2[k];
        JobLayout& L = b->lay[k];
        auto take = [&](size_t& off, size_t bytes) { size_t r = off; off = align_up(off + bytes, 128); return r; };
        L.sites1 = take(in_off, sizeof(float) * n * L.N1e);
        L.sites2 = take(in_off, sizeof(float) * n * L.N2e);
        L.Mf = take(in_off, sizeof(float) * n * n);
        L.pif = take(in_off, sizeof(float) * n);
        L.p1_ptr = take(in_off, 4 * (a.N + 1)); L.p1_idx = take(in_off, 4 * L.np1); L.p1_val = take(in_off, 4 * L.np1);
        L.p1_alt = take(in_off, 4 * L.np1); L.p1_rep = take(in_off, L.np1);
        L.p2_ptr = take(in_off, 4 * (bb.N + 1)); L.p2_idx = take(in_off, 4 * L.np2); L.p2_val = take(in_off, 4 * L.np2);
        L.p2_alt = take(in_off, 4 * L.np2); L.p2_rep = take(in_off, L.np2);
        L.ni1 = take(in_off, sizeof(NodeInfo) * a.N); L.ni2 = take(in_off, sizeof(NodeInfo) * bb.N);
        L.colinfo = take(in_off, 2 * (size_t)bb.N); L.blkstart = take(in_off, 2 * ((size_t)L.nb + 1));
        L.scol = take(in_off, 2 * (size_t)bb.N);
        L.sstart = take(in_off, 4 * ((size_t)L.nstrips + 1)); L.colstrip = take(in_off, L.nstrips ? 4 * (size_t)bb.N : 0);
        L.specmask = take(in_off, L.sgen ? 4 * (size_t)L.nstrips : 0); L.srcmask = take(in_off, L.sgen ? 4 * (size_t)L.nstrips : 0);
        L.colsrc_off = take(in_off, L.sgen ? 2 * (size_t)bb.N : 0);
        L.ncidx1_off = take(in_off, L.sgen ? 4 * (size_t)a.N : 0); L.ncidx2_off = take(in_off, L.sgen ? 4 * (size_t)bb.N : 0);
        L.rinfo_off = take(in_off, L.sgen ? sizeof(RowInfo) * (size_t)a.N : 0);
        L.wv_rows = take(in_off, L.wave ? sizeof(WaveRow) * (size_t)L.wv_tr : 0); L.wv_cols = take(in_off, L.wave ? sizeof(WaveCol) * (size_t)L.wv_ncols : 0);
        L.wv_ci = take(in_off, L.wave ? sizeof(WaveCI) * (size_t)bb.N : 0); L.wv_rowslot = take(in_off, L.wave ? 2 * (size_t)a.N : 0);
        L.wv_xi = take(in_off, L.wave ? 4 * (size_t)L.wv_nx : 0); L.wv_xv = take(in_off, L.wave ? 4 * (size_t)L.wv_nx : 0);
        L.wv_adj1 = take(in_off, L.wave ? 4 * (size_t)a.N : 0);
        L.map1 = take(out_off, 4 * L.cap); L.map2 = take(out_off, 4 * L.cap); L.result = take(out_off, 32);
        b->max_N1 = std::max(b->max_N1, a.N); b->max_N2 = std::max(b->max_N2, bb.N);
    }
    // sub-batches by scratch budget (each job's scratch offsets are relative to its sub-batch)
    const uint64_t budget = pgm_ctx_scratch(c);
    std::vector<size_t> scratch_base(n_jobs);
    {
        int begin = 0; size_t used = 0;
        for (int k = 0; k < n_jobs; ++k) {
            if (k > begin && used + scratch_need[k] > budget) { b->sub.push_back({begin, k}); begin = k; used = 0; }
            scratch_base[k] = used; used += scratch_need[k];
            b->scratch_bytes_used = std::max(b->scratch_bytes_used, used);
        }
        if (n_jobs > begin) b->sub.push_back({begin, n_jobs});
    }
    int rc;
    if ((rc = pgm_ctx_take_host(c, std::max<size_t>(in_off, 1), b->h_in)) || (rc = pgm_ctx_take_dev(c, std::max<size_t>(in_off, 1), b->d_in)) ||
        (rc = pgm_ctx_take_dev(c, std::max<size_t>(out_off, 1), b->d_out)) || (rc = pgm_ctx_take_host(c, std::max<size_t>(out_off, 1), b->h_out)) ||
        (rc = pgm_ctx_take_dev(c, std::max<size_t>(b->scratch_bytes_used, 1), b->d_scratch)) ||
        (rc = pgm_ctx_take_dev(c, std::max<size_t>((sizeof(GraphJobDev) + sizeof(int)) * n_jobs, 1), b->d_jobs))) { pgm_graph_batch_destroy(c, b); return rc; }
    b->out_bytes = out_off;
    const double ms_b = ms_since(t_b); auto t_c = now();
    char* hin = b->h_in.as<char>();
    char* din = b->d_in.as<char>(); char* dsc = b->d_scratch.as<char>(); char* dout = b->d_out.as<char>();
    // ---- pass C (parallel): pack.  Packing a job (double -> float sites, predecessor lists, node infos, DynProgScores) touches
    // only that job's slices of the arenas ----
    auto pack_job = [&](int k) {
        const pgm_graph* gs[2] = {&g1[k], &g2[k]};
        const JobLayout& L = b->lay[k];
        GraphJobDev& J = b->jobs[k];
        memset(&J, 0, sizeof(J));
        J.N1 = g1[k].N; J.N2 = g2[k].N; J.pitch = L.pitch; J.nb = L.nb; J.layout = L.layout; J.bands = L.bands; J.band_rows = L.band_rows; J.cap = L.cap;
        J.N1e = L.N1e; J.N2e = L.N2e; J.coded1 = L.coded1; J.coded2 = L.coded2; J.nstrips = L.nstrips; J.C = L.C;
        J.ldS = (L.N2e + 3) & ~3;
        dynprog_scores(g1[k], g2[k], models[k], *opts, J.sc);
        const float rinit = J.sc[8], rext = J.sc[9];
        for (int w = 0; w < 2; ++w) {                               // emission inputs: the graph's columns, or its n + 1 class columns
            const pgm_graph& g = *gs[w];
            float* s = (float*)(hin + (w == 0 ? L.sites1 : L.sites2));
            if (w == 0 ? L.coded1 : L.coded2) {
                const float uf = (float)(w == 0 ? L.uniform1 : L.uniform2);
                memset(s, 0, sizeof(float) * n * (n + 3));
                for (int a = 0; a < n; ++a) s[(size_t)n * (1 + a) + a] = 1.0f;          // pseudo-node 1 + a: one-hot state a
                for (int a = 0; a < n; ++a) s[(size_t)n * (1 + n) + a] = uf;            // pseudo-node 1 + n: the uniform column
            } else {
                for (size_t i = 0; i < (size_t)n * g.N; ++i) s[i] = (float)g.sites[i];
            }
        }
        float* Mf = (float*)(hin + L.Mf); float* pif = (float*)(hin + L.pif);
        for (int i = 0; i < n * n; ++i) Mf[i] = (float)models[k].M[i];
        for (int i = 0; i < n; ++i) pif[i] = (float)models[k].pi[i];
        const size_t ptr_off[2] = {L.p1_ptr, L.p2_ptr}, idx_off[2] = {L.p1_idx, L.p2_idx}, val_off[2] = {L.p1_val, L.p2_val},
                     alt_off[2] = {L.p1_alt, L.p2_alt}, rep_off[2] = {L.p1_rep, L.p2_rep};
        uint64_t deg_sum[2] = {0, 0};
        for (int w = 0; w < 2; ++w) {
            const pgm_graph& g = *gs[w];
            int* ptr = (int*)(hin + ptr_off[w]); int* idx = (int*)(hin + idx_off[w]);
            float* val = (float*)(hin + val_off[w]); float* alt = (float*)(hin + alt_off[w]);
            uint8_t* rep = (uint8_t*)(hin + rep_off[w]);
            int q = 0;
            for (int node = 0; node < g.N; ++node) {
                ptr[node] = q;
                for (int e = g.e_rowptr[node]; e < g.e_rowptr[node + 1]; ++e) {       // Graph.h:223-231
                    const float cst = g.e_val[e];
                    const float v = cst == 0 ? FINF : cst + 10000.0f;
                    idx[q] = g.e_col[e]; val[q] = v; alt[q] = v; rep[q] = 0; ++q;
                }
                if (g.r_rowptr)
                    for (int e = g.r_rowptr[node]; e < g.r_rowptr[node + 1]; ++e) {   // Graph.h:232-238
                        const uint32_t u = g.r_units[e];
                        float v, va;
                        if (u == 0) { v = FINF; va = FINF; }
                        else {
                            const float um1 = (float)(u - 1);
                            const float mm = rext * um1; v = rinit + mm;
                            const float ma = FINF * um1; va = FINF + ma;      // markAlternativePath: (inf, inf)
                        }
                        idx[q] = g.r_col[e]; val[q] = v; alt[q] = va; rep[q] = 1; ++q;
                    }
                if (node >= 1 && node <= g.N - 2) deg_sum[w] += (uint64_t)(q - ptr[node]);
            }
            ptr[g.N] = q;
            NodeInfo* ni = (NodeInfo*)(hin + (w == 0 ? L.ni1 : L.ni2));
            const bool coded = w == 0 ? L.coded1 : L.coded2;
            const std::vector<uint8_t>& cls = w == 0 ? L.cls1 : L.cls2;
            for (int node = 0; node < g.N; ++node) {
                const int cnt = ptr[node + 1] - ptr[node];
                const int kept = (w == 0 ? L.keep1 : L.keep2)[node];                         // value-store row / column, or -1
                const int sclass = coded ? 1 + (int)cls[node] : 0;                           // S row / column of a class-coded node (pseudo-node index)
                ni[node].ptr = ptr[node];
                ni[node].cnt = cnt | ((kept + 1) << 8) | (sclass << 24) | ((cnt && rep[ptr[node]]) ? (int)0x80000000 : 0);   // NI_CNT / NI_KEEP / NI_CLASS / first pred is a repeat edge
                ni[node].idx0 = cnt ? idx[ptr[node]] : 0; ni[node].val0 = cnt ? val[ptr[node]] : 0.f;
            }
        }
        {   // column blocks of graph 2, from the node infos just built (same rule as the analysis pass)
            const NodeInfo* ni = (const NodeInfo*)(hin + L.ni2);
            uint16_t* ci = (uint16_t*)(hin + L.colinfo); uint16_t* bs = (uint16_t*)(hin + L.blkstart);
            uint16_t* sc = (uint16_t*)(hin + L.scol);
            int blk = -1;
            for (int x = 0; x <= g2[k].N - 2; ++x) {
                const bool chain = x == 0 || NI_CHAIN(ni[x], x);
                if ((x & (GF_K - 1)) == 0 || !chain) { ++blk; bs[blk] = (uint16_t)x; }
                ci[x] = (uint16_t)((blk << GF_LOGK) | (x & (GF_K - 1)));
            }
            bs[blk + 1] = (uint16_t)(g2[k].N - 1);
            ci[g2[k].N - 1] = 0;
            for (int x = 0; x < g2[k].N; ++x) sc[x] = (uint16_t)(L.coded2 ? NI_CLASS(ni[x].cnt) : x);
        }
        if (L.nstrips) {                                            // strips of the strip kernels
            int* ss = (int*)(hin + L.sstart); uint32_t* cs = (uint32_t*)(hin + L.colstrip);
            const int cols = g2[k].N - 1;
            if (!L.sgen) {                                          // chain kernel: C columns each
                for (int j = 0; j <= L.nstrips; ++j) ss[j] = std::min(j * L.C, cols);
                for (int x = 0; x < g2[k].N; ++x) cs[x] = (uint32_t)((x / L.C) * 32 + x % L.C);
            } else {                                                // general kernel: a non-chain column starts a strip
                int j = -1, len = 0;
                for (int x = 0; x < cols; ++x) {
                    if (x == 0 || len == L.C || L.ncidx2[x] >= 0) { ++j; ss[j] = x; len = 0; }
                    cs[x] = (uint32_t)(j * 32 + len);
                    ++len;
                }
                ss[j + 1] = cols; cs[cols] = 0;
            }
        }
        if (L.sgen) {
            uint32_t* sm = (uint32_t*)(hin + L.specmask); uint32_t* rm = (uint32_t*)(hin + L.srcmask);
            const int* ss = (const int*)(hin + L.sstart);
            for (int j = 0; j < L.nstrips; ++j) {
                uint32_t a_ = 0, b_ = 0;
                for (int x = ss[j]; x < ss[j + 1]; ++x) {
                    if (L.ncidx2[x] >= 0) a_ |= 1u << (x - ss[j]);
                    if (L.colsrc[x] >= 0) b_ |= 1u << (x - ss[j]);
                }
                sm[j] = a_; rm[j] = b_;
            }
            memcpy(hin + L.colsrc_off, L.colsrc.data(), 2 * (size_t)g2[k].N);
            memcpy(hin + L.ncidx1_off, L.ncidx1.data(), 4 * (size_t)g1[k].N);
            memcpy(hin + L.ncidx2_off, L.ncidx2.data(), 4 * (size_t)g2[k].N);
            const NodeInfo* ni = (const NodeInfo*)(hin + L.ni1);
            const int* pidx = (const int*)(hin + L.p1_idx); const float* pval = (const float*)(hin + L.p1_val);
            RowInfo* ri = (RowInfo*)(hin + L.rinfo_off);
            for (int y = 0; y < g1[k].N; ++y) {
                const int cnt = NI_CNT(ni[y].cnt);
                RowInfo r; r.cnt = ni[y].cnt; r.idx0 = ni[y].idx0; r.val0 = ni[y].val0; r.ptr = ni[y].ptr; r.ncidx = L.ncidx1[y];
                r.idx1 = cnt > 1 ? pidx[ni[y].ptr + 1] : 0; r.val1 = cnt > 1 ? pval[ni[y].ptr + 1] : 0.f;
                const int k0 = cnt > 0 ? L.keep1[r.idx0] + 1 : 0, k1 = cnt > 1 ? L.keep1[r.idx1] + 1 : 0;
                r.kq01 = k0 | (k1 << 16);
                ri[y] = r;
            }
        }
        if (L.wave) {                                               // role-sorted anti-diagonal kernel: thread records
            const int N1 = g1[k].N, N2 = g2[k].N;
            const NodeInfo* ni1 = (const NodeInfo*)(hin + L.ni1); const NodeInfo* ni2 = (const NodeInfo*)(hin + L.ni2);
            const int* p1i = (const int*)(hin + L.p1_idx); const float* p1v = (const float*)(hin + L.p1_val);
            const int* p2i = (const int*)(hin + L.p2_idx); const float* p2v = (const float*)(hin + L.p2_val);
            WaveRow* wr = (WaveRow*)(hin + L.wv_rows); WaveCol* wc = (WaveCol*)(hin + L.wv_cols); WaveCI* ci = (WaveCI*)(hin + L.wv_ci);
            uint16_t* slot = (uint16_t*)(hin + L.wv_rowslot); int* xi = (int*)(hin + L.wv_xi); float* xv = (float*)(hin + L.wv_xv);
            float* adj1 = (float*)(hin + L.wv_adj1);
            adj1[0] = FINF; adj1[N1 - 1] = FINF;
            for (int t = 0; t < L.wv_tr; ++t) wr[t] = WaveRow{-1, 0, 0.f, 0, 0.f, 0, 0.f, 0};
            int tc = 0, ts = L.wv_chain_pad, nx = 0;
            slot[N1 - 1] = 0;
            for (int y = 0; y <= N1 - 2; ++y) {
                const bool chain = y == 0 || NI_CHAIN(ni1[y], y);
                WaveRow r{y, 0, FINF, 0, 0.f, 0, 0.f, nx};                  // row 0: no predecessor row, Y and M come out as -inf
                if (y >= 1 && chain) r.ey_adj = ni1[y].val0;
                if (!chain) {
                    r.ey_adj = FINF;                                // no edge from y-1: its candidates come out as -inf
                    bool adj = false;
                    const int cnt = NI_CNT(ni1[y].cnt);
                    for (int q = 0; q < cnt; ++q) {
                        const int yp = p1i[ni1[y].ptr + q]; const float ey = p1v[ni1[y].ptr + q];
                        if (yp == y - 1 && !adj) { adj = true; r.ey_adj = ey; continue; }
                        if (r.nextra == 0) { r.yp0 = yp; r.ey0 = ey; } else if (r.nextra == 1) { r.yp1 = yp; r.ey1 = ey; }
                        xi[nx] = yp; xv[nx] = ey; ++nx; ++r.nextra;
                    }
                }
                const int t = chain ? tc++ : ts++;
                wr[t] = r; slot[y] = (uint16_t)t;
                if (y >= 1) adj1[y] = r.ey_adj;
            }
            int nc = 0;
            for (int x = 0; x < N2; ++x) {
                const bool spec = x >= 1 && x <= N2 - 2 && !NI_CHAIN(ni2[x], x);
                const bool lspec = x >= 2 && x <= N2 - 2 && !NI_CHAIN(ni2[x - 1], x - 1);
                ci[x].ex_adj = (x >= 1 && x <= N2 - 2 && !spec) ? ni2[x].val0 : 0.f;
                ci[x].flags = (spec ? 1 : 0) | (lspec ? 2 : 0);
                if (spec) {
                    const int cnt = NI_CNT(ni2[x].cnt), ptr = ni2[x].ptr;
                    WaveCol c{x, cnt, 0, 0.f, 0, 0.f, ptr, FINF};
                    if (cnt > 0) { c.xp0 = p2i[ptr]; c.ex0 = p2v[ptr]; }
                    if (cnt > 1) { c.xp1 = p2i[ptr + 1]; c.ex1 = p2v[ptr + 1]; }
                    for (int q = 0; q < cnt; ++q) if (p2i[ptr + q] == x - 1) c.ex_adj_col = p2v[ptr + q];
                    wc[nc++] = c;
                }
            }
        }
        J.wave = L.wave; J.wv_tr = L.wv_tr; J.wv_ncols = L.wv_ncols; J.wv_nx = L.wv_nx; J.wv_chain_pad = L.wv_chain_pad; J.wv_tcols = L.wv_tcols;
        J.wv_adj1 = (const float*)(din + L.wv_adj1);
        J.wv_rows = (const WaveRow*)(din + L.wv_rows); J.wv_cols = (const WaveCol*)(din + L.wv_cols); J.wv_ci = (const WaveCI*)(din + L.wv_ci);
        J.wv_rowslot = (const uint16_t*)(din + L.wv_rowslot); J.wv_xi = (const int*)(din + L.wv_xi); J.wv_xv = (const float*)(din + L.wv_xv);
        J.dbg = env.dbg; J.sgen = L.sgen; J.nsrc = L.nsrc; J.ring_rows = L.ring_rows; J.hist_rows = L.hist_rows; J.ldRI = L.ldRI; J.ldCI = L.ldCI;
        J.specmask = (const uint32_t*)(din + L.specmask); J.srcmask = (const uint32_t*)(din + L.srcmask);
        J.colsrc = (const int16_t*)(din + L.colsrc_off); J.ncidx1 = (const int*)(din + L.ncidx1_off); J.ncidx2 = (const int*)(din + L.ncidx2_off);
        J.rinfo = (const RowInfo*)(din + L.rinfo_off);
        J.RowIdx = (void*)(dsc + scratch_base[k] + L.RowIdx); J.ColIdx = (void*)(dsc + scratch_base[k] + L.ColIdx);
        b->cell_updates[k] = deg_sum[0] * deg_sum[1];
        char* sb = dsc + scratch_base[k];
        J.sites1 = (const float*)(din + L.sites1); J.sites2 = (const float*)(din + L.sites2);
        J.Mf = (const float*)(din + L.Mf); J.pif = (const float*)(din + L.pif);
        J.p1_ptr = (const int*)(din + L.p1_ptr); J.p1_idx = (const int*)(din + L.p1_idx); J.p1_val = (const float*)(din + L.p1_val);
        J.p1_rep = (const uint8_t*)(din + L.p1_rep); J.p1_alt = (const float*)(din + L.p1_alt);
        J.p2_ptr = (const int*)(din + L.p2_ptr); J.p2_idx = (const int*)(din + L.p2_idx); J.p2_val = (const float*)(din + L.p2_val);
        J.p2_rep = (const uint8_t*)(din + L.p2_rep); J.p2_alt = (const float*)(din + L.p2_alt);
        J.ni1 = (const NodeInfo*)(din + L.ni1); J.ni2 = (const NodeInfo*)(din + L.ni2);
        J.colinfo = (const uint16_t*)(din + L.colinfo); J.blkstart = (const uint16_t*)(din + L.blkstart);
        J.scol = (const uint16_t*)(din + L.scol);
        J.sstart = (const int*)(din + L.sstart); J.colstrip = (const uint32_t*)(din + L.colstrip);
        J.np1 = L.np1; J.np2 = L.np2;
        J.S = (float*)(sb + L.S); J.MXYW = (float4*)(sb + L.MXYW); J.D = (void*)(sb + L.D); J.dbits = L.dbits; J.Tbuf = (float*)(sb + L.Tbuf);
        J.Vrow = (float4*)(sb + L.Vrow); J.Vcol = (float4*)(sb + L.Vcol); J.ldVr = L.ldVr; J.ldVc = L.ldVc;
        J.u = (float*)(sb + L.u); J.v = (float*)(sb + L.v);
        J.Aimg = (float*)(sb + L.Aimg); J.Bimg = (float*)(sb + L.Bimg);
        J.alt_score = (float*)(sb + L.alt_score); J.alt_prev = (int*)(sb + L.alt_prev); J.progress = (int*)(sb + L.progress);
        J.map1 = (int*)(dout + L.map1); J.map2 = (int*)(dout + L.map2); J.result = (int*)(dout + L.result);
    };
    // The workers pack job after job (their slices of the pinned arena lie in job order); the calling thread meanwhile sends
    // every finished prefix to the device in pieces of >= 8 MB, so that the upload overlaps the packing instead of following it
    cudaError_t up_err = cudaSuccess;
    {
        unsigned hw = std::thread::hardware_concurrency();
        const int nthreads = (int)std::max(1u, std::min<unsigned>(std::min<unsigned>(hw ? hw : 1u, 16u), (unsigned)(n_jobs / 8 + 1)));
        if (nthreads <= 1 || in_off < ((size_t)16 << 20)) {
            parallel_jobs(n_jobs, pack_job);
            if (n_jobs) up_err = cudaMemcpyAsync(b->d_in.p, b->h_in.p, in_off, cudaMemcpyHostToDevice, st);
        } else {
            std::unique_ptr<std::atomic<unsigned char>[]> done(new std::atomic<unsigned char>[n_jobs]);
            for (int k = 0; k < n_jobs; ++k) done[k].store(0, std::memory_order_relaxed);
            std::atomic<int> next(0);
            std::vector<std::thread> pool;
            for (int t = 0; t < nthreads; ++t)
                pool.emplace_back([&] { for (int k = next.fetch_add(1); k < n_jobs; k = next.fetch_add(1)) { pack_job(k); done[k].store(1, std::memory_order_release); } });
            int ready = 0; size_t sent = 0;
            while (sent < in_off && up_err == cudaSuccess) {
                while (ready < n_jobs && done[ready].load(std::memory_order_acquire)) ++ready;
                const size_t upto = ready == n_jobs ? in_off : b->lay[ready].sites1;     // job k's slices start at its sites1
                if (upto > sent && (upto - sent >= ((size_t)8 << 20) || ready == n_jobs)) {
                    up_err = cudaMemcpyAsync(din + sent, hin + sent, upto - sent, cudaMemcpyHostToDevice, st);
                    sent = upto;
                } else std::this_thread::yield();
            }
            for (auto& th : pool) th.join();
        }
    }
    const double ms_c = ms_since(t_c); auto t_d = now();
    for (int k = 0; k < n_jobs; ++k) { b->total_updates += b->cell_updates[k]; b->total_cells += (uint64_t)b->jobs[k].N1 * b->jobs[k].N2; }
    {   // fill schedule: per kernel flavour, block-size classes (blocked: column blocks per row rounded up to warps;
        // diagonal: graph-1 rows; strip: columns per strip); longest sweeps first
        b->order.resize(n_jobs);
        static const int dcls[] = {192, 320, 448, 544, 672, 832, 1024, 1 << 30};
        auto class_of = [&](const GraphJobDev& J) {
            if (J.layout == GF_LAYOUT_STRIP) return J.sgen ? 400 + J.C + 32 * ((J.nstrips + 31) / 32) : 300 + J.C;
            if (J.wave) {      // by graph-1 length and by block size (a launch's block is its largest job's: a job with many non-chain
                               // columns must not make every other job of its launch carry hundreds of idle threads through the barriers)
                int q = 0; while (J.N1 - 1 > dcls[q]) ++q;
                const int lw = b->lay[&J - b->jobs.data()].wv_L, thr = J.wv_tr + ru32(J.wv_tcols * lw);
                return 500 + 64 * lw + 16 * (thr <= 512 ? 0 : (thr <= 768 ? 1 : 2)) + q;
            }
            if (J.bands > 1) return 200 + J.band_rows / 32;                   // banded anti-diagonal sweeps: their own launches, one per band height
            if (J.layout == GF_LAYOUT_DIAG) { int q = 0; while (J.N1 - 1 > dcls[q]) ++q; return 100 + q; }
            const int w = (J.nb + 31) / 32; return w <= 4 ? w : (w <= 6 ? 6 : (w <= 8 ? 8 : 16));
        };
        for (size_t si = 0; si < b->sub.size(); ++si) {
            const auto& sbr = b->sub[si];
            std::vector<int> idx;
            for (int k = sbr.first; k < sbr.second; ++k) idx.push_back(k - sbr.first);
            // classes ordered by their longest sweep (estimated, microseconds), longest first: earlier launches get the higher
            // stream priority, so the few long latency-bound sweeps near the root start at once and are not starved of issue slots
            auto est_us = [&](const GraphJobDev& J) -> double {
                if (J.layout == GF_LAYOUT_STRIP) return (double)(J.N1 + J.nstrips) * (J.sgen ? 1.2 : 0.6);
                if (J.wave) return (double)(J.N1 + J.N2) * 1.0;
                if (J.layout == GF_LAYOUT_DIAG) return (double)(J.N1 + J.N2) * (J.bands > 1 ? 8.0 : 1.8);
                return (double)(J.N1 + J.nb) * 1.2;
            };
            std::map<int, double> class_est;
            for (int k = sbr.first; k < sbr.second; ++k) { double& e = class_est[class_of(b->jobs[k])]; e = std::max(e, est_us(b->jobs[k])); }
            std::stable_sort(idx.begin(), idx.end(), [&](int a, int bq) {
                const GraphJobDev &A = b->jobs[sbr.first + a], &B = b->jobs[sbr.first + bq];
                const int ca = class_of(A), cb = class_of(B);
                if (ca != cb) { const double ea = class_est[ca], eb = class_est[cb]; return ea != eb ? ea > eb : ca > cb; }
                return est_us(A) > est_us(B);
            });
            std::copy(idx.begin(), idx.end(), b->order.begin() + sbr.first);
            size_t p = 0;
            while (p < idx.size()) {
                const GraphJobDev& J0 = b->jobs[sbr.first + idx[p]];
                const int q = class_of(J0);
                pgm_graph_batch::FillLaunch f{(int)si, (int)p, 0, J0.layout, 0, 0, 0, 0, 0, 1, J0.band_rows, J0.dbits, J0.C, 0, 0, J0.sgen, (J0.nstrips + 31) / 32, 0};
                f.wave = J0.wave; f.wL = b->lay[sbr.first + idx[p]].wv_L;
                while (p < idx.size() && class_of(b->jobs[sbr.first + idx[p]]) == q) {
                    const GraphJobDev& J = b->jobs[sbr.first + idx[p]];
                    if (J.wave && f.count > 0) {                    // the launch's shared memory is sized by the maxima over its jobs: every job fits
                        // on its own (checked in the analysis pass), the combination may not -- then the job starts the next launch
                        const size_t s8 = wave_smem_bytes(8, std::max(f.max_n1, J.N1), std::max(f.max_n2, J.N2), std::max(f.max_tr, J.wv_tr),
                                                          std::max(f.max_np1, J.np1), std::max(f.max_nx, J.wv_nx));
                        if (s8 > 227 * 1024) break;
                    }
                    f.max_nb = std::max(f.max_nb, J.nb); f.max_n1 = std::max(f.max_n1, J.N1); f.max_n2 = std::max(f.max_n2, J.N2);
                    f.max_np2 = std::max(f.max_np2, J.np2); f.max_np = std::max(f.max_np, J.np1 + J.np2);
                    f.max_bands = std::max(f.max_bands, J.bands);
                    f.max_n1e = std::max(f.max_n1e, J.N1e); f.max_n2e = std::max(f.max_n2e, J.N2e);
                    f.ring_bytes = std::max(f.ring_bytes, (size_t)J.nsrc * J.ring_rows * 8 + (size_t)J.hist_rows * (size_t)(J.C + 1) * ((J.nstrips + 31) & ~31) * 8);
                    if (J.wave) {
                        f.wthreads = std::max(f.wthreads, J.wv_tr + ru32(J.wv_tcols * f.wL)); f.max_tr = std::max(f.max_tr, J.wv_tr);
                        f.max_np1 = std::max(f.max_np1, J.np1); f.max_nx = std::max(f.max_nx, J.wv_nx);
                    }
                    ++f.count; ++p;
                }
                if (f.wave) {       // ring depth: 16 anti-diagonals when that does not cost resident CTAs the launch can use
                    const size_t s8 = wave_smem_bytes(8, f.max_n1, f.max_n2, f.max_tr, f.max_np1, f.max_nx), s16 = wave_smem_bytes(16, f.max_n1, f.max_n2, f.max_tr, f.max_np1, f.max_nx);
                    const int by_thr = std::max(1, 2048 / std::max(f.wthreads, 32));
                    const int c8 = std::min<int>(by_thr, (int)((227 * 1024) / (s8 + 1024))), c16 = std::min<int>(by_thr, (int)((227 * 1024) / (s16 + 1024)));
                    // (a batch with more jobs than SMs: the other classes' CTAs want to share the SM, so the smaller ring)
                    f.wring = (s16 <= 227 * 1024 && (c16 >= c8 || (n_jobs <= 148 && f.count <= 148 * std::max(c16, 1)))) ? 16 : 8;
                    if (env.wave_ring == 8 || (env.wave_ring == 16 && s16 <= 227 * 1024)) f.wring = env.wave_ring;
                }
                b->fills.push_back(f);
            }
        }
    }
    const double ms_d = ms_since(t_d); auto t_e = now();
    if (n_jobs) {
        cudaError_t e = up_err;
        if (e == cudaSuccess) e = cudaMemcpyAsync(b->d_jobs.as<char>() + sizeof(GraphJobDev) * n_jobs, b->order.data(), sizeof(int) * n_jobs, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(b->d_jobs.p, b->jobs.data(), sizeof(GraphJobDev) * n_jobs, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
            pgm_set_error("CUDA error %s while uploading the graph batch: %s", cudaGetErrorName(e), cudaGetErrorString(e));
            pgm_graph_batch_destroy(c, b);
            return PGM_ERR_CUDA;
        }
    }
    if (host_prof)
        fprintf(stderr, "[host] create: analysis %.2f ms, offsets + arenas %.2f ms, pack %.2f ms, schedule %.2f ms, upload %.2f ms (%.1f MB in, %.1f MB scratch)\n",
                ms_a, ms_b, ms_c, ms_d, ms_since(t_e), in_off / 1e6, b->scratch_bytes_used / 1e6);
    *out = b;
    return 0;
}

// Schedule of one sub-batch.  Every size class of the fill schedule is an independent pipeline emission -> fill ->
// traceback on its own stream, so the few long sweeps near the root start at once and overlap the bulk's emission and
// traceback instead of every phase waiting for the slowest job of the previous one.  PGM_GRAPH_PIPELINE=0 runs the three
// phases one after the other over the whole sub-batch (what bench.py uses for the per-kernel breakdown).
int pgm_graph_batch_run(pgm_ctx* c, pgm_graph_batch* b) {
    if (!c || !b) { pgm_set_error("NULL argument"); return PGM_ERR_ARG; }
    PGM_CUDA(cudaSetDevice(pgm_ctx_device(c)));
    cudaStream_t st = pgm_ctx_stream(c);
    pgm_ctx_reset_timing(c);
    const char* env = getenv("PGM_GRAPH_PIPELINE");
    const bool pipelined = !(env && atoi(env) == 0);
    const int pipe_mode = env ? atoi(env) : 1;
    const bool skip_trace = getenv("PGM_SKIP_TRACE") != nullptr;       // timing experiments on fill kernels whose results are wrong on purpose
    const GraphJobDev* dj = b->d_jobs.as<GraphJobDev>();
    const int* d_order_all = (const int*)(b->d_jobs.as<char>() + sizeof(GraphJobDev) * b->n_jobs);
    // status words (emission failure flag, end-node pick) start from zero on every run
    if (b->out_bytes) PGM_CUDA(cudaMemsetAsync(b->d_out.p, 0, b->out_bytes, st));
    for (size_t si = 0; si < b->sub.size(); ++si) {
        const auto& sb = b->sub[si];
        const int nj = sb.second - sb.first;
        int mx1 = 0, mx2 = 0;
        for (int k = sb.first; k < sb.second; ++k) { mx1 = std::max(mx1, b->jobs[k].N1e); mx2 = std::max(mx2, b->jobs[k].N2e); }
        std::vector<const pgm_graph_batch::FillLaunch*> fl;
        for (auto& f : b->fills) if (f.sub == (int)si) fl.push_back(&f);
        int rc = 0; void* tok;
        auto emit = [&](cudaStream_t sq, int count, const int* ord, int m1, int m2) -> int {
            const int rc_e = b->emission_mode == PGM_EMISSION_TENSOR ? launch_emission_tensor(sq, b->n_states, count, dj + sb.first, ord, m1, m2)
                                                                     : launch_emission_exact(sq, b->n_states, count, dj + sb.first, ord, m1, m2);
            if (rc_e) return rc_e;
            // fault injection for the tests: PGM_INJECT_EMISSION_FAIL=<job> marks that job's emission as failed, exactly as
            // the tensor kernel does when a bounded mbarrier wait times out (result[4]); the job must then fail, not align
            if (const char* e = getenv("PGM_INJECT_EMISSION_FAIL")) {
                const int kf = atoi(e);
                if (kf >= sb.first && kf < sb.second) {
                    static const int err = PGM_ERR_CUDA;
                    PGM_CUDA(cudaMemcpyAsync(b->jobs[kf].result + 4, &err, sizeof(int), cudaMemcpyHostToDevice, sq));
                }
            }
            return 0;
        };
        auto fill = [&](cudaStream_t sq, const pgm_graph_batch::FillLaunch* f) -> int {
            const int* ord = d_order_all + sb.first + f->offset;
            if (f->layout == GF_LAYOUT_STRIP)
                return f->sgen ? launch_graph_fill_strip_gen(sq, f->C, f->dbits, f->swarps, f->ring_bytes, f->count, dj + sb.first, ord)
                               : launch_graph_fill_strip(sq, f->C, f->count, dj + sb.first, ord);
            if (f->wave) return launch_graph_fill_wave(sq, f->wL, f->wring, f->wthreads, f->count, dj + sb.first, ord, f->max_n1, f->max_n2, f->max_tr, f->max_np1, f->max_nx);
            if (f->max_bands > 1) {
                for (int jq = 0; jq < f->count; ++jq) {             // zero the bands' progress flags and the band ticket
                    const GraphJobDev& Jq = b->jobs[sb.first + b->order[sb.first + f->offset + jq]];
                    PGM_CUDA(cudaMemsetAsync(Jq.progress, 0, sizeof(int) * (Jq.bands + 1), sq));
                }
                return launch_graph_fill_banded(sq, f->count, dj + sb.first, ord, f->max_bands, 148, f->band_rows, f->max_n2, f->max_np);
            }
            return f->layout == GF_LAYOUT_DIAG ? launch_graph_fill_diag(sq, f->count, dj + sb.first, ord, f->max_n1, f->max_n2, f->max_np)
                                               : launch_graph_fill(sq, f->count, dj + sb.first, ord, f->max_nb, f->max_n2, f->max_np2);
        };
        // side streams: the first class stays on the main stream, the others fork from it and are joined at the end
        SideStreams& ss = side_streams(pgm_ctx_device(c));
        const size_t n_side = fl.size() > 1 ? std::min<size_t>(fl.size(), 16) : 0;
        auto fork = [&]() -> int {
            if (!n_side) return 0;
            if (!ss.fork) PGM_CUDA(cudaEventCreateWithFlags(&ss.fork, cudaEventDisableTiming));
            while (ss.streams.size() < n_side) {
                // classes are ordered longest sweep first: earlier side streams get the higher priority, so the long sweeps' CTAs
                // are placed before the bulk's whenever an SM frees up
                int lo = 0, hi = 0; cudaDeviceGetStreamPriorityRange(&lo, &hi);
                const int prio = std::min(lo, hi + (int)ss.streams.size());
                cudaStream_t s2; PGM_CUDA(cudaStreamCreateWithPriority(&s2, cudaStreamNonBlocking, prio)); ss.streams.push_back(s2);
                cudaEvent_t e2; PGM_CUDA(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming)); ss.join.push_back(e2);
            }
            PGM_CUDA(cudaEventRecord(ss.fork, st));
            for (size_t q = 0; q < n_side; ++q) PGM_CUDA(cudaStreamWaitEvent(ss.streams[q], ss.fork, 0));
            return 0;
        };
        auto join = [&]() -> int {
            for (size_t q = 0; q < n_side; ++q) {
                PGM_CUDA(cudaEventRecord(ss.join[q], ss.streams[q]));
                PGM_CUDA(cudaStreamWaitEvent(st, ss.join[q], 0));
            }
            return 0;
        };
        auto stream_of = [&](size_t q) { return !n_side ? st : ss.streams[q % n_side]; };
        if (pipe_mode >= 2) {
            // emissions of every class first, back to back on the main stream (throughput kernels: they get the whole GPU instead
            // of the SMs the long sweeps leave over); then one fill -> traceback pipeline per class.  Mode 3: a class's pipeline
            // starts as soon as ITS emission is done
            std::vector<cudaEvent_t>& evs = ss.emit_done;
            while (pipe_mode == 3 && evs.size() < fl.size()) { cudaEvent_t e2; PGM_CUDA(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming)); evs.push_back(e2); }
            if (pipe_mode == 3 && (rc = fork())) return rc;               // (the side streams then wait for their own class's emission only)
            for (size_t q = 0; q < fl.size() && !rc; ++q) {
                const int* ord = d_order_all + sb.first + fl[q]->offset;
                pgm_ctx_time_begin_on(c, 3, st, &tok); rc = emit(st, fl[q]->count, ord, fl[q]->max_n1e, fl[q]->max_n2e); pgm_ctx_time_end(c, tok);
                if (!rc && pipe_mode == 3) PGM_CUDA(cudaEventRecord(evs[q], st));
            }
            if (rc) return rc;
            if (pipe_mode != 3 && (rc = fork())) return rc;
            for (size_t q = 0; q < fl.size() && !rc; ++q) {
                cudaStream_t sq = stream_of(q);
                const int* ord = d_order_all + sb.first + fl[q]->offset;
                if (pipe_mode == 3 && sq != st) PGM_CUDA(cudaStreamWaitEvent(sq, evs[q], 0));
                pgm_ctx_time_begin_on(c, 4, sq, &tok); rc = fill(sq, fl[q]); pgm_ctx_time_end(c, tok);
                if (rc) break;
                pgm_ctx_time_begin_on(c, 5, sq, &tok);
                if (skip_trace) { pgm_ctx_time_end(c, tok); continue; }
                rc = fl[q]->layout == GF_LAYOUT_STRIP ? launch_graph_trace_packed(sq, fl[q]->count, dj + sb.first, ord) : launch_graph_trace(sq, fl[q]->count, dj + sb.first, ord);
                pgm_ctx_time_end(c, tok);
            }
            if (rc) return rc;
            if ((rc = join())) return rc;
        } else if (pipelined) {
            if ((rc = fork())) return rc;
            // The emission of the BULK classes (more jobs than SMs) goes out first: these are short throughput kernels, and launched
            // behind the long sweeps they only get the SMs those leave over, which delays the bulk's own fill (measured on the bench
            // batch: 2.71-2.80 ms against 2.90-2.95 ms; PGM_NO_BULK_FIRST=1 for the A/B run)
            std::vector<char> emitted(fl.size(), 0);
            const bool bulk_first = !getenv("PGM_NO_BULK_FIRST");
            for (size_t q = 0; q < fl.size() && !rc && bulk_first; ++q)
                if (fl[q]->count > 148) {
                    cudaStream_t sq = stream_of(q);
                    const int* ord = d_order_all + sb.first + fl[q]->offset;
                    pgm_ctx_time_begin_on(c, 3, sq, &tok); rc = emit(sq, fl[q]->count, ord, fl[q]->max_n1e, fl[q]->max_n2e); pgm_ctx_time_end(c, tok);
                    emitted[q] = 1;
                }
            for (size_t q = 0; q < fl.size() && !rc; ++q) {
                cudaStream_t sq = stream_of(q);
                const int* ord = d_order_all + sb.first + fl[q]->offset;
                if (!emitted[q]) { pgm_ctx_time_begin_on(c, 3, sq, &tok); rc = emit(sq, fl[q]->count, ord, fl[q]->max_n1e, fl[q]->max_n2e); pgm_ctx_time_end(c, tok); }
                if (rc) break;
                pgm_ctx_time_begin_on(c, 4, sq, &tok); rc = fill(sq, fl[q]); pgm_ctx_time_end(c, tok);
                if (rc) break;
                pgm_ctx_time_begin_on(c, 5, sq, &tok);
                if (skip_trace) { pgm_ctx_time_end(c, tok); continue; }
                rc = fl[q]->layout == GF_LAYOUT_STRIP ? launch_graph_trace_packed(sq, fl[q]->count, dj + sb.first, ord) : launch_graph_trace(sq, fl[q]->count, dj + sb.first, ord);
                pgm_ctx_time_end(c, tok);
            }
            if (rc) return rc;
            if ((rc = join())) return rc;
        } else {
            pgm_ctx_time_begin(c, 3, &tok);
            rc = emit(st, nj, nullptr, mx1, mx2);
            pgm_ctx_time_end(c, tok);
            if (rc) return rc;
            pgm_ctx_time_begin(c, 4, &tok);
            if (!(rc = fork())) {
                for (size_t q = 0; q < fl.size() && !rc; ++q) rc = fill(stream_of(q), fl[q]);
                if (!rc) rc = join();
            }
            pgm_ctx_time_end(c, tok);
            if (rc) return rc;
            pgm_ctx_time_begin(c, 5, &tok);
            if (!skip_trace) rc = launch_graph_trace(st, nj, dj + sb.first, nullptr);        // each traceback kernel skips the other layouts' jobs
            if (!rc && !skip_trace) rc = launch_graph_trace_packed(st, nj, dj + sb.first, nullptr);
            pgm_ctx_time_end(c, tok);
            if (rc) return rc;
        }
    }
    pgm_timing* t = pgm_ctx_timing(c);
    t->graph_cell_updates += b->total_updates;
    t->graph_cells += b->total_cells;
    return 0;
}

int pgm_graph_batch_fetch(pgm_ctx* c, pgm_graph_batch* b, pgm_alignment* out) {
    if (!c || !b || (b->n_jobs && !out)) { pgm_set_error("NULL argument"); return PGM_ERR_ARG; }
    PGM_CUDA(cudaSetDevice(pgm_ctx_device(c)));
    cudaStream_t st = pgm_ctx_stream(c);
    if (b->n_jobs == 0) return 0;
    auto t0f = std::chrono::steady_clock::now();
    PGM_CUDA(cudaMemcpyAsync(b->h_out.p, b->d_out.p, b->out_bytes, cudaMemcpyDeviceToHost, st));
    PGM_CUDA(cudaStreamSynchronize(st));
    const double ms_d2h = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0f).count();
    const char* ho = b->h_out.as<char>();
    int worst = 0;
    parallel_jobs(b->n_jobs, [&](int k) {                        // every job writes its own pgm_alignment and mapping buffers
        const JobLayout& L = b->lay[k];
        const int* res = (const int*)(ho + L.result);
        pgm_alignment& A = out[k];
        A.status = res[0]; A.length = res[1]; A.n_tr_indels = (uint32_t)res[2]; A.score = ((const float*)res)[3];
        A.cell_updates = b->cell_updates[k];
        if (A.status == 0 && A.length > A.capacity) A.status = PGM_ERR_CAPACITY;
        if (A.status == 0) {
            const int* m1 = (const int*)(ho + L.map1); const int* m2 = (const int*)(ho + L.map2);
            for (int i = 0; i < A.length; ++i) {                 // device wrote end -> start (std::reverse, GraphAlign.h:512-513)
                A.mapping1[i] = m1[A.length - 1 - i];
                A.mapping2[i] = m2[A.length - 1 - i];
            }
        }
    });
    for (int k = 0; k < b->n_jobs; ++k) {
        const pgm_alignment& A = out[k];
        if (A.status != 0 && worst == 0) {
            worst = A.status;
            pgm_set_error(A.status == PGM_ERR_BACKTRACK ? "job %d: backtracking failed"
                          : (A.status == PGM_ERR_CAPACITY ? "job %d: mapping capacity too small" : "job %d: the emission kernel failed (tensor pipeline timeout)"), k);
        }
    }
    if (getenv("PGM_HOST_PROFILE"))
        fprintf(stderr, "[host] fetch: wait + D2H %.2f ms (%.1f MB), unpack %.2f ms\n", ms_d2h, b->out_bytes / 1e6,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0f).count() - ms_d2h);
    return worst;
}

void pgm_graph_batch_destroy(pgm_ctx* c, pgm_graph_batch* b) {
    if (!b) return;
    if (c) { cudaSetDevice(pgm_ctx_device(c)); cudaStreamSynchronize(pgm_ctx_stream(c)); }
    pgm_ctx_give_dev(c, b->d_in); pgm_ctx_give_dev(c, b->d_scratch); pgm_ctx_give_dev(c, b->d_out); pgm_ctx_give_dev(c, b->d_jobs);
    pgm_ctx_give_host(c, b->h_in); pgm_ctx_give_host(c, b->h_out);
    delete b;
}

int pgm_align_graphs_batch(pgm_ctx* c, int n, int n_jobs, const pgm_graph* g1, const pgm_graph* g2,
                           const pgm_dp_model* models, const pgm_dp_opts* opts, pgm_alignment* out) {
    pgm_graph_batch* b = nullptr;
    int rc = pgm_graph_batch_create(c, n, n_jobs, g1, g2, models, opts, &b);
    if (rc) return rc;
    rc = pgm_graph_batch_run(c, b);
    if (!rc) rc = pgm_graph_batch_fetch(c, b, out);
    pgm_graph_batch_destroy(c, b);
    return rc;
}

int pgm_emission_scores(pgm_ctx* c, int n, const pgm_graph* g1, const pgm_graph* g2, const pgm_dp_model* model,
                        float match_init, int emission_mode, float* S) {
    if (!c || !g1 || !g2 || !model || !S) { pgm_set_error("NULL argument"); return PGM_ERR_ARG; }
    pgm_dp_opts o{}; o.indel_rate = 0.0093359375; o.end_indel_prob = 0.12; o.repeat_rate = 0.1; o.repeatext_prob = 0.3;
    o.emission_mode = emission_mode;
    pgm_graph_batch* b = nullptr;
    t_full_emission = true;
    int rc = pgm_graph_batch_create(c, n, 1, g1, g2, model, &o, &b);
    t_full_emission = false;
    if (rc) return rc;
    cudaStream_t st = pgm_ctx_stream(c);
    b->jobs[0].sc[2] = match_init;
    cudaMemcpyAsync(b->d_jobs.p, b->jobs.data(), sizeof(GraphJobDev), cudaMemcpyHostToDevice, st);
    if (emission_mode == PGM_EMISSION_TENSOR) rc = launch_emission_tensor(st, n, 1, b->d_jobs.as<GraphJobDev>(), nullptr, g1->N, g2->N);
    else rc = launch_emission_exact(st, n, 1, b->d_jobs.as<GraphJobDev>(), nullptr, g1->N, g2->N);
    if (!rc) {
        const GraphJobDev& J = b->jobs[0];
        std::vector<float> sk((size_t)J.N1 * J.ldS);
        cudaError_t e = cudaMemcpyAsync(sk.data(), J.S, sk.size() * 4, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { pgm_set_error("CUDA error: %s", cudaGetErrorString(e)); rc = PGM_ERR_CUDA; }
        else
            for (int x = 0; x < J.N2; ++x)
                for (int y = 0; y < J.N1; ++y) S[y + (size_t)J.N1 * x] = sk[(size_t)y * J.ldS + x];
    }
    pgm_graph_batch_destroy(c, b);
    return rc;
}

} // extern "C"
