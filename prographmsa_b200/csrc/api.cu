// api.cu -- C-ABI glue (include/pgm_b200.h): context, error reporting, timing, and the host-side
// scheduling of HP-2 (pair descriptors, length-binned chunks sized to the scratch budget).
#include "internal.h"
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <cstdlib>

static thread_local char g_err[1024] = "";
void pgm_set_error(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int nw_pick_C(int max_L1);

struct TimedSection { cudaEvent_t a = nullptr, b = nullptr; int kind = 0; };

struct pgm_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    uint64_t scratch_bytes = 16ull << 30;          // of 180 GB HBM; pgm_set_scratch_bytes changes it
    bool timing_on = false;
    bool cached_symmetric = false, ml_no_merge = false;   // PGM_ML_NO_MERGE=1: keep (i,j) and (j,i) counts separate (A-B runs)
    bool force_nw32 = false;             // PGM_NW_FORCE_32BIT=1: never use the 16-bit forward flavour (tests / A-B runs)
    pgm_timing timing{};
    std::vector<TimedSection> sections;       // pending (resolved at sync)
    std::vector<cudaEvent_t> event_pool;
    // model cache
    DevBuf d_model; std::vector<double> cachedQ; int cached_n = 0;
    // HP-2 scratch
    std::vector<NwTask> h_tasks, h_ftasks; std::vector<int8_t> h_tab8;
    DevBuf d_off, d_codes, d_descs, d_tasks, d_ftasks, d_tab, d_tb, d_colbuf, d_list, d_list_len, d_gaps, d_dist, d_var,
           d_counts, d_scores, d_misc, d_rows, d_seqlen;
    DevBuf d_full;                                  // two N x N matrices (Prealigned all-pairs result, scattered on the device)
    DevBuf d_pre_rows, d_pre_counts, d_pre_gaps;    // Prealigned: pitched rows, one chunk's uint16 counts and gap openings
    DevBuf d_tb2, d_list2;                          // second traceback / list set: chunk k+1's forward overlaps chunk k's walk + ML
    cudaStream_t stream2 = nullptr;                 // walk + ML of the chunk pipeline
    cudaEvent_t ev_fwd[2] = {nullptr, nullptr}, ev_wm[2] = {nullptr, nullptr};
    HostPinned h_stage;
    // buffer pool for HP-1 batches: cudaMalloc / cudaMallocHost of GB-sized arenas per call would dominate e2e
    std::vector<DevBuf> pool_dev; std::vector<HostPinned> pool_host;
};

// ---- timing ----------------------------------------------------------------------------------------
enum { K_NW_FWD = 0, K_NW_WALK, K_ML, K_EMIT, K_GFILL, K_GTRACE };

static cudaEvent_t get_event(pgm_ctx* c) {
    if (!c->event_pool.empty()) { cudaEvent_t e = c->event_pool.back(); c->event_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
struct ScopedKernelTimer {
    pgm_ctx* c; TimedSection s; cudaStream_t st;
    ScopedKernelTimer(pgm_ctx* ctx, int kind, cudaStream_t on = nullptr) : c(ctx), st(on ? on : ctx->stream) {
        s.kind = kind;
        if (c->timing_on) { s.a = get_event(c); s.b = get_event(c); cudaEventRecord(s.a, st); }
    }
    ~ScopedKernelTimer() {
        if (c->timing_on) { cudaEventRecord(s.b, st); c->sections.push_back(s); }
    }
};
static void resolve_timing(pgm_ctx* c) {
    const bool timeline = getenv("PGM_TIMELINE") != nullptr && !c->sections.empty();     // profiling aid: when each timed section began and ended
    for (auto& s : c->sections) {
        float ms = 0; cudaEventElapsedTime(&ms, s.a, s.b);
        if (timeline) {
            float t0 = 0; cudaEventElapsedTime(&t0, c->sections.front().a, s.a);
            static const char* names[] = {"nw_forward", "nw_walk", "ml", "emission", "fill", "trace"};
            fprintf(stderr, "[timeline] %-10s start %8.3f ms  duration %8.3f ms\n", names[s.kind], t0, ms);
        }
        switch (s.kind) {
            case K_NW_FWD: c->timing.nw_forward_ms += ms; c->timing.nw_forward_launches++; break;
            case K_NW_WALK: c->timing.nw_walk_ms += ms; c->timing.nw_walk_launches++; break;
            case K_ML: c->timing.ml_ms += ms; c->timing.ml_launches++; break;
            case K_EMIT: c->timing.emission_ms += ms; c->timing.emission_launches++; break;
            case K_GFILL: c->timing.graph_fill_ms += ms; c->timing.graph_fill_launches++; break;
            case K_GTRACE: c->timing.graph_trace_ms += ms; c->timing.graph_trace_launches++; break;
        }
        c->event_pool.push_back(s.a); c->event_pool.push_back(s.b);
    }
    c->sections.clear();
}

extern "C" {

const char* pgm_last_error(void) { return g_err; }
int pgm_abi_version(void) { return 1; }

int pgm_create(int device, pgm_ctx** out) {
    if (!out) { pgm_set_error("pgm_create: out is NULL"); return PGM_ERR_ARG; }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        pgm_set_error("no CUDA device available (%s); libpgm_b200 has no CPU fallback", cudaGetErrorString(e));
        cudaGetLastError();
        return PGM_ERR_CUDA;
    }
    if (device < 0 || device >= count) { pgm_set_error("device %d out of range (0..%d)", device, count - 1); return PGM_ERR_ARG; }
    PGM_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop; PGM_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        pgm_set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
        return PGM_ERR_CUDA;
    }
    pgm_ctx* c = new pgm_ctx();
    c->device = device;
    { const char* e = getenv("PGM_NW_FORCE_32BIT"); c->force_nw32 = e && e[0] == '1'; }
    { const char* e = getenv("PGM_ML_NO_MERGE"); c->ml_no_merge = e && e[0] == '1'; }
    PGM_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    *out = c;
    return 0;
}

void pgm_destroy(pgm_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    resolve_timing(c);
    for (DevBuf* b : {&c->d_model, &c->d_off, &c->d_ftasks, &c->d_codes, &c->d_descs, &c->d_tasks, &c->d_tab, &c->d_tb, &c->d_colbuf, &c->d_list,
                      &c->d_list_len, &c->d_gaps, &c->d_dist, &c->d_var, &c->d_counts, &c->d_scores, &c->d_misc, &c->d_rows,
                      &c->d_seqlen, &c->d_tb2, &c->d_list2, &c->d_pre_rows, &c->d_pre_counts, &c->d_pre_gaps, &c->d_full}) b->release();
    if (c->stream2) cudaStreamDestroy(c->stream2);
    for (int k = 0; k < 2; ++k) { if (c->ev_fwd[k]) cudaEventDestroy(c->ev_fwd[k]); if (c->ev_wm[k]) cudaEventDestroy(c->ev_wm[k]); }
    c->h_stage.release();
    for (auto& b : c->pool_dev) b.release();
    for (auto& b : c->pool_host) b.release();
    for (auto e : c->event_pool) cudaEventDestroy(e);
    cudaStreamDestroy(c->stream);
    delete c;
}

int pgm_sync(pgm_ctx* c) {
    PGM_CUDA(cudaSetDevice(c->device));
    PGM_CUDA(cudaStreamSynchronize(c->stream));
    resolve_timing(c);
    return 0;
}

int pgm_set_scratch_bytes(pgm_ctx* c, uint64_t bytes) {
    if (bytes < (64ull << 20)) { pgm_set_error("scratch budget must be >= 64 MiB"); return PGM_ERR_ARG; }
    c->scratch_bytes = bytes; return 0;
}
void* pgm_stream(pgm_ctx* c) { return (void*)c->stream; }
int pgm_enable_timing(pgm_ctx* c, int on) { c->timing_on = on != 0; return 0; }
int pgm_get_timing(pgm_ctx* c, pgm_timing* out) {
    int rc = pgm_sync(c); if (rc) return rc;
    *out = c->timing; return 0;
}

} // extern "C"

static void reset_timing(pgm_ctx* c) { resolve_timing(c); memset(&c->timing, 0, sizeof(c->timing)); }

// ---- model upload -----------------------------------------------------------------------------------
static int prepare_model(pgm_ctx* c, const pgm_ml_params* prm, MlDev& ml) {
    const int n = prm->n_states;
    if (n < 2 || n > 63) { pgm_set_error("n_states %d out of range (2..63)", n); return PGM_ERR_ARG; }
    const size_t nn = (size_t)n * n;
    const bool need_basis = prm->mldist_flag || prm->mldist_gap_flag;
    if (need_basis && (!prm->Q || !prm->pi)) { pgm_set_error("Q / pi are required for ML distances"); return PGM_ERR_ARG; }
    int rc = c->d_model.reserve(sizeof(double) * (2 * nn + n));
    if (rc) return rc;
    if (need_basis && !(c->cached_n == n && c->cachedQ.size() == nn && memcmp(c->cachedQ.data(), prm->Q, nn * 8) == 0)) {
        std::vector<double> buf(2 * nn + n);
        rc = pgm_eigen_basis(prm->Q, n, buf.data(), buf.data() + 2 * nn, buf.data() + nn);
        if (rc) return rc;
        // is P_ij(t) / P_ji(t) independent of t?  (true for reversible Q; ECM is only reversible to ~1e-6)
        {
            const double* V = buf.data(); const double* Vi = buf.data() + nn;
            double amax = 0, dev = 0;
            for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) for (int k = 0; k < n; ++k) amax = std::max(amax, std::fabs(V[i * n + k] * Vi[k * n + j]));
            for (int i = 0; i < n && dev <= 1e-11 * amax; ++i)
                for (int j = i + 1; j < n; ++j) {
                    double num = 0, den = 0;                         // least-squares ratio c_ij over k
                    for (int k = 0; k < n; ++k) { const double a = V[i * n + k] * Vi[k * n + j], b = V[j * n + k] * Vi[k * n + i]; num += a * b; den += b * b; }
                    const double cij = den > 0 ? num / den : 0;
                    for (int k = 0; k < n; ++k) dev = std::max(dev, std::fabs(V[i * n + k] * Vi[k * n + j] - cij * V[j * n + k] * Vi[k * n + i]));
                    if (!(cij > 0)) dev = amax;
                }
            c->cached_symmetric = dev <= 1e-11 * amax;
        }
        PGM_CUDA(cudaMemcpyAsync(c->d_model.p, buf.data(), buf.size() * 8, cudaMemcpyHostToDevice, c->stream));
        PGM_CUDA(cudaStreamSynchronize(c->stream));
        c->cachedQ.assign(prm->Q, prm->Q + nn); c->cached_n = n;
    }
    ml.n = n;
    ml.V = c->d_model.as<double>(); ml.Vi = ml.V + nn; ml.sigma = ml.V + 2 * nn;
    ml.mldist_flag = prm->mldist_flag; ml.mldist_gap_flag = prm->mldist_gap_flag;
    ml.symmetric = (need_basis && c->cached_symmetric && !c->ml_no_merge) ? 1 : 0;
    ml.indel_rate = prm->indel_rate; ml.min_dist = prm->min_dist; ml.max_dist = prm->max_dist;
    ml.cutoff_dist = prm->cutoff_dist; ml.DIST_MAX = prm->DIST_MAX; ml.VAR_MAX = prm->VAR_MAX; ml.VAR_MIN = prm->VAR_MIN;
    return 0;
}

static inline void pair_from_index_host(int64_t p, int64_t N, int64_t& i, int64_t& j) {
    const double b = 2.0 * N - 1.0;
    int64_t ii = (int64_t)std::floor((b - std::sqrt(b * b - 8.0 * (double)p)) * 0.5);
    if (ii < 0) ii = 0;
    while (ii > 0 && ii * N - ii * (ii + 1) / 2 > p) --ii;
    while ((ii + 1) * N - (ii + 1) * (ii + 2) / 2 <= p) ++ii;
    i = ii; j = p - (ii * N - ii * (ii + 1) / 2) + ii + 1;
}

// ---- HP-2a: NW + ML over a pair range, device-resident buffers -------------------------------------------
static int nw_ml_range_dev(pgm_ctx* c, const pgm_ml_params* prm, const uint8_t* d_codes, const uint32_t* offsets, int N,
                           const int32_t* score_tab, int go, int ge, int64_t pb, int64_t pe,
                           double* d_dist, double* d_var, int32_t* d_counts, uint32_t* d_gaps_user, int32_t* d_scores) {
    PGM_CUDA(cudaSetDevice(c->device));
    const int64_t total_pairs = (int64_t)N * (N - 1) / 2;
    if (N < 0 || pb < 0 || pe < pb || pe > total_pairs) { pgm_set_error("pair range [%lld,%lld) invalid for N=%d", (long long)pb, (long long)pe, N); return PGM_ERR_ARG; }
    if (!prm || !offsets || !score_tab) { pgm_set_error("NULL argument"); return PGM_ERR_ARG; }
    const int64_t np = pe - pb;
    if (np == 0) return 0;
    if (np > 0x7fffffffll) { pgm_set_error("pair range too large for one call (%lld); shard it", (long long)np); return PGM_ERR_ARG; }
    MlDev ml; int rc = prepare_model(c, prm, ml); if (rc) return rc;
    const int n = prm->n_states, n1 = n + 1;

    // score table -> int8 (reference tables are within [-4, 11])
    std::vector<int8_t>& tab8 = c->h_tab8; tab8.assign((size_t)n1 * n1, 0);
    for (size_t i = 0; i < tab8.size(); ++i) {
        if (score_tab[i] < -32 || score_tab[i] > 31) { pgm_set_error("score table entry %d outside [-32, 31] (the profile stores 4*score+3 as int8)", score_tab[i]); return PGM_ERR_ARG; }
        tab8[i] = (int8_t)score_tab[i];
    }
    rc = c->d_tab.reserve(tab8.size()); if (rc) return rc;
    PGM_CUDA(cudaMemcpyAsync(c->d_tab.p, tab8.data(), tab8.size(), cudaMemcpyHostToDevice, c->stream));

    // per-call outputs that live across chunks
    rc = c->d_list_len.reserve((size_t)np * 4); if (rc) return rc;
    uint32_t* d_gaps = d_gaps_user;
    if (!d_gaps) { rc = c->d_gaps.reserve((size_t)np * 4); if (rc) return rc; d_gaps = c->d_gaps.as<uint32_t>(); }
    rc = c->d_misc.reserve(64); if (rc) return rc;
    PGM_CUDA(cudaMemsetAsync(c->d_misc.p, 0, 64, c->stream));
    int* d_err = c->d_misc.as<int>();
    unsigned long long* d_evals = (unsigned long long*)(c->d_misc.as<char>() + 16);

    int max_len = 0;
    for (int k = 0; k < N; ++k) max_len = std::max<int>(max_len, (int)(offsets[k + 1] - offsets[k]));
    if (max_len > 60000) { pgm_set_error("sequence length %d exceeds the supported 60000", max_len); return PGM_ERR_ARG; }
    const int max_bands = std::max(1, (max_len + 32 * 16 - 1) / (32 * 16));   // only C = 16 is ever multi-band
    const uint32_t colstride = (uint32_t)((max_len + 2 + 31) & ~31);

    // offsets on the device: pairs are derived from them inside the kernels (no per-pair descriptors)
    rc = c->d_off.reserve((size_t)(N + 1) * 4); if (rc) return rc;
    PGM_CUDA(cudaMemcpyAsync(c->d_off.p, offsets, (size_t)(N + 1) * 4, cudaMemcpyHostToDevice, c->stream));

    // 16-bit flavour constants (hp2_nw.cu): potential slope H, bias, and the value bound that decides eligibility
    int smax = 0, smin = 0;
    for (size_t i = 0; i < tab8.size(); ++i) { smax = std::max(smax, (int)tab8[i]); smin = std::min(smin, (int)tab8[i]); }
    int u16_h = std::max(std::max(-(4 * go + 1), -4 * ge), 0); u16_h = (u16_h + 3) & ~3;
    const int u16_bias = 16 * (std::abs(go) + std::abs(ge)) + 64;
    const bool u16_table_ok = (4 * smin + 3 + 2 * u16_h >= 0) && go <= 0 && ge <= 0 && !c->force_nw32;
    // tasks = up to 32 consecutive pairs of one row; chunks = runs of tasks that fit the traceback budget
    // Chunks can also be the unit of a forward | walk + ML pipeline (two traceback sets, two streams; PGM_NW_PIPELINE=<chunks>).
    // It is wired, bit-identical and OFF by default: on the 500 x 400 matrix 2 / 4 / 8 chunks took 8.2 / 8.9 / 9.9 ms against
    // 7.5 ms in one chunk -- the walk's warps take issue slots from the issue-bound forward kernel and its random reads
    // queue behind the forward kernel's traceback stream (walk 0.7 -> 2.8-3.9 ms when overlapped).
    uint64_t tb_budget_words = std::max<uint64_t>(c->scratch_bytes / 4, 1u << 20);
    bool pipeline_chunks = false;
    {
        double est_cells = 0;                                            // pairs * (mean length + 1)^2
        const double mean_len = N > 0 ? (double)offsets[N] / N : 0.0;
        est_cells = (double)np * (mean_len + 1.0) * (mean_len + 1.0);
        const double est_tb_bytes = est_cells * 0.36;                    // packed traceback, 16-bit flavour: ~0.34 B per cell
        const char* e = getenv("PGM_NW_PIPELINE");
        const int want = (e && atoi(e) > 1) ? atoi(e) : 1;
        if (want >= 2) {
            pipeline_chunks = true;
            const uint64_t per = (uint64_t)(est_tb_bytes / want / 4.0) + (1u << 20);
            tb_budget_words = std::max<uint64_t>(std::min<uint64_t>(per, c->scratch_bytes / 8), 1u << 20);   // two sets share the budget
        }
    }
    std::vector<NwTask>& tasks = c->h_tasks;
    std::vector<size_t> chunk_end;                      // task index where each chunk ends
    std::vector<uint64_t> chunk_tb, chunk_list, chunk_cells;
    tasks.clear();
    // small pair ranges (a shard of a strongly-scaled matrix): smaller tasks keep >= ~2 waves of forward CTAs, and the walk
    // runs 8 instead of 32 independent chains per warp (a warp-step waits for the slowest lane's DRAM access)
    const int64_t task_pairs = np >= 38000 ? NW_TASK_PAIRS : (np >= 19000 ? 16 : 8);
    int walk_lanes = np >= 70000 ? 32 : 8;
    if (const char* e = getenv("PGM_WALK_LANES")) { const int v = atoi(e); if (v == 8 || v == 32) walk_lanes = v; }
    int u16_max_rows = 0;                               // longest row sequence of any 16-bit task (sizes the residue staging)
    {
        int64_t i, j; pair_from_index_host(pb, N, i, j);
        int64_t p = pb;
        uint64_t tb_words = 0, list_entries = 0, cells = 0;
        auto len_of = [&](int64_t k) -> uint32_t { return offsets[k + 1] - offsets[k]; };
        // can rows with at most Lcol columns run the 16-bit flavour against the row sequences [j0, j1)?
        // shortcut: if even the longest sequence against the longest sequence fits 16 bits, no per-block scan is needed
        uint32_t g_lmax = 0, g_lmin = 0xffffffffu;
        for (int k = 0; k < N; ++k) { g_lmax = std::max(g_lmax, len_of(k)); g_lmin = std::min(g_lmin, len_of(k)); }
        const bool u16_all = u16_table_ok && g_lmin >= 1 && g_lmax <= 512 &&
                             (uint64_t)4 * smax * g_lmax + (uint64_t)u16_h * 2 * g_lmax + u16_bias + 4 * smax + 2 * u16_h + 64 < 65536;
        auto u16_ok = [&](uint32_t Lcol_min, uint32_t Lcol_max, int64_t j0, int64_t j1, uint32_t& l2max) -> bool {
            if (u16_all) { l2max = g_lmax; return true; }
            l2max = 0;
            if (!u16_table_ok || Lcol_min < 1 || Lcol_max > 512) return false;
            uint32_t l2min = 0xffffffffu;
            for (int64_t jq = j0; jq < j1; ++jq) { const uint32_t l = len_of(jq); l2max = std::max(l2max, l); l2min = std::min(l2min, l); }
            const uint64_t bound = (uint64_t)4 * smax * std::min<uint32_t>(Lcol_max, l2max) + (uint64_t)u16_h * (Lcol_max + l2max) + u16_bias +
                                   4 * smax + 2 * u16_h + 64;
            return l2min >= 1 && bound < 65536;
        };
        auto close_chunk_if_full = [&](uint64_t w, uint64_t span) {
            if (tb_words > 0 && (tb_words + w > tb_budget_words || list_entries + span > 0xf0000000ull)) {
                chunk_end.push_back(tasks.size()); chunk_tb.push_back(tb_words); chunk_list.push_back(list_entries);
                chunk_cells.push_back(cells);
                tb_words = 0; list_entries = 0; cells = 0;
            }
        };
        auto tb_need = [&](uint32_t L1, uint32_t Cflags, int64_t j0, int64_t j1) -> uint64_t {
            const uint32_t C = Cflags & 0xffu, nb = (L1 + 32u * C - 1u) / (32u * C);
            return nw_tb_words((Cflags & NW_MODE_U16) != 0, nb, offsets[j1] - offsets[j0], (uint64_t)(j1 - j0));
        };
        auto push_task = [&](int64_t row, int64_t j0, int64_t j1, int64_t out_base, uint32_t Cflags) {
            const uint32_t L1 = len_of(row);
            const uint64_t span = offsets[j1] - offsets[j0];
            NwTask t; t.i = (uint32_t)row; t.j0 = (uint32_t)j0; t.j1 = (uint32_t)j1; t.out_base = (uint32_t)out_base;
            t.tb_base = tb_words; t.list_base = (uint32_t)list_entries; t.C = Cflags;
            tasks.push_back(t);
            tb_words += tb_need(L1, Cflags, j0, j1); list_entries += span;
            cells += (uint64_t)(L1 + 1) * (span + (uint64_t)(j1 - j0));          // sum (L1+1)(L2+1)
        };
        while (p < pe) {
            const int64_t row_end = std::min<int64_t>(N, j + (pe - p));          // exclusive j bound within this row
            const uint32_t L1 = len_of(i);
            // twin rows: row i runs to its end and row i+1 is covered completely -> both rows share their j blocks
            const bool twin_rows = row_end == N && i + 2 < N && j <= i + 2 && p + (N - j) + (N - i - 2) <= pe;
            if (twin_rows) {
                const uint32_t L1b = len_of(i + 1);
                const int64_t pA = p - pb, pB = pA + (N - j);                    // packed index of (i, j) and of (i+1, i+2)
                for (int64_t j0 = j; j0 < N; j0 += task_pairs) {
                    const int64_t j1 = std::min<int64_t>(N, j0 + task_pairs), jb0 = std::max<int64_t>(j0, i + 2);
                    uint32_t l2max;
                    if (jb0 < j1 && u16_ok(std::min(L1, L1b), std::max(L1, L1b), j0, j1, l2max)) {
                        const uint32_t C = (uint32_t)nw_pick_C((int)std::max(L1, L1b));
                        close_chunk_if_full(tb_need(L1, C | NW_MODE_U16, j0, j1) + tb_need(L1b, C | NW_MODE_U16, jb0, j1), offsets[j1] - offsets[j0] + offsets[j1] - offsets[jb0]);
                        push_task(i, j0, j1, pA + (j0 - j), C | NW_MODE_U16 | NW_TWIN_A);
                        push_task(i + 1, jb0, j1, pB + (jb0 - (i + 2)), C | NW_MODE_U16 | NW_TWIN_B);
                        u16_max_rows = std::max<int>(u16_max_rows, (int)l2max);
                    } else {
                        for (int h = 0; h < 2; ++h) {
                            const int64_t row = i + h, ja = h ? jb0 : j0;
                            if (ja >= j1) continue;
                            const uint32_t Lr = len_of(row);
                            uint32_t C = (uint32_t)nw_pick_C((int)Lr);
                            if (u16_ok(Lr, Lr, ja, j1, l2max)) { C |= NW_MODE_U16; u16_max_rows = std::max<int>(u16_max_rows, (int)l2max); }
                            close_chunk_if_full(tb_need(Lr, C, ja, j1), offsets[j1] - offsets[ja]);
                            push_task(row, ja, j1, h ? pB + (ja - (i + 2)) : pA + (ja - j), C);
                        }
                    }
                }
                p += (N - j) + (N - i - 2);
                i += 2; j = i + 1;
                continue;
            }
            for (int64_t j0 = j; j0 < row_end; j0 += task_pairs) {
                const int64_t j1 = std::min<int64_t>(row_end, j0 + task_pairs);
                uint32_t C = (uint32_t)nw_pick_C((int)L1), l2max;              // per row: keeps all 32 lanes busy
                if (u16_ok(L1, L1, j0, j1, l2max)) { C |= NW_MODE_U16; u16_max_rows = std::max<int>(u16_max_rows, (int)l2max); }
                close_chunk_if_full(tb_need(L1, C, j0, j1), offsets[j1] - offsets[j0]);
                push_task(i, j0, j1, p - pb, C);
                p += j1 - j0;
            }
            ++i; j = i + 1;
        }
        chunk_end.push_back(tasks.size()); chunk_tb.push_back(tb_words); chunk_list.push_back(list_entries);
        chunk_cells.push_back(cells);
    }
    uint64_t max_tb = 1, max_list = 1; size_t max_tasks = 1, prev = 0;
    for (size_t k = 0; k < chunk_end.size(); ++k) {
        max_tb = std::max(max_tb, chunk_tb[k]); max_list = std::max(max_list, chunk_list[k]);
        max_tasks = std::max(max_tasks, chunk_end[k] - prev); prev = chunk_end[k];
    }
    if ((rc = c->d_tasks.reserve(tasks.size() * sizeof(NwTask)))) return rc;

    if ((rc = c->d_tb.reserve(max_tb * 4))) return rc;
    if ((rc = c->d_list.reserve(max_list * 2))) return rc;
    pipeline_chunks = pipeline_chunks && chunk_end.size() >= 2;
    if (pipeline_chunks) {
        if ((rc = c->d_tb2.reserve(max_tb * 4))) return rc;
        if ((rc = c->d_list2.reserve(max_list * 2))) return rc;
        if (!c->stream2) {
            PGM_CUDA(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
            for (int k = 0; k < 2; ++k) {
                PGM_CUDA(cudaEventCreateWithFlags(&c->ev_fwd[k], cudaEventDisableTiming));
                PGM_CUDA(cudaEventCreateWithFlags(&c->ev_wm[k], cudaEventDisableTiming));
            }
        }
    }
    if ((rc = c->d_colbuf.reserve(max_bands > 1 ? max_tasks * 8 * 2 * colstride * 4 : 256))) return rc;
    PGM_CUDA(cudaMemcpyAsync(c->d_tasks.p, tasks.data(), tasks.size() * sizeof(NwTask), cudaMemcpyHostToDevice, c->stream));
    size_t t0 = 0;
    for (size_t k = 0; k < chunk_end.size(); ++k) {
        const int ntasks = (int)(chunk_end[k] - t0);
        const NwTask* dt = c->d_tasks.as<NwTask>() + t0;
        const int set = pipeline_chunks ? (int)(k & 1) : 0;
        uint32_t* tb = (set ? c->d_tb2 : c->d_tb).as<uint32_t>();
        uint16_t* list = (set ? c->d_list2 : c->d_list).as<uint16_t>();
        cudaStream_t s_wm = pipeline_chunks ? c->stream2 : c->stream;
        if (pipeline_chunks && k >= 2) PGM_CUDA(cudaStreamWaitEvent(c->stream, c->ev_wm[set], 0));   // the set's previous walk + ML are done
        {
            ScopedKernelTimer t(c, K_NW_FWD);
            rc = launch_nw_forward(c->stream, ntasks, n, max_bands, d_codes, c->d_off.as<uint32_t>(), dt, c->d_tab.as<int8_t>(),
                                   go, ge, tb, c->d_colbuf.as<int32_t>(), colstride, d_scores, u16_h, u16_bias,
                                   u16_max_rows);
            if (rc) return rc;
        }
        if (pipeline_chunks) {
            PGM_CUDA(cudaEventRecord(c->ev_fwd[set], c->stream));
            PGM_CUDA(cudaStreamWaitEvent(s_wm, c->ev_fwd[set], 0));
        }
        {
            ScopedKernelTimer t(c, K_NW_WALK, s_wm);
            rc = launch_nw_walk(s_wm, ntasks, n, d_codes, c->d_off.as<uint32_t>(), dt, tb,
                                list, c->d_list_len.as<uint32_t>(), d_gaps, d_err, walk_lanes);
            if (rc) return rc;
        }
        {
            ScopedKernelTimer t(c, K_ML, s_wm);
            rc = launch_ml_from_lists(s_wm, ml, ntasks, c->d_off.as<uint32_t>(), dt, list,
                                      c->d_list_len.as<uint32_t>(), d_gaps, d_dist, d_var, d_counts, d_evals, (int)task_pairs);
            if (rc) return rc;
        }
        if (pipeline_chunks) PGM_CUDA(cudaEventRecord(c->ev_wm[set], s_wm));
        c->timing.nw_cells += chunk_cells[k];
        c->timing.tb_bytes += chunk_tb[k] * 4;
        t0 = chunk_end[k];
    }
    if (pipeline_chunks) {                                               // join: the main stream continues after the last walk + ML
        PGM_CUDA(cudaStreamWaitEvent(c->stream, c->ev_wm[0], 0));
        PGM_CUDA(cudaStreamWaitEvent(c->stream, c->ev_wm[1], 0));
    }
    // error flag + evaluation counter come back with the sync of whoever consumes the results
    struct { int err; int pad[3]; unsigned long long evals; } h;
    PGM_CUDA(cudaMemcpyAsync(&h, c->d_misc.p, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    PGM_CUDA(cudaStreamSynchronize(c->stream));
    c->timing.ml_evaluations += h.evals;
    if (h.err) { pgm_set_error("error while backtracking"); return PGM_ERR_BACKTRACK; }
    return 0;
}

// full[i][j] = full[j][i] = packed[p(i,j)], zero diagonal: one thread per matrix element (both triangles read the packed arrays)
__global__ void scatter_pairs_kernel(const double* __restrict__ dp, const double* __restrict__ vp, int N, double* __restrict__ dist, double* __restrict__ var) {
    const size_t total = (size_t)N * N;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(k / N), q = (int)(k - (size_t)r * N);
        const int i = r < q ? r : q, j = r < q ? q : r;
        double d = 0.0, v = 0.0;
        if (i != j) { const size_t p = (size_t)i * N - (size_t)i * (i + 1) / 2 + (size_t)(j - i - 1); d = dp[p]; v = vp[p]; }
        dist[k] = d; var[k] = v;
    }
}

static void scatter_pairs(const double* dp, const double* vp, int N, double* dist, double* var) {
    for (size_t k = 0; k < (size_t)N * N; ++k) dist[k] = var[k] = 0;
    size_t p = 0;
    for (int i = 0; i < N; ++i)
        for (int j = i + 1; j < N; ++j, ++p) {
            dist[(size_t)i * N + j] = dist[(size_t)j * N + i] = dp[p];
            var[(size_t)i * N + j] = var[(size_t)j * N + i] = vp[p];
        }
}

extern "C" {

int pgm_nw_ml_pair_range_dev(pgm_ctx* c, const pgm_ml_params* prm, const uint8_t* d_codes, const uint32_t* offsets, int N,
                             const int32_t* score_tab, int go, int ge, int64_t pb, int64_t pe,
                             double* d_dist, double* d_var, int32_t* d_counts, uint32_t* d_gaps, int32_t* d_scores) {
    if (!c) { pgm_set_error("ctx is NULL"); return PGM_ERR_ARG; }
    reset_timing(c);
    return nw_ml_range_dev(c, prm, d_codes, offsets, N, score_tab, go, ge, pb, pe, d_dist, d_var, d_counts, d_gaps, d_scores);
}

int pgm_nw_ml_pair_range(pgm_ctx* c, const pgm_ml_params* prm, const uint8_t* codes, const uint32_t* offsets, int N,
                         const int32_t* score_tab, int go, int ge, int64_t pb, int64_t pe,
                         double* dist, double* var, int32_t* counts, uint32_t* gaps, int32_t* scores) {
    if (!c || !prm) { pgm_set_error("ctx/prm is NULL"); return PGM_ERR_ARG; }
    PGM_CUDA(cudaSetDevice(c->device));
    reset_timing(c);
    if (N < 0 || !offsets) { pgm_set_error("bad sequence set"); return PGM_ERR_ARG; }
    const int64_t np = pe - pb;
    if (np < 0) { pgm_set_error("pair range invalid"); return PGM_ERR_ARG; }
    if (np == 0) return 0;
    const size_t total = offsets[N], nn = (size_t)prm->n_states * prm->n_states;
    int rc;
    if ((rc = c->d_codes.reserve(std::max<size_t>(total, 1)))) return rc;
    if ((rc = c->d_dist.reserve(np * 8))) return rc;
    if ((rc = c->d_var.reserve(np * 8))) return rc;
    if (counts && (rc = c->d_counts.reserve(np * nn * 4))) return rc;
    if (gaps && (rc = c->d_gaps.reserve(np * 4))) return rc;
    if (scores && (rc = c->d_scores.reserve(np * 4))) return rc;
    PGM_CUDA(cudaMemcpyAsync(c->d_codes.p, codes, total, cudaMemcpyHostToDevice, c->stream));
    rc = nw_ml_range_dev(c, prm, c->d_codes.as<uint8_t>(), offsets, N, score_tab, go, ge, pb, pe, c->d_dist.as<double>(),
                         c->d_var.as<double>(), counts ? c->d_counts.as<int32_t>() : nullptr,
                         gaps ? c->d_gaps.as<uint32_t>() : nullptr, scores ? c->d_scores.as<int32_t>() : nullptr);
    if (rc) return rc;
    PGM_CUDA(cudaMemcpyAsync(dist, c->d_dist.p, np * 8, cudaMemcpyDeviceToHost, c->stream));
    PGM_CUDA(cudaMemcpyAsync(var, c->d_var.p, np * 8, cudaMemcpyDeviceToHost, c->stream));
    if (counts) PGM_CUDA(cudaMemcpyAsync(counts, c->d_counts.p, np * nn * 4, cudaMemcpyDeviceToHost, c->stream));
    if (gaps) PGM_CUDA(cudaMemcpyAsync(gaps, c->d_gaps.p, np * 4, cudaMemcpyDeviceToHost, c->stream));
    if (scores) PGM_CUDA(cudaMemcpyAsync(scores, c->d_scores.p, np * 4, cudaMemcpyDeviceToHost, c->stream));
    return pgm_sync(c);
}

int pgm_nw_ml_all_pairs(pgm_ctx* c, const pgm_ml_params* prm, const uint8_t* codes, const uint32_t* offsets, int N,
                        const int32_t* score_tab, int go, int ge, double* dist, double* var) {
    if (N < 0) { pgm_set_error("N < 0"); return PGM_ERR_ARG; }
    const int64_t np = (int64_t)N * (N - 1) / 2;
    std::vector<double> dp(std::max<int64_t>(np, 1)), vp(std::max<int64_t>(np, 1));
    int rc = pgm_nw_ml_pair_range(c, prm, codes, offsets, N, score_tab, go, ge, 0, np, dp.data(), vp.data(), nullptr, nullptr, nullptr);
    if (rc) return rc;
    scatter_pairs(dp.data(), vp.data(), N, dist, var);
    return 0;
}

// ---- HP-2b: prealigned ------------------------------------------------------------------------------------
static int prealigned_range_dev(pgm_ctx* c, const pgm_ml_params* prm, const uint8_t* d_rows, int N, int L, int64_t pb, int64_t pe,
                                double* d_dist, double* d_var, int32_t* d_counts, uint32_t* d_gaps) {
    PGM_CUDA(cudaSetDevice(c->device));
    const int64_t total_pairs = (int64_t)N * (N - 1) / 2;
    if (N < 0 || L < 0 || pb < 0 || pe < pb || pe > total_pairs) { pgm_set_error("pair range invalid"); return PGM_ERR_ARG; }
    if (pe == pb) return 0;
    MlDev ml; int rc = prepare_model(c, prm, ml); if (rc) return rc;
    if ((rc = c->d_misc.reserve(64))) return rc;
    PGM_CUDA(cudaMemsetAsync(c->d_misc.p, 0, 64, c->stream));
    unsigned long long* d_evals = (unsigned long long*)(c->d_misc.as<char>() + 16);
    {
        ScopedKernelTimer t(c, K_ML);
        const char* e = getenv("PGM_PREALIGNED_FUSED");
        if ((e && e[0] == '1') || L > 65535) {                         // (16-bit counters)  the round-1 flavour: column scan inside the Newton kernel (A-B runs)
            rc = launch_ml_from_rows(c->stream, ml, N, L, pb, pe - pb, d_rows, d_dist, d_var, d_counts, d_gaps, d_evals);
            if (rc) return rc;
        } else {
            // residue lists per row, then list merge (thread per pair) and Newton kernel per chunk of pairs; the chunk's counts (lim x lim uint16 per pair) live in
            // scratch: 4 Mi pairs = 3.4 GB for proteins
            const int lim = ml.n < 20 ? ml.n : 20;
            const int64_t np = pe - pb, chunk = std::min<int64_t>(np, (int64_t)4 << 20);
            if ((rc = c->d_pre_rows.reserve((size_t)std::max(N, 1) * prealigned_pitch(L) * 4))) return rc;      // residue lists (uint32 entries)
            if ((rc = c->d_pre_counts.reserve((size_t)chunk * lim * lim * 2))) return rc;
            if ((rc = c->d_pre_gaps.reserve((size_t)chunk * 4))) return rc;
            if ((rc = launch_prealigned_repitch(c->stream, d_rows, N, L, c->d_pre_rows.as<uint8_t>()))) return rc;
            const size_t nn = (size_t)ml.n * ml.n;
            for (int64_t q = 0; q < np; q += chunk) {
                const int64_t m = std::min<int64_t>(chunk, np - q);
                rc = launch_ml_from_rows_split(c->stream, ml, N, L, pb + q, m, c->d_pre_rows.as<uint8_t>(), c->d_pre_counts.as<uint16_t>(),
                                               c->d_pre_gaps.as<uint32_t>(), d_dist + q, d_var + q, d_counts ? d_counts + (size_t)q * nn : nullptr,
                                               d_gaps ? d_gaps + q : nullptr, d_evals);
                if (rc) return rc;
            }
        }
    }
    unsigned long long ev = 0;
    PGM_CUDA(cudaMemcpyAsync(&ev, d_evals, 8, cudaMemcpyDeviceToHost, c->stream));
    PGM_CUDA(cudaStreamSynchronize(c->stream));
    c->timing.ml_evaluations += ev;
    return 0;
}

int pgm_prealigned_ml_pair_range_dev(pgm_ctx* c, const pgm_ml_params* prm, const uint8_t* d_rows, int N, int L, int64_t pb, int64_t pe,
                                     double* d_dist, double* d_var, int32_t* d_counts, uint32_t* d_gaps) {
    if (!c || !prm) { pgm_set_error("ctx/prm is NULL"); return PGM_ERR_ARG; }
    reset_timing(c);
    return prealigned_range_dev(c, prm, d_rows, N, L, pb, pe, d_dist, d_var, d_counts, d_gaps);
}

int pgm_prealigned_ml_pair_range(pgm_ctx* c, const pgm_ml_params* prm, const uint8_t* rows, int N, int L, int64_t pb, int64_t pe,
                                 double* dist, double* var, int32_t* counts, uint32_t* gaps) {
    if (!c || !prm) { pgm_set_error("ctx/prm is NULL"); return PGM_ERR_ARG; }
    PGM_CUDA(cudaSetDevice(c->device));
    reset_timing(c);
    const int64_t np = pe - pb;
    if (np < 0 || N < 0 || L < 0) { pgm_set_error("bad arguments"); return PGM_ERR_ARG; }
    if (np == 0) return 0;
    const size_t nn = (size_t)prm->n_states * prm->n_states;
    int rc;
    if ((rc = c->d_rows.reserve(std::max<size_t>((size_t)N * L, 1)))) return rc;
    if ((rc = c->d_dist.reserve(np * 8))) return rc;
    if ((rc = c->d_var.reserve(np * 8))) return rc;
    if (counts && (rc = c->d_counts.reserve(np * nn * 4))) return rc;
    if (gaps && (rc = c->d_gaps.reserve(np * 4))) return rc;
    PGM_CUDA(cudaMemcpyAsync(c->d_rows.p, rows, (size_t)N * L, cudaMemcpyHostToDevice, c->stream));
    rc = prealigned_range_dev(c, prm, c->d_rows.as<uint8_t>(), N, L, pb, pe, c->d_dist.as<double>(), c->d_var.as<double>(),
                              counts ? c->d_counts.as<int32_t>() : nullptr, gaps ? c->d_gaps.as<uint32_t>() : nullptr);
    if (rc) return rc;
    PGM_CUDA(cudaMemcpyAsync(dist, c->d_dist.p, np * 8, cudaMemcpyDeviceToHost, c->stream));
    PGM_CUDA(cudaMemcpyAsync(var, c->d_var.p, np * 8, cudaMemcpyDeviceToHost, c->stream));
    if (counts) PGM_CUDA(cudaMemcpyAsync(counts, c->d_counts.p, np * nn * 4, cudaMemcpyDeviceToHost, c->stream));
    if (gaps) PGM_CUDA(cudaMemcpyAsync(gaps, c->d_gaps.p, np * 4, cudaMemcpyDeviceToHost, c->stream));
    return pgm_sync(c);
}

int pgm_prealigned_ml_all_pairs(pgm_ctx* c, const pgm_ml_params* prm, const uint8_t* rows, int N, int L, double* dist, double* var) {
    if (!c || !prm) { pgm_set_error("ctx/prm is NULL"); return PGM_ERR_ARG; }
    if (N < 0 || L < 0) { pgm_set_error("bad arguments"); return PGM_ERR_ARG; }
    PGM_CUDA(cudaSetDevice(c->device));
    reset_timing(c);
    const int64_t np = (int64_t)N * (N - 1) / 2;
    const size_t full = (size_t)N * N;
    if (np == 0) { for (size_t k = 0; k < full; ++k) dist[k] = var[k] = 0; return 0; }
    // the packed (dist, var) of the pairs are spread into the two symmetric N x N matrices on the device and come back as two copies:
    // at N = 5000 the host-side scatter (two 200 MB matrices written with stride N) took longer than all the kernels together
    int rc;
    if ((rc = c->d_rows.reserve(std::max<size_t>((size_t)N * L, 1)))) return rc;
    if ((rc = c->d_dist.reserve(np * 8))) return rc;
    if ((rc = c->d_var.reserve(np * 8))) return rc;
    if ((rc = c->d_full.reserve(2 * full * 8))) return rc;
    PGM_CUDA(cudaMemcpyAsync(c->d_rows.p, rows, (size_t)N * L, cudaMemcpyHostToDevice, c->stream));
    rc = prealigned_range_dev(c, prm, c->d_rows.as<uint8_t>(), N, L, 0, np, c->d_dist.as<double>(), c->d_var.as<double>(), nullptr, nullptr);
    if (rc) return rc;
    double* fd = c->d_full.as<double>(); double* fv = fd + full;
    scatter_pairs_kernel<<<148 * 8, 256, 0, c->stream>>>(c->d_dist.as<double>(), c->d_var.as<double>(), N, fd, fv);
    PGM_CUDA(cudaGetLastError());
    PGM_CUDA(cudaMemcpyAsync(dist, fd, full * 8, cudaMemcpyDeviceToHost, c->stream));
    PGM_CUDA(cudaMemcpyAsync(var, fv, full * 8, cudaMemcpyDeviceToHost, c->stream));
    return pgm_sync(c);
}

int pgm_angle_all_pairs(pgm_ctx* c, int n, int K, const uint8_t* codes, const uint32_t* offsets, int N, int ml_flag,
                        double* dist, double* var) {
    if (!c || !offsets || N < 0 || K < 1 || n < 2) { pgm_set_error("bad arguments"); return PGM_ERR_ARG; }
    PGM_CUDA(cudaSetDevice(c->device));
    reset_timing(c);
    if (N == 0) return 0;
    long long dims = 1;
    for (int k = 0; k < K; ++k) { dims *= n; if (dims > 8192) { pgm_set_error("n_states^K = %lld exceeds 8192", dims); return PGM_ERR_ARG; } }
    for (int k = 0; k < N; ++k) if (offsets[k + 1] - offsets[k] > 65535u) { pgm_set_error("sequence longer than 65535"); return PGM_ERR_ARG; }
    const size_t total = offsets[N];
    int rc;
    if ((rc = c->d_codes.reserve(std::max<size_t>(total, 1))) || (rc = c->d_off.reserve((size_t)(N + 1) * 4)) ||
        (rc = c->d_counts.reserve((size_t)N * dims * 2)) || (rc = c->d_seqlen.reserve((size_t)N * 8)) ||
        (rc = c->d_dist.reserve((size_t)N * N * 8)) || (rc = c->d_var.reserve((size_t)N * N * 8))) return rc;
    PGM_CUDA(cudaMemcpyAsync(c->d_codes.p, codes, total, cudaMemcpyHostToDevice, c->stream));
    PGM_CUDA(cudaMemcpyAsync(c->d_off.p, offsets, (size_t)(N + 1) * 4, cudaMemcpyHostToDevice, c->stream));
    rc = launch_angle(c->stream, n, K, (int)dims, N, c->d_codes.as<uint8_t>(), c->d_off.as<uint32_t>(), c->d_counts.as<uint16_t>(),
                      c->d_seqlen.as<double>(), ml_flag, c->d_dist.as<double>(), c->d_var.as<double>());
    if (rc) return rc;
    PGM_CUDA(cudaMemcpyAsync(dist, c->d_dist.p, (size_t)N * N * 8, cudaMemcpyDeviceToHost, c->stream));
    PGM_CUDA(cudaMemcpyAsync(var, c->d_var.p, (size_t)N * N * 8, cudaMemcpyDeviceToHost, c->stream));
    return pgm_sync(c);
}

int pgm_ml_distances(pgm_ctx* c, const pgm_ml_params* prm, int64_t np, const int32_t* counts, const uint32_t* gaps,
                     const double* seqlen, double* dist, double* var) {
    if (!c || !prm) { pgm_set_error("ctx/prm is NULL"); return PGM_ERR_ARG; }
    PGM_CUDA(cudaSetDevice(c->device));
    reset_timing(c);
    if (np < 0) { pgm_set_error("n_pairs < 0"); return PGM_ERR_ARG; }
    if (np == 0) return 0;
    MlDev ml; int rc = prepare_model(c, prm, ml); if (rc) return rc;
    const size_t nn = (size_t)prm->n_states * prm->n_states;
    if (!counts || !gaps || !seqlen || !dist || !var) { pgm_set_error("NULL argument"); return PGM_ERR_ARG; }
    for (size_t q = 0; q < (size_t)np * nn; ++q)                  // the kernel packs (index << 20 | count) into one word
        if (counts[q] < 0 || counts[q] >= (1 << 20)) { pgm_set_error("counts[%zu] = %d is outside [0, 2^20)", q, counts[q]); return PGM_ERR_ARG; }
    if ((rc = c->d_counts.reserve(np * nn * 4))) return rc;
    if ((rc = c->d_gaps.reserve(np * 4))) return rc;
    if ((rc = c->d_seqlen.reserve(np * 8))) return rc;
    if ((rc = c->d_dist.reserve(np * 8))) return rc;
    if ((rc = c->d_var.reserve(np * 8))) return rc;
    if ((rc = c->d_misc.reserve(64))) return rc;
    PGM_CUDA(cudaMemsetAsync(c->d_misc.p, 0, 64, c->stream));
    PGM_CUDA(cudaMemcpyAsync(c->d_counts.p, counts, np * nn * 4, cudaMemcpyHostToDevice, c->stream));
    PGM_CUDA(cudaMemcpyAsync(c->d_gaps.p, gaps, np * 4, cudaMemcpyHostToDevice, c->stream));
    PGM_CUDA(cudaMemcpyAsync(c->d_seqlen.p, seqlen, np * 8, cudaMemcpyHostToDevice, c->stream));
    {
        ScopedKernelTimer t(c, K_ML);
        rc = launch_ml_from_counts(c->stream, ml, np, c->d_counts.as<int32_t>(), c->d_gaps.as<uint32_t>(), c->d_seqlen.as<double>(),
                                   c->d_dist.as<double>(), c->d_var.as<double>(), (unsigned long long*)(c->d_misc.as<char>() + 16));
        if (rc) return rc;
    }
    PGM_CUDA(cudaMemcpyAsync(dist, c->d_dist.p, np * 8, cudaMemcpyDeviceToHost, c->stream));
    PGM_CUDA(cudaMemcpyAsync(var, c->d_var.p, np * 8, cudaMemcpyDeviceToHost, c->stream));
    unsigned long long ev = 0;
    PGM_CUDA(cudaMemcpyAsync(&ev, c->d_misc.as<char>() + 16, 8, cudaMemcpyDeviceToHost, c->stream));
    rc = pgm_sync(c);
    c->timing.ml_evaluations += ev;
    return rc;
}

} // extern "C"

// ---- accessors for the HP-1 host code (hp1_host.cu) -----------------------------------------------------
cudaStream_t pgm_ctx_stream(pgm_ctx* c) { return c->stream; }
int pgm_ctx_device(pgm_ctx* c) { return c->device; }
uint64_t pgm_ctx_scratch(pgm_ctx* c) { return c->scratch_bytes; }
pgm_timing* pgm_ctx_timing(pgm_ctx* c) { return &c->timing; }
void pgm_ctx_reset_timing(pgm_ctx* c) { reset_timing(c); }
void pgm_ctx_time_begin(pgm_ctx* c, int kind, void** tok) { *tok = new ScopedKernelTimer(c, kind); }
void pgm_ctx_time_begin_on(pgm_ctx* c, int kind, cudaStream_t st, void** tok) { *tok = new ScopedKernelTimer(c, kind, st); }
void pgm_ctx_time_end(pgm_ctx* c, void* tok) { (void)c; delete (ScopedKernelTimer*)tok; }

// ---- buffer pool (HP-1 arenas) ----------------------------------------------------------------------
template <class B> static int pool_take(std::vector<B>& pool, size_t bytes, B& out) {
    int best = -1;
    for (int i = 0; i < (int)pool.size(); ++i)
        if (pool[i].cap >= bytes && (best < 0 || pool[i].cap < pool[best].cap)) best = i;
    if (best >= 0) { out = pool[best]; pool.erase(pool.begin() + best); return 0; }
    if (!pool.empty()) {                       // too small: drop the largest one instead of hoarding memory
        int big = 0;
        for (int i = 1; i < (int)pool.size(); ++i) if (pool[i].cap > pool[big].cap) big = i;
        pool[big].release(); pool.erase(pool.begin() + big);
    }
    out = B();
    return out.reserve(bytes);
}
int pgm_ctx_take_dev(pgm_ctx* c, size_t bytes, DevBuf& out) { return pool_take(c->pool_dev, bytes, out); }
int pgm_ctx_take_host(pgm_ctx* c, size_t bytes, HostPinned& out) { return pool_take(c->pool_host, bytes, out); }
void pgm_ctx_give_dev(pgm_ctx* c, DevBuf& b) { if (b.p) { if (c && c->pool_dev.size() < 16) c->pool_dev.push_back(b); else b.release(); } b = DevBuf(); }
void pgm_ctx_give_host(pgm_ctx* c, HostPinned& b) { if (b.p) { if (c && c->pool_host.size() < 16) c->pool_host.push_back(b); else b.release(); } b = HostPinned(); }
