/*
 * oracle.cpp -- CPU parity oracle (TEST INFRASTRUCTURE ONLY; see oracle.h).
 *
 * Each function restates one piece of the reference and cites it.  Written from the algorithm
 * description in SURVEY.md Appendix A; data structures and control flow are our own (flat arrays,
 * CSR, no Eigen).  Build: `make -C oracle` (g++ -O2 -ffp-contract=off: the reference's float
 * arithmetic is SSE2 mul/add without contraction).
 */
#include "oracle.h"
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#include <limits>

namespace {

constexpr float FINF = std::numeric_limits<float>::infinity();

/* ------------------------------------------------------------------------------------------ */
/* HP-2                                                                                        */
/* ------------------------------------------------------------------------------------------ */

// C = A * B, row-major n*n
void matmul(const double* A, const double* B, double* C, int n) {
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double s = 0;
            for (int k = 0; k < n; ++k) s += A[i * n + k] * B[k * n + j];
            C[i * n + j] = s;
        }
}

// Taylor scaling-and-squaring expm; stands in for V*diag(exp(sigma t))*Vi (ModelFactory.h:54)
void expm(const double* Q, int n, double t, double* P) {
    const int nn = n * n;
    double norm = 0;
    for (int i = 0; i < n; ++i) {
        double r = 0;
        for (int j = 0; j < n; ++j) r += std::fabs(Q[i * n + j] * t);
        norm = std::max(norm, r);
    }
    int s = 0;
    while (norm > 0.25) { norm *= 0.5; ++s; }
    const double scale = std::ldexp(t, -s);
    std::vector<double> A(nn), term(nn), tmp(nn);
    for (int i = 0; i < nn; ++i) A[i] = Q[i] * scale;
    for (int i = 0; i < nn; ++i) { P[i] = 0; term[i] = 0; }
    for (int i = 0; i < n; ++i) { P[i * n + i] = 1; term[i * n + i] = 1; }
    for (int k = 1; k <= 20; ++k) {
        matmul(term.data(), A.data(), tmp.data(), n);
        for (int i = 0; i < nn; ++i) { term[i] = tmp[i] / k; P[i] += term[i]; }
    }
    for (int k = 0; k < s; ++k) {
        matmul(P, P, tmp.data(), n);
        std::memcpy(P, tmp.data(), sizeof(double) * nn);
    }
}

// ModelFactory.h:103-127 (ML branch) -- only the distance clamp matters to HP-2
double parse_distance_ml(double d, const orc_ml_params* p) {
    d = (0.0 < d) ? d : 0.0;                     // std::max(0.0, d); NaN -> 0
    return std::max(std::min(d, p->max_dist), p->min_dist);
}

struct distvar { double dist, var; };

// DistanceFactoryML.h:66-136
distvar compute_ml_dist(const int32_t* counts, uint32_t gaps, double seqlen, double dist0, double var0,
                        const orc_ml_params* prm, int* n_evals) {
    const int n = prm->n, nn = n * n;
    double dist_min = 0, dist_max = INFINITY;
    double dist = dist0, var = var0;
    double delta = 1;
    unsigned iteration = 0;
    std::vector<double> p(nn), pp(nn), ppp(nn);
    while (std::fabs(delta) > 1e-5) {
        if (iteration > 20) {
            if (dist_max == INFINITY) { dist = prm->DIST_MAX; var = prm->VAR_MAX; }
            else { dist = dist0; var = var0; }
            break;
        }
        expm(prm->Q, n, parse_distance_ml(dist, prm), p.data());
        matmul(prm->Q, p.data(), pp.data(), n);
        matmul(prm->Q, pp.data(), ppp.data(), n);
        if (n_evals) ++*n_evals;
        double f = 0, ff = 0;
        for (int j = 0; j < n; ++j)          // Eigen .sum() walks column-major
            for (int i = 0; i < n; ++i) {
                const int k = i * n + j;
                const double c = counts[k];
                f += c * pp[k] / p[k];
                ff += (c * (ppp[k] * p[k] - pp[k] * pp[k])) / (p[k] * p[k]);
            }
        if (prm->mldist_gap_flag) {
            const double grate = prm->indel_rate * seqlen * dist;
            f += (-grate + gaps) / dist;
            // sic: `-gaps` negates an UNSIGNED index_t (DistanceFactoryML.h:102) -> 2^32-gaps for gaps>0,
            // which turns the Newton step into pure bisection; reproduced, and confirmed against the binary
            ff += (double)(uint32_t)(0u - gaps) / (dist * dist);
        }
        var = -1.0 / ff;
        if (f > 0) dist_min = std::max(dist_min, dist);
        else dist_max = std::min(dist_max, dist);
        double new_dist = dist - f / ff;
        if (!(new_dist < dist_max && new_dist > dist_min)) {
            const double upper = (dist_max == INFINITY) ? dist * 3 : dist_max;
            new_dist = (upper + dist_min) / 2.0;
        }
        delta = 1.0 - new_dist / dist;
        dist = new_dist;
        ++iteration;
    }
    return {dist, var};
}

// DistanceFactoryML.h:138-190
distvar compute_distance(const int32_t* counts, uint32_t gaps, double seqlen, const orc_ml_params* prm, int* n_evals) {
    const int n = prm->n;
    double ident = 0, total = 0;
    for (int i = 0; i < n; ++i) ident += counts[i * n + i];
    for (int i = 0; i < n * n; ++i) total += counts[i];
    double dist0 = 1.0 - ident / total;
    double dist, var;
    if (prm->mldist_flag || prm->mldist_gap_flag) {
        if (total == 0 || dist0 > 0.85) { dist = dist0 = prm->DIST_MAX; var = prm->VAR_MAX; }
        else { dist = dist0 = -std::log(1.0 - dist0 - 0.2 * dist0 * dist0); var = dist / total; }
        if (total > 0 && ident != total) {
            distvar dv = compute_ml_dist(counts, gaps, seqlen, dist, var, prm, n_evals);
            dist = dv.dist; var = dv.var;
        }
    } else {
        if (total == 0) { dist = dist0 = 1.0; var = prm->VAR_MAX; }
        else { dist = dist0; var = dist0 / total; }
    }
    if (!(dist < prm->DIST_MAX)) { dist = prm->DIST_MAX; var = prm->VAR_MAX; }
    if (dist > prm->cutoff_dist) dist = prm->cutoff_dist;
    if (var < prm->VAR_MIN) var = prm->VAR_MIN;
    if (!(var < prm->VAR_MAX)) var = prm->VAR_MAX;
    return {dist, var};
}

/* ------------------------------------------------------------------------------------------ */
/* HP-1                                                                                        */
/* ------------------------------------------------------------------------------------------ */

// one predecessor as Graph::PredIterator yields it (Graph.h:172-248)
struct Pred { int idx; float val; bool rep; uint32_t units; };

// regular edges ascending, then repeat edges ascending (Graph.h:196,241-247);
// value(): Graph.h:223-239
void preds_of(const orc_graph* g, int node, float repeat_init, float repeat_ext, std::vector<Pred>& out) {
    out.clear();
    for (int k = g->e_rowptr[node]; k < g->e_rowptr[node + 1]; ++k) {
        const float c = g->e_val[k];
        out.push_back({g->e_col[k], c == 0 ? FINF : c + 10000.0f, false, 0});
    }
    if (g->r_rowptr)
        for (int k = g->r_rowptr[node]; k < g->r_rowptr[node + 1]; ++k) {
            const uint32_t c = g->r_units[k];
            float v;
            if (c == 0) v = FINF;
            else { const float m = repeat_ext * (float)(c - 1); v = repeat_init + m; }
            out.push_back({g->r_col[k], v, true, c});
        }
}

// GraphAlign.h:56-96, evaluated bottom-up (all edges point to smaller indices)
double average_alignment_length(const orc_graph* g) {
    if (g->N == 0) return 0;
    std::vector<double> cache(g->N, -2.0);
    cache[0] = 0;
    for (int cur = 1; cur < g->N; ++cur) {
        double sum = 0; unsigned paths = 0;
        for (int k = g->e_rowptr[cur]; k < g->e_rowptr[cur + 1]; ++k) {
            const float c = g->e_val[k];
            const float v = c == 0 ? FINF : c + 10000.0f;
            if (v == 0.0f) {
                const double res = cache[g->e_col[k]];
                if (res >= 0.0) { sum += res + 1.0; ++paths; }
            }
        }
        cache[cur] = paths > 0 ? sum / paths : -2.0;
    }
    return cache[g->N - 1];
}

inline double LOG2(double x) { return std::log(x) / std::log(2); }   // GraphAlign.h:46-47 (USE_LS_LOG)

struct Scores { float gap_init, gap_extend, match_init, end_match, end_gap, end_skip, start_gap, start_init, repeat_init, repeat_ext; };

// GraphAlign.h:98-143
Scores dynprog_scores(const orc_graph* g1, const orc_graph* g2, const orc_dp_model* m, const orc_dp_opts* o) {
    const double l1 = average_alignment_length(g1), l2 = average_alignment_length(g2);
    const double eps = m->epsilon, delta = m->delta, E = o->end_indel_prob;
    const double exp_length = std::max(l1, l2) * std::exp(m->distance * o->indel_rate * (eps / (1.0 - eps) + 1.0));
    const double nu = 2.0 / (2.0 + l1 + l2);
    double tau = 1.0 / (1.0 + exp_length);
    if (eps + tau >= 1.0) tau = (1.0 - eps) / 2.0;
    Scores s;
    s.gap_init = (float)LOG2(delta * (1.0 - eps - tau) / (1.0 - nu));
    s.gap_extend = (float)LOG2(eps / (1.0 - nu));
    s.match_init = (float)LOG2((1.0 - 2.0 * delta) * (1.0 - tau) / (1.0 - nu) / (1.0 - nu));
    s.end_skip = (float)LOG2(tau);
    if (E >= 0 && E <= 1) {
        s.end_match = (float)LOG2(tau * (1.0 - E) / (1.0 - 2.0 * delta) / (1.0 - tau));
        s.end_gap = (float)LOG2(tau * E / 2.0 / (1.0 - eps - tau) / delta);
        s.start_gap = (float)LOG2(E / 2.0 * (1.0 - eps - tau) / (1.0 - E) / (1.0 - nu));
        s.start_init = (float)LOG2((1.0 - tau) * (1.0 - E));
    } else {
        s.end_match = (float)LOG2(tau / (1.0 - tau));
        s.end_gap = (float)LOG2(tau / (1.0 - eps - tau));
        s.start_gap = (float)LOG2(delta * (1.0 - eps - tau) / (1.0 - nu));
        s.start_init = (float)LOG2(1.0 - tau);
    }
    const double repeat_prob = 1.0 - std::exp(-m->distance * o->repeat_rate);
    s.repeat_init = (float)-LOG2(std::min<double>(1, repeat_prob / (1 - repeat_prob) * (1 - o->repeatext_prob)));
    s.repeat_ext = (float)-LOG2(std::min<double>(1, std::max<double>(0, o->repeatext_prob)));
    return s;
}

// ls_log.h:22-59, vector body for index < len&~3, scalar tail (ls_log.h:7-19) after
void ls_log_add(float* data, float add, unsigned len) {
    const float a = 2.44247459618085927548717403238913328776812604856113966238617812902399112761292613763080658235564;
    const float b = -4.2040783745848554315883301529007786406310628696382695994938550046831869207082846248658671;
    const float c = -0.72123729809042963774358701619456664388406302428056983119308906451199556380646306;
    const unsigned alen = len & ~3u;
    for (unsigned i = 0; i < alen; ++i) {
        uint32_t bits; std::memcpy(&bits, data + i, 4);
        const float e = (float)((int32_t)(bits >> 23) - 126);
        uint32_t mb = ((bits << 9) >> 9) ^ 0x3f000000u;
        float d; std::memcpy(&d, &mb, 4);
        float vd = b / (d - c);
        const float ve = e + a;
        vd = vd + add;
        data[i] = vd + ve;
    }
    for (unsigned i = alen; i < len; ++i) {
        int e; const float d = std::frexp(data[i], &e);
        data[i] = ((a + b / (d - c)) + (float)e) + add;
    }
}

// The two normalisers g1^T pi and pi^T g2 (GraphAlign.h:154) are Eigen matrix-vector products.  Pinning against the
// reference binary on a 1000-sequence family (tools/make_golden_large.py) showed that they are NOT summed sequentially
// like the GEMM: the binary reproduces only with Eigen's SSE2 row-major gemv order -- four lane sums over
// k = lane, lane+4, ..., the horizontal add (l0 + l2) + (l1 + l3), then any n % 4 tail sequentially.  The difference
// is one ulp of S on non-one-hot profile columns and decides exact score ties (e.g. which N of "NN" takes the gap).
static float dot_gemv(const float* a, const float* b, int n) {
    float lane[4] = {0, 0, 0, 0};
    const int n4 = n & ~3;
    for (int k = 0; k < n4; k += 4) for (int l = 0; l < 4; ++l) lane[l] = lane[l] + a[k + l] * b[k + l];
    float s = (lane[0] + lane[2]) + (lane[1] + lane[3]);
    for (int k = n4; k < n; ++k) s = s + a[k] * b[k];
    return s;
}
// GraphAlign.h:145-163: float casts, S = (g1^T (M^T g2)) ./ ((g1^T pi)(pi^T g2)), then ls_log_add
void precompute_scores(const orc_graph* g1, const orc_graph* g2, const orc_dp_model* m, float match_init, float* S) {
    const int n = m->n, N1 = g1->N, N2 = g2->N;
    std::vector<float> Mf(n * n), pif(n), T((size_t)n * N2), u(N1), v(N2);
    for (int i = 0; i < n * n; ++i) Mf[i] = (float)m->M[i];
    for (int i = 0; i < n; ++i) pif[i] = (float)m->pi[i];
    for (int x = 0; x < N2; ++x) {
        for (int a = 0; a < n; ++a) {               // T(a,x) = sum_b M(b,a) g2(b,x)
            float s = 0;
            for (int b = 0; b < n; ++b) s = s + Mf[b * n + a] * (float)g2->sites[b + (size_t)n * x];
            T[a + (size_t)n * x] = s;
        }
        std::vector<float> col(n);
        for (int b = 0; b < n; ++b) col[b] = (float)g2->sites[b + (size_t)n * x];
        v[x] = dot_gemv(col.data(), pif.data(), n);
    }
    for (int y = 0; y < N1; ++y) {
        std::vector<float> col(n);
        for (int a = 0; a < n; ++a) col[a] = (float)g1->sites[a + (size_t)n * y];
        u[y] = dot_gemv(col.data(), pif.data(), n);
    }
    for (int x = 0; x < N2; ++x)
        for (int y = 0; y < N1; ++y) {
            float s = 0;
            for (int a = 0; a < n; ++a) s = s + (float)g1->sites[a + (size_t)n * y] * T[a + (size_t)n * x];
            S[y + (size_t)N1 * x] = s / (u[y] * v[x]);
        }
    ls_log_add(S, match_init, (unsigned)((size_t)N1 * N2));
}

// GraphAlign.h:165-198
void mark_alternative_path(int start, int end, const orc_graph* g, std::vector<int32_t>& mapping, std::vector<int32_t>& other) {
    const int len = end - start + 1;
    std::vector<float> score(len, -FINF);
    std::vector<int> prev(len, -1);
    std::vector<Pred> pr;
    score[0] = 0;
    for (int i = 1; i < len; ++i) {
        preds_of(g, i + start, FINF, FINF, pr);
        for (const Pred& p : pr)
            if (p.idx >= start && p.idx <= end) {
                const int i2 = p.idx - start;
                if (score[i] <= score[i2] - p.val) { score[i] = score[i2] - p.val; prev[i] = i2; }
            }
    }
    if (score[len - 1] > -FINF) {
        int i = prev[len - 1];
        while (i != 0) { mapping.push_back(i + start); other.push_back(-1); i = prev[i]; }
    }
}

} // namespace

extern "C" {

void orc_expm(const double* Q, int n, double t, double* P) { expm(Q, n, t, P); }

int orc_nw_align_pair(const uint8_t* s1, int L1, const uint8_t* s2, int L2, int n, const int32_t* tab,
                      int go, int ge, int32_t* counts, uint32_t* gaps_out, int32_t* score_out) {
    const int minfty = -10000, W1 = L1 + 1, n1 = n + 1;
    std::vector<int> W((size_t)(L2 + 1) * W1), X((size_t)(L2 + 1) * W1), Y((size_t)(L2 + 1) * W1);
    auto at = [W1](int y, int x) { return (size_t)y * W1 + x; };
    auto sc = [&](int y, int x) { return tab[s2[y - 1] + n1 * s1[x - 1]]; };   // scoring_matrix(s2,s1), col-major
    W[0] = 0; X[0] = 0; Y[0] = 0;
    for (int x = 1; x <= L1; ++x) { X[at(0, x)] = W[at(0, x)] = go + (x - 1) * ge; Y[at(0, x)] = minfty; }
    for (int y = 1; y <= L2; ++y) { Y[at(y, 0)] = W[at(y, 0)] = go + (y - 1) * ge; X[at(y, 0)] = minfty; }
    for (int y = 1; y <= L2; ++y)
        for (int x = 1; x <= L1; ++x) {
            const int d = W[at(y - 1, x - 1)] + sc(y, x);
            const int xx = std::max(X[at(y, x - 1)] + ge, W[at(y, x - 1)] + go);
            const int yy = std::max(Y[at(y - 1, x)] + ge, W[at(y - 1, x)] + go);
            X[at(y, x)] = xx; Y[at(y, x)] = yy;
            W[at(y, x)] = std::max(std::max(xx, yy), d);
        }
    std::memset(counts, 0, sizeof(int32_t) * n * n);
    uint32_t gaps = 0;
    bool o1 = false, o2 = false;
    for (int y = L2, x = L1; y != 0 && x != 0;) {
        if (W[at(y, x)] == W[at(y - 1, x - 1)] + sc(y, x)) {
            if (s1[x - 1] < n && s2[y - 1] < n) ++counts[s1[x - 1] * n + s2[y - 1]];
            o1 = o2 = false; --x; --y;
        } else if (W[at(y, x)] == X[at(y, x)]) {
            if (!o1) ++gaps;
            o1 = true; o2 = false; --x;
        } else if (W[at(y, x)] == Y[at(y, x)]) {
            if (!o2) ++gaps;
            o1 = false; o2 = true; --y;
        } else return -1;
    }
    *gaps_out = gaps;
    if (score_out) *score_out = W[at(L2, L1)];
    return 0;
}

void orc_prealigned_counts(const uint8_t* r1, const uint8_t* r2, int L, int n, int32_t* counts, uint32_t* gaps_out) {
    std::memset(counts, 0, sizeof(int32_t) * n * n);
    uint32_t gaps = 0; bool o1 = false, o2 = false;
    const int lim = std::min(20, n);                       // hard-coded 20 (DistanceFactoryPrealigned.h:65)
    for (int k = 0; k < L; ++k) {
        const bool g1 = r1[k] == 254, g2 = r2[k] == 254;
        if (!g1 && !g2) {
            if (r1[k] < lim && r2[k] < lim) ++counts[r1[k] * n + r2[k]];
            o1 = o2 = false;
        } else if (g1 && g2) {
        } else if (!g1 && !o1) { ++gaps; o1 = true; o2 = false; }
        else if (!g2 && !o2) { ++gaps; o1 = false; o2 = true; }
    }
    *gaps_out = gaps;
}

void orc_compute_distance(const int32_t* counts, uint32_t gaps, double seqlen, const orc_ml_params* p,
                          double* dist, double* var, int* n_evals) {
    if (n_evals) *n_evals = 0;
    distvar dv = compute_distance(counts, gaps, seqlen, p, n_evals);
    *dist = dv.dist; *var = dv.var;
}

int orc_nw_ml_all_pairs(const uint8_t* codes, const uint32_t* off, int N, int n, const int32_t* tab,
                        int go, int ge, const orc_ml_params* p, double* dist, double* var) {
    std::vector<int32_t> counts(n * n);
    for (size_t i = 0; i < (size_t)N * N; ++i) dist[i] = var[i] = 0;
    for (int i = 0; i < N; ++i)
        for (int j = i + 1; j < N; ++j) {
            const int L1 = off[i + 1] - off[i], L2 = off[j + 1] - off[j];
            uint32_t gaps;
            if (orc_nw_align_pair(codes + off[i], L1, codes + off[j], L2, n, tab, go, ge, counts.data(), &gaps, nullptr)) return -1;
            distvar dv = compute_distance(counts.data(), gaps, (L1 + L2) / 2.0, p, nullptr);
            dist[(size_t)i * N + j] = dist[(size_t)j * N + i] = dv.dist;
            var[(size_t)i * N + j] = var[(size_t)j * N + i] = dv.var;
        }
    return 0;
}

int orc_prealigned_ml_all_pairs(const uint8_t* rows, int N, int L, int n, const orc_ml_params* p, double* dist, double* var) {
    std::vector<int32_t> counts(n * n);
    for (size_t i = 0; i < (size_t)N * N; ++i) dist[i] = var[i] = 0;
    for (int i = 0; i < N; ++i)
        for (int j = i + 1; j < N; ++j) {
            uint32_t gaps;
            orc_prealigned_counts(rows + (size_t)i * L, rows + (size_t)j * L, L, n, counts.data(), &gaps);
            distvar dv = compute_distance(counts.data(), gaps, (L + L) / 2.0, p, nullptr);
            dist[(size_t)i * N + j] = dist[(size_t)j * N + i] = dv.dist;
            var[(size_t)i * N + j] = var[(size_t)j * N + i] = dv.var;
        }
    return 0;
}

void orc_angle_all_pairs(const uint8_t* codes, const uint32_t* off, int N, int n, int K, int ml_flag, double* dist, double* var) {
    size_t dims = 1;
    for (int k = 0; k < K; ++k) dims *= n;
    std::vector<double> cnt((size_t)N * dims, 0.0), inv(N), A((size_t)N * dims);
    for (int i = 0; i < N; ++i) {
        const int L = off[i + 1] - off[i];
        std::vector<int> chars(K, -1);
        for (int j = 0; j < L; ++j) {
            for (int k = 1; k < K; ++k) chars[k - 1] = chars[k];
            const int c = codes[off[i] + j];
            chars[K - 1] = (c >= n) ? -1 : c;
            long index = 0;
            for (int k = 0; k < K; ++k) { if (chars[k] == -1) { index = -1; break; } index = index * n + chars[k]; }
            if (index != -1) cnt[(size_t)i * dims + index] += 1;
        }
        double ss = 0;
        for (size_t k = 0; k < dims; ++k) ss += cnt[(size_t)i * dims + k] * cnt[(size_t)i * dims + k];
        inv[i] = 1.0 / std::sqrt(ss);
        for (size_t k = 0; k < dims; ++k) A[(size_t)i * dims + k] = inv[i] * cnt[(size_t)i * dims + k];   // D * C2^T
    }
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            double s = 0;
            for (size_t k = 0; k < dims; ++k) s += A[(size_t)i * dims + k] * cnt[(size_t)j * dims + k];       // (D C2^T) C2
            double d = s * inv[j];                                                                                // ... D
            d = -1.0 * std::log((d * d + 0.4) / 1.4);
            if (!ml_flag) { d = std::exp(d); d = -0.5 * (5.0 * d - std::sqrt(45.0 * d * d - 20.0 * d)) * (1.0 / d); }
            const double li = off[i + 1] - off[i], lj = off[j + 1] - off[j];
            double v = (1.0 / ((li + lj) / 2)) * d;
            v = std::max(v, 1e-5);
            dist[(size_t)i * N + j] = d; var[(size_t)i * N + j] = v;
        }
}

// mergeGraphs, profile part (GraphAlign.h:569-620): P * g as a sequential-k mul/add per output state, Hadamard product,
// L2 normalisation (skipped for an all-zero profile)
void orc_merge_profiles(int n, const double* sites1, const double* sites2, const double* P1, const double* P2,
                        int len, const int32_t* rm1, const int32_t* rm2, const uint8_t* matched, double* out) {
    std::vector<double> a(n), b(n);
    auto matvec = [&](const double* P, const double* g, std::vector<double>& r) {
        for (int i = 0; i < n; ++i) { double s = 0; for (int k = 0; k < n; ++k) s = s + P[i * n + k] * g[k]; r[i] = s; }
    };
    for (int i = 0; i < len; ++i) {
        double* o = out + (size_t)i * n;
        if (!matched[i]) {
            matvec(P1, rm1[i] >= 0 ? sites1 + (size_t)rm1[i] * n : sites2 + (size_t)rm2[i] * n, a);
        } else if (rm1[i] >= 0 && rm2[i] >= 0) {
            matvec(P1, sites1 + (size_t)rm1[i] * n, a); matvec(P2, sites2 + (size_t)rm2[i] * n, b);
            for (int k = 0; k < n; ++k) a[k] = a[k] * b[k];
        } else if (rm1[i] >= 0) matvec(P1, sites1 + (size_t)rm1[i] * n, a);
        else matvec(P2, sites2 + (size_t)rm2[i] * n, a);
        double ss = 0;
        for (int k = 0; k < n; ++k) ss = ss + a[k] * a[k];
        const double nrm = std::sqrt(ss);
        for (int k = 0; k < n; ++k) o[k] = nrm == 0 ? a[k] : a[k] / nrm;
    }
}

double orc_average_alignment_length(const orc_graph* g) { return average_alignment_length(g); }

void orc_dynprog_scores(const orc_graph* g1, const orc_graph* g2, const orc_dp_model* m, const orc_dp_opts* o, float out[10]) {
    const Scores s = dynprog_scores(g1, g2, m, o);
    std::memcpy(out, &s, sizeof(float) * 10);
}

void orc_ls_log_add(float* data, float add, unsigned len) { ls_log_add(data, add, len); }

void orc_precompute_scores(const orc_graph* g1, const orc_graph* g2, const orc_dp_model* m, float match_init, float* S) {
    precompute_scores(g1, g2, m, match_init, S);
}

int orc_align_graphs(const orc_graph* g1, const orc_graph* g2, const orc_dp_model* m, const orc_dp_opts* o,
                     float* score_out, uint32_t* n_tr_out, int32_t* map1, int32_t* map2, int cap, int* len_out,
                     float* dumpM, float* dumpX, float* dumpY, float* dumpW, float* dumpS, uint64_t* cell_updates) {
    const Scores s = dynprog_scores(g1, g2, m, o);
    const int N1 = g1->N, N2 = g2->N;
    const size_t sz = (size_t)N1 * N2;
    std::vector<float> S(sz);
    precompute_scores(g1, g2, m, s.match_init, S.data());
    std::vector<float> M(sz, -FINF), X(sz, -FINF), Y(sz, -FINF), W(sz, -FINF);
    auto at = [N1](int y, int x) { return (size_t)y + (size_t)N1 * x; };
    std::vector<std::vector<Pred>> P1(N1), P2(N2);
    for (int y = 0; y < N1; ++y) preds_of(g1, y, s.repeat_init, s.repeat_ext, P1[y]);
    for (int x = 0; x < N2; ++x) preds_of(g2, x, s.repeat_init, s.repeat_ext, P2[x]);
    uint64_t updates = 0;

    // boundary (GraphAlign.h:212-234)
    W[at(0, 0)] = s.start_init;
    for (int y = 1; y < N1 - 1; ++y) {
        float Sy = -FINF;
        for (const Pred& p : P1[y])
            Sy = std::max(Sy, std::max(Y[at(p.idx, 0)] + s.gap_extend, W[at(p.idx, 0)] + s.start_gap) - p.val);
        Y[at(y, 0)] = Sy; W[at(y, 0)] = Sy;
    }
    for (int x = 1; x < N2 - 1; ++x) {
        float Sx = -FINF;
        for (const Pred& p : P2[x])
            Sx = std::max(Sx, std::max(X[at(0, p.idx)] + s.gap_extend, W[at(0, p.idx)] + s.start_gap) - p.val);
        X[at(0, x)] = Sx; W[at(0, x)] = Sx;
    }
    // fill (GraphAlign.h:238-260)
    for (int y = 1; y < N1 - 1; ++y)
        for (int x = 1; x < N2 - 1; ++x) {
            float m_ = -FINF, x_ = -FINF, y_ = -FINF, w_ = -FINF;
            const float Syx = S[at(y, x)];
            for (const Pred& py : P1[y])
                for (const Pred& px : P2[x]) {
                    const int yp = py.idx, xp = px.idx;
                    const float Sm = W[at(yp, xp)] + Syx - py.val - px.val;
                    const float Sx = std::max(X[at(y, xp)] + s.gap_extend, W[at(y, xp)] + s.gap_init) - px.val;
                    const float Sy = std::max(Y[at(yp, x)] + s.gap_extend, W[at(yp, x)] + s.gap_init) - py.val;
                    const float Sw = std::max(Sm, std::max(Sx, Sy));
                    m_ = std::max(m_, Sm); x_ = std::max(x_, Sx); y_ = std::max(y_, Sy); w_ = std::max(w_, Sw);
                    ++updates;
                }
            M[at(y, x)] = m_; X[at(y, x)] = x_; Y[at(y, x)] = y_; W[at(y, x)] = w_;
        }
    // end node (GraphAlign.h:262-280)
    float Wend = -FINF;
    for (const Pred& py : P1[N1 - 1])
        for (const Pred& px : P2[N2 - 1]) {
            const int yp = py.idx, xp = px.idx;
            if (xp == 0 && yp == 0) Wend = std::max(s.end_skip - py.val - px.val, Wend);
            else {
                Wend = std::max(X[at(yp, xp)] + s.end_gap - py.val - px.val, Wend);
                Wend = std::max(Y[at(yp, xp)] + s.end_gap - py.val - px.val, Wend);
                Wend = std::max(M[at(yp, xp)] + s.end_match - py.val - px.val, Wend);
            }
        }
    if (dumpM) std::memcpy(dumpM, M.data(), sz * 4);
    if (dumpX) std::memcpy(dumpX, X.data(), sz * 4);
    if (dumpY) std::memcpy(dumpY, Y.data(), sz * 4);
    if (dumpW) std::memcpy(dumpW, W.data(), sz * 4);
    if (dumpS) std::memcpy(dumpS, S.data(), sz * 4);
    if (cell_updates) *cell_updates = updates;

    // traceback (GraphAlign.h:283-521)
    enum State { State_m, State_x, State_y } cur_state = State_m;
    float cur_score = -FINF;
    int y = N1 - 1, x = N2 - 1;
    uint32_t n_tr = 0;
    std::vector<int32_t> m1, m2;
    m1.reserve(N1 + N2); m2.reserve(N1 + N2);
    m1.push_back(N1 - 1); m2.push_back(N2 - 1);
    bool trx = false, try_ = false;
    float best = FINF;
    for (const Pred& py : P1[N1 - 1])
        for (const Pred& px : P2[N2 - 1]) {
            const int yp = py.idx, xp = px.idx;
            float c = std::fabs(Wend - (M[at(yp, xp)] + s.end_match - py.val - px.val));
            if (best > c) { best = c; trx = px.rep; try_ = py.rep; cur_score = M[at(yp, xp)]; cur_state = State_m; y = yp; x = xp; }
            c = std::fabs(Wend - (Y[at(yp, xp)] + s.end_gap - py.val - px.val));
            if (best > c) { best = c; trx = px.rep; try_ = py.rep; cur_score = Y[at(yp, xp)]; cur_state = State_y; y = yp; x = xp; }
            c = std::fabs(Wend - (X[at(yp, xp)] + s.end_gap - py.val - px.val));
            if (best > c) { best = c; trx = px.rep; try_ = py.rep; cur_score = X[at(yp, xp)]; cur_state = State_x; y = yp; x = xp; }
            if (xp == 0 && yp == 0) {
                c = std::fabs(Wend - (s.end_skip - py.val - px.val));
                if (best > c) { best = c; trx = px.rep; try_ = py.rep; y = yp; x = xp; }
            }
        }
    n_tr += (trx ? 1 : 0) + (try_ ? 1 : 0);
    if (try_) mark_alternative_path(y, N1 - 1, g1, m1, m2);
    if (trx) mark_alternative_path(x, N2 - 1, g2, m2, m1);
    auto emit = [&]() {
        if (x != 0 || y != 0) {
            if (cur_state == State_m) { m1.push_back(y); m2.push_back(x); }
            else if (cur_state == State_x) { m1.push_back(-1); m2.push_back(x); }
            else { m1.push_back(y); m2.push_back(-1); }
        }
    };
    emit();
    float next_score = FINF; State next_state = State_m; int next_x = -1, next_y = -1;
    int rc = 0;
    auto resolve_w = [&]() -> bool {       // landing on W: M, then Y, then X (GraphAlign.h:399-411)
        if (next_x != 0 || next_y != 0) {
            const size_t k = at(next_y, next_x);
            if (W[k] == M[k]) { next_score = M[k]; next_state = State_m; }
            else if (W[k] == Y[k]) { next_score = Y[k]; next_state = State_y; }
            else if (W[k] == X[k]) { next_score = X[k]; next_state = State_x; }
            else return false;
        }
        return true;
    };
    size_t guard = 0;
    while (x != 0 || y != 0) {
        if (++guard > (size_t)4 * (N1 + N2) + 16) { rc = -2; break; }
        best = FINF;
        if (cur_state == State_y) {
            for (const Pred& py : P1[y]) {
                const int yp = py.idx;
                float c = std::fabs(cur_score - (Y[at(yp, x)] + s.gap_extend - py.val));
                if (best > c) { best = c; trx = false; try_ = py.rep; next_x = x; next_y = yp; next_score = Y[at(yp, x)]; next_state = State_y; }
                c = std::fabs(cur_score - (W[at(yp, x)] + s.gap_init - py.val));
                if (best > c) { best = c; trx = false; try_ = py.rep; next_x = x; next_y = yp; if (!resolve_w()) { rc = -1; } }
            }
        }
        if (cur_state == State_x) {
            for (const Pred& px : P2[x]) {
                const int xp = px.idx;
                float c = std::fabs(cur_score - (X[at(y, xp)] + s.gap_extend - px.val));
                if (best > c) { best = c; trx = px.rep; try_ = false; next_x = xp; next_y = y; next_score = X[at(y, xp)]; next_state = State_x; }
                c = std::fabs(cur_score - (W[at(y, xp)] + s.gap_init - px.val));
                if (best > c) { best = c; trx = px.rep; try_ = false; next_x = xp; next_y = y; if (!resolve_w()) { rc = -1; } }
            }
        }
        if (cur_state == State_m) {
            for (const Pred& py : P1[y])
                for (const Pred& px : P2[x]) {
                    const int yp = py.idx, xp = px.idx;
                    const float c = std::fabs(cur_score - (W[at(yp, xp)] + S[at(y, x)] - py.val - px.val));
                    if (best > c) { best = c; trx = px.rep; try_ = py.rep; next_y = yp; next_x = xp; if (!resolve_w()) { rc = -1; } }
                }
        }
        if (rc) break;
        n_tr += (trx ? 1 : 0) + (try_ ? 1 : 0);
        if (try_) mark_alternative_path(next_y, y, g1, m1, m2);
        if (trx) mark_alternative_path(next_x, x, g2, m2, m1);
        x = next_x; y = next_y; cur_state = next_state; cur_score = next_score;
        emit();
    }
    m1.push_back(0); m2.push_back(0);
    std::reverse(m1.begin(), m1.end());
    std::reverse(m2.begin(), m2.end());
    *score_out = Wend; *n_tr_out = n_tr;
    *len_out = (int)m1.size();
    if ((int)m1.size() > cap) return -3;
    std::memcpy(map1, m1.data(), m1.size() * 4);
    std::memcpy(map2, m2.data(), m2.size() * 4);
    return rc;
}

} // extern "C"
