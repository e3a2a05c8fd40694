// host_graph.cpp -- native HOST glue around the graph DP (no GPU work): the graph surgery the reference does in C++
// between two alignGraphs calls.  Behaviour is restated from
//   CleanedGraph.h:38-146       keep the nodes on a finite-cost start->end path, drop stored-0 edges and 0-unit repeats
//   GraphAlign.h:569-620        which child nodes each node of the merged graph comes from ("merge plan")
//   GraphAlign.h:622-727        edges of the merged graph: homologous-path / skip edges of cost 0, child edges with the
//                               unused-edge penalty -log2(altsplice + (1-altsplice)(1-support)), repeat edges mapped through
//   Graph.h:81-100              stored edge value min(cost, 1e4) - 1e4 in float; entries sorted by (to, from)
// The Python mirrors in prographmsa_b200/graph.py are the readable specification; tests compare the two array by array.
// All arithmetic on edge values is float32, the penalty is computed in double and rounded once, as in the reference.
#include "internal.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

namespace {

const float kMaxEdge = 10000.0f;
inline float edge_value(float stored) { return stored == 0.0f ? std::numeric_limits<float>::infinity() : stored + kMaxEdge; }   // Graph.h:223-231
inline float store_value(float cost) { return std::min(cost, kMaxEdge) - kMaxEdge; }                                            // Graph.h:85

struct Trip { int to, from; float v; uint32_t u; };

}  // namespace

extern "C" {

// CleanedGraph(original, repeatext_prob).  Outputs are caller-allocated with the INPUT's sizes (N, edges, repeats):
//   keep[new_N]      index of every kept node in the original graph (CleanedGraph::outmapping)
//   e_* / r_*        CSR of the cleaned graph (e_rowptr / r_rowptr have new_N + 1 entries)
int pgm_host_clean_graph(const pgm_graph* g, double repeatext_prob, int32_t* keep, int* new_N,
                         int32_t* e_rowptr, int32_t* e_col, float* e_val, int32_t* r_rowptr, int32_t* r_col, uint32_t* r_units) {
    if (!g || !keep || !new_N || !e_rowptr || !r_rowptr || g->N < 2) { pgm_set_error("pgm_host_clean_graph: bad arguments"); return PGM_ERR_ARG; }
    const int N = g->N;
    const bool has_rep = g->r_rowptr != nullptr;
    // a predecessor entry "passes" when its cost is not +inf.  Regular edge: stored value != 0.  Repeat edge of u units:
    // u != 0 and 0 + repeat_ext*(u-1) != inf; with repeat_ext = inf (repeatext_prob == 0) that is inf*0 = NaN for u = 1,
    // which is != inf like in the reference, and inf for u > 1.
    auto rep_pass = [&](uint32_t u) { return u != 0 && !(repeatext_prob == 0 && u > 1); };
    std::vector<char> fw(N, 0), bw(N, 0);
    std::vector<int> mapping(N, -1);
    fw[0] = 1; bw[N - 1] = 1; mapping[0] = 0;
    for (int to = N - 1; to >= 0; --to) {
        if (!bw[to]) continue;
        for (int k = g->e_rowptr[to]; k < g->e_rowptr[to + 1]; ++k) if (g->e_val[k] != 0.0f) bw[g->e_col[k]] = 1;
        if (has_rep) for (int k = g->r_rowptr[to]; k < g->r_rowptr[to + 1]; ++k) if (rep_pass(g->r_units[k])) bw[g->r_col[k]] = 1;
    }
    int nd = 1;
    for (int to = 1; to < N; ++to) {
        bool reach = false;
        for (int k = g->e_rowptr[to]; k < g->e_rowptr[to + 1] && !reach; ++k) reach = g->e_val[k] != 0.0f && fw[g->e_col[k]];
        if (has_rep) for (int k = g->r_rowptr[to]; k < g->r_rowptr[to + 1] && !reach; ++k) reach = rep_pass(g->r_units[k]) && fw[g->r_col[k]];
        if (reach) { fw[to] = 1; if (bw[to]) mapping[to] = nd++; }
    }
    if (!bw[0] || !fw[N - 1]) { pgm_set_error("pgm_host_clean_graph: no finite start->end path"); return PGM_ERR_ARG; }
    *new_N = nd;
    // mapping is monotone, so rows stay sorted by from-node
    int ne = 0, nr = 0, row = 0;
    e_rowptr[0] = 0; r_rowptr[0] = 0;
    for (int to = 0; to < N; ++to) {
        if (mapping[to] < 0) continue;
        keep[row] = to;
        for (int k = g->e_rowptr[to]; k < g->e_rowptr[to + 1]; ++k) {
            const int x = mapping[g->e_col[k]];
            if (x != -1 && g->e_val[k] < 0.0f) { e_col[ne] = x; e_val[ne] = g->e_val[k]; ++ne; }
        }
        if (has_rep) for (int k = g->r_rowptr[to]; k < g->r_rowptr[to + 1]; ++k) {
            const int x = mapping[g->r_col[k]];
            if (x != -1 && g->r_units[k] > 0) { r_col[nr] = x; r_units[nr] = g->r_units[k]; ++nr; }
        }
        ++row;
        e_rowptr[row] = ne; r_rowptr[row] = nr;
    }
    return 0;
}

// merge plan (GraphAlign.h:569-620, loop structure): for the alignment columns (mapping1[k], mapping2[k]), -1 = gap, emit
// the child indices of every merged node, skipped child nodes included (matched = 0).  Outputs hold up to N1 + N2 entries.
int pgm_host_merge_plan(const int64_t* mapping1, const int64_t* mapping2, int len, int32_t* rm1, int32_t* rm2, uint8_t* matched, int* out_len) {
    if ((len && (!mapping1 || !mapping2)) || !rm1 || !rm2 || !matched || !out_len) { pgm_set_error("pgm_host_merge_plan: bad arguments"); return PGM_ERR_ARG; }
    int i1 = 0, i2 = 0, n = 0;
    for (int k = 0; k < len; ++k) {
        const int k1 = (int)mapping1[k], k2 = (int)mapping2[k];
        if (k1 != -1) { while (i1 != k1) { rm1[n] = i1; rm2[n] = -1; matched[n] = 0; ++n; ++i1; } ++i1; }
        if (k2 != -1) { while (i2 != k2) { rm1[n] = -1; rm2[n] = i2; matched[n] = 0; ++n; ++i2; } ++i2; }
        if (k1 == -1 && k2 == -1) { pgm_set_error("error in mapping"); return PGM_ERR_ARG; }
        rm1[n] = k1; rm2[n] = k2; matched[n] = 1; ++n;
    }
    *out_len = n;
    return 0;
}

// Edges of the merged graph (GraphAlign.h:622-727).  len merged nodes with plan (rm1, rm2, matched).
// Outputs: CSR with len + 1 row pointers; capacity E1 + E2 + 4*len edges, R1 + R2 repeats.
int pgm_host_merge_edges(const pgm_graph* g1, const pgm_graph* g2, int len, const int32_t* rm1, const int32_t* rm2, const uint8_t* matched,
                         double support1, double support2, double altsplice_prob,
                         int32_t* e_rowptr, int32_t* e_col, float* e_val, int* n_edges,
                         int32_t* r_rowptr, int32_t* r_col, uint32_t* r_units, int* n_repeats) {
    if (!g1 || !g2 || len < 0 || (len && (!rm1 || !rm2 || !matched)) || !e_rowptr || !r_rowptr || !n_edges || !n_repeats) {
        pgm_set_error("pgm_host_merge_edges: bad arguments"); return PGM_ERR_ARG;
    }
    std::vector<Trip> ed, rp;
    ed.reserve((size_t)g1->e_rowptr[g1->N] + g2->e_rowptr[g2->N] + 4 * (size_t)len);
    auto add = [&](int from, int to, float cost) { ed.push_back(Trip{to, from, cost, 0}); };
    int last_xy = 0, last_x = 0, last_y = 0, last_mapped = 0;
    for (int i = 1; i < len; ++i) {
        if (!matched[i]) continue;
        add(last_mapped, i, 0.0f); last_mapped = i;
        if (rm1[i] != -1 && rm2[i] != -1) { if (last_xy != i - 1) add(last_xy, i, 0.0f); last_xy = i; }
        if (rm1[i] != -1) { if (last_y != i - 1) add(last_y, i, 0.0f); last_y = i; }
        if (rm2[i] != -1) { if (last_x != i - 1) add(last_x, i, 0.0f); last_x = i; }
    }
    std::vector<int> inv1(g1->N, 0), inv2(g2->N, 0);
    for (int i = 0; i < len; ++i) { if (rm1[i] != -1) inv1[rm1[i]] = i; if (rm2[i] != -1) inv2[rm2[i]] = i; }
    auto LOG2 = [](double x) { return std::log(x) / std::log(2.0); };
    const float pen[2] = {(float)-LOG2(altsplice_prob + (1.0 - altsplice_prob) * (1.0 - support1)),
                          (float)-LOG2(altsplice_prob + (1.0 - altsplice_prob) * (1.0 - support2))};
    const pgm_graph* gs[2] = {g1, g2};
    const std::vector<int>* invs[2] = {&inv1, &inv2};
    for (int w = 0; w < 2; ++w) {
        const pgm_graph* g = gs[w];
        const std::vector<int>& inv = *invs[w];
        const float half = pen[w] / 2.0f;
        for (int to = 0; to < g->N; ++to) {
            for (int k = g->e_rowptr[to]; k < g->e_rowptr[to + 1]; ++k) {
                const int frm = g->e_col[k];
                const float val = edge_value(g->e_val[k]);
                // sic: the reference indexes is_matched[] (merged-node flags) with CHILD indices (:692-695,711-714)
                const bool mf = frm < len && matched[frm], mt = to < len && matched[to];
                add(inv[frm], inv[to], (mf && mt) ? val + pen[w] : ((mf || mt) ? val + half : val));
            }
            if (g->r_rowptr)
                for (int k = g->r_rowptr[to]; k < g->r_rowptr[to + 1]; ++k)
                    rp.push_back(Trip{inv[to], inv[g->r_col[k]], 0.0f, g->r_units[k]});
        }
    }
    // std::map semantics: one entry per (to, from), the minimum wins; rows ordered by to, entries by from.
    // stable_sort keeps insertion order among equal keys, so "minimum" is evaluated in the reference's order.
    auto by_key = [](const Trip& a, const Trip& b) { return a.to != b.to ? a.to < b.to : a.from < b.from; };
    std::stable_sort(ed.begin(), ed.end(), by_key);
    std::stable_sort(rp.begin(), rp.end(), by_key);
    int ne = 0, nr = 0;
    {
        size_t p = 0;
        for (int to = 0; to < len; ++to) {
            e_rowptr[to] = ne;
            while (p < ed.size() && ed[p].to == to) {
                float c = ed[p].v; const int from = ed[p].from; ++p;
                while (p < ed.size() && ed[p].to == to && ed[p].from == from) { if (ed[p].v < c) c = ed[p].v; ++p; }
                e_col[ne] = from; e_val[ne] = store_value(c); ++ne;
            }
        }
        e_rowptr[len] = ne;
    }
    {
        size_t p = 0;
        for (int to = 0; to < len; ++to) {
            r_rowptr[to] = nr;
            while (p < rp.size() && rp[p].to == to) {
                uint32_t u = rp[p].u; const int from = rp[p].from; ++p;
                while (p < rp.size() && rp[p].to == to && rp[p].from == from) { if (rp[p].u < u) u = rp[p].u; ++p; }
                r_col[nr] = from; r_units[nr] = u; ++nr;
            }
        }
        r_rowptr[len] = nr;
    }
    *n_edges = ne; *n_repeats = nr;
    return 0;
}

}  // extern "C"
