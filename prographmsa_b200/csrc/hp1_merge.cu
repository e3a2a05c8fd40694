// hp1_merge.cu -- N2: the profile part of mergeGraphs (reference GraphAlign.h:569-620).
//
// After alignGraphs, every node of the merged graph gets the profile
//     normalise( (P1 g1[k1]) o (P2 g2[k2]) )        both children contribute (an aligned column)
//     normalise(  P1 g1[k1] )  /  normalise( P2 g2[k2] )   only one child has a node on the path
//     normalise(  P1 g[k]   )                        a node the alignment path skipped -- P1 even for graph 2 (:591)
// with P = Model::P of the child's branch.  All merges of one guide-tree level are one launch: a warp per merged
// node, lane a accumulates output state a with sequential-k FP64 mul/add (the library is built with --fmad=false, so
// this is bit-identical to the oracle's plain C loop), then a shuffle reduction for the norm.
// Edge construction (std::map logic in the reference) stays on the host.
#include "internal.h"
#include <algorithm>
#include <vector>

cudaStream_t pgm_ctx_stream(pgm_ctx* c);
int pgm_ctx_device(pgm_ctx* c);
int pgm_ctx_take_dev(pgm_ctx* c, size_t bytes, DevBuf& out);
int pgm_ctx_take_host(pgm_ctx* c, size_t bytes, HostPinned& out);
void pgm_ctx_give_dev(pgm_ctx* c, DevBuf& b);
void pgm_ctx_give_host(pgm_ctx* c, HostPinned& b);

namespace {

struct MergeJobDev {
    const double *sites1, *sites2, *P1, *P2;
    const int32_t *rm1, *rm2;
    const uint8_t* matched;
    double* out;
    int len, node_base;                 // merged nodes; index of this job's first node in the launch
};

// sequential sum over k of P[a][k] * g[k] for output state a
__device__ __forceinline__ double mv_row(const double* __restrict__ P, const double* __restrict__ g, int n, int a) {
    double s = 0;
    for (int k = 0; k < n; ++k) s = s + P[a * n + k] * g[k];
    return s;
}

__global__ void __launch_bounds__(256)
merge_profiles_kernel(int n, int n_jobs, int total_nodes, const MergeJobDev* __restrict__ jobs, const int* __restrict__ node_job) {
    const int node = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (node >= total_nodes) return;
    const MergeJobDev J = jobs[node_job[node]];
    const int i = node - J.node_base;
    const int k1 = J.rm1[i], k2 = J.rm2[i];
    const bool m = J.matched[i] != 0;
    double ss = 0;
    double vals[2] = {0, 0};                                            // states lane and lane + 32 (n <= 64)
    for (int q = 0; q < 2; ++q) {
        const int a = lane + 32 * q;
        if (a < n) {
            double v;
            if (!m) v = mv_row(J.P1, k1 >= 0 ? J.sites1 + (size_t)k1 * n : J.sites2 + (size_t)k2 * n, n, a);
            else if (k1 >= 0 && k2 >= 0) v = mv_row(J.P1, J.sites1 + (size_t)k1 * n, n, a) * mv_row(J.P2, J.sites2 + (size_t)k2 * n, n, a);
            else if (k1 >= 0) v = mv_row(J.P1, J.sites1 + (size_t)k1 * n, n, a);
            else v = mv_row(J.P2, J.sites2 + (size_t)k2 * n, n, a);
            vals[q] = v;
        }
    }
    // the norm is summed in state order 0..n-1 like the oracle: lane 0 walks the values of the other lanes
    for (int a = 0; a < n; ++a) {
        const double v = __shfl_sync(0xffffffffu, a < 32 ? vals[0] : vals[1], a & 31);
        ss = ss + v * v;
    }
    const double nrm = sqrt(ss);
    for (int q = 0; q < 2; ++q) {
        const int a = lane + 32 * q;
        if (a < n) J.out[(size_t)i * n + a] = nrm == 0 ? vals[q] : vals[q] / nrm;
    }
}

} // namespace

extern "C" int pgm_merge_profiles_batch(pgm_ctx* c, int n, int n_jobs, const pgm_merge_job* jobs) {
    if (!c || n_jobs < 0 || (n_jobs > 0 && !jobs)) { pgm_set_error("bad arguments"); return PGM_ERR_ARG; }
    if (n < 2 || n > 64) { pgm_set_error("n_states %d out of range (2..64)", n); return PGM_ERR_ARG; }
    if (n_jobs == 0) return 0;
    PGM_CUDA(cudaSetDevice(pgm_ctx_device(c)));
    cudaStream_t st = pgm_ctx_stream(c);
    // one upload arena: per job sites1, sites2, P1, P2 (doubles), rm1, rm2 (int32), matched (bytes); then descriptors
    auto al = [](size_t v) { return (v + 127) / 128 * 128; };
    size_t in_bytes = 0, out_bytes = 0; long long total_nodes = 0;
    for (int k = 0; k < n_jobs; ++k) {
        const pgm_merge_job& j = jobs[k];
        if (j.N1 < 0 || j.N2 < 0 || j.len < 0 || !j.P1 || !j.P2 || (j.len && (!j.rm1 || !j.rm2 || !j.matched || !j.out))) { pgm_set_error("merge job %d: bad arguments", k); return PGM_ERR_ARG; }
        in_bytes += al(8ull * n * j.N1) + al(8ull * n * j.N2) + 2 * al(8ull * n * n) + 2 * al(4ull * j.len) + al(j.len);
        out_bytes += al(8ull * n * j.len);
        total_nodes += j.len;
    }
    if (total_nodes == 0) return 0;
    if (total_nodes > 0x7fffffffll) { pgm_set_error("too many merged nodes in one call"); return PGM_ERR_ARG; }
    const size_t desc_off = in_bytes, map_off = desc_off + al(sizeof(MergeJobDev) * n_jobs);
    const size_t up_bytes = map_off + al(4ull * total_nodes);
    HostPinned h_in, h_out; DevBuf d_in, d_out;
    int rc;
    if ((rc = pgm_ctx_take_host(c, up_bytes, h_in)) || (rc = pgm_ctx_take_dev(c, up_bytes, d_in)) ||
        (rc = pgm_ctx_take_host(c, out_bytes, h_out)) || (rc = pgm_ctx_take_dev(c, out_bytes, d_out))) {
        pgm_ctx_give_host(c, h_in); pgm_ctx_give_dev(c, d_in); pgm_ctx_give_host(c, h_out); pgm_ctx_give_dev(c, d_out);
        return rc;
    }
    char* hi = h_in.as<char>(); char* di = d_in.as<char>(); char* dout = d_out.as<char>();
    MergeJobDev* hd = (MergeJobDev*)(hi + desc_off);
    int* hmap = (int*)(hi + map_off);
    size_t off = 0, ooff = 0; int node_base = 0;
    std::vector<size_t> out_offs(n_jobs);
    for (int k = 0; k < n_jobs; ++k) {
        const pgm_merge_job& j = jobs[k];
        MergeJobDev& d = hd[k];
        auto put = [&](const void* src, size_t bytes) { const size_t r = off; if (bytes && src) memcpy(hi + off, src, bytes); off += al(bytes); return r; };
        d.sites1 = (const double*)(di + put(j.sites1, 8ull * n * j.N1));
        d.sites2 = (const double*)(di + put(j.sites2, 8ull * n * j.N2));
        d.P1 = (const double*)(di + put(j.P1, 8ull * n * n));
        d.P2 = (const double*)(di + put(j.P2, 8ull * n * n));
        d.rm1 = (const int32_t*)(di + put(j.rm1, 4ull * j.len));
        d.rm2 = (const int32_t*)(di + put(j.rm2, 4ull * j.len));
        d.matched = (const uint8_t*)(di + put(j.matched, (size_t)j.len));
        for (int i = 0; i < j.len; ++i) {
            const int a = j.rm1[i], b = j.rm2[i];
            if ((a < 0 && b < 0) || a >= j.N1 || b >= j.N2) {
                pgm_set_error("merge job %d: node %d maps to (%d, %d)", k, i, a, b);
                pgm_ctx_give_host(c, h_in); pgm_ctx_give_dev(c, d_in); pgm_ctx_give_host(c, h_out); pgm_ctx_give_dev(c, d_out);
                return PGM_ERR_ARG;
            }
            hmap[node_base + i] = k;
        }
        d.out = (double*)(dout + ooff); out_offs[k] = ooff; ooff += al(8ull * n * j.len);
        d.len = j.len; d.node_base = node_base; node_base += j.len;
    }
    PGM_CUDA(cudaMemcpyAsync(di, hi, up_bytes, cudaMemcpyHostToDevice, st));
    merge_profiles_kernel<<<(unsigned)((total_nodes + 7) / 8), 256, 0, st>>>(n, n_jobs, (int)total_nodes, (const MergeJobDev*)(di + desc_off),
                                                                             (const int*)(di + map_off));
    PGM_CUDA(cudaGetLastError());
    PGM_CUDA(cudaMemcpyAsync(h_out.p, d_out.p, out_bytes, cudaMemcpyDeviceToHost, st));
    PGM_CUDA(cudaStreamSynchronize(st));
    for (int k = 0; k < n_jobs; ++k) memcpy(jobs[k].out, h_out.as<char>() + out_offs[k], 8ull * n * jobs[k].len);
    pgm_ctx_give_host(c, h_in); pgm_ctx_give_dev(c, d_in); pgm_ctx_give_host(c, h_out); pgm_ctx_give_dev(c, d_out);
    return 0;
}
