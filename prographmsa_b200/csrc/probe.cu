// probe.cu -- FP64 FMA throughput of the device, measured in place (the ML kernel's roofline is the
// box's measured FP64 peak, which MEASURED_PEAKS.json does not carry).  Measurement infrastructure, not product code.
#include "internal.h"

struct pgm_ctx;
cudaStream_t pgm_ctx_stream(pgm_ctx* c);
int pgm_ctx_device(pgm_ctx* c);

__global__ void __launch_bounds__(256)
fp64_fma_probe_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 12345.678) out[0] = s;                    // keeps the chains alive
}

extern "C" int pgm_probe_fp64(pgm_ctx* c, double* tflops) {
    if (!c || !tflops) { pgm_set_error("NULL argument"); return PGM_ERR_ARG; }
    PGM_CUDA(cudaSetDevice(pgm_ctx_device(c)));
    cudaStream_t st = pgm_ctx_stream(c);
    int n_sms = 148; cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, pgm_ctx_device(c));
    double* d = nullptr; PGM_CUDA(cudaMalloc(&d, 64));
    const int blocks = n_sms * 8, iters = 8192;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    fp64_fma_probe_kernel<<<blocks, 256, 0, st>>>(d, 64, 0.999999, 1e-9);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0, st);
        fp64_fma_probe_kernel<<<blocks, 256, 0, st>>>(d, iters, 0.999999, 1e-9);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { pgm_set_error("CUDA error in the FP64 probe: %s", cudaGetErrorString(e)); return PGM_ERR_CUDA; }
    *tflops = 2.0 * 8.0 * (double)iters * 256.0 * blocks / (best * 1e-3) / 1e12;
    return 0;
}
