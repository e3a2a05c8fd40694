// hp2_angle.cu -- k-mer angle distances (the default initial-distance path when -a
// is absent).  Replaces DistanceFactoryAngle<A,K>::computePwDistances (reference DistanceFactoryAngle.h:54-131):
//   counts(i, kmer) over valid K-mers;  cos(i,j) = <c_i, c_j> / (|c_i| |c_j|);
//   d = -ln((cos^2 + 0.4) / 1.4)   [+ the non-ML transform d <- -0.5 (5e - sqrt(45 e^2 - 20 e)) / e, e = exp(d)]
//   var = d / ((L_i + L_j)/2), floored at 1e-5.
// The Gram matrix is an exact integer contraction (int32 accumulate), tiled through shared memory; only the
// epilogue is floating point (FP64).  The reference scales the count vectors before the product, so results
// agree to rounding (~1e-16 relative), not bit for bit.
#include "internal.h"

// one CTA per sequence: histogram of valid K-mers in shared memory -> counts[i][dims] (uint16) + squared norm
__global__ void __launch_bounds__(256)
angle_count_kernel(const uint8_t* __restrict__ codes, const uint32_t* __restrict__ off, int n, int K, int dims,
                   uint16_t* __restrict__ counts, double* __restrict__ norm2) {
    extern __shared__ uint32_t hist[];
    const int i = blockIdx.x;
    for (int k = threadIdx.x; k < dims; k += blockDim.x) hist[k] = 0;
    __syncthreads();
    const uint8_t* s = codes + off[i];
    const int L = (int)(off[i + 1] - off[i]);
    for (int j = K - 1 + (int)threadIdx.x; j < L; j += blockDim.x) {
        int index = 0; bool ok = true;
        for (int k = 0; k < K; ++k) {
            const int c = s[j - (K - 1) + k];
            if (c >= n) { ok = false; break; }
            index = index * n + c;
        }
        if (ok) atomicAdd(&hist[index], 1u);
    }
    __syncthreads();
    __shared__ double red[256];
    double acc = 0;
    for (int k = threadIdx.x; k < dims; k += blockDim.x) {
        const uint32_t c = hist[k];
        counts[(size_t)i * dims + k] = (uint16_t)c;
        acc += (double)c * (double)c;
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) { if ((int)threadIdx.x < st) red[threadIdx.x] += red[threadIdx.x + st]; __syncthreads(); }
    if (threadIdx.x == 0) norm2[i] = red[0];
}

// 32 x 32 tile of the Gram matrix per CTA (16 x 16 threads, 2 x 2 outputs each), dims in chunks of 64
__global__ void __launch_bounds__(256)
angle_gram_kernel(const uint16_t* __restrict__ counts, const double* __restrict__ norm2, const uint32_t* __restrict__ off,
                  int N, int dims, int ml_flag, double* __restrict__ dist, double* __restrict__ var) {
    __shared__ uint16_t A[32][66], B[32][66];
    const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    long long acc[2][2] = {{0, 0}, {0, 0}};
    for (int k0 = 0; k0 < dims; k0 += 64) {
        for (int t = threadIdx.x; t < 32 * 64; t += 256) {
            const int r = t >> 6, k = t & 63;
            A[r][k] = (i0 + r < N && k0 + k < dims) ? counts[(size_t)(i0 + r) * dims + k0 + k] : (uint16_t)0;
            B[r][k] = (j0 + r < N && k0 + k < dims) ? counts[(size_t)(j0 + r) * dims + k0 + k] : (uint16_t)0;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < 64; ++k) {
            // counts reach 65534 (sequences up to 65535 long): the product needs 32 unsigned bits, the sum 64
            const uint32_t a0 = A[ty][k], a1 = A[ty + 16][k], b0 = B[tx][k], b1 = B[tx + 16][k];
            acc[0][0] += (long long)(a0 * b0); acc[0][1] += (long long)(a0 * b1); acc[1][0] += (long long)(a1 * b0); acc[1][1] += (long long)(a1 * b1);
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int i = i0 + ty + 16 * a, j = j0 + tx + 16 * b;
            if (i >= N || j >= N) continue;
            double c = ((double)acc[a][b] * (1.0 / sqrt(norm2[i]))) * (1.0 / sqrt(norm2[j]));
            double d = -1.0 * log((c * c + 0.4) / 1.4);
            if (!ml_flag) { d = exp(d); d = -0.5 * (5.0 * d - sqrt(45.0 * d * d - 20.0 * d)) * (1.0 / d); }
            const double li = (double)(off[i + 1] - off[i]), lj = (double)(off[j + 1] - off[j]);
            double v = (1.0 / ((li + lj) / 2)) * d;
            v = fmax(v, 1e-5);
            dist[(size_t)i * N + j] = d; var[(size_t)i * N + j] = v;
        }
}

int launch_angle(cudaStream_t st, int n, int K, int dims, int N, const uint8_t* d_codes, const uint32_t* d_off,
                 uint16_t* d_counts, double* d_norm2, int ml_flag, double* d_dist, double* d_var) {
    if (N == 0) return 0;
    angle_count_kernel<<<N, 256, (size_t)dims * 4, st>>>(d_codes, d_off, n, K, dims, d_counts, d_norm2);
    PGM_CUDA(cudaGetLastError());
    dim3 grid((N + 31) / 32, (N + 31) / 32);
    angle_gram_kernel<<<grid, 256, 0, st>>>(d_counts, d_norm2, d_off, N, dims, ml_flag, d_dist, d_var);
    PGM_CUDA(cudaGetLastError());
    return 0;
}
