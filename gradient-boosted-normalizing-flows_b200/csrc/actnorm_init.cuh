// Data-dependent ActNorm initialisation (models/layers.py:473-486): bias = -mean_B(x), logs = log(scale / (sqrt(mean_B((x +
// bias)^2)) + 1e-6)).  Two streaming column reductions over x[B, D] with fp64 accumulators (HBM-bound; x is read twice, as the
// reference does, because the second moment is taken about the ROUNDED fp32 mean).
#pragma once
#include "common.cuh"

namespace gbnf {

constexpr int kAnThreads = 256;

// acc[c] += sum_b f(x[b, c] + shift[c]),  f = identity (SQ = false) or square (SQ = true).  Thread t owns column t % Dp of the
// rows t / Dp, t / Dp + rows_per_pass, ...  (Dp = D rounded up to a power of two <= 256: a warp reads whole rows, coalesced).
template <bool SQ>
__global__ void __launch_bounds__(kAnThreads) actnorm_colsum_kernel(const float* __restrict__ x, long long B, int D, int Dp,
                                                                    const float* __restrict__ shift, double* __restrict__ acc) {
  __shared__ double part[kAnThreads];
  const int c = threadIdx.x % Dp, r0 = threadIdx.x / Dp, rpb = kAnThreads / Dp;
  double s = 0.0;
  if (c < D) {
    const float sh = SQ ? shift[c] : 0.f;
    for (long long b = (long long)blockIdx.x * rpb + r0; b < B; b += (long long)gridDim.x * rpb) {
      const float v = __ldg(x + b * D + c) + sh;          // fp32 add, as `sample + bias` in the reference
      s += SQ ? (double)(v * v) : (double)v;              // fp32 square, as `** 2`
    }
  }
  part[threadIdx.x] = s;
  __syncthreads();
  if (r0 == 0 && c < D) {
    for (int r = 1; r < rpb; ++r) s += part[r * Dp + c];
    atomicAdd(acc + c, s);
  }
}
__global__ void actnorm_bias_kernel(const double* __restrict__ acc, long long B, int D, float* __restrict__ bias) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < D) bias[c] = -(float)(acc[c] / (double)B);
}
__global__ void actnorm_logs_kernel(const double* __restrict__ acc, long long B, int D, float scale, float* __restrict__ logs) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < D) logs[c] = logf(scale / (sqrtf((float)(acc[c] / (double)B)) + 1e-6f));
}

}  // namespace gbnf
