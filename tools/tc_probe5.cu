// Hardware probe #5: tanh.approx.f16x2 -- throughput (elements per clock per SM) and accuracy against the current epilogue
// path (tanh.approx.f32 followed by round-to-fp16).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/tc_probe5.cu -o tools/bin/tc_probe5
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cstdint>
#include <cuda_fp16.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t tanh_h2(uint32_t v) { uint32_t r; asm volatile("tanh.approx.f16x2 %0, %1;" : "=r"(r) : "r"(v)); return r; }
__device__ __forceinline__ float tanh_f32(float v) { float r; asm volatile("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
__device__ __forceinline__ uint32_t pack(float a, float b) { __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }

template <int MODE>
__global__ void rate(uint32_t* out, int iters) {
  uint32_t v[8];
  float f[8];
  for (int i = 0; i < 8; ++i) { v[i] = 0x3c003800u + threadIdx.x + i; f[i] = 0.001f * (threadIdx.x + i); }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) v[i] = tanh_h2(v[i]) + 0x00010001u;
      else { f[i] = tanh_f32(f[i]) + 0.25f; }
    }
  }
  uint32_t s = 0; for (int i = 0; i < 8; ++i) s += v[i] + __float_as_uint(f[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void acc(const float* x, int n, float* y_h2, float* y_f32) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t r = tanh_h2(pack(x[i], x[i]));
  __half2 h = *reinterpret_cast<__half2*>(&r);
  y_h2[i] = __low2float(h);
  y_f32[i] = __half2float(__float2half_rn(tanh_f32(x[i])));
}

int main() {
  CK(cudaSetDevice(0));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int nsm = prop.multiProcessorCount;
  uint32_t* out; CK(cudaMalloc(&out, (size_t)nsm * 8 * 256 * 4));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  int clk; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
  for (int mode = 0; mode < 2; ++mode) {
    const int iters = 8192;
    if (mode == 0) rate<0><<<nsm * 8, 256>>>(out, 64); else rate<1><<<nsm * 8, 256>>>(out, 64);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    if (mode == 0) rate<0><<<nsm * 8, 256>>>(out, iters); else rate<1><<<nsm * 8, 256>>>(out, iters);
    CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    double ops = (double)nsm * 8 * 256 * 8 * iters;
    printf("%s: %.2f MUFU ops per ns per SM = %.2f tanh per ns per SM (%.3f ms)\n", mode == 0 ? "tanh.approx.f16x2" : "tanh.approx.f32  ",
           ops / ms / 1e6 / nsm, ops * (mode == 0 ? 2 : 1) / ms / 1e6 / nsm, ms);
  }
  // accuracy on [-8, 8]
  const int n = 1 << 20;
  std::vector<float> hx(n);
  for (int i = 0; i < n; ++i) hx[i] = -8.f + 16.f * (i + 0.5f) / n;
  float *dx, *d1, *d2; CK(cudaMalloc(&dx, n * 4)); CK(cudaMalloc(&d1, n * 4)); CK(cudaMalloc(&d2, n * 4));
  CK(cudaMemcpy(dx, hx.data(), n * 4, cudaMemcpyHostToDevice));
  acc<<<n / 256, 256>>>(dx, n, d1, d2); CK(cudaDeviceSynchronize());
  std::vector<float> y1(n), y2(n);
  CK(cudaMemcpy(y1.data(), d1, n * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(y2.data(), d2, n * 4, cudaMemcpyDeviceToHost));
  double m1 = 0, m2 = 0, r1 = 0, r2 = 0, s1 = 0, s2 = 0;
  for (int i = 0; i < n; ++i) {
    double t = tanh((double)hx[i]);
    double a = fabs(y1[i] - t), b = fabs(y2[i] - t);
    m1 = fmax(m1, a); m2 = fmax(m2, b); s1 += a * a; s2 += b * b;
    if (fabs(t) > 1e-3) { r1 = fmax(r1, a / fabs(t)); r2 = fmax(r2, b / fabs(t)); }
  }
  printf("accuracy vs tanh (x in [-8,8], 2^20 points): f16x2 path max abs %.3e rms %.3e max rel %.3e | f32+round path max abs %.3e rms %.3e max rel %.3e\n",
         m1, sqrt(s1 / n), r1, m2, sqrt(s2 / n), r2);
  printf("PROBE5 DONE\n");
  return 0;
}
