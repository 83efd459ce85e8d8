// Standalone hardware probe (not part of the product): validates the UMMA descriptor conventions used by
// csrc/tc_ptx.cuh on a real B200 and measures three rates the coupling-kernel design depends on:
//   (1) tcgen05.mma kind::f16 throughput with operands resident in shared memory,
//   (2) L2 -> shared bulk-TMA bandwidth per SM when every SM streams the same (L2-resident) weight blob,
//   (3) MUFU throughput for tanh.approx.f32 and for the ex2+rcp formulation.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 tools/tc_probe.cu -o tools/bin/tc_probe
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_fp16.h>
#include "../gradient-boosted-normalizing-flows_b200/csrc/tc_ptx.cuh"

using namespace gbnf::ptx;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, int version) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)version << 46;
  return d;
}

// A: K/16 slabs of [16 row groups][2][8][8] halves (4096 B); B: K/16 slabs of [N/8][2][8][8] halves (N*32 B)
__global__ void __launch_bounds__(192, 1) umma_probe(const __half* __restrict__ A, const __half* __restrict__ B, float* __restrict__ D,
                                                      int N, int K, uint32_t lbo, uint32_t sbo, int version, int iters, int* err) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t full_bar, done_bar;
  __shared__ uint32_t tmem_base;
  const int ks = K / 16;
  const uint32_t a_bytes = ks * 4096u, b_bytes = ks * (uint32_t)N * 32u;
  unsigned char* sA = smem;
  unsigned char* sB = smem + a_bytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&full_bar, 1); mbar_init(&done_bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&full_bar, a_bytes + b_bytes);
    tma_bulk_g2s(sA, A, a_bytes, &full_bar);
    tma_bulk_g2s(sB, B, b_bytes, &full_bar);
  }
  if (warp == 1 && lane == 0) {
    mbar_wait(&full_bar, 0, err, 1);
    tc_fence_after();
    const uint32_t idesc = make_idesc_f16(128, N);
    for (int it = 0; it < iters; ++it) {
      for (int k = 0; k < ks; ++k) {
        uint64_t da = make_desc(smem_u32(sA + k * 4096), lbo, sbo, version);
        uint64_t db = make_desc(smem_u32(sB + (size_t)k * N * 32), lbo, sbo, version);
        umma_f16(tbase, da, db, idesc, (it > 0 || k > 0) ? 1u : 0u);
      }
    }
    umma_commit(&done_bar);
  }
  if (warp >= 2) {
    mbar_wait(&done_bar, 0, err, 2);
    tc_fence_after();
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    if (D != nullptr && blockIdx.x == 0) {
      for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tbase + ((uint32_t)(quad * 32) << 16) + c0, r);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) D[(size_t)row * N + c0 + j] = __uint_as_float(r[j]);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}

// every CTA streams the same `bytes`-long blob through a 4 x 16 KB ring, `passes` times
__global__ void __launch_bounds__(64, 1) l2_stream_probe(const unsigned char* __restrict__ blob, size_t bytes, int passes, int* err,
                                                          unsigned long long* sink, int groups, int depth) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t full[4];
  constexpr uint32_t CH = 16384;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&full[i], 1); fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    const size_t nch = bytes / CH;
    const size_t total = nch * passes;
    // rotate the starting chunk per CTA so that the CTAs do not all hit the same L2 slice at the same time
    size_t issued = 0, waited = 0;
    uint32_t phase[4] = {0, 0, 0, 0};
    // groups == 0: every CTA starts at its own pseudo-random chunk; groups >= 1: CTAs of the same group walk the blob in
    // lock-step (groups == 1: all 148 CTAs request the same 16 KB chunk at the same time, as the coupling kernel does)
    const size_t start = groups == 0 ? (blockIdx.x * 7919u) % nch : ((size_t)(blockIdx.x % groups) * nch) / groups;
    while (waited < total) {
      while (issued < total && issued - waited < (size_t)depth) {
        int s = issued & 3;
        mbar_arrive_expect_tx(&full[s], CH);
        tma_bulk_g2s(smem + s * CH, blob + ((start + issued) % nch) * CH, CH, &full[s]);
        ++issued;
      }
      int s = waited & 3;
      mbar_wait(&full[s], phase[s], err, 3);
      phase[s] ^= 1;
      ++waited;
    }
    if (sink) atomicAdd(sink, (unsigned long long)smem[0]);
  }
}

template <int MODE>
__global__ void mufu_probe(float* out, int iters) {
  float v[8];
  for (int i = 0; i < 8; ++i) v[i] = 0.001f * (threadIdx.x + i) - 0.5f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) { asm volatile("tanh.approx.f32 %0, %0;" : "+f"(v[i])); v[i] += 0.25f; }
      else {
        float e, r;
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v[i] * 2.8853900817779268f));
        asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
        v[i] = fmaf(-2.0f, r, 1.25f);
      }
    }
  }
  float s = 0; for (int i = 0; i < 8; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static void pack_canonical(const std::vector<float>& M, int rows, int K, std::vector<__half>& out) {
  out.assign((size_t)rows * K, __float2half(0.f));
  for (int r = 0; r < rows; ++r)
    for (int k = 0; k < K; ++k) {
      size_t slab = k / 16, kc = (k / 8) % 2, e = k % 8, grp = r / 8, rr = r % 8;
      out[slab * (size_t)rows * 16 + grp * 128 + kc * 64 + rr * 8 + e] = __float2half(M[(size_t)r * K + k]);
    }
}

int main() {
  int dev = 0; CK(cudaSetDevice(dev));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
  printf("device %s sm_%d%d, %d SMs\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  int* err; CK(cudaMalloc(&err, 4)); CK(cudaMemset(err, 0, 4));
  const int nsm = prop.multiProcessorCount;

  // ---- (0) descriptor convention / correctness ------------------------------------------------------------
  for (int N : {64, 256}) {
    const int K = 64;
    std::vector<float> A(128 * K), B((size_t)N * K);
    srand(1);
    for (auto& v : A) v = __half2float(__float2half((rand() / (float)RAND_MAX - 0.5f)));
    for (auto& v : B) v = __half2float(__float2half((rand() / (float)RAND_MAX - 0.5f)));
    std::vector<__half> Ap, Bp; pack_canonical(A, 128, K, Ap); pack_canonical(B, N, K, Bp);
    __half *dA, *dB; float* dD;
    CK(cudaMalloc(&dA, Ap.size() * 2)); CK(cudaMalloc(&dB, Bp.size() * 2)); CK(cudaMalloc(&dD, 128 * N * 4));
    CK(cudaMemcpy(dA, Ap.data(), Ap.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, Bp.data(), Bp.size() * 2, cudaMemcpyHostToDevice));
    size_t smem = (K / 16) * (4096 + N * 32);
    CK(cudaFuncSetAttribute(umma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    struct { uint32_t lbo, sbo; int ver; } conv[] = {{128, 256, 1}, {256, 128, 1}, {128, 256, 0}};
    for (auto cv : conv) {
      CK(cudaMemset(dD, 0, 128 * N * 4));
      umma_probe<<<1, 192, smem>>>(dA, dB, dD, N, K, cv.lbo, cv.sbo, cv.ver, 1, err);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("N=%d lbo=%u sbo=%u ver=%d: launch failed: %s\n", N, cv.lbo, cv.sbo, cv.ver, cudaGetErrorString(e)); return 3; }
      std::vector<float> D(128 * N);
      CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
      double maxerr = 0;
      for (int r = 0; r < 128; ++r) for (int n = 0; n < N; ++n) {
        double ref = 0; for (int k = 0; k < K; ++k) ref += (double)A[r * K + k] * B[(size_t)n * K + k];
        maxerr = fmax(maxerr, fabs(ref - D[r * N + n]));
      }
      printf("UMMA M=128 N=%d K=%d lbo=%u sbo=%u version=%d : max abs err %.3e  %s\n", N, K, cv.lbo, cv.sbo, cv.ver, maxerr,
             maxerr < 1e-3 ? "MATCH" : "mismatch");
    }
    // ---- (1) MMA throughput, operands resident in smem, all SMs
    if (N == 256) {
      const int iters = 4000;
      cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
      umma_probe<<<nsm, 192, smem>>>(dA, dB, nullptr, N, K, 128, 256, 1, 100, err); CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0));
      umma_probe<<<nsm, 192, smem>>>(dA, dB, nullptr, N, K, 128, 256, 1, iters, err);
      CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      double fl = 2.0 * 128 * N * K * (double)iters * nsm;
      printf("MMA rate (smem-resident, M128 N256 K16 x %d per SM): %.1f TFLOP/s (%.3f ms)\n", iters * K / 16, fl / ms / 1e9, ms);
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
  }
  // ---- (2) L2 -> smem bulk TMA bandwidth ------------------------------------------------------------------
  {
    const size_t bytes = 24u << 20;
    unsigned char* blob; CK(cudaMalloc(&blob, bytes)); CK(cudaMemset(blob, 1, bytes));
    unsigned long long* sink; CK(cudaMalloc(&sink, 8));
    CK(cudaFuncSetAttribute(l2_stream_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    l2_stream_probe<<<nsm, 64, 65536>>>(blob, bytes, 1, err, sink, 0, 4); CK(cudaDeviceSynchronize());
    const int passes = 4;
    for (int depth : {4, 2}) for (int groups : {0, 1, 2, 4, 8, 37}) {
      CK(cudaEventRecord(e0));
      l2_stream_probe<<<nsm, 64, 65536>>>(blob, bytes, passes, err, sink, groups, depth);
      CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      double tot = (double)bytes * passes * nsm;
      printf("L2->smem bulk TMA, %d CTAs, ring %dx16KB, %zu MB blob, lock-step groups=%d: %.1f GB/s total, %.1f GB/s per SM (%.3f ms)\n", nsm,
             depth, bytes >> 20, groups, tot / ms / 1e6, tot / ms / 1e6 / nsm, ms);
    }
    cudaFree(blob);
  }
  // ---- (3) MUFU rates ---------------------------------------------------------------------------------------
  {
    float* out; CK(cudaMalloc(&out, (size_t)nsm * 8 * 256 * 4));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int iters = 4096;
    for (int mode = 0; mode < 2; ++mode) {
      if (mode == 0) mufu_probe<0><<<nsm * 8, 256>>>(out, 16); else mufu_probe<1><<<nsm * 8, 256>>>(out, 16);
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0));
      if (mode == 0) mufu_probe<0><<<nsm * 8, 256>>>(out, iters); else mufu_probe<1><<<nsm * 8, 256>>>(out, iters);
      CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      double n = (double)nsm * 8 * 256 * 8 * iters;
      printf("%s: %.1f G tanh/s total, %.2f per ns per SM (%.3f ms)\n", mode == 0 ? "tanh.approx.f32" : "tanh via ex2+rcp", n / ms / 1e6,
             n / ms / 1e6 / nsm, ms);
    }
  }
  int herr = 0; CK(cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost));
  printf("error flag: %d\nPROBE DONE\n", herr);
  return 0;
}
