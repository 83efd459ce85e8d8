// Standalone hardware probe #8 (not part of the product; written at the end of round 1, NOT yet run on a GPU - it only
// compiles so far).  Question for round 2: does a CTA pair (tcgen05 cta_group::2) halve the weight stream per SM for the coupling
// GEMMs, and at what MMA rate?
//   One 2-CTA cluster computes D[256 x 128] = A[256 x K] * B[128 x K]^T with ONE instruction stream issued by the leader:
//   each CTA holds ITS 128 rows of A and HALF of B (64 of the 128 n-rows) in its own shared memory, at the same offsets, in the
//   canonical K-major no-swizzle layout the product kernels use; each CTA receives its 128 rows of D in its own TMEM.
//   The probe (1) checks D against a CPU product for both CTAs - i.e. checks the assumption about which half of B lives where -
//   and (2) times a dependent chain of 64 such MMAs (K = 16 each) on the leader.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 tools/tc_probe8.cu -o tools/bin/tc_probe8
// Run  : timeout 60 tools/bin/tc_probe8
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_fp16.h>
#include <cooperative_groups.h>
#include "../gradient-boosted-normalizing-flows_b200/csrc/tc_ptx.cuh"

namespace cg = cooperative_groups;
using namespace gbnf::ptx;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

constexpr int kM = 256, kN = 128, kK = 64, kSlabs = kK / 16;

// canonical K-major no-swizzle image of a [rows x 16] k-slab: LBO = 128 (between the two 8-element k-chunks), SBO = 256
__host__ __device__ inline uint32_t slab_off(int row, int k) { return (uint32_t)((row >> 3) * 256 + ((k >> 3) & 1) * 128 + (row & 7) * 16 + (k & 7) * 2); }

__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma2_commit(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}

// A: [256][K] row-major fp16 (global), B: [128][K] row-major fp16 (global), D: [256][128] fp32 (global)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) pair_mma_probe(const __half* __restrict__ A, const __half* __restrict__ B,
                                                                                   float* __restrict__ D, long long* cycles, int* err) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t done_bar, time_bar;
  __shared__ uint32_t tmem_base;
  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t rank = cluster.block_rank();
  unsigned char* sA = smem;                                   // kSlabs x [128 x 16]  = kSlabs * 4096 B
  unsigned char* sB = smem + kSlabs * 4096;                   // kSlabs x [ 64 x 16]  = kSlabs * 2048 B  (this CTA's half of N)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 128 * kK; i += blockDim.x) {
    const int r = i / kK, k = i % kK;
    *reinterpret_cast<__half*>(sA + (k >> 4) * 4096 + slab_off(r, k & 15)) = A[(size_t)(rank * 128 + r) * kK + k];
  }
  for (int i = threadIdx.x; i < 64 * kK; i += blockDim.x) {
    const int n = i / kK, k = i % kK;
    *reinterpret_cast<__half*>(sB + (k >> 4) * 2048 + slab_off(n, k & 15)) = B[(size_t)(rank * 64 + n) * kK + k];
  }
  if (threadIdx.x == 0) { mbar_init(&done_bar, 1); mbar_init(&time_bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc2(&tmem_base, 128);
  fence_proxy_async_smem();
  tc_fence_before();
  cluster.sync();
  tc_fence_after();
  const uint32_t tbase = tmem_base;
  const uint32_t idesc = make_idesc_f16(kM, kN);
  const uint64_t a_desc = make_smem_desc(smem_u32(sA)), b_desc = make_smem_desc(smem_u32(sB));

  // (1) one product, K = 64
  if (rank == 0 && warp == 0) {
    if (elect_one()) {
      for (int s = 0; s < kSlabs; ++s)
        umma2_f16(tbase, a_desc + (uint64_t)(s * (4096 >> 4)), b_desc + (uint64_t)(s * (2048 >> 4)), idesc, s > 0 ? 1u : 0u);
      umma2_commit(&done_bar, (uint16_t)0b11);
    }
    __syncwarp();
  }
  mbar_wait(&done_bar, 0u, err, 80);
  tc_fence_after();
  {
    const uint32_t lane_base = tbase + ((uint32_t)(warp * 32) << 16);
    const int row = rank * 128 + warp * 32 + lane;
    for (int c = 0; c < kN; c += 32) {
      uint32_t r[32];
      tmem_ld32(lane_base + (uint32_t)c, r);
      tmem_ld_wait();
      for (int i = 0; i < 32; ++i) D[(size_t)row * kN + c + i] = __uint_as_float(r[i]);
    }
  }
  tc_fence_before();
  cluster.sync();
  tc_fence_after();

  // (2) rate: 64 dependent MMAs (same accumulator), commit, wait
  if (rank == 0 && warp == 0) {
    const long long t0 = clock64();
    if (elect_one()) {
      for (int i = 0; i < 64; ++i)
        umma2_f16(tbase, a_desc + (uint64_t)((i & 3) * (4096 >> 4)), b_desc + (uint64_t)((i & 3) * (2048 >> 4)), idesc, 1u);
      umma2_commit(&time_bar, (uint16_t)0b11);
    }
    __syncwarp();
    const long long t1 = clock64();
    mbar_wait(&time_bar, 0u, err, 81);
    const long long t2 = clock64();
    if (lane == 0) { cycles[0] = t1 - t0; cycles[1] = t2 - t0; }
  } else {
    mbar_wait(&time_bar, 0u, err, 82);
  }
  tc_fence_before();
  cluster.sync();
  if (warp == 0) tmem_dealloc2(tbase, 128);
}

int main() {
  CK(cudaSetDevice(0));
  std::vector<__half> hA((size_t)kM * kK), hB((size_t)kN * kK);
  std::vector<float> fA(hA.size()), fB(hB.size());
  srand(7);
  for (size_t i = 0; i < hA.size(); ++i) { fA[i] = (float)((rand() % 9) - 4); hA[i] = __float2half(fA[i]); }
  for (size_t i = 0; i < hB.size(); ++i) { fB[i] = (float)((rand() % 7) - 3); hB[i] = __float2half(fB[i]); }
  __half *dA, *dB; float* dD; long long* dC; int* dE;
  CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dD, (size_t)kM * kN * 4));
  CK(cudaMalloc(&dC, 2 * sizeof(long long))); CK(cudaMalloc(&dE, sizeof(int)));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0, (size_t)kM * kN * 4)); CK(cudaMemset(dC, 0, 2 * sizeof(long long))); CK(cudaMemset(dE, 0, sizeof(int)));
  const int smem = kSlabs * (4096 + 2048);
  CK(cudaFuncSetAttribute(pair_mma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  pair_mma_probe<<<2, 128, smem>>>(dA, dB, dD, dC, dE);
  CK(cudaDeviceSynchronize());
  std::vector<float> hD((size_t)kM * kN);
  long long hc[2]; int he = 0;
  CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hc, dC, sizeof(hc), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&he, dE, sizeof(int), cudaMemcpyDeviceToHost));
  int bad[2] = {0, 0};
  for (int r = 0; r < kM; ++r)
    for (int n = 0; n < kN; ++n) {
      float ref = 0.f;
      for (int k = 0; k < kK; ++k) ref += fA[(size_t)r * kK + k] * fB[(size_t)n * kK + k];
      if (std::fabs(ref - hD[(size_t)r * kN + n]) > 1e-3f) ++bad[r / 128];
    }
  printf("# cta_group::2, M = 256 (128 rows per CTA), N = 128 (64 n-rows of B per CTA), K = %d\n", kK);
  printf("mismatches: CTA 0 rows %d / %d, CTA 1 rows %d / %d   (0 / 0 confirms the operand split)\n", bad[0], 128 * kN, bad[1], 128 * kN);
  printf("64 dependent MMAs (K = 16): issue %lld cycles, until complete %lld cycles = %.1f cycles per MMA for 256 x 128 x 16\n", hc[0], hc[1],
         (double)hc[1] / 64.0);
  printf("error flag: %d\n", he);
  return 0;
}
