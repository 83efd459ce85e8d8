// Standalone hardware probe #7 (not part of the product; written at the end of round 1, NOT yet run on a GPU - it only
// compiles so far).  Questions the h = 1024 tensor-core path depends on:
//   (A) where does a cta_group::1, M = 64 tcgen05.mma put its accumulator rows in TMEM?  (A 64-row half tile would let two
//       chains share TMEM if - and only if - its accumulator occupies 64 lanes x N columns or 128 lanes x N/2 columns.)
//       Method: D = A x B with A[r][0] = r + 1, B[n][0] = 1 (K = 16), then dump all 128 lanes x 64 columns.
//   (B) distributed shared memory between the two CTAs of a cluster: cycles for CTA 1 to push 64 KB into CTA 0's shared
//       memory with ONE cp.async.bulk.shared::cluster.shared::cta (completion on CTA 0's mbarrier), and for CTA 0 to pull the
//       same bytes with ld.shared::cluster.v4 from 128 threads.  (The 2-CTA hidden-unit split exchanges 64 KB of partial sums
//       per layer-2 chunk: it needs >= ~32 B/clk to hide behind a 3.1 k-cycle chunk.)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 tools/tc_probe7.cu -o tools/bin/tc_probe7
// Run  : timeout 60 tools/bin/tc_probe7
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>
#include <cooperative_groups.h>
#include "../gradient-boosted-normalizing-flows_b200/csrc/tc_ptx.cuh"

namespace cg = cooperative_groups;
using namespace gbnf::ptx;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

// byte offset of element (row, k) in a canonical K-major, no-swizzle operand image with LBO = 128, SBO = 256 (K = 16 slab)
__host__ __device__ inline uint32_t canon_off(int row, int k) { return (uint32_t)((row >> 3) * 256 + ((k >> 3) & 1) * 128 + (row & 7) * 16 + (k & 7) * 2); }

// ---- (A) ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) m64_layout_probe(float* __restrict__ dump /* [128 lanes][64 cols] */, int* err) {
  __shared__ __align__(1024) unsigned char sA[64 * 32];     // 64 rows x 16 k fp16
  __shared__ __align__(1024) unsigned char sB[64 * 32];     // 64 n    x 16 k fp16
  __shared__ uint64_t done_bar;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 64 * 16; i += blockDim.x) {
    const int r = i >> 4, k = i & 15;
    *reinterpret_cast<__half*>(sA + canon_off(r, k)) = __float2half(k == 0 ? (float)(r + 1) : 0.f);
    *reinterpret_cast<__half*>(sB + canon_off(r, k)) = __float2half(k == 0 ? 1.f : 0.f);
  }
  if (threadIdx.x == 0) { mbar_init(&done_bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_base, 128);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base;
  // zero the whole 128 x 64 window first so that untouched lanes read back as exact zeros
  {
    uint32_t z[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) z[i] = 0u;
    const uint32_t lane_base = tbase + ((uint32_t)(warp * 32) << 16);
    tmem_st32(lane_base, z);
    tmem_st32(lane_base + 32u, z);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) {
    if (elect_one()) {
      umma_f16(tbase, make_smem_desc(smem_u32(sA)), make_smem_desc(smem_u32(sB)), make_idesc_f16(64, 64), 0u);
      umma_commit(&done_bar);
    }
    __syncwarp();
  }
  mbar_wait(&done_bar, 0u, err, 70);
  tc_fence_after();
  {
    uint32_t r[32];
    const uint32_t lane_base = tbase + ((uint32_t)(warp * 32) << 16);
    const int row = warp * 32 + lane;
    tmem_ld32(lane_base, r);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) dump[row * 64 + i] = __uint_as_float(r[i]);
    tmem_ld32(lane_base + 32u, r);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) dump[row * 64 + 32 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 128);
}

// ---- (B) ------------------------------------------------------------------------------------------------------------------
constexpr uint32_t kXferBytes = 64 * 1024;

__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(cta_rank));
  return r;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) dsmem_probe(long long* cycles /* [2] */, float* check, int* err) {
  extern __shared__ __align__(1024) unsigned char buf[];    // kXferBytes
  __shared__ uint64_t bar;
  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t rank = cluster.block_rank();
  float* f = reinterpret_cast<float*>(buf);
  for (uint32_t i = threadIdx.x; i < kXferBytes / 4; i += blockDim.x) f[i] = (rank == 1) ? (float)i : -1.f;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  fence_proxy_async_smem();
  cluster.sync();

  // (B1) push: one elected thread of CTA 1 issues a bulk copy own smem -> CTA 0's smem, completion on CTA 0's mbarrier
  long long t0 = 0;
  if (rank == 0 && threadIdx.x == 0) { mbar_arrive_expect_tx(&bar, kXferBytes); }
  cluster.sync();
  t0 = clock64();
  if (rank == 1 && threadIdx.x == 0) {
    const uint32_t dst = map_to_cta(smem_u32(buf), 0), rbar = map_to_cta(smem_u32(&bar), 0);
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "r"(smem_u32(buf)), "r"(kXferBytes), "r"(rbar)
                 : "memory");
  }
  if (rank == 0) {
    mbar_wait(&bar, 0u, err, 71);
    if (threadIdx.x == 0) { cycles[0] = clock64() - t0; check[0] = f[12345]; }
  }
  cluster.sync();

  // (B2) pull: the 128 threads of CTA 0 read CTA 1's buffer with 128-bit remote loads and accumulate
  if (rank == 0) {
    const uint32_t src = map_to_cta(smem_u32(buf), 1);
    float acc = 0.f;
    __syncthreads();
    const long long t1 = clock64();
    for (uint32_t off = threadIdx.x * 16u; off < kXferBytes; off += blockDim.x * 16u) {
      float4 v;
      asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(src + off));
      acc += (v.x + v.y) + (v.z + v.w);
    }
    __syncthreads();
    const long long t2 = clock64();
    if (threadIdx.x == 0) cycles[1] = t2 - t1;
    if (acc == 12345.678f) check[1] = acc;       // keep the loads alive
  }
  cluster.sync();                                 // CTA 1 must stay resident while CTA 0 reads its shared memory
}

int main() {
  int dev = 0;
  CK(cudaSetDevice(dev));
  int* err; CK(cudaMalloc(&err, sizeof(int))); CK(cudaMemset(err, 0, sizeof(int)));

  // (A)
  float* dump; CK(cudaMalloc(&dump, 128 * 64 * sizeof(float)));
  m64_layout_probe<<<1, 128>>>(dump, err);
  CK(cudaDeviceSynchronize());
  std::vector<float> h(128 * 64);
  CK(cudaMemcpy(h.data(), dump, h.size() * sizeof(float), cudaMemcpyDeviceToHost));
  printf("# (A) cta_group::1 M=64 N=64: TMEM lane -> accumulator row (value - 1) seen in column 0 / 32 / 63 (0 = untouched)\n");
  for (int lane = 0; lane < 128; ++lane) {
    const float a = h[lane * 64], b = h[lane * 64 + 32], c = h[lane * 64 + 63];
    if (a != 0.f || b != 0.f || c != 0.f) printf("lane %3d: col0 row %3.0f  col32 row %3.0f  col63 row %3.0f\n", lane, a - 1, b - 1, c - 1);
  }
  int touched = 0;
  for (int lane = 0; lane < 128; ++lane) { bool t = false; for (int c = 0; c < 64; ++c) t |= h[lane * 64 + c] != 0.f; touched += t; }
  printf("lanes touched: %d of 128\n", touched);

  // (B)
  long long* cyc; CK(cudaMalloc(&cyc, 2 * sizeof(long long))); CK(cudaMemset(cyc, 0, 2 * sizeof(long long)));
  float* chk; CK(cudaMalloc(&chk, 2 * sizeof(float))); CK(cudaMemset(chk, 0, 2 * sizeof(float)));
  CK(cudaFuncSetAttribute(dsmem_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kXferBytes));
  dsmem_probe<<<2, 128, kXferBytes>>>(cyc, chk, err);
  CK(cudaDeviceSynchronize());
  long long hc[2]; float hk[2]; int herr = 0;
  CK(cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hk, chk, sizeof(hk), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&herr, err, sizeof(int), cudaMemcpyDeviceToHost));
  printf("# (B) DSMEM, 64 KB between the CTAs of a 2-cluster\n");
  printf("push (cp.async.bulk shared::cta -> shared::cluster): %lld cycles = %.1f B/clk (check %s)\n", hc[0], (double)kXferBytes / (double)hc[0],
         hk[0] == 12345.f ? "ok" : "MISMATCH");
  printf("pull (ld.shared::cluster.v4, 128 threads)          : %lld cycles = %.1f B/clk\n", hc[1], (double)kXferBytes / (double)hc[1]);
  printf("error flag: %d\n", herr);
  return 0;
}
