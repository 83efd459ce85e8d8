// Hardware probe #6: why does a tcgen05.mma cost ~35 cycles more in the coupling kernel than in a tight loop?
// N = 128 / 64, A from TMEM, B slabs in shared memory.  Variants:
//   0: tight loop over 8 precomputed descriptors (as tc_probe2)        -> reference
//   1: B walks 32 distinct slabs (128 KB / 64 KB), descriptor = base + i * step computed in the loop, loop NOT unrolled
//   2: as 1, unrolled by 8 (what ptxas makes of the kernel's issue blocks)
//   3: as 2, but blocks of 8 MMAs separated by tcgen05.commit to an mbarrier (as the kernel's ring stages)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/tc_probe6.cu -o tools/bin/tc_probe6
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_fp16.h>
#include "../gradient-boosted-normalizing-flows_b200/csrc/tc_ptx.cuh"
using namespace gbnf::ptx;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

template <int N, int VARIANT>
__global__ void __launch_bounds__(64, 1) probe(long long* out, int* err, int iters) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar_done, bar_stage;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar_done, 1); mbar_init(&bar_stage, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tbase = tmem_base;
  if (warp == 1) {
    const uint64_t bd0 = make_smem_desc(smem_u32(smem));
    const uint32_t idesc = make_idesc_f16(128, N);
    const uint32_t bstep = (uint32_t)N * 2u;             // (N * 32 B) >> 4
    const uint32_t d = tbase + 256u, a0 = tbase;
    const long long t0 = clock64();
    if (elect_one()) {
      if (VARIANT == 0) {
        uint64_t db[8]; uint32_t ta[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { db[k] = bd0 + (uint64_t)(k * bstep); ta[k] = a0 + 8u * k; }
        for (int it = 0; it < iters * 4; ++it) {
#pragma unroll
          for (int k = 0; k < 8; ++k) umma_f16_ts(d, ta[k], db[k], idesc, 1u);
        }
      } else if (VARIANT == 1) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll 1
          for (int i = 0; i < 32; ++i) umma_f16_ts(d, a0 + 8u * i, bd0 + (uint64_t)(i * bstep), idesc, 1u);
        }
      } else {
        for (int it = 0; it < iters; ++it) {
#pragma unroll 1
          for (int s = 0; s < 4; ++s) {
            const uint64_t bs = bd0 + (uint64_t)(s * 8 * bstep);
            const uint32_t as = a0 + 64u * s;
#pragma unroll
            for (int i = 0; i < 8; ++i) umma_f16_ts(d, as + 8u * i, bs + (uint64_t)(i * bstep), idesc, 1u);
            if (VARIANT == 3) umma_commit(&bar_stage);
          }
        }
      }
      umma_commit(&bar_done);
    }
    __syncwarp();
    mbar_wait(&bar_done, 0, err, 2);
    if (lane == 0) out[0] = clock64() - t0;
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}

template <int N, int V> void run(const char* name, long long* d, int* err) {
  const int iters = 500;
  CK(cudaFuncSetAttribute(probe<N, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  probe<N, V><<<1, 64, 160 * 1024>>>(d, err, iters);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: failed %s\n", name, cudaGetErrorString(e)); exit(3); }
  long long h; CK(cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost));
  printf("N=%3d %-70s: %.1f cycles per MMA\n", N, name, (double)h / (iters * 32.0));
}

int main() {
  CK(cudaSetDevice(0));
  long long* d; CK(cudaMalloc(&d, 64)); int* err; CK(cudaMalloc(&err, 4)); CK(cudaMemset(err, 0, 4));
  run<128, 0>("tight loop, 8 precomputed descriptors", d, err);
  run<128, 1>("32 distinct B slabs, descriptor arithmetic in a rolled loop", d, err);
  run<128, 2>("32 distinct B slabs, blocks of 8 unrolled", d, err);
  run<128, 3>("32 distinct B slabs, blocks of 8 unrolled + commit per block", d, err);
  run<64, 0>("tight loop, 8 precomputed descriptors", d, err);
  run<64, 1>("32 distinct B slabs, descriptor arithmetic in a rolled loop", d, err);
  run<64, 2>("32 distinct B slabs, blocks of 8 unrolled", d, err);
  run<64, 3>("32 distinct B slabs, blocks of 8 unrolled + commit per block", d, err);
  int herr; CK(cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost));
  printf("error flag %d\nPROBE6 DONE\n", herr);
  return 0;
}
