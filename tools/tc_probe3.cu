// Hardware probe #3: cost of the control primitives on a single latency-bound warp (the MMA-issuer role):
// tcgen05.fence::after_thread_sync, elect.sync, __syncwarp, mbarrier.try_wait on a completed phase, tcgen05.commit,
// and the issue time of a block of 8 tcgen05.mma (N=64, A from TMEM) -- does the issuing thread block until they execute?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/tc_probe3.cu -o tools/bin/tc_probe3
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include "../gradient-boosted-normalizing-flows_b200/csrc/tc_ptx.cuh"
using namespace gbnf::ptx;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__global__ void __launch_bounds__(64, 1) prim_probe(long long* out, int* err) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar_done, bar_c[4];
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar_done, 1); for (int i = 0; i < 4; ++i) mbar_init(&bar_c[i], i == 2 ? 4 : 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tbase = tmem_base;
  if (warp == 1) {
    const int N = 1000;
    long long t0, t1;
    // (1) tcgen05.fence::after_thread_sync
    t0 = clock64(); for (int i = 0; i < N; ++i) tc_fence_after(); t1 = clock64(); if (lane == 0) out[0] = (t1 - t0);
    // (2) elect + syncwarp
    unsigned acc = 0;
    t0 = clock64(); for (int i = 0; i < N; ++i) { if (elect_one()) acc += i; __syncwarp(); } t1 = clock64(); if (lane == 0) out[1] = (t1 - t0);
    // (3) try_wait on a completed phase (bar_done: complete phase 0 first)
    if (lane == 0) mbar_arrive(&bar_done);
    __syncwarp();
    t0 = clock64(); for (int i = 0; i < N; ++i) { if (lane == 0) mbar_wait(&bar_done, 0, err, 1); __syncwarp(); } t1 = clock64(); if (lane == 0) out[2] = (t1 - t0);
    // (3b) all lanes poll
    t0 = clock64(); for (int i = 0; i < N; ++i) { mbar_wait(&bar_done, 0, err, 1); } t1 = clock64(); if (lane == 0) out[3] = (t1 - t0);
    // (4) tcgen05.commit alone (no MMAs outstanding), then wait for it
    t0 = clock64();
    for (int i = 0; i < N; ++i) { if (elect_one()) umma_commit(&bar_c[0]); __syncwarp(); mbar_wait(&bar_c[0], i & 1, err, 2); }
    t1 = clock64(); if (lane == 0) out[4] = (t1 - t0);
    // (5) issue 8 MMAs (N=64, A from TMEM, B = garbage smem) + commit: time to ISSUE vs time to COMPLETE
    const uint64_t bd = make_smem_desc(smem_u32(smem));
    const uint32_t idesc = make_idesc_f16(128, 64);
    long long iss = 0, comp = 0;
    for (int it = 0; it < 200; ++it) {
      t0 = clock64();
      if (elect_one()) {
#pragma unroll
        for (int i = 0; i < 8; ++i) umma_f16_ts(tbase + 64u, tbase + 256u + 8u * i, bd + (uint64_t)(i * 128), idesc, 1u);
        umma_commit(&bar_c[1]);
      }
      __syncwarp();
      t1 = clock64();
      mbar_wait(&bar_c[1], it & 1, err, 3);
      iss += t1 - t0; comp += clock64() - t0;
    }
    if (lane == 0) { out[5] = iss; out[6] = comp; }
    // (6) same, 4 blocks back to back before waiting (queue depth)
    iss = 0; comp = 0;
    for (int it = 0; it < 200; ++it) {
      t0 = clock64();
      for (int b = 0; b < 4; ++b) {
        if (elect_one()) {
#pragma unroll
          for (int i = 0; i < 8; ++i) umma_f16_ts(tbase + 64u, tbase + 256u + 8u * i, bd + (uint64_t)(i * 128), idesc, 1u);
          umma_commit(&bar_c[2]);
        }
        __syncwarp();
        if (b < 3) { /* consume the phase later */ }
      }
      t1 = clock64();
      mbar_wait(&bar_c[2], it & 1, err, 4);
      iss += t1 - t0; comp += clock64() - t0;
    }
    if (lane == 0) { out[7] = iss; out[8] = comp; out[9] = acc; }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}

int main() {
  CK(cudaSetDevice(0));
  long long* d; CK(cudaMalloc(&d, 16 * 8)); CK(cudaMemset(d, 0, 128));
  int* err; CK(cudaMalloc(&err, 4)); CK(cudaMemset(err, 0, 4));
  CK(cudaFuncSetAttribute(prim_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  prim_probe<<<1, 64, 65536>>>(d, err);
  CK(cudaDeviceSynchronize());
  long long h[16]; CK(cudaMemcpy(h, d, 128, cudaMemcpyDeviceToHost));
  printf("tcgen05.fence::after_thread_sync        : %.1f cycles\n", h[0] / 1000.0);
  printf("elect.sync + __syncwarp                 : %.1f cycles\n", h[1] / 1000.0);
  printf("lane-0 try_wait (done) + __syncwarp     : %.1f cycles\n", h[2] / 1000.0);
  printf("all-lane try_wait (done)                : %.1f cycles\n", h[3] / 1000.0);
  printf("commit (idle pipe) + wait               : %.1f cycles\n", h[4] / 1000.0);
  printf("8 x MMA N64 + commit: issue %.1f cycles, until complete %.1f cycles\n", h[5] / 200.0, h[6] / 200.0);
  printf("4 x (8 MMA + commit): issue %.1f cycles, until complete %.1f cycles\n", h[7] / 200.0, h[8] / 200.0);
  int herr; CK(cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost));
  printf("error flag %d\nPROBE3 DONE\n", herr);
  return 0;
}
