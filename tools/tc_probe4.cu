// Hardware probe #4: per-stage cost of the MMA-issuer loop of coupling_tc2.cuh in isolation.  A fake producer warp
// recycles ring slots (wait empty -> arrive full, no data movement), the issuer runs acquire + 8 x tcgen05.mma (N=64, A from
// TMEM) + commit per stage.  Variants isolate the cost of each ingredient.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/tc_probe4.cu -o tools/bin/tc_probe4
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include "../gradient-boosted-normalizing-flows_b200/csrc/tc_ptx.cuh"
using namespace gbnf::ptx;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

constexpr int NST = 8;
// VARIANT bit 0: pretest (test_wait one stage ahead) instead of a blocking wait at use
//         bit 1: commit every stage to empty[slot] (else the producer is paced by a plain arrive from the issuer)
//         bit 2: two stages per elect block (16 MMAs between checks)
//         bit 3: lane-0 polling + __syncwarp instead of all-lane polling
template <int VARIANT>
__global__ void __launch_bounds__(96, 1) issue_probe(long long* out, int* err, int stages) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t full[NST], empty[NST];
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { for (int i = 0; i < NST; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); } fence_mbar_init(); }
  if (warp == 2) tmem_alloc(&tmem_base, 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tbase = tmem_base;
  if (warp == 0) {            // fake producer
    int slot = 0; uint32_t par = 0;
    for (int s = 0; s < stages; ++s) {
      if (lane == 0) { mbar_wait(&empty[slot], par ^ 1u, err, 10); mbar_arrive(&full[slot]); }
      __syncwarp();
      if (++slot == NST) { slot = 0; par ^= 1u; }
    }
  } else if (warp == 1) {     // issuer
    int nslot = 0, slot = 0; uint32_t npar = 0;
    const uint64_t ring_desc = make_smem_desc(smem_u32(smem));
    const uint32_t idesc = make_idesc_f16(128, 64);
    uint32_t full_ok = 0;
    auto acquire = [&]() -> uint64_t {
      slot = nslot;
      if (VARIANT & 1) { if (!full_ok) mbar_wait(&full[slot], npar, err, 21); }
      else if (VARIANT & 8) { if (lane == 0) mbar_wait(&full[slot], npar, err, 21); __syncwarp(); }
      else mbar_wait(&full[slot], npar, err, 21);
      tc_fence_after();
      if (++nslot == NST) { nslot = 0; npar ^= 1u; }
      if (VARIANT & 1) full_ok = mbar_test_wait(&full[nslot], npar) ? 1u : 0u;
      return ring_desc + (uint64_t)((uint32_t)slot * 1024u);
    };
    const long long t0 = clock64();
    const int step = (VARIANT & 4) ? 2 : 1;
    for (int s = 0; s < stages; s += step) {
      const uint64_t bd = acquire();
      const int slot0 = slot;
      uint64_t bd2 = 0;
      if (VARIANT & 4) bd2 = acquire();
      if (elect_one()) {
        const uint32_t d = tbase + 64u, at = tbase + 256u;
        umma_f16_ts(d, at, bd, idesc, 1u);
#pragma unroll
        for (int i = 1; i < 8; ++i) umma_f16_ts(d, at + 8u * i, bd + (uint64_t)(i * 128), idesc, 1u);
        if (VARIANT & 2) umma_commit(&empty[slot0]); else mbar_arrive(&empty[slot0]);
        if (VARIANT & 4) {
#pragma unroll
          for (int i = 0; i < 8; ++i) umma_f16_ts(d, at + 8u * i, bd2 + (uint64_t)(i * 128), idesc, 1u);
          if (VARIANT & 2) umma_commit(&empty[slot]); else mbar_arrive(&empty[slot]);
        }
      }
      __syncwarp();
    }
    const long long t1 = clock64();
    if (lane == 0) out[0] = t1 - t0;
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 2) tmem_dealloc(tbase, 512);
}

template <int V> void run(const char* name, long long* d, int* err) {
  const int stages = 4096;
  CK(cudaFuncSetAttribute(issue_probe<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, NST * 16384));
  issue_probe<V><<<1, 96, NST * 16384>>>(d, err, stages);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: failed %s\n", name, cudaGetErrorString(e)); exit(3); }
  long long h; CK(cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost));
  printf("%-72s: %.1f cycles per stage (8 MMA N64 = 256 cycles of tensor work)\n", name, (double)h / stages);
}

int main() {
  CK(cudaSetDevice(0));
  long long* d; CK(cudaMalloc(&d, 64)); int* err; CK(cudaMalloc(&err, 4)); CK(cudaMemset(err, 0, 4));
  run<0>("blocking all-lane wait, arrive (no commit)", d, err);
  run<8>("blocking lane-0 wait + syncwarp, arrive", d, err);
  run<2>("blocking all-lane wait, tcgen05.commit -> empty", d, err);
  run<3>("pretest (test_wait one stage ahead), commit", d, err);
  run<1>("pretest, arrive", d, err);
  run<7>("pretest, commit, two stages per elect block", d, err);
  run<6>("blocking wait, commit, two stages per elect block", d, err);
  int herr; CK(cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost));
  printf("error flag %d\nPROBE4 DONE\n", herr);
  return 0;
}
