// Standalone hardware probe #2 (not part of the product).  Questions the pipelined coupling kernel depends on:
//   (A) tcgen05.mma with the A operand in TENSOR MEMORY (written by tcgen05.st as packed fp16, lane = row,
//       column c = K elements 2c, 2c+1): correctness of the layout assumption.
//   (B) cycles per tcgen05.mma for N = 256 / 128 / 64 with A in shared memory vs A in TMEM (single SM, clock64).
//   (C) L2 -> shared bulk-TMA bandwidth as a function of the number of streaming CTAs (per-SM port vs chip limit),
//       and with 2-CTA cluster multicast (each CTA fetches half of every stage and multicasts it to both).
//   (D) tcgen05.ld throughput (8 warps draining a 128 x 512 fp32 accumulator repeatedly).
//   (E) an epilogue-shaped loop: tcgen05.ld -> tanh.approx -> fp16 pack -> st.shared (canonical A image), 8 and 16 warps.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 tools/tc_probe2.cu -o tools/bin/tc_probe2
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_fp16.h>
#include "../gradient-boosted-normalizing-flows_b200/csrc/tc_ptx.cuh"

using namespace gbnf::ptx;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\tselp.b32 %0, 1, 0, px;\n\t}" : "=r"(pred));
  return pred != 0;
}
// ---- (A)+(B): A (128 x K, row-major fp16 in global) goes to smem canonical AND to TMEM; B canonical slabs in smem ----------
// mode 0: A from smem, mode 1: A from TMEM.  D columns [0, N).  A-in-TMEM lives at columns [256, 256 + K/2).
__global__ void __launch_bounds__(192, 1) mma_probe(const __half* __restrict__ Arow, const __half* __restrict__ Acan,
                                                     const __half* __restrict__ B, float* __restrict__ D, int N, int K, int mode,
                                                     int iters, long long* cycles, int* err, int nacc) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t full_bar, done_bar, a_bar;
  __shared__ uint32_t tmem_base;
  const int ks = K / 16;
  const uint32_t a_bytes = ks * 4096u, b_bytes = ks * (uint32_t)N * 32u;
  unsigned char* sA = smem;
  unsigned char* sB = smem + a_bytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&full_bar, 1); mbar_init(&done_bar, 1); mbar_init(&a_bar, 128); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&full_bar, a_bytes + b_bytes);
    tma_bulk_g2s(sA, Acan, a_bytes, &full_bar);
    tma_bulk_g2s(sB, B, b_bytes, &full_bar);
  }
  if (warp >= 2) {
    // write A into TMEM: thread = row; 8 columns (16 K elements) per store
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    for (int s = 0; s < ks; ++s) {
      uint32_t r[8];
      const uint32_t* src = reinterpret_cast<const uint32_t*>(Arow + (size_t)row * K + s * 16);
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = src[j];
      tmem_st8(tbase + ((uint32_t)(quad * 32) << 16) + 256u + 8u * s, r);
    }
    tmem_st_wait();
    tc_fence_before();
    mbar_arrive(&a_bar);
  }
  if (warp == 1) {
    mbar_wait(&full_bar, 0, err, 1);
    mbar_wait(&a_bar, 0, err, 4);
    tc_fence_after();
    const uint32_t idesc = make_idesc_f16(128, N);
    uint64_t da[8], db[8];
    uint32_t ta[8], dc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      da[k] = make_smem_desc(smem_u32(sA + k * 4096));
      db[k] = make_smem_desc(smem_u32(sB + (size_t)k * N * 32));
      ta[k] = tbase + 256u + 8u * k;
      dc[k] = tbase + (uint32_t)((k % nacc) * N);
    }
    const long long t0 = clock64();
    if (elect_one()) {
    if (mode == 0) {
#pragma unroll
      for (int k = 0; k < 8; ++k) umma_f16(dc[k], da[k], db[k], idesc, k >= nacc ? 1u : 0u);
      for (int it = 1; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_f16(dc[k], da[k], db[k], idesc, 1u);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) umma_f16_ts(dc[k], ta[k], db[k], idesc, k >= nacc ? 1u : 0u);
      for (int it = 1; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) umma_f16_ts(dc[k], ta[k], db[k], idesc, 1u);
      }
    }
    umma_commit(&done_bar);
    }
    __syncwarp();
    mbar_wait(&done_bar, 0, err, 2);
    if (cycles && lane == 0) cycles[blockIdx.x] = clock64() - t0;
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

// ---- (C) L2 -> smem streaming, optional 2-CTA multicast ---------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(bar)), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s_mc(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
      : "memory");
}

constexpr uint32_t CH = 16384;
template <int NSLOT>
__global__ void __launch_bounds__(64, 1) l2_stream(const unsigned char* __restrict__ blob, size_t bytes, int passes, int* err,
                                                    unsigned long long* sink, long long* cycles) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t full[NSLOT];
  if (threadIdx.x == 0) { for (int i = 0; i < NSLOT; ++i) mbar_init(&full[i], 1); fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    const size_t nch = bytes / CH, total = nch * passes;
    size_t issued = 0, waited = 0;
    const long long t0 = clock64();
    while (waited < total) {
      while (issued < total && issued - waited < (size_t)NSLOT) {
        int s = issued % NSLOT;
        mbar_arrive_expect_tx(&full[s], CH);
        tma_bulk_g2s(smem + s * CH, blob + (issued % nch) * CH, CH, &full[s]);
        ++issued;
      }
      int s = waited % NSLOT;
      mbar_wait(&full[s], (waited / NSLOT) & 1, err, 3);
      ++waited;
    }
    if (cycles) cycles[blockIdx.x] = clock64() - t0;
    if (sink) atomicAdd(sink, (unsigned long long)smem[0]);
  }
}

// cluster of CS CTAs; every stage (16 KB) is fetched in CS pieces, piece r by CTA r, multicast to all CTAs of the cluster.
template <int CS>
__global__ void __launch_bounds__(64, 1) l2_stream_mc(const unsigned char* __restrict__ blob, size_t bytes, int passes, int* err,
                                                       unsigned long long* sink, long long* cycles) {
  constexpr int NSLOT = 4;
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t full[NSLOT], empty[NSLOT];
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) {
    for (int i = 0; i < NSLOT; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], CS); }
    fence_mbar_init();
  }
  cluster_sync();
  if (threadIdx.x == 0) {
    const size_t nch = bytes / CH, total = nch * passes;
    size_t issued = 0, waited = 0;
    const uint32_t piece = CH / CS;
    const uint16_t mask = (uint16_t)((1u << CS) - 1u);
    const long long t0 = clock64();
    while (waited < total) {
      while (issued < total && issued - waited < (size_t)NSLOT) {
        int s = issued % NSLOT;
        if (issued >= NSLOT) mbar_wait(&empty[s], ((issued / NSLOT) - 1) & 1, err, 5);   // every CTA consumed the previous fill
        mbar_arrive_expect_tx(&full[s], CH);
        tma_bulk_g2s_mc(smem + s * CH + rank * piece, blob + (issued % nch) * CH + rank * piece, piece, &full[s], mask);
        ++issued;
      }
      int s = waited % NSLOT;
      mbar_wait(&full[s], (waited / NSLOT) & 1, err, 3);
      for (uint32_t c = 0; c < CS; ++c) mbar_arrive_remote(&empty[s], c);
      ++waited;
    }
    if (cycles) cycles[blockIdx.x] = clock64() - t0;
    if (sink) atomicAdd(sink, (unsigned long long)smem[0]);
  }
  cluster_sync();
}

// ---- (D)/(E) TMEM drain and epilogue-shaped loop -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_h2(float a, float b) { __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ float tanh_fast(float v) { float r; asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }

// MODE 0: ld only; 1: ld + tanh + pack + st.shared (A image); 2: same, two loads in flight (software pipelined);
// 3: ld + tanh + pack + tcgen05.st back to TMEM (A operand in TMEM)
template <int MODE>
__global__ void __launch_bounds__(512, 1) epi_probe(int nwarps, int iters, float* out, long long* cycles) {
  extern __shared__ __align__(1024) unsigned char smem[];   // 128 KB A image
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base;
  float acc = 0.f;
  const long long t0 = clock64();
  if (warp < nwarps) {
    const int quad = warp & 3, part = warp >> 2, nparts = nwarps >> 2;
    const int row = quad * 32 + lane;
    const int cols = 512 / nparts;
    const uint32_t lane_base = tbase + ((uint32_t)(quad * 32) << 16);
    for (int it = 0; it < iters; ++it) {
      if (MODE == 2) {
        uint32_t ra[32], rb[32];
        tmem_ld32(lane_base + part * cols, ra);
        for (int c0 = part * cols; c0 < (part + 1) * cols; c0 += 64) {
          tmem_ld_wait();
          tmem_ld32(lane_base + c0 + 32, rb);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = tanh_fast(__uint_as_float(ra[8 * q + e]));
            const uint32_t off = (uint32_t)(((c0 + 8 * q) >> 4) * 4096 + (row >> 3) * 256 + (((c0 + 8 * q) >> 3) & 1) * 128 + (row & 7) * 16);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(smem + off)), "r"(pack_h2(v[0], v[1])),
                         "r"(pack_h2(v[2], v[3])), "r"(pack_h2(v[4], v[5])), "r"(pack_h2(v[6], v[7])) : "memory");
          }
          tmem_ld_wait();
          if (c0 + 64 < (part + 1) * cols) tmem_ld32(lane_base + c0 + 64, ra);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = tanh_fast(__uint_as_float(rb[8 * q + e]));
            const int cc = c0 + 32 + 8 * q;
            const uint32_t off = (uint32_t)((cc >> 4) * 4096 + (row >> 3) * 256 + ((cc >> 3) & 1) * 128 + (row & 7) * 16);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(smem + off)), "r"(pack_h2(v[0], v[1])),
                         "r"(pack_h2(v[2], v[3])), "r"(pack_h2(v[4], v[5])), "r"(pack_h2(v[6], v[7])) : "memory");
          }
        }
      } else {
        for (int c0 = part * cols; c0 < (part + 1) * cols; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(lane_base + c0, r);
          tmem_ld_wait();
          if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < 32; ++j) acc += __uint_as_float(r[j]);
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float v[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = tanh_fast(__uint_as_float(r[8 * q + e]));
              if (MODE == 1) {
                const int cc = c0 + 8 * q;
                const uint32_t off = (uint32_t)((cc >> 4) * 4096 + (row >> 3) * 256 + ((cc >> 3) & 1) * 128 + (row & 7) * 16);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(smem + off)), "r"(pack_h2(v[0], v[1])),
                             "r"(pack_h2(v[2], v[3])), "r"(pack_h2(v[4], v[5])), "r"(pack_h2(v[6], v[7])) : "memory");
              } else {
                // 4 packed columns per 8 elements, written in place over already-read accumulator columns
                asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(lane_base + (uint32_t)((c0 >> 1) + 4 * q)),
                             "r"(pack_h2(v[0], v[1])), "r"(pack_h2(v[2], v[3])), "r"(pack_h2(v[4], v[5])), "r"(pack_h2(v[6], v[7])) : "memory");
              }
            }
            if (MODE == 3) tmem_st_wait();
          }
        }
      }
    }
  }
  const long long t1 = clock64();
  if (out) out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && cycles) cycles[blockIdx.x] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}

static void pack_canonical(const std::vector<float>& M, int rows, int K, std::vector<__half>& out) {
  out.assign((size_t)rows * K, __float2half(0.f));
  for (int r = 0; r < rows; ++r)
    for (int k = 0; k < K; ++k) {
      size_t slab = k / 16, kc = (k / 8) % 2, e = k % 8, grp = r / 8, rr = r % 8;
      out[slab * (size_t)rows * 16 + grp * 128 + kc * 64 + rr * 8 + e] = __float2half(M[(size_t)r * K + k]);
    }
}

int main(int argc, char** argv) {
  CK(cudaSetDevice(0));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int nsm = prop.multiProcessorCount;
  printf("device %s sm_%d%d, %d SMs\n", prop.name, prop.major, prop.minor, nsm);
  int* err; CK(cudaMalloc(&err, 4)); CK(cudaMemset(err, 0, 4));
  long long* dcyc; CK(cudaMalloc(&dcyc, 8 * 1024));
  std::vector<long long> hcyc(1024);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));

  // ---- (A) correctness, (B) cycles per MMA ----
  CK(cudaFuncSetAttribute(mma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (int N : {64, 128, 256}) {
    const int K = 128;
    std::vector<float> A(128 * K), B((size_t)N * K);
    srand(7);
    for (auto& v : A) v = __half2float(__float2half((rand() / (float)RAND_MAX - 0.5f)));
    for (auto& v : B) v = __half2float(__float2half((rand() / (float)RAND_MAX - 0.5f)));
    std::vector<__half> Ar(128 * K), Ap, Bp;
    for (size_t i = 0; i < A.size(); ++i) Ar[i] = __float2half(A[i]);
    pack_canonical(A, 128, K, Ap); pack_canonical(B, N, K, Bp);
    __half *dAr, *dA, *dB; float* dD;
    CK(cudaMalloc(&dAr, Ar.size() * 2)); CK(cudaMalloc(&dA, Ap.size() * 2)); CK(cudaMalloc(&dB, Bp.size() * 2)); CK(cudaMalloc(&dD, 128 * N * 4));
    CK(cudaMemcpy(dAr, Ar.data(), Ar.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dA, Ap.data(), Ap.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, Bp.data(), Bp.size() * 2, cudaMemcpyHostToDevice));
    const size_t smem = (K / 16) * (4096 + N * 32);
    for (int mode = 0; mode < 2; ++mode) {
      CK(cudaMemset(dD, 0, 128 * N * 4));
      mma_probe<<<1, 192, smem>>>(dAr, dA, dB, dD, N, K, mode, 1, nullptr, err, 1);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mma_probe N=%d mode=%d failed: %s\n", N, mode, cudaGetErrorString(e)); return 3; }
      std::vector<float> D(128 * N);
      CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
      double maxerr = 0;
      for (int r = 0; r < 128; ++r) for (int n = 0; n < N; ++n) {
        double ref = 0; for (int k = 0; k < K; ++k) ref += (double)A[r * K + k] * B[(size_t)n * K + k];
        maxerr = fmax(maxerr, fabs(ref - D[r * N + n]));
      }
      printf("UMMA N=%d K=%d A-from-%s: max abs err %.3e %s\n", N, K, mode ? "TMEM" : "smem", maxerr, maxerr < 1e-3 ? "MATCH" : "mismatch");
      for (int nacc : {1, 2, 4}) {
        if (nacc * N > (mode ? 256 : 512)) continue;
        const int iters = 2000;
        mma_probe<<<nsm, 192, smem>>>(dAr, dA, dB, nullptr, N, K, mode, iters, dcyc, err, nacc);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(hcyc.data(), dcyc, 8 * nsm, cudaMemcpyDeviceToHost));
        double cyc = (double)hcyc[0] / ((double)iters * (K / 16));
        printf("  %d accumulators round-robin: %.1f cycles per MMA (M128 N%d K16) -> %.0f MAC/clk/SM\n", nacc, cyc, N, 128.0 * N * 16 / cyc);
      }
    }
    cudaFree(dAr); cudaFree(dA); cudaFree(dB); cudaFree(dD);
  }

  const bool only_mma = argc > 1;
  // ---- (C) L2 -> smem ----
  if (!only_mma) {
    const size_t bytes = 24u << 20;
    unsigned char* blob; CK(cudaMalloc(&blob, bytes)); CK(cudaMemset(blob, 1, bytes));
    unsigned long long* sink; CK(cudaMalloc(&sink, 8));
    CK(cudaFuncSetAttribute(l2_stream<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * CH));
    CK(cudaFuncSetAttribute(l2_stream<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * CH));
    CK(cudaFuncSetAttribute(l2_stream_mc<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * CH));
    CK(cudaFuncSetAttribute(l2_stream_mc<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * CH));
    l2_stream<4><<<nsm, 64, 4 * CH>>>(blob, bytes, 1, err, sink, nullptr); CK(cudaDeviceSynchronize());   // warm L2
    for (int ctas : {1, 2, 8, 32, 74, 148}) {
      for (int slots : {4, 8}) {
        CK(cudaEventRecord(e0));
        if (slots == 4) l2_stream<4><<<ctas, 64, 4 * CH>>>(blob, bytes, 2, err, sink, dcyc);
        else            l2_stream<8><<<ctas, 64, 8 * CH>>>(blob, bytes, 2, err, sink, dcyc);
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        CK(cudaMemcpy(hcyc.data(), dcyc, 8 * ctas, cudaMemcpyDeviceToHost));
        printf("L2->smem unicast  %3d CTAs ring %dx16KB: %.1f GB/s per SM, %.1f B/clk per SM (%.3f ms)\n", ctas, slots,
               2.0 * bytes / ms / 1e6, 2.0 * bytes / (double)hcyc[0], ms);
      }
    }
    for (int cs : {2, 4}) {
      CK(cudaEventRecord(e0));
      {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((nsm / cs) * cs); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = 4 * CH;
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        if (cs == 2) CK(cudaLaunchKernelEx(&cfg, l2_stream_mc<2>, (const unsigned char*)blob, bytes, 2, err, sink, dcyc));
        else         CK(cudaLaunchKernelEx(&cfg, l2_stream_mc<4>, (const unsigned char*)blob, bytes, 2, err, sink, dcyc));
      }
      CK(cudaEventRecord(e1));
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("multicast cs=%d failed: %s\n", cs, cudaGetErrorString(e)); break; }
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      CK(cudaMemcpy(hcyc.data(), dcyc, 8, cudaMemcpyDeviceToHost));
      printf("L2->smem multicast cluster=%d, %d CTAs ring 4x16KB: %.1f GB/s delivered per SM, %.1f B/clk per SM (%.3f ms)\n", cs,
             cs == 2 ? (nsm / 2) * 2 : (nsm / 4) * 4, 2.0 * bytes / ms / 1e6, 2.0 * bytes / (double)hcyc[0], ms);
    }
    cudaFree(blob);
  }

  // ---- (D)/(E) ----
  if (!only_mma) {
    float* out; CK(cudaMalloc(&out, (size_t)nsm * 512 * 4));
    CK(cudaFuncSetAttribute(epi_probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    CK(cudaFuncSetAttribute(epi_probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    CK(cudaFuncSetAttribute(epi_probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    CK(cudaFuncSetAttribute(epi_probe<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    const int iters = 200;
    for (int mode = 0; mode < 4; ++mode)
      for (int nw : {4, 8, 16}) {
        if (mode == 0) epi_probe<0><<<nsm, 512, 128 * 1024>>>(nw, iters, out, dcyc);
        if (mode == 1) epi_probe<1><<<nsm, 512, 128 * 1024>>>(nw, iters, out, dcyc);
        if (mode == 2) epi_probe<2><<<nsm, 512, 128 * 1024>>>(nw, iters, out, dcyc);
        if (mode == 3) epi_probe<3><<<nsm, 512, 128 * 1024>>>(nw, iters, out, dcyc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("epi_probe mode %d failed: %s\n", mode, cudaGetErrorString(e)); return 4; }
        CK(cudaMemcpy(hcyc.data(), dcyc, 8, cudaMemcpyDeviceToHost));
        const char* nm[] = {"tcgen05.ld only", "ld+tanh+pack+st.shared", "ld+tanh+pack+st.shared (2 loads in flight)", "ld+tanh+pack+tcgen05.st"};
        printf("epilogue probe [%s] %2d warps: %.0f cycles per 128x512 tile (%.2f elem/clk)\n", nm[mode], nw, (double)hcyc[0] / iters,
               65536.0 * iters / (double)hcyc[0]);
      }
  }
  int herr = 0; CK(cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost));
  printf("error flag: %d\nPROBE2 DONE\n", herr);
  return 0;
}
