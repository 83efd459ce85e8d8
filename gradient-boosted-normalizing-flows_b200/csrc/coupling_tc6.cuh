// Two-chain kernel for RealNVP components of hidden width 256 with tanh networks (BASELINE configurations 1 and 2), forward
// direction: TWO COMPONENTS per CTA, each with its own half of tensor memory.
//
// These configurations are MUFU bound (4 h tanh per row and step = 8.2 k cycles per 128-row step; 6.0 k of GEMM), but a single
// component is a serial chain gather -> L1 -> tanh -> L2 -> tanh -> last layer (x 2 networks) -> coupling, in which the epilogue
// warps and the tensor pipe mostly wait for each other: 20 k cycles per step in coupling_tc2_kernel, 18.6 k with the two networks of
// a step interleaved (coupling_tc5_kernel: each network has ONE accumulator slot, so its layer 2 is a serial chain again).
// Components are independent, and at h = 256 one network needs only 256 tensor-memory columns (A1 128 + one accumulator slot 128).
// This kernel therefore runs two chains X / Y (the two halves of a work unit's components on the same 128 rows, own z tile, A0,
// bias / table buffers, shift and piece buffers), chain c in TMEM columns [256 c, 256 c + 256), and ALL THREE roles walk one global
// op list that alternates between the chains with chain Y half a step behind: while the epilogue warps activate a chunk of one
// chain (or transform / gather for it), the tensor pipe multiplies for the other.
//
//   per chain and step, 12 ops: for network n in (t, s): L1(q=0) L1(q=1) L2(j=0) L3(j=0) L2(j=1) L3(j=1)
//     L1(q) : A0 x W1 chunk q -> slot; epilogue: bias + tanh -> fp16 pairs -> A1 quarter q
//     L2(j) : A1 x W2 chunk j (two 32 KB ring stages) -> slot; epilogue: bias + tanh -> fp16 pairs packed in place
//     L3(j) : packed piece x its last-layer k-slabs -> slot[64, 64 + Np3); epilogue: piece 0 parks the product in shared memory,
//             piece 1 adds it: network t -> the shift goes to shared memory, network s -> the coupling z2' = t + z2 exp(s)
//             (transformations.py:575-577), then everything between two passes of the chain (log q at a component's end, x reload,
//             gather of the next pass, staging)
//   Ops are decoded at run time from one copy of the code (instruction footprint).  Results equal coupling_tc5_kernel's
//   bit for bit (same operands, same summation order).  Launch rule as for coupling_tc4_kernel: every work unit needs the same,
//   even number of components; other launches run coupling_tc5_kernel / coupling_tc2_kernel on the same packed image.
#pragma once
#include "coupling_tc4.cuh"
#include "coupling_tc5.cuh"

namespace gbnf {

struct Tc6Misc {
  uint64_t full[kT2MaxStages];
  uint64_t empty[kT2MaxStages];
  uint64_t a0r[2];       // per chain: epilogue -> MMA : A0 gathered                                   (16 arrivals)
  uint64_t a1r[2];       //            epilogue -> MMA : layer-1 chunk packed, slot drained             (16 arrivals)
  uint64_t sr[2];        //            epilogue -> MMA : last-layer A piece packed in the slot          (16 arrivals)
  uint64_t l3r[2];       //            epilogue -> MMA : piece product read, slot free                  (16 arrivals)
  uint64_t l1f[2];       //            MMA -> epilogue : layer-1 chunk accumulated                      (commit)
  uint64_t l2f[2];       //            MMA -> epilogue : layer-2 chunk accumulated                      (commit)
  uint64_t l3f[2];       //            MMA -> epilogue : piece product complete                         (commit)
  uint64_t w3full[2];    //            producer -> MMA : the network's last-layer weights landed
  uint64_t w3empty[2];   //            MMA -> producer : ... may be overwritten
  uint32_t tmem_base;
  uint32_t last_flag;
  float coef[kMaxComponents];
};
static_assert(sizeof(Tc6Misc) <= kT2MiscBytes, "misc region too small");

inline bool tc6_eligible(const ModelDims& md, const std::vector<StepDesc>& steps) {
  if (!tc2_eligible(md, steps) || steps.empty()) return false;
  if (md.h != 256 || md.nnets != 2 || md.kind != GBNF_KIND_REALNVP || md.act != GBNF_ACT_TANH) return false;
  const StepDesc& s0 = steps[0];
  for (const StepDesc& s : steps)
    for (int n = 0; n < 2; ++n)
      if (s.layer[n][0].Kp != s0.layer[0][0].Kp || s.layer[n][2].Np != s0.layer[0][2].Np) return false;
  return s0.layer[0][0].Kp <= 32 && s0.layer[0][2].Np <= 64;
}

struct Tc6Offsets { uint32_t zs_stride, a0_stride, w3_stride, px_stride; };
__host__ __device__ inline Tc6Offsets t6_offsets(int Dv, int k0p, int np3) {
  Tc6Offsets o;
  o.zs_stride = ((uint32_t)(kTcRows * Dv * 4) + 127u) & ~127u;
  o.a0_stride = (uint32_t)(k0p >> 4) * 4096u;
  o.w3_stride = 16u * (uint32_t)np3 * 32u;
  o.px_stride = (uint32_t)kTcRows * (uint32_t)np3 * 4u;          // one [np3][128 rows] float buffer
  return o;
}

// shared memory: zs[2] | A0[2] | part | misc | bias[2 chains][2 buffers] | tables[2][2] | w3[2] | piece[2] | shift[2] | ring
inline bool tc6_make_plan(const ModelDims& md, const std::vector<StepDesc>& steps, TcPlan* p) {
  const int k0p = steps[0].layer[0][0].Kp, np3 = steps[0].layer[0][2].Np;
  const Tc6Offsets f = t6_offsets(md.Dv, k0p, np3);
  auto al = [](uint32_t v) { return (v + 127u) & ~127u; };
  int out_max = 1;
  for (const StepDesc& s : steps) out_max = std::max(out_max, s.out_dim);
  p->K0p = k0p; p->out_max = out_max;
  uint32_t o = 0;
  p->off_zs = o;   o = al(o + 2 * f.zs_stride);
  p->off_a0 = o;   o = al(o + 2 * f.a0_stride);
  p->off_a1 = o;   o = al(o + 6 * kTcRows * 4);
  p->off_misc = o; o = al(o + kT2MiscBytes);
  p->off_bias = o; o = al(o + 4 * t5_bias_floats(np3) * 4);
  p->off_tab = o;  o = al(o + 4 * 2 * kEpPad * 16);
  p->off_w3 = o;   o = al(o + 2 * f.w3_stride);
  p->off_sh = o;   o = al(o + 4 * f.px_stride);                   // piece[0], piece[1], shift[0], shift[1]
  p->off_ring = o;
  const uint32_t limit = 227 * 1024;
  if (o + 3 * kT2StageBytes > limit) return false;
  p->nst = std::min<int>(6, (limit - o) / kT2StageBytes);
  p->smem_bytes = o + (size_t)p->nst * kT2StageBytes;
  p->tmem_cols = 512;
  return true;
}

constexpr int kT6Ops = 12;      // ops per chain and pass
constexpr int kT6Lag = 6;       // chain 1 runs this many ops behind chain 0

template <int TANH_MODE>
__global__ void __launch_bounds__(kT2Threads, 1) coupling_tc6_kernel(CouplingArgs a, TcPlan plan) {
  extern __shared__ __align__(1024) unsigned char smem[];
  asm volatile(".reg .pred t6_p_full;" ::);
  const ModelDims& md = a.md;
  const int D = md.D, Dv = md.Dv, K = md.K;
  const int k0p = __ldg(&a.steps[a.c0 * md.K].layer[0][0].Kp), np3 = __ldg(&a.steps[a.c0 * md.K].layer[0][2].Np);
  const int k0s = k0p >> 4;
  const Tc6Offsets F = t6_offsets(Dv, k0p, np3);
  unsigned char* const zs_b = smem + plan.off_zs;
  unsigned char* const A0_b = smem + plan.off_a0;
  float* const part_s = reinterpret_cast<float*>(smem + plan.off_a1);
  float* const bias_s = reinterpret_cast<float*>(smem + plan.off_bias);
  float4* const tab_s = reinterpret_cast<float4*>(smem + plan.off_tab);
  Tc6Misc* const misc = reinterpret_cast<Tc6Misc*>(smem + plan.off_misc);
  unsigned char* const ring = smem + plan.off_ring;
  unsigned char* const w3buf = smem + plan.off_w3;
  unsigned char* const px_b = smem + plan.off_sh;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nst = plan.nst;
  const uint32_t bfl = t5_bias_floats(np3);
  const int my_units = (a.num_units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int npass = my_units * (a.comps_per_unit >> 1) * K;      // passes per chain
  const int nops = npass * kT6Ops;                               // ops per chain
  const int gops = npass > 0 ? 2 * (nops + kT6Lag) : 0;          // global op slots (even: chain 0, odd: chain 1)

  if (threadIdx.x == 0) {
    for (int i = 0; i < nst; ++i) { ptx::mbar_init(&misc->full[i], 1); ptx::mbar_init(&misc->empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&misc->a0r[i], 16); ptx::mbar_init(&misc->a1r[i], 16); ptx::mbar_init(&misc->sr[i], 16); ptx::mbar_init(&misc->l3r[i], 16);
      ptx::mbar_init(&misc->l1f[i], 1); ptx::mbar_init(&misc->l2f[i], 1); ptx::mbar_init(&misc->l3f[i], 1);
      ptx::mbar_init(&misc->w3full[i], 1); ptx::mbar_init(&misc->w3empty[i], 1);
    }
    ptx::fence_mbar_init();
    if (a.G_ll != nullptr) mixture_coefficients(a.rho, a.n_mix, a.skip_c, a.mix_mode, misc->coef);
  }
  if (warp == 1) ptx::tmem_alloc(&misc->tmem_base, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tbase = misc->tmem_base;
  const __half* wb = reinterpret_cast<const __half*>(a.wblob);

  // decode global slot gi -> (chain, op index within the chain); false: the chain has no op in this slot
  auto decode = [&](int gi, int& chain, int& o) -> bool {
    chain = gi & 1;
    o = (gi >> 1) - (chain ? kT6Lag : 0);
    return o >= 0 && o < nops;
  };

  if (warp == 0) {
    // ===================================== TMA producer =====================================================
    uint32_t par = 0;
    int slot = 0;
    uint32_t nw3[2] = {0u, 0u};                      // last-layer weight loads done so far per chain (phase of w3empty)
    auto push = [&](const __half* src, uint32_t bytes) {
      t2_wait(&misc->empty[slot], par ^ 1u, a.error_flag, 10, lane);
      if (ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(&misc->full[slot], bytes);
        ptx::tma_bulk_g2s(ring + (size_t)slot * kT2StageBytes, src, bytes, &misc->full[slot]);
      }
      __syncwarp();
      if (++slot == nst) { slot = 0; par ^= 1u; }
    };
    T4Pos pX, pY;
    pX.u = pY.u = (int)blockIdx.x;
    t4_pos_unit(a, 0, pX); t4_pos_unit(a, 1, pY);
#pragma unroll 1
    for (int gi = 0; gi < gops; ++gi) {
      int chain, o;
      if (!decode(gi, chain, o)) continue;
      const int r = o % kT6Ops, n = r / 6, t = r - 6 * n;
      if (t != 0 && t != 2 && t != 4) { if (r == kT6Ops - 1) { if (chain) t4_pos_next(a, 1, pY); else t4_pos_next(a, 0, pX); } continue; }
      const T4Pos& p = chain ? pY : pX;
      const StepDesc* sd = a.steps + (p.c * K + p.k);
      if (t == 0) {
        push(wb + __ldg(&sd->layer[n][0].w_off), (uint32_t)(2 * k0s) * 4096u);             // W1 of both layer-1 chunks
        const uint32_t k3 = chain ? nw3[1] : nw3[0];
        t2_wait(&misc->w3empty[chain], (k3 & 1u) ^ 1u, a.error_flag, 11, lane);
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(&misc->w3full[chain], F.w3_stride);
          ptx::tma_bulk_g2s(w3buf + (size_t)chain * F.w3_stride, wb + __ldg(&sd->layer[n][2].w_off), F.w3_stride, &misc->w3full[chain]);
        }
        __syncwarp();
        if (chain) ++nw3[1]; else ++nw3[0];
      } else {
        const __half* base = wb + __ldg(&sd->layer[n][1].w_off) + (size_t)(t == 2 ? 0 : 32768);   // chunk j: [16 k-slabs][128 x 16]
        push(base, 32768u);
        push(base + (size_t)8 * 2048, 32768u);
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================================
    int nslot = 0, slot = 0;
    uint32_t npar = 0;
    uint32_t ph_a0 = 0, ph_a1 = 0, ph_sr = 0, ph_l3r = 0, ph_w3 = 0;     // phase bits, bit c = chain c
    uint32_t first = 3u;                                                  // bit c: chain c has not issued its first L1 yet
    const uint64_t a0_desc = ptx::make_smem_desc(ptx::smem_u32(A0_b));
    const uint64_t ring_desc = ptx::make_smem_desc(ptx::smem_u32(ring));
    const uint64_t w3_desc = ptx::make_smem_desc(ptx::smem_u32(w3buf));
    const uint32_t idesc_128 = ptx::make_idesc_f16(128, 128);
    const uint32_t idesc_o = ptx::make_idesc_f16(128, np3);
    const uint32_t b3_step = (uint32_t)np3 * 2u;
    uint64_t l1_desc[2] = {0, 0};
    int l1_slot[2] = {0, 0};
    auto test_full = [&](uint64_t* bar, uint32_t par) {
      asm volatile("mbarrier.test_wait.parity.shared::cta.b64 t6_p_full, [%0], %1;" ::"r"(ptx::smem_u32(bar)), "r"(par) : "memory");
    };
    test_full(&misc->full[0], 0u);
    auto acquire = [&]() -> uint64_t {
      slot = nslot;
      uint32_t full_ok;
      asm volatile("selp.u32 %0, 1, 0, t6_p_full;" : "=r"(full_ok));
      if (!full_ok) ptx::mbar_wait(&misc->full[slot], npar, a.error_flag, 21);
      ptx::tc_fence_after();
      if (++nslot == nst) { nslot = 0; npar ^= 1u; }
      test_full(&misc->full[nslot], npar);
      return ring_desc + (uint64_t)((uint32_t)slot * (kT2StageBytes >> 4));
    };
    auto wait_bit = [&](uint64_t* bar, uint32_t& bits, int c, int code) {
      ptx::mbar_wait(bar, (bits >> c) & 1u, a.error_flag, code);
      bits ^= 1u << c;
      ptx::tc_fence_after();
    };
#pragma unroll 1
    for (int gi = 0; gi < gops; ++gi) {
      int chain, o;
      if (!decode(gi, chain, o)) continue;
      const int r = o % kT6Ops, n = r / 6, t = r - 6 * n;
      const uint32_t a1c = tbase + 256u * chain, slc = a1c + 128u;
      if (t <= 1) {
        // ---- layer-1 chunk t -> slot ----
        if (t == 0) {
          const uint64_t dsc = acquire();
          if (chain) { l1_desc[1] = dsc; l1_slot[1] = slot; } else { l1_desc[0] = dsc; l1_slot[0] = slot; }
          if (n == 0) wait_bit(&misc->a0r[chain], ph_a0, chain, 20);                     // this pass's A0 gathered
          if ((first >> chain) & 1u) first &= ~(1u << chain);
          else wait_bit(&misc->l3r[chain], ph_l3r, chain, 26);                           // previous network's last piece product read
        } else {
          wait_bit(&misc->a1r[chain], ph_a1, chain, 22);                                 // chunk 0 packed, slot drained
        }
        const uint64_t ld = chain ? l1_desc[1] : l1_desc[0];
        const int ls = chain ? l1_slot[1] : l1_slot[0];
        if (ptx::elect_one()) {
          const uint64_t ad = a0_desc + (uint64_t)((uint32_t)chain * (F.a0_stride >> 4));
          const uint64_t bq = ld + (uint64_t)(t * k0s * 256);
          for (int i = 0; i < k0s; ++i) ptx::umma_f16(slc, ad + (uint64_t)(i * 256), bq + (uint64_t)(i * 256), idesc_128, i > 0 ? 1u : 0u);
          ptx::umma_commit(&misc->l1f[chain]);
          if (t == 1) ptx::umma_commit(&misc->empty[ls]);
        }
        __syncwarp();
      } else if (t == 2 || t == 4) {
        // ---- layer-2 chunk j -> slot (two ring stages) ----
        if (t == 2) wait_bit(&misc->a1r[chain], ph_a1, chain, 22);                       // A1 complete, slot drained
        else wait_bit(&misc->l3r[chain], ph_l3r, chain, 26);                             // piece 0's product read: slot free
#pragma unroll
        for (int y = 0; y < 2; ++y) {
          const uint64_t bd = acquire();
          const int cur = slot;
          if (ptx::elect_one()) {
            const uint32_t at = a1c + 64u * y;
#pragma unroll
            for (int i = 0; i < 8; ++i) ptx::umma_f16_ts(slc, at + 8u * i, bd + (uint64_t)(i * 256), idesc_128, (y > 0 || i > 0) ? 1u : 0u);
            ptx::umma_commit(&misc->empty[cur]);
            if (y == 1) ptx::umma_commit(&misc->l2f[chain]);
          }
          __syncwarp();
        }
      } else {
        // ---- last-layer piece j: packed piece in slot[0, 64) x its k-slabs -> slot[64, 64 + np3) ----
        const int j = (t == 3) ? 0 : 1;
        wait_bit(&misc->sr[chain], ph_sr, chain, 23);
        if (j == 0) wait_bit(&misc->w3full[chain], ph_w3, chain, 24);
        if (ptx::elect_one()) {
          const uint64_t bd = w3_desc + (uint64_t)((uint32_t)chain * (F.w3_stride >> 4)) + (uint64_t)((uint32_t)(8 * j) * b3_step);
          for (int i = 0; i < 8; ++i) ptx::umma_f16_ts(slc + 64u, slc + 8u * i, bd + (uint64_t)((uint32_t)i * b3_step), idesc_o, i > 0 ? 1u : 0u);
          ptx::umma_commit(&misc->l3f[chain]);
          if (j == 1) ptx::umma_commit(&misc->w3empty[chain]);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================================== epilogue / elementwise warps =====================================
    const int et = threadIdx.x - 64;
    const int warp_e = et >> 5;
    const int quad = warp & 3;
    const int g = warp_e >> 2;
    const int row = quad * 32 + lane;
    const uint32_t lane_base = tbase + ((uint32_t)(quad * 32) << 16);
    const int dq = (D + 3) >> 2;
    const int h0col = min(D, g * dq), h1col = min(D, (g + 1) * dq);
    const int c0 = g * 16;
    uint32_t ph_l1f = 0, ph_l2f = 0, ph_l3f = 0;
    float lsumX = 0.f, lsumY = 0.f;
    uint32_t ngX = 0, ngY = 0;               // passes of the chain gathered so far
    T4Pos pX, pY;                            // the pass of the chain that was gathered last
    pX.u = pY.u = (int)blockIdx.x;
    t4_pos_unit(a, 0, pX); t4_pos_unit(a, 1, pY);
    float* const part = part_s;
    float* const part2 = part + 3 * kTcRows;

    // staging of pass `sd` of a chain into its buffer `buf`: biases of both networks + the step's two tables
    auto stage = [&](int chain, const StepDesc* sd, uint32_t buf) {
      const float* bsrc = a.fblob + __ldg(&sd->layer[0][0].b_off);
      float* bdst = bias_s + (2 * chain + buf) * bfl;
      for (int i = et; i < (int)(bfl >> 2); i += kT2EpiThreads) ptx::cp_async16(bdst + 4 * i, bsrc + 4 * i);
      if (et >= kT2EpiThreads - 2 * kEpPad) {
        const int i = et - (kT2EpiThreads - 2 * kEpPad);
        ptx::cp_async16(tab_s + (2 * chain + buf) * (2 * kEpPad) + i, reinterpret_cast<const float4*>(a.fblob + __ldg(&sd->ep_off)) + i);
      }
    };
    // x rows -> z tile of the chain (every thread loads the columns it owns), then the gather of pass index `ng` -> A0 of the chain
    auto start_pass = [&](int chain, const T4Pos& p, uint32_t ng) {
      float* zrow = reinterpret_cast<float*>(zs_b + chain * F.zs_stride) + row * Dv;
      if (p.k == 0) {
        const long long gr = (long long)(p.u / a.split) * kTcRows + row;
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = (h0col + i < h1col && gr < a.B) ? __ldg(a.x + gr * D + h0col + i) : 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) if (h0col + i < h1col) zrow[h0col + i] = v[i];
        if (g == 0) for (int pc = D; pc < Dv; ++pc) zrow[pc] = 0.f;
        t2_quad_bar(quad);
      }
      const StepDesc* sd = a.steps + (p.c * K + p.k);
      const int in_dim = __ldg(&sd->in_dim);
      const float4* tab1 = tab_s + (2 * chain + (ng & 1u)) * (2 * kEpPad);
      unsigned char* A0 = A0_b + chain * F.a0_stride;
      const int nch = k0p >> 3;
      for (int chn = g; chn < nch; chn += 4) {
        float4 t[8];
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) t[e] = tab1[chn * 8 + e];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = zrow[__float_as_int(t[e].w)];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = (v[e] + t[e].x) * t[e].y + t[e].z;
#pragma unroll
        for (int e = 0; e < 8; ++e) if (chn * 8 + e < in_dim) zrow[__float_as_int(t[e].w)] = v[e];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = (chn * 8 + e < in_dim) ? v[e] : 0.f;
        {
          const float mx = fmax_nan(fmax_nan(fmax_nan(fabsf(v[0]), fabsf(v[1])), fmax_nan(fabsf(v[2]), fabsf(v[3]))),
                                    fmax_nan(fmax_nan(fabsf(v[4]), fabsf(v[5])), fmax_nan(fabsf(v[6]), fabsf(v[7]))));
          if (!(mx <= 65504.f)) *reinterpret_cast<volatile int*>(a.error_flag + 2) = 1;
        }
        st_shared_v4(A0 + a_chunk_off(row, chn * 8), pack_half2(v[0], v[1]), pack_half2(v[2], v[3]), pack_half2(v[4], v[5]),
                     pack_half2(v[6], v[7]));
      }
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      t2_warp_arrive(&misc->a0r[chain], lane);
    };
    // end of a component: log q for this row; chain 1 at the end of a unit: the tile's logsumexp
    auto comp_end = [&](int chain, const T4Pos& p, const float* zrow, float lsum) {
      const int tile = p.u / a.split;
      const long long gr = (long long)tile * kTcRows + row;
      const int c = p.c;
      const float2 cconst = __ldg(reinterpret_cast<const float2*>(a.fblob + a.cc_off) + c);
      float q = 0.f;
      if (md.base == GBNF_BASE_STD_NORMAL) {
        for (int pc = h0col; pc < h1col; ++pc) { const float d = zrow[pc]; q = fmaf(d * d, 0.5f, q); }
      } else {
        const float* bm = a.fblob + __ldg(&a.comps[c].base_off);
        const float* bi = bm + Dv;
        for (int pc = h0col; pc < h1col; ++pc) { const float d = zrow[pc] - __ldg(bm + pc); q = fmaf(d * d, __ldg(bi + pc), q); }
      }
      if (g > 0) { part[(g - 1) * kTcRows + row] = q; part2[(g - 1) * kTcRows + row] = lsum; }
      t2_quad_bar(quad);
      if (g == 0) {
        q += part[row] + part[kTcRows + row] + part[2 * kTcRows + row];
        const float ldj_tot = (lsum + part2[row] + part2[kTcRows + row] + part2[2 * kTcRows + row]) + cconst.x;
        const float lq = (cconst.y - q) + ldj_tot;
        if (gr < a.B) {
          if (a.logq) a.logq[gr * a.ld_logq + (c - a.c0)] = lq;
          for (int qq = 0; qq < a.n_peers; ++qq) a.logq_peers[qq][(long long)(a.peer_col0 + (c - a.c0)) * a.peer_ld + gr] = lq;
          if (a.ldj_out) a.ldj_out[gr] = ldj_tot;
        }
        if (a.G_ll != nullptr && c < a.n_mix) __stcg(a.lse_terms + ((long long)tile * kTcRows + row) * a.n_mix + c, misc->coef[c] + lq);
      }
      t2_quad_bar(quad);                               // the partial sums are dead before anybody writes them again
      if (chain == 1 && c + 1 == p.cend && a.G_ll != nullptr) {
        bool last = true;
        if (a.split > 1) {
          if (g == 0) __threadfence();
          t2_epi_bar();
          if (et == 0) misc->last_flag = (atomicAdd(a.tile_ctr + tile, 1u) == (unsigned)(a.split - 1)) ? 1u : 0u;
          t2_epi_bar();
          last = misc->last_flag != 0u;
          if (last && et == 0) a.tile_ctr[tile] = 0u;
          if (last && g == 0) __threadfence();
        }
        if (last && g == 0 && gr < a.B) {
          const float* tv = a.lse_terms + ((long long)tile * kTcRows + row) * a.n_mix;
          float M = -INFINITY;
          bool has_nan = false;
          for (int i = 0; i < a.n_mix; ++i) { const float t = __ldcg(tv + i); M = fmaxf(M, t); has_nan |= (t != t); }
          float S = 0.f;
          for (int i = 0; i < a.n_mix; ++i) { const float t = __ldcg(tv + i); if (t != -INFINITY) S += expf(t - M); }
          a.G_ll[gr] = has_nan ? __int_as_float(0x7fc00000) : (M == INFINITY || M == -INFINITY) ? M : M + logf(S);
        }
      }
    };

    // ---- prologue: staging of the first two passes of both chains, then x tile + gather of the first passes ----
    if (npass > 0) {
#pragma unroll 1
      for (int chain = 0; chain < 2; ++chain) {
        T4Pos p = chain ? pY : pX;
        stage(chain, a.steps + (p.c * K + p.k), 0u);
        if (npass > 1) { t4_pos_next(a, chain, p); stage(chain, a.steps + (p.c * K + p.k), 1u); }
      }
      ptx::cp_async_wait_all();
      t2_epi_bar();
      start_pass(0, pX, 0u); ngX = 1u;
      start_pass(1, pY, 0u); ngY = 1u;
    }
#pragma unroll 1
    for (int gi = 0; gi < gops; ++gi) {
      int chain, o;
      if (!decode(gi, chain, o)) continue;
      const int r = o % kT6Ops, n = r / 6, t = r - 6 * n;
      const uint32_t pidx = (uint32_t)(o / kT6Ops);                                   // pass index of the chain
      const uint32_t a1c = lane_base + 256u * chain, slc = a1c + 128u;
      const float* bias_c = bias_s + (2 * chain + (pidx & 1u)) * bfl + n * (512 + np3);  // b1 | b2 | b3 of this network
      if (t <= 1) {
        // ---- layer-1 chunk t: accumulator -> bias + tanh -> fp16 pairs -> A1 quarter t ----
        t2_wait(&misc->l1f[chain], (ph_l1f >> chain) & 1u, a.error_flag, 30, lane);
        ph_l1f ^= 1u << chain;
        ptx::tc_fence_after();
        uint32_t pk[16];
        {
          uint32_t rr[32];
          ptx::tmem_ld32(slc + (uint32_t)g * 32u, rr);
          ptx::tmem_ld_wait();
          t2_act_pack32<1, TANH_MODE>(rr, bias_c + t * 128 + g * 32, pk, a.error_flag);
        }
        ptx::tmem_st16(a1c + 64u * t + (uint32_t)g * 16u, pk);
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        t2_warp_arrive(&misc->a1r[chain], lane);
      } else if (t == 2 || t == 4) {
        // ---- layer-2 chunk j: accumulator -> bias + tanh -> fp16 pairs packed in place ----
        const int j = (t == 2) ? 0 : 1;
        t2_wait(&misc->l2f[chain], (ph_l2f >> chain) & 1u, a.error_flag, 31, lane);
        ph_l2f ^= 1u << chain;
        ptx::tc_fence_after();
        uint32_t pk[16];
        {
          uint32_t rr[32];
          ptx::tmem_ld32(slc + (uint32_t)g * 32u, rr);
          ptx::tmem_ld_wait();
          t2_act_pack32<1, TANH_MODE>(rr, bias_c + 256 + j * 128 + g * 32, pk, a.error_flag);
        }
        t2_quad_bar(quad);
        ptx::tmem_st16(slc + (uint32_t)g * 16u, pk);
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        t2_warp_arrive(&misc->sr[chain], lane);
      } else {
        // ---- last-layer piece j: product -> (piece 0) parked in shared memory | (piece 1) summed; t network: shift parked,
        //      s network: coupling transform and everything up to the chain's next gather ----
        const int j = (t == 3) ? 0 : 1;
        t2_wait(&misc->l3f[chain], (ph_l3f >> chain) & 1u, a.error_flag, 32, lane);
        ph_l3f ^= 1u << chain;
        ptx::tc_fence_after();
        float* const piece = reinterpret_cast<float*>(px_b + chain * F.px_stride) + c0 * kTcRows + row;          // [col][row]
        float* const shift = reinterpret_cast<float*>(px_b + (2 + chain) * F.px_stride) + c0 * kTcRows + row;
        uint32_t r3[16];
        if (c0 < np3) {
          ptx::tmem_ld16(slc + 64u + (uint32_t)c0, r3);
          ptx::tmem_ld_wait();
        }
        ptx::tc_fence_before();
        t2_warp_arrive(&misc->l3r[chain], lane);
        if (c0 < np3) {
          if (j == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) piece[i * kTcRows] = __uint_as_float(r3[i]);
          } else if (n == 0) {
            const float* b3 = bias_c + 512;
#pragma unroll
            for (int i = 0; i < 16; ++i) shift[i * kTcRows] = (piece[i * kTcRows] + __uint_as_float(r3[i])) + b3[c0 + i];
          }
        }
        if (j == 1 && n == 1) {
          T4Pos p = chain ? pY : pX;
          uint32_t ng = chain ? ngY : ngX;
          float lsum = chain ? lsumY : lsumX;
          float* zrow = reinterpret_cast<float*>(zs_b + chain * F.zs_stride) + row * Dv;
          const StepDesc* sd = a.steps + (p.c * K + p.k);
          const int out_dim = __ldg(&sd->out_dim);
          if (c0 < np3) {
            const float4* tab2 = tab_s + (2 * chain + (pidx & 1u)) * (2 * kEpPad) + kEpPad;
            const float* b3 = bias_c + 512;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              float4 tt[8];
              float z[8];
#pragma unroll
              for (int jj = 0; jj < 8; ++jj) tt[jj] = tab2[c0 + half * 8 + jj];
#pragma unroll
              for (int jj = 0; jj < 8; ++jj) z[jj] = zrow[__float_as_int(tt[jj].w)];
#pragma unroll
              for (int jj = 0; jj < 8; ++jj) {
                const int jc = c0 + half * 8 + jj;
                const int i = half * 8 + jj;
                const float sc = (piece[i * kTcRows] + __uint_as_float(r3[i])) + b3[jc];
                const float zn = (z[jj] + tt[jj].x) * tt[jj].y + tt[jj].z;
                z[jj] = shift[i * kTcRows] + zn * __expf(sc);                              // transformations.py:575
                lsum += (jc < out_dim) ? sc : 0.f;                                          // transformations.py:577
              }
#pragma unroll
              for (int jj = 0; jj < 8; ++jj) if (c0 + half * 8 + jj < out_dim) zrow[__float_as_int(tt[jj].w)] = z[jj];
            }
          }
          ptx::cp_async_wait_all();                    // (the staging issued at this chain's previous gather)
          t2_epi_bar();                                // z2 visible to the row's other threads; staged buffers published
          if (p.k == K - 1) { comp_end(chain, p, zrow, lsum); lsum = 0.f; }
          if ((int)ng < npass) {
            t4_pos_next(a, chain, p);
            start_pass(chain, p, ng);
            if ((int)ng + 1 < npass) {                 // the pass after it goes into the buffers this pass just stopped using
              T4Pos p2 = p;
              t4_pos_next(a, chain, p2);
              stage(chain, a.steps + (p2.c * K + p2.k), (ng + 1u) & 1u);
            }
            ++ng;
          }
          if (chain) { pY = p; ngY = ng; lsumY = lsum; } else { pX = p; ngX = ng; lsumX = lsum; }
        }
      }
    }
    ptx::tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tbase, 512);
  }
}

inline cudaError_t tc6_configure() {
  cudaError_t e = cudaFuncSetAttribute(coupling_tc6_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(coupling_tc6_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  return e;
}

inline int tc6_launch(const CouplingArgs& a, const TcPlan& p, int grid, cudaStream_t st) {
  if (p.tanh_mode == 0) coupling_tc6_kernel<0><<<grid, kT2Threads, p.smem_bytes, st>>>(a, p);
  else                  coupling_tc6_kernel<1><<<grid, kT2Threads, p.smem_bytes, st>>>(a, p);
  return 0;
}

}  // namespace gbnf
