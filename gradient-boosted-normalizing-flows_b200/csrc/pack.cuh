// Parameter re-tiling: raw reference tensors (device) -> the handle's packed blobs (device).
#pragma once
#include "common.cuh"

namespace gbnf {

// Device copy of gbnf_step_params plus what the meta kernel needs.
struct PackMetaArgs {
  const gbnf_step_params* steps;   // device array [K]
  const StepDesc* sdesc;           // device array [K] (this component's slice)
  CompDesc cdesc;
  ModelDims md;
  int flip_init;
  const float* base_mean;          // logical order, may be null
  const float* base_scale;
  float* fblob;
  int* iblob;
  float* wblob_f32;                // invconv matrices (fp32 path only)
};

// Composes the permutations / flips of one component and writes, per step, the elementwise affine constants in
// physical column order, the gather indices of z1 / z2, and the component's data-independent log-det constant.
//   Glow   : ActNorm1d (models/layers.py:488-518) then Permute1d (models/layers.py:661-668) then split (utilities.py:153).
//   RealNVP: eval-mode BatchNorm (models/layers.py:349-358) then flip/split (models/transformations.py:568-576).
// Tiny and sequential by nature (K * D elements): one thread.
__global__ void pack_meta_kernel(PackMetaArgs a) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int D = a.md.D, Dv = a.md.Dv, h0 = D / 2, h1 = D - h0;
  int sigma[kMaxD], tmp[kMaxD];
  for (int j = 0; j < D; ++j) sigma[j] = j;
  double ldj_const = 0.0;
  for (int k = 0; k < a.md.K; ++k) {
    const gbnf_step_params& sp = a.steps[k];
    const StepDesc& sd = a.sdesc[k];
    float* add = a.fblob + sd.vec_off;
    float* mul = add + Dv;
    float* off = mul + Dv;
    int* idx1 = a.iblob + sd.idx_off;
    int* idx2 = idx1 + sd.in_dim;
    for (int p = 0; p < Dv; ++p) { add[p] = 0.f; mul[p] = 1.f; off[p] = 0.f; }
    if (a.md.kind == GBNF_KIND_GLOW) {
      for (int j = 0; j < D; ++j) {
        add[sigma[j]] = sp.an_bias[j];
        mul[sigma[j]] = expf(sp.an_logs[j]);
        ldj_const += (double)sp.an_logs[j];
      }
      if (sd.has_invconv) {
        // InvertibleConv1x1 on a feature vector (models/layers.py:781-796 with h = w = 1): z_new[i] = sum_j W[i][j] y[j], y in the
        // LOGICAL order (y[j] sits at physical column sigma[j]).  Forward matrix M[k = sigma[j]][n = i] = W[i][j]; the result is
        // stored in logical order, i.e. the physical layout becomes the identity.  Inverse: y[j] = sum_i Winv[j][i] z[i], written
        // back to physical column sigma[j]: Minv[k = i][n = sigma[j]] = Winv[j][i].
        float* M = a.wblob_f32 + sd.icw_off;
        float* Mi = a.wblob_f32 + sd.icwinv_off;
        for (int i = 0; i < sd.ic_Kp * sd.ic_Np; ++i) { M[i] = 0.f; Mi[i] = 0.f; }
        for (int i = 0; i < D; ++i)
          for (int j = 0; j < D; ++j) {
            M[sigma[j] * sd.ic_Np + i] = sp.invconv_w[i * D + j];
            if (sp.invconv_winv != nullptr) Mi[i * sd.ic_Np + sigma[j]] = sp.invconv_winv[j * D + i];
          }
        for (int n = 0; n < sd.ic_Np; ++n) a.fblob[sd.icb_off + n] = 0.f;
        ldj_const += (double)sp.invconv_logdet[0];
        for (int j = 0; j < D; ++j) sigma[j] = j;
      } else {
        for (int j = 0; j < D; ++j) tmp[j] = sigma[(int)sp.perm[j]];
        for (int j = 0; j < D; ++j) sigma[j] = tmp[j];
      }
      for (int j = 0; j < h0; ++j) idx1[j] = sigma[j];
      for (int j = 0; j < h1; ++j) idx2[j] = sigma[h0 + j];
    } else {
      if (sp.bn_mean != nullptr) {
        for (int j = 0; j < D; ++j) {
          double var = (double)sp.bn_var[j] + 1e-5;
          add[sigma[j]] = -sp.bn_mean[j];
          mul[sigma[j]] = (float)(exp((double)sp.bn_log_gamma[j]) / sqrt(var));
          off[sigma[j]] = sp.bn_beta[j];
          ldj_const += (double)sp.bn_log_gamma[j] - 0.5 * log(var);
        }
      }
      const bool flipped = ((k + a.flip_init) % 2) > 0;
      if (!flipped) {
        for (int j = 0; j < h0; ++j) idx1[j] = sigma[j];
        for (int j = 0; j < h1; ++j) idx2[j] = sigma[h0 + j];
      } else {
        for (int j = 0; j < h1; ++j) idx1[j] = sigma[h0 + j];
        for (int j = 0; j < h0; ++j) idx2[j] = sigma[j];
        for (int j = 0; j < h1; ++j) tmp[j] = sigma[h0 + j];      // output is [z1, z2'] for both flips
        for (int j = 0; j < h0; ++j) tmp[h1 + j] = sigma[j];
        for (int j = 0; j < D; ++j) sigma[j] = tmp[j];
      }
    }
  }
  // gather-order, padded copies of the step tables (pipelined tensor-core kernel)
  for (int k = 0; k < a.md.K; ++k) {
    const StepDesc& sd = a.sdesc[k];
    const float* add = a.fblob + sd.vec_off;
    const float* mul = add + Dv;
    const float* off = mul + Dv;
    const int* idx1 = a.iblob + sd.idx_off;
    const int* idx2 = idx1 + sd.in_dim;
    float4* t1 = reinterpret_cast<float4*>(a.fblob + sd.ep_off);     // {add, mul, off, column (int bits)} in z1 order
    float4* t2 = t1 + kEpPad;                                         // same in z2 order
    float4* u1 = t2 + kEpPad;                                         // INVERSE affine in the same form: x = (y + -off) * (1 / mul) + -add
    float4* u2 = u1 + kEpPad;
    for (int j = 0; j < kEpPad; ++j) {
      const bool v1 = j < sd.in_dim, v2 = j < sd.out_dim;
      const int c1 = v1 ? idx1[j] : D, c2 = v2 ? idx2[j] : D;
      t1[j] = make_float4(v1 ? add[c1] : 0.f, v1 ? mul[c1] : 0.f, v1 ? off[c1] : 0.f, __int_as_float(c1));
      t2[j] = make_float4(v2 ? add[c2] : 0.f, v2 ? mul[c2] : 0.f, v2 ? off[c2] : 0.f, __int_as_float(c2));
      u1[j] = make_float4(v1 ? -off[c1] : 0.f, v1 ? 1.0f / mul[c1] : 0.f, v1 ? -add[c1] : 0.f, __int_as_float(c1));
      u2[j] = make_float4(v2 ? -off[c2] : 0.f, v2 ? 1.0f / mul[c2] : 0.f, v2 ? -add[c2] : 0.f, __int_as_float(c2));
    }
  }
  int* sig_out = a.iblob + a.cdesc.sigma_off;
  for (int j = 0; j < D; ++j) sig_out[j] = sigma[j];
  a.fblob[a.cdesc.const_off] = (float)ldj_const;
  a.fblob[a.cdesc.const_off + 1] = 1.f;
  // base density constants in physical order
  float* bm = a.fblob + a.cdesc.base_off;
  float* bi = bm + Dv;
  double c0 = -0.5 * (double)D * 1.8378770664093453;   // -D/2 log(2 pi)
  for (int p = 0; p < Dv; ++p) { bm[p] = 0.f; bi[p] = 0.f; }
  if (a.md.base == GBNF_BASE_DIAG_NORMAL && a.base_mean != nullptr) {
    for (int j = 0; j < D; ++j) {
      double s = (double)a.base_scale[j];
      bm[sigma[j]] = a.base_mean[j];
      bi[sigma[j]] = (float)(1.0 / (2.0 * s * s));
      c0 -= log(s);
    }
  } else {
    for (int j = 0; j < D; ++j) bi[sigma[j]] = 0.5f;
  }
  bm[2 * Dv] = (float)c0;
  a.fblob[a.cdesc.dense_off] = (float)ldj_const;
  a.fblob[a.cdesc.dense_off + 1] = (float)c0;
}

// fp32 path: dst[k][n] = W[n][k] (zero padded).
__global__ void pack_weight_fp32_kernel(const float* __restrict__ W, const float* __restrict__ b, LayerDesc ld,
                                        float* __restrict__ wblob, float* __restrict__ fblob) {
  const long long total = (long long)ld.Kp * ld.Np;
  float* dst = wblob + ld.w_off;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int k = (int)(i / ld.Np), n = (int)(i % ld.Np);
    dst[i] = (k < ld.K_in && n < ld.N_out) ? W[(long long)n * ld.K_in + k] : 0.f;
  }
  if (blockIdx.x == 0)
    for (int n = threadIdx.x; n < ld.Np; n += blockDim.x) fblob[ld.b_off + n] = (n < ld.N_out) ? b[n] : 0.f;
}

// f16 path: UMMA canonical K-major no-swizzle k-slabs (see LayerDesc).  Sets *overflow if |w| is not finite in fp16.
__global__ void pack_weight_f16_kernel(const float* __restrict__ W, const float* __restrict__ b, LayerDesc ld,
                                       __half* __restrict__ wblob, float* __restrict__ fblob, int* overflow) {
  const long long total = (long long)ld.Kp * ld.Np;
  __half* dst = wblob + ld.w_off;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    // decode destination index -> (n, k): [N chunk][k-slab][width/8 row groups][2 k-chunks][8 rows][8 halves], chunk widths
    // alternating NC, NC2
    const int nce = ld.NC, nco = ld.NC2 ? ld.NC2 : ld.NC;
    const long long pair_elems = (long long)(nce + nco) * ld.Kp;
    const long long pr = i / pair_elems;
    long long r = i % pair_elems;
    int w = nce, col0 = (int)pr * (nce + nco);
    if (r >= (long long)nce * ld.Kp) { r -= (long long)nce * ld.Kp; w = nco; col0 += nce; }
    const int slab = (int)(r / (w * 16)), rr = (int)(r % (w * 16));
    const int grp = rr / 128, kc = (rr % 128) / 64, row = (rr % 64) / 8, e = rr % 8;
    const int n = col0 + grp * 8 + row, k = slab * 16 + kc * 8 + e;
    float v = (k < ld.K_in && n < ld.N_out) ? W[(long long)n * ld.K_in + k] : 0.f;
    __half hv = __float2half_rn(v);
    if (__hisinf(hv) || __hisnan(hv)) *overflow = 1;
    dst[i] = hv;
  }
  if (blockIdx.x == 0)
    for (int n = threadIdx.x; n < ld.Np; n += blockDim.x) fblob[ld.b_off + n] = (n < ld.N_out) ? b[n] : 0.f;
}

}  // namespace gbnf
