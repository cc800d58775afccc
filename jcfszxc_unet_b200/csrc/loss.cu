// Segmentation head + loss, fused: OutConv (1x1, C -> 1, bias) + sigmoid + BCEWithLogits + dice sums in
// ONE pass over the last activation (132 B/pixel at C = 64), and the matching backward that rebuilds
// dL/dlogit from (logit, label, global sums) and emits the activation gradient plus dW/db partials.
// Reference semantics replaced: OutConv.forward (UNetFamily/utils/unet_parts.py:73-79),
// train.py:264-278 (sigmoid, BCEWithLogitsLoss, dice_loss, alpha = 0.5) and utils/dice_score.py:13-59.
// The three dice sums are produced separately from the finalize step so that data-parallel ranks can
// all-reduce them (the reference's dice is ONE ratio over the whole batch).
#include "host_common.cuh"
#include "kernels.cuh"
#include "ptx.cuh"
#include "reduce2.cuh"

namespace unetk {

namespace {

constexpr int kThreads = 256;
constexpr float kClampLo = 1e-7f, kClampHi = 1.0f - 1e-7f;  // dice_loss clamp, dice_score.py:56
constexpr double kDiceEps = 1e-5;                            // dice_score.py:32

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}

__device__ __forceinline__ float group_sum(float v, int lpp) {
  for (int o = lpp >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// sums layout (double[4]): 0 = sum bce, 1 = sum p*y, 2 = sum p, 3 = sum y      (p clamped to [1e-7, 1-1e-7])
__global__ void __launch_bounds__(kThreads)
head_loss_fwd_kernel(const __nv_bfloat16* __restrict__ x, int64_t ld, const float* __restrict__ w,
                     const float* __restrict__ bias, const float* __restrict__ labels, float* __restrict__ logits,
                     int post_sigmoid, int64_t npix, int C, float* __restrict__ partial) {
  pdl_trigger();
  pdl_wait();
  const int lpp = C >> 3;                 // lanes per pixel (power of two <= 32)
  const int gpb = kThreads / lpp;         // pixel groups per block
  const int sub = threadIdx.x % lpp, grp = threadIdx.x / lpp;
  float wv[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) wv[j] = __ldg(w + sub * 8 + j);
  const float b = bias ? __ldg(bias) : 0.f;
  float s_bce = 0.f, s_py = 0.f, s_p = 0.f, s_y = 0.f;
  // uniform trip count per block (all lanes of a pixel group iterate together); kU pixels per group are loaded
  // before any is reduced, so every thread keeps kU 16-byte loads in flight (the first version had one: 2 TB/s)
  constexpr int kU = 4;
  const int64_t step = static_cast<int64_t>(gridDim.x) * gpb;
  for (int64_t base = static_cast<int64_t>(blockIdx.x) * gpb; base < npix; base += kU * step) {
    uint4 v[kU];
    float yv[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int64_t pix = base + u * step + grp;
      const bool ok = pix < npix;
      v[u] = ok ? __ldg(reinterpret_cast<const uint4*>(x + pix * ld + sub * 8)) : make_uint4(0, 0, 0, 0);
      yv[u] = (ok && sub == 0 && labels != nullptr) ? __ldg(labels + pix) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int64_t pix = base + u * step + grp;
      float f[8], dot = 0.f;
      unpack8(v[u], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) dot = fmaf(f[j], wv[j], dot);
      dot = group_sum(dot, lpp);
      if (sub == 0 && pix < npix) {
        // post_sigmoid: the model itself ends in nn.Sigmoid (ResUNet.py:47-50, UNetPP.py:105-106) and train.py
        // still feeds that output to BCEWithLogits / sigmoid+dice, so the loss sees sigmoid(z) as its "logit"
        const float z = post_sigmoid ? 1.f / (1.f + __expf(-(dot + b))) : dot + b;
        logits[pix] = z;
        if (labels != nullptr) {
          const float y = yv[u];
          s_bce += fmaxf(z, 0.f) - z * y + log1pf(__expf(-fabsf(z)));
          float p = 1.f / (1.f + __expf(-z));
          p = fminf(fmaxf(p, kClampLo), kClampHi);
          s_py = fmaf(p, y, s_py);
          s_p += p;
          s_y += y;
        }
      }
    }
  }
  __shared__ float red[4][kThreads / 32];
  s_bce = warp_sum(s_bce); s_py = warp_sum(s_py); s_p = warp_sum(s_p); s_y = warp_sum(s_y);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = s_bce; red[1][warp] = s_py; red[2][warp] = s_p; red[3][warp] = s_y; }
  __syncthreads();
  if (threadIdx.x < 4) {
    float s = 0.f;
    for (int i = 0; i < kThreads / 32; ++i) s += red[threadIdx.x][i];
    partial[static_cast<size_t>(blockIdx.x) * 4 + threadIdx.x] = s;
  }
}

__global__ void loss_sums_kernel(const float* __restrict__ partial, int nblk, double* __restrict__ sums) {
  pdl_trigger();
  pdl_wait();
  const int k = threadIdx.x;
  if (k >= 4) return;
  double s = 0.0;
  for (int b = 0; b < nblk; ++b) s += partial[static_cast<size_t>(b) * 4 + k];
  sums[k] = s;
}

// out[0]=loss out[1]=bce out[2]=dice out[3]=1/npix  out[4]=cA out[5]=cB  (d dice / d p_i = y_i*cA - cB)
__global__ void loss_finalize_kernel(const double* __restrict__ sums, double npix_total, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  if (threadIdx.x != 0) return;
  const double bce = sums[0] / npix_total;
  const double inter = 2.0 * sums[1];
  double sets = sums[2] + sums[3];
  double cA, cB;
  if (sets < kDiceEps) {  // empty-mask branch (dice_score.py:35): sets_sum := inter -> dice == 1, zero gradient
    sets = inter;
    cA = 0.0; cB = 0.0;
  } else {
    cA = 2.0 / (sets + kDiceEps);
    cB = (inter + kDiceEps) / ((sets + kDiceEps) * (sets + kDiceEps));
  }
  const double dice = (inter + kDiceEps) / (sets + kDiceEps);
  out[0] = static_cast<float>(0.5 * bce + 0.5 * (1.0 - dice));
  out[1] = static_cast<float>(bce);
  out[2] = static_cast<float>(dice);
  out[3] = static_cast<float>(1.0 / npix_total);
  out[4] = static_cast<float>(cA);
  out[5] = static_cast<float>(cB);
}

// dz = gscale * [ 0.5 * (sigmoid(z) - y) / Npix  -  0.5 * (y*cA - cB) * p(1-p) * 1[clamp inactive] ]
// dx[pix][c] = dz * w[c];  partial[blk][c] = sum dz * x[pix][c];  partial[blk][C] = sum dz
__global__ void __launch_bounds__(kThreads, 4)   // head_grid launches 4 blocks per SM: keep them all resident
head_loss_bwd_kernel(const __nv_bfloat16* __restrict__ x, int64_t ld, const float* __restrict__ w,
                     const float* __restrict__ labels, const float* __restrict__ logits,
                     const float* __restrict__ fin, const float* __restrict__ dlogits, float gscale, int post_sigmoid,
                     __nv_bfloat16* __restrict__ dx, int64_t dx_ld, int64_t npix, int C,
                     float* __restrict__ partial) {
  pdl_trigger();
  pdl_wait();
  const int lpp = C >> 3;
  const int gpb = kThreads / lpp;
  const int sub = threadIdx.x % lpp, grp = threadIdx.x / lpp;
  float wv[8], acc[8] = {};
#pragma unroll
  for (int j = 0; j < 8; ++j) wv[j] = __ldg(w + sub * 8 + j);
  float inv_n = 0.f, cA = 0.f, cB = 0.f;
  if (dlogits == nullptr) { inv_n = __ldg(fin + 3); cA = __ldg(fin + 4); cB = __ldg(fin + 5); }
  float s_dz = 0.f;
  constexpr int kU = 4;
  const int64_t step = static_cast<int64_t>(gridDim.x) * gpb;
  for (int64_t p0 = static_cast<int64_t>(blockIdx.x) * gpb + grp; p0 < npix; p0 += kU * step) {
    uint4 v[kU];
    float dzv[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int64_t pix = p0 + u * step;
      const bool ok = pix < npix;
      v[u] = ok ? __ldg(reinterpret_cast<const uint4*>(x + pix * ld + sub * 8)) : make_uint4(0, 0, 0, 0);
      float dz = 0.f;
      if (ok) {
        if (dlogits != nullptr) {  // gradient handed in by autograd (loss computed outside the library)
          dz = gscale * __ldg(dlogits + pix);
        } else {
          const float z = __ldg(logits + pix), y = __ldg(labels + pix);
          const float p = 1.f / (1.f + __expf(-z));
          const float inside = (p >= kClampLo && p <= kClampHi) ? 1.f : 0.f;
          dz = gscale * (0.5f * inv_n * (p - y) - 0.5f * (y * cA - cB) * p * (1.f - p) * inside);
        }
        if (post_sigmoid) {  // chain through the model's own output sigmoid: logits[] holds its value
          const float o = __ldg(logits + pix);
          dz *= o * (1.f - o);
        }
      }
      dzv[u] = dz;
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int64_t pix = p0 + u * step;
      if (pix >= npix) break;
      const float dz = dzv[u];
      float f[8], o[8];
      unpack8(v[u], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) { acc[j] = fmaf(dz, f[j], acc[j]); o[j] = dz * wv[j]; }
      uint4 w4;
      w4.x = pack_bf16x2(o[0], o[1]); w4.y = pack_bf16x2(o[2], o[3]);
      w4.z = pack_bf16x2(o[4], o[5]); w4.w = pack_bf16x2(o[6], o[7]);
      *reinterpret_cast<uint4*>(dx + pix * dx_ld + sub * 8) = w4;
      if (sub == 0) s_dz += dz;
    }
  }
  extern __shared__ float red[];  // [gpb][C + 1]
#pragma unroll
  for (int j = 0; j < 8; ++j) red[grp * (C + 1) + sub * 8 + j] = acc[j];
  if (sub == 0) red[grp * (C + 1) + C] = s_dz;
  __syncthreads();
  for (int i = threadIdx.x; i <= C; i += kThreads) {
    float s = 0.f;
    for (int g = 0; g < gpb; ++g) s += red[g * (C + 1) + i];
    partial[static_cast<size_t>(blockIdx.x) * (C + 1) + i] = s;
  }
}

__global__ void head_bwd_finalize_kernel(const float* __restrict__ partial, int nblk, int C, float* dw, float* db,
                                         int accumulate) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > C) return;
  double s = 0.0;
  for (int b = 0; b < nblk; ++b) s += partial[static_cast<size_t>(b) * (C + 1) + i];
  float* o = (i < C) ? dw + i : db;
  if (o) *o = accumulate ? *o + static_cast<float>(s) : static_cast<float>(s);
}

int head_grid(int64_t npix, int C) {
  const int gpb = kThreads / (C / 8);
  int64_t b = (npix + gpb * 8 - 1) / (gpb * 8);
  const int64_t cap = static_cast<int64_t>(num_sms()) * 4;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}
bool bn_head_c_ok(int C) { return C == 32 || C == 64; }   // the widths in front of OutConv in the U-Net family
bool head_c_ok(int C) { return C == 8 || C == 16 || C == 32 || C == 64 || C == 128 || C == 256; }


// ------------------------------------------------------------------ last BatchNorm + ReLU folded into the head
// The activation in front of the head, a = relu(bf16(raw*scale + shift)), has the head as its ONLY consumer and its
// gradient is rank one (dz[pix] * w[c]), so neither `a` nor d(a) has to exist in HBM:
//   forward : logits straight from the conv output `raw`                       (2 B/elem instead of 4 + 2)
//   backward: one pass over `raw` gives dz, dW/db of the head AND the two BatchNorm backward sums; a second pass
//             writes d(raw)                                                     (2 + 4 B/elem instead of 4 + 4 + 6)
// Same arithmetic as bn_apply -> head_fwd and head_bwd -> bn_bwd_reduce/apply (bf16 rounding of a and of dz*w kept).
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

__device__ __forceinline__ float head_dz(const float* __restrict__ labels, const float* __restrict__ logits,
                                         const float* __restrict__ dlogits, float gscale, int post_sigmoid,
                                         float inv_n, float cA, float cB, int64_t pix) {
  float dz;
  if (dlogits != nullptr) {
    dz = gscale * __ldg(dlogits + pix);
  } else {
    const float z = __ldg(logits + pix), y = __ldg(labels + pix);
    const float p = 1.f / (1.f + __expf(-z));
    const float inside = (p >= kClampLo && p <= kClampHi) ? 1.f : 0.f;
    dz = gscale * (0.5f * inv_n * (p - y) - 0.5f * (y * cA - cB) * p * (1.f - p) * inside);
  }
  if (post_sigmoid) {
    const float o = __ldg(logits + pix);
    dz *= o * (1.f - o);
  }
  return dz;
}

// Thread layout of the two kernels below (LPP = C/8 lanes per pixel, GW = 32/LPP pixels per warp and load round,
// kU = 8 load rounds in flight): a warp iteration covers S = 8*GW pixel "slots" s = u*GW + j.  The per-PIXEL math
// (logit -> loss terms, or logit/label -> dz: exp, log1p, reciprocal, ~100 instructions) is done ONCE per pixel with
// all 32 lanes busy — lane L owns slots L, L+32, ... — and moved to / from the lanes that hold the pixel's channels
// with shuffles.  (The first version evaluated it in every lane of a pixel group, i.e. at 4 live results per 32
// lanes, and ran at 1.2-2.1 TB/s: issue-bound.)
template <int LPP>
__global__ void __launch_bounds__(kThreads, 2)
bn_head_fwd_kernel(const __nv_bfloat16* __restrict__ raw, int64_t ld, const float* __restrict__ scale,
                   const float* __restrict__ shift, int relu, const float* __restrict__ w,
                   const float* __restrict__ bias, const float* __restrict__ labels, float* __restrict__ logits,
                   int post_sigmoid, int64_t npix, float* __restrict__ partial) {
  pdl_trigger();
  pdl_wait();
  constexpr int GW = 32 / LPP, kU = 8, R = (kU * GW) / 32, gpb = kThreads / LPP;
  static_assert(R >= 1, "C > 64 is not instantiated");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane % LPP, grp = warp * GW + lane / LPP;
  float wv[8], sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    wv[j] = __ldg(w + sub * 8 + j); sc[j] = __ldg(scale + sub * 8 + j); sh[j] = __ldg(shift + sub * 8 + j);
  }
  const float b = bias ? __ldg(bias) : 0.f;
  float s_bce = 0.f, s_py = 0.f, s_p = 0.f, s_y = 0.f;
  const int64_t step = static_cast<int64_t>(gridDim.x) * gpb;
  for (int64_t base = static_cast<int64_t>(blockIdx.x) * gpb; base < npix; base += kU * step) {
    uint4 v[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int64_t pix = base + u * step + grp;
      v[u] = pix < npix ? __ldg(reinterpret_cast<const uint4*>(raw + pix * ld + sub * 8)) : make_uint4(0, 0, 0, 0);
    }
    // the rows of the NEXT iteration go to L2 now (one lane per pixel row): with 16 warps per SM and a long compute
    // phase per iteration the kernel was latency-bound on these loads (profiles/r01_ncu_stem_head.txt)
    if (sub == 0) {
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int64_t pix = base + (kU + u) * step + grp;
        if (pix < npix) prefetch_l2(raw + pix * ld);
      }
    }
    float mine[R];
#pragma unroll
    for (int r = 0; r < R; ++r) mine[r] = 0.f;
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      float f[8], dot = 0.f;
      unpack8(v[u], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float z = bf16_round(fmaf(f[j], sc[j], sh[j]));
        dot = fmaf(relu ? fmaxf(z, 0.f) : z, wv[j], dot);
      }
      dot = group_sum(dot, LPP);
      const float t = __shfl_sync(0xffffffffu, dot, (lane % GW) * LPP);
#pragma unroll
      for (int r = 0; r < R; ++r)
        if ((lane + 32 * r) / GW == u) mine[r] = t;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int slot = lane + 32 * r;
      const int64_t pix = base + (slot / GW) * step + warp * GW + slot % GW;
      if (pix < npix) {
        const float z = post_sigmoid ? 1.f / (1.f + __expf(-(mine[r] + b))) : mine[r] + b;
        logits[pix] = z;
        if (labels != nullptr) {
          const float y = __ldg(labels + pix);
          s_bce += fmaxf(z, 0.f) - z * y + log1pf(__expf(-fabsf(z)));
          float p = 1.f / (1.f + __expf(-z));
          p = fminf(fmaxf(p, kClampLo), kClampHi);
          s_py = fmaf(p, y, s_py);
          s_p += p;
          s_y += y;
        }
      }
    }
  }
  __shared__ float red[4][kThreads / 32];
  s_bce = warp_sum(s_bce); s_py = warp_sum(s_py); s_p = warp_sum(s_p); s_y = warp_sum(s_y);
  if (lane == 0) { red[0][warp] = s_bce; red[1][warp] = s_py; red[2][warp] = s_p; red[3][warp] = s_y; }
  __syncthreads();
  if (threadIdx.x < 4) {
    float s = 0.f;
    for (int i = 0; i < kThreads / 32; ++i) s += red[threadIdx.x][i];
    partial[static_cast<size_t>(blockIdx.x) * 4 + threadIdx.x] = s;
  }
}

// partial[blk][0][c] = sum dz*a   (head dW)          partial[blk][1][c] = sum gm          (BN: S0)
// partial[blk][2][c] = sum gm*(raw - mean) (BN: S1)  partial[blk][3C]   = sum dz          (head db)
//   a = act(bf16(raw*scale+shift)),  gm = mask * bf16(dz*w[c]);   dz[pix] is kept (fp32) for the apply pass.
template <int LPP>
__global__ void __launch_bounds__(kThreads, 2)
bn_head_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ raw, int64_t ld, const float* __restrict__ scale,
                          const float* __restrict__ shift, const float* __restrict__ mean, int relu,
                          const float* __restrict__ w, const float* __restrict__ labels,
                          const float* __restrict__ logits, const float* __restrict__ fin,
                          const float* __restrict__ dlogits, float gscale, int post_sigmoid,
                          float* __restrict__ dz_out, int64_t npix, float* __restrict__ partial) {
  pdl_trigger();
  pdl_wait();
  constexpr int GW = 32 / LPP, kU = 8, R = (kU * GW) / 32, gpb = kThreads / LPP, C = 8 * LPP;
  static_assert(R >= 1, "C > 64 is not instantiated");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane % LPP, jg = lane / LPP, grp = warp * GW + jg;
  float wv[8], sc[8], sh[8], mu[8], acc[3][8] = {};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    wv[j] = __ldg(w + sub * 8 + j); sc[j] = __ldg(scale + sub * 8 + j);
    sh[j] = __ldg(shift + sub * 8 + j); mu[j] = __ldg(mean + sub * 8 + j);
  }
  float inv_n = 0.f, cA = 0.f, cB = 0.f;
  if (dlogits == nullptr) { inv_n = __ldg(fin + 3); cA = __ldg(fin + 4); cB = __ldg(fin + 5); }
  float s_dz = 0.f;
  const int64_t step = static_cast<int64_t>(gridDim.x) * gpb;
  for (int64_t base = static_cast<int64_t>(blockIdx.x) * gpb; base < npix; base += kU * step) {
    uint4 v[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int64_t pix = base + u * step + grp;
      v[u] = pix < npix ? __ldg(reinterpret_cast<const uint4*>(raw + pix * ld + sub * 8)) : make_uint4(0, 0, 0, 0);
    }
    // the rows of the NEXT iteration go to L2 now (one lane per pixel row): with 16 warps per SM and a long compute
    // phase per iteration the kernel was latency-bound on these loads (profiles/r01_ncu_stem_head.txt)
    if (sub == 0) {
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int64_t pix = base + (kU + u) * step + grp;
        if (pix < npix) prefetch_l2(raw + pix * ld);
      }
    }
    float dzl[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int slot = lane + 32 * r;
      const int64_t pix = base + (slot / GW) * step + warp * GW + slot % GW;
      dzl[r] = 0.f;
      if (pix < npix) {
        dzl[r] = head_dz(labels, logits, dlogits, gscale, post_sigmoid, inv_n, cA, cB, pix);
        dz_out[pix] = dzl[r];
        s_dz += dzl[r];
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      // slot u*GW + jg lives in lane (u*GW + jg) % 32 of round (u*GW) / 32; out-of-range pixels carry dz = 0, raw = 0
      const float dz = __shfl_sync(0xffffffffu, dzl[(u * GW) / 32], (u * GW + jg) & 31);
      float f[8];
      unpack8(v[u], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float z = bf16_round(fmaf(f[j], sc[j], sh[j]));
        const bool on = !relu || z > 0.f;
        const float a = on ? z : 0.f;
        const float gm = on ? bf16_round(dz * wv[j]) : 0.f;
        acc[0][j] = fmaf(dz, a, acc[0][j]);
        acc[1][j] += gm;
        acc[2][j] = fmaf(gm, f[j] - mu[j], acc[2][j]);
      }
    }
  }
  extern __shared__ float red[];  // [gpb][3C] + [8 warps]
  constexpr int row = 3 * C;
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) red[grp * row + k * C + sub * 8 + j] = acc[k][j];
  s_dz = warp_sum(s_dz);
  if (lane == 0) red[gpb * row + warp] = s_dz;
  __syncthreads();
  for (int i = threadIdx.x; i <= row; i += kThreads) {
    float s = 0.f;
    if (i < row) {
      for (int g = 0; g < gpb; ++g) s += red[g * row + i];
    } else {
      for (int g = 0; g < kThreads / 32; ++g) s += red[gpb * row + g];
    }
    partial[static_cast<size_t>(blockIdx.x) * (row + 1) + i] = s;
  }
}

// i < C: dw[i];  C <= i < 3C: sums[i - C] (double, the BatchNorm backward sums S0 | S1);  i == 3C: db
__global__ void bn_head_bwd_sums_kernel(const float* __restrict__ partial, int nblk, int C, float* dw, float* db,
                                        int accumulate, double* __restrict__ sums) {
  pdl_trigger();
  pdl_wait();
  const int row = 3 * C + 1;
  const int i = blockIdx.x * kSum2Lanes + threadIdx.x;
  const bool valid = i < row;
  const double s = sliced_ordered_sum(partial, nblk, valid, [&](int b) { return static_cast<size_t>(b) * row + i; });
  if (!valid || threadIdx.y != 0) return;
  if (i >= C && i < 3 * C) {
    sums[i - C] = s;
  } else {
    float* o = (i < C) ? (dw ? dw + i : nullptr) : db;
    if (o) *o = accumulate ? *o + static_cast<float>(s) : static_cast<float>(s);
  }
}

// d(raw) = scale*gm + K1*raw + K0  (coef = [K0 | K1], see bn_bwd_finalize_kernel in elementwise.cu)
__global__ void __launch_bounds__(kThreads, 2)
bn_head_bwd_apply_kernel(const __nv_bfloat16* __restrict__ raw, int64_t ld, const float* __restrict__ scale,
                         const float* __restrict__ shift, int relu, const float* __restrict__ w,
                         const float* __restrict__ dz_in, const float* __restrict__ coef,
                         __nv_bfloat16* __restrict__ draw, int64_t draw_ld, int64_t npix, int C) {
  pdl_trigger();
  pdl_wait();
  const int lpp = C >> 3;
  const int gpb = kThreads / lpp;
  const int sub = threadIdx.x % lpp, grp = threadIdx.x / lpp;
  float wv[8], sc[8], sh[8], k0[8], k1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    wv[j] = __ldg(w + sub * 8 + j); sc[j] = __ldg(scale + sub * 8 + j); sh[j] = __ldg(shift + sub * 8 + j);
    k0[j] = __ldg(coef + sub * 8 + j); k1[j] = __ldg(coef + C + sub * 8 + j);
  }
  constexpr int kU = 8;
  const int64_t step = static_cast<int64_t>(gridDim.x) * gpb;
  for (int64_t p0 = static_cast<int64_t>(blockIdx.x) * gpb + grp; p0 < npix; p0 += kU * step) {
    uint4 v[kU];
    float dzv[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int64_t pix = p0 + u * step;
      const bool ok = pix < npix;
      v[u] = ok ? __ldg(reinterpret_cast<const uint4*>(raw + pix * ld + sub * 8)) : make_uint4(0, 0, 0, 0);
      dzv[u] = ok ? __ldg(dz_in + pix) : 0.f;
    }
    if (sub == 0) {   // next iteration's rows -> L2
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int64_t pix = p0 + (kU + u) * step;
        if (pix < npix) prefetch_l2(raw + pix * ld);
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int64_t pix = p0 + u * step;
      if (pix >= npix) break;
      float f[8], o[8];
      unpack8(v[u], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float z = bf16_round(fmaf(f[j], sc[j], sh[j]));
        const float gm = (!relu || z > 0.f) ? bf16_round(dzv[u] * wv[j]) : 0.f;
        o[j] = fmaf(sc[j], gm, fmaf(k1[j], f[j], k0[j]));
      }
      uint4 w4;
      w4.x = pack_bf16x2(o[0], o[1]); w4.y = pack_bf16x2(o[2], o[3]);
      w4.z = pack_bf16x2(o[4], o[5]); w4.w = pack_bf16x2(o[6], o[7]);
      *reinterpret_cast<uint4*>(draw + pix * draw_ld + sub * 8) = w4;
    }
  }
}

// one resident wave of 2 blocks per SM (register budget of the backward kernels)
int bn_head_grid(int64_t npix, int C) {
  const int gpb = kThreads / (C / 8);
  int64_t b = (npix + gpb * 8 - 1) / (gpb * 8);
  const int64_t cap = static_cast<int64_t>(num_sms()) * 2;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

}  // namespace

size_t head_partial_floats(int64_t npix, int C) { return static_cast<size_t>(head_grid(npix, C)) * (C + 1 > 4 ? C + 1 : 4); }

int head_loss_fwd_run(const void* x, int64_t ld, const float* w, const float* bias, const float* labels, float* logits,
                      int post_sigmoid, int64_t npix, int C, float* partial, double* sums, cudaStream_t s) {
  UNETK_CHECK(head_c_ok(C), -1, "head: C=%d must be a power of two in [8,256]", C);
  const int grid = head_grid(npix, C);
  UNETK_CUDA(launch_pdl(head_loss_fwd_kernel, dim3(grid), dim3(kThreads), 0, s, static_cast<const __nv_bfloat16*>(x), ld, w, bias, labels, logits,
                                                post_sigmoid, npix, C, partial));
  UNETK_LAUNCHED();
  if (labels != nullptr) {
    UNETK_CHECK(sums != nullptr, -1, "head_loss_fwd: sums is null");
    UNETK_CUDA(launch_pdl(loss_sums_kernel, dim3(1), dim3(32), 0, s, partial, grid, sums));
    UNETK_LAUNCHED();
  }
  return 0;
}

int loss_finalize_run(const double* sums, double npix_total, float* out, cudaStream_t s) {
  UNETK_CUDA(launch_pdl(loss_finalize_kernel, dim3(1), dim3(32), 0, s, sums, npix_total, out));
  UNETK_LAUNCHED();
  return 0;
}

int head_loss_bwd_run(const void* x, int64_t ld, const float* w, const float* labels, const float* logits,
                      const float* fin, const float* dlogits, float gscale, int post_sigmoid, void* dx, int64_t dx_ld,
                      float* dw, float* db, int accumulate, int64_t npix, int C, float* partial, cudaStream_t s) {
  UNETK_CHECK(dlogits != nullptr || (labels && logits && fin), -1, "head_bwd: need dlogits or (labels, logits, fin)");
  UNETK_CHECK(!post_sigmoid || logits != nullptr, -1, "head_bwd: post_sigmoid needs the forward's output");
  UNETK_CHECK(head_c_ok(C), -1, "head: C=%d must be a power of two in [8,256]", C);
  const int grid = head_grid(npix, C);
  const int gpb = kThreads / (C / 8);
  const size_t smem = static_cast<size_t>(gpb) * (C + 1) * sizeof(float);
  UNETK_CUDA(launch_pdl(head_loss_bwd_kernel, dim3(grid), dim3(kThreads), smem, s, static_cast<const __nv_bfloat16*>(x), ld, w, labels, logits, fin,
                                                   dlogits, gscale, post_sigmoid, static_cast<__nv_bfloat16*>(dx), dx_ld,
                                                   npix, C, partial));
  UNETK_LAUNCHED();
  UNETK_CUDA(launch_pdl(head_bwd_finalize_kernel, dim3((C + 1 + 127) / 128), dim3(128), 0, s, partial, grid, C, dw, db, accumulate));
  UNETK_LAUNCHED();
  return 0;
}

size_t bn_head_partial_floats(int64_t npix, int C) {
  const size_t a = static_cast<size_t>(bn_head_grid(npix, C)) * (3 * C + 1);
  const size_t b = head_partial_floats(npix, C);
  return a > b ? a : b;
}

int bn_head_fwd_run(const void* raw, int64_t ld, const float* scale, const float* shift, int relu, const float* w,
                    const float* bias, const float* labels, float* logits, int post_sigmoid, int64_t npix, int C,
                    float* partial, double* sums, cudaStream_t s) {
  UNETK_CHECK(bn_head_c_ok(C), -1, "bn_head: C=%d must be 32 or 64", C);
  const int grid = bn_head_grid(npix, C);
  if (C == 64)
    UNETK_CUDA(launch_pdl(bn_head_fwd_kernel<8>, dim3(grid), dim3(kThreads), 0, s, static_cast<const __nv_bfloat16*>(raw),
                          ld, scale, shift, relu, w, bias, labels, logits, post_sigmoid, npix, partial));
  else
    UNETK_CUDA(launch_pdl(bn_head_fwd_kernel<4>, dim3(grid), dim3(kThreads), 0, s, static_cast<const __nv_bfloat16*>(raw),
                          ld, scale, shift, relu, w, bias, labels, logits, post_sigmoid, npix, partial));
  UNETK_LAUNCHED();
  if (labels != nullptr) {
    UNETK_CHECK(sums != nullptr, -1, "bn_head_fwd: sums is null");
    UNETK_CUDA(launch_pdl(loss_sums_kernel, dim3(1), dim3(32), 0, s, partial, grid, sums));
    UNETK_LAUNCHED();
  }
  return 0;
}

int bn_head_bwd_reduce_run(const void* raw, int64_t ld, const float* scale, const float* shift, const float* mean,
                           int relu, const float* w, const float* labels, const float* logits, const float* fin,
                           const float* dlogits, float gscale, int post_sigmoid, float* dz, float* dw, float* db,
                           int accumulate, double* sums, int64_t npix, int C, float* partial, cudaStream_t s) {
  UNETK_CHECK(dlogits != nullptr || (labels && logits && fin), -1, "bn_head_bwd: need dlogits or (labels, logits, fin)");
  UNETK_CHECK(!post_sigmoid || logits != nullptr, -1, "bn_head_bwd: post_sigmoid needs the forward's output");
  UNETK_CHECK(bn_head_c_ok(C), -1, "bn_head: C=%d must be 32 or 64", C);
  const int grid = bn_head_grid(npix, C);
  const int gpb = kThreads / (C / 8);
  const size_t smem = (static_cast<size_t>(gpb) * 3 * C + kThreads / 32) * sizeof(float);
  if (C == 64)
    UNETK_CUDA(launch_pdl(bn_head_bwd_reduce_kernel<8>, dim3(grid), dim3(kThreads), smem, s,
                          static_cast<const __nv_bfloat16*>(raw), ld, scale, shift, mean, relu, w, labels, logits, fin,
                          dlogits, gscale, post_sigmoid, dz, npix, partial));
  else
    UNETK_CUDA(launch_pdl(bn_head_bwd_reduce_kernel<4>, dim3(grid), dim3(kThreads), smem, s,
                          static_cast<const __nv_bfloat16*>(raw), ld, scale, shift, mean, relu, w, labels, logits, fin,
                          dlogits, gscale, post_sigmoid, dz, npix, partial));
  UNETK_LAUNCHED();
  UNETK_CUDA(launch_pdl(bn_head_bwd_sums_kernel, dim3((3 * C + 1 + kSum2Lanes - 1) / kSum2Lanes),
                        dim3(kSum2Lanes, kSum2Slices), 0, s, partial, grid, C, dw, db, accumulate, sums));
  UNETK_LAUNCHED();
  return 0;
}

int bn_head_bwd_apply_run(const void* raw, int64_t ld, const float* scale, const float* shift, int relu, const float* w,
                          const float* dz, const float* coef, void* draw, int64_t draw_ld, int64_t npix, int C,
                          cudaStream_t s) {
  UNETK_CHECK(bn_head_c_ok(C), -1, "bn_head: C=%d must be 32 or 64", C);
  UNETK_CUDA(launch_pdl(bn_head_bwd_apply_kernel, dim3(bn_head_grid(npix, C)), dim3(kThreads), 0, s,
                        static_cast<const __nv_bfloat16*>(raw), ld, scale, shift, relu, w, dz, coef,
                        static_cast<__nv_bfloat16*>(draw), draw_ld, npix, C));
  UNETK_LAUNCHED();
  return 0;
}

}  // namespace unetk
