// HBM-bound kernels of the U-Net step: BatchNorm statistics / finalize / apply(+ReLU)(+MaxPool),
// their backward (with the max-pool scatter and the skip-gradient add folded in), MaxPool2d(2) with
// PyTorch-exact argmax, per-channel column sums (bias gradients).
// All of them stream NHWC bf16 as 16-byte vectors (8 channels per thread), reduce with registers ->
// shared memory -> per-block partials, and finish with an ordered (deterministic) second stage.
// Reference semantics replaced: nn.BatchNorm2d + nn.ReLU(inplace) + nn.MaxPool2d(2) inside
// DoubleConv/Down (UNetFamily/utils/unet_parts.py:24-31,42-44) and their autograd backward.
#include <cstdlib>
#include "reduce2.cuh"
#include "fastdiv.cuh"
#include "host_common.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

namespace unetk {

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  return u;
}
__device__ __forceinline__ uint4 ldg16(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void stg16(__nv_bfloat16* p, const uint4& v) { *reinterpret_cast<uint4*>(p) = v; }
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// Thread layout shared by the per-channel reductions: cg = C/8 channel groups, ppb = kThreads/cg unit lanes.
struct Lanes {
  int cg, ppb, g, pl;
  bool active;
  __device__ Lanes(int C) {
    cg = C >> 3;
    ppb = kThreads / cg;
    if (ppb < 1) ppb = 1;
    g = threadIdx.x % cg;
    pl = threadIdx.x / cg;
    active = pl < ppb;
  }
};

// Sum K*8 per-thread accumulators over the unit lanes of the block and write partial[blockIdx][k][C].
template <int K>
__device__ __forceinline__ void block_reduce_store(float (&acc)[K][8], const Lanes& L, int C, float* partial) {
  extern __shared__ float red[];  // [ppb][K][C]
  if (L.active) {
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) red[(L.pl * K + k) * C + L.g * 8 + j] = acc[k][j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * C; i += kThreads) {
    float s = 0.f;
    for (int pl = 0; pl < L.ppb; ++pl) s += red[pl * K * C + i];
    partial[static_cast<size_t>(blockIdx.x) * K * C + i] = s;
  }
}

// ------------------------------------------------------------------ forward statistics
__global__ void __launch_bounds__(kThreads) stats_kernel(const __nv_bfloat16* __restrict__ x, int64_t ld,
                                                         int64_t npix, int C, float* __restrict__ partial) {
  pdl_trigger();
  pdl_wait();
  Lanes L(C);
  float acc[2][8] = {};
  if (L.active) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * L.ppb;
    int64_t pix = static_cast<int64_t>(blockIdx.x) * L.ppb + L.pl;
    for (; pix + 3 * stride < npix; pix += 4 * stride) {
      uint4 u[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) u[i] = ldg16(x + (pix + i * stride) * ld + L.g * 8);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float f[8];
        unpack8(u[i], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) { acc[0][j] += f[j]; acc[1][j] = fmaf(f[j], f[j], acc[1][j]); }
      }
    }
    for (; pix < npix; pix += stride) {
      float f[8];
      unpack8(ldg16(x + pix * ld + L.g * 8), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) { acc[0][j] += f[j]; acc[1][j] = fmaf(f[j], f[j], acc[1][j]); }
    }
  }
  block_reduce_store<2>(acc, L, C, partial);
}

// One thread per channel: ordered sum over the block partials in double, then the BN bookkeeping of
// nn.BatchNorm2d (biased variance to normalise, unbiased for running_var, momentum 0.1).
// Ordered (deterministic) second stage: sums[k][c] = sum over block partials, in double.
__global__ void chan_sums_kernel(const float* __restrict__ partial, int nblk, int C, double* __restrict__ sums) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * kSum2Lanes + threadIdx.x;
  const bool valid = i < 2 * C;
  const double s = sliced_ordered_sum(partial, nblk, valid, [&](int b) { return static_cast<size_t>(b) * 2 * C + i; });
  if (valid && threadIdx.y == 0) sums[i] = s;
}

__global__ void bn_finalize_kernel(const double* __restrict__ sums, int C, double count,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                   float momentum, float* running_mean, float* running_var,
                                   long long* num_batches_tracked, float* __restrict__ scale,
                                   float* __restrict__ shift, float* __restrict__ mean_out,
                                   float* __restrict__ invstd_out) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && num_batches_tracked != nullptr) *num_batches_tracked += 1;
  if (c >= C) return;
  const double s = sums[c], ss = sums[C + c];
  const double mean = s / count;
  double var = ss / count - mean * mean;
  if (var < 0.0) var = 0.0;
  const float invstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  scale[c] = g * invstd;
  shift[c] = b - static_cast<float>(mean) * g * invstd;
  mean_out[c] = static_cast<float>(mean);
  invstd_out[c] = invstd;
  if (running_mean != nullptr) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * static_cast<float>(mean);
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * static_cast<float>(unbiased);
  }
}

// Eval mode: fold running statistics into (scale, shift).
__global__ void bn_eval_fold_kernel(int C, const float* gamma, const float* beta, float eps, const float* rm,
                                    const float* rv, float* scale, float* shift, float* mean_out,
                                    float* invstd_out) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float invstd = 1.f / sqrtf(rv[c] + eps);
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  scale[c] = g * invstd;
  shift[c] = b - rm[c] * g * invstd;
  mean_out[c] = rm[c];
  invstd_out[c] = invstd;
}

// eval mode, for the conv epilogue fold: scale = gamma * invstd, shift = beta - mean * scale + conv_bias * scale
__global__ void bn_eval_fold_bias_kernel(int C, const float* gamma, const float* beta, float eps, const float* rm,
                                         const float* rv, const float* conv_bias, float* scale, float* shift) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float invstd = 1.f / sqrtf(rv[c] + eps);
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  const float sc = g * invstd;
  scale[c] = sc;
  shift[c] = fmaf(conv_bias ? conv_bias[c] - rm[c] : -rm[c], sc, b);
}

__global__ void colsum_finalize_kernel(const float* __restrict__ partial, int nblk, int C, float* __restrict__ out,
                                       int accumulate) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * kSum2Lanes + threadIdx.x;
  const bool valid = c < C;
  const double s = sliced_ordered_sum(partial, nblk, valid, [&](int b) { return (static_cast<size_t>(b) * 2 + 0) * C + c; });
  if (valid && threadIdx.y == 0) out[c] = accumulate ? out[c] + static_cast<float>(s) : static_cast<float>(s);
}

// ------------------------------------------------------------------ forward apply
// act = relu(bf16(raw*scale + shift)); optional fused 2x2 max-pool of act (first-max-wins like ATen).
__device__ __forceinline__ void bn_relu8(const uint4& raw, const float* sc, const float* sh, float* a, bool relu) {
  float f[8];
  unpack8(raw, f);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float z = bf16_round(fmaf(f[j], sc[j], sh[j]));
    a[j] = relu ? fmaxf(z, 0.f) : z;
  }
}

// Extra destinations of the activation (NestedUNet: a node is a member of up to four torch.cat's, UNetPP.py:80-97):
// written from the registers of the BatchNorm pass instead of re-reading the tensor once per copy.
struct Copies {
  __nv_bfloat16* p[3]; int64_t ld[3]; int n;
};

// Threads keep ONE channel group for the whole kernel and walk pixels with a constant pointer step (no division
// in the loop); four 16-byte loads per tensor per thread are in flight per iteration.
// RES: out = act(bn(raw)) + res  (the `x + x1` of Recurrent_block / RRCNN_block / ResidualConv,
// unet_parts.py:128,146,475: a bf16 add of two bf16 tensors under autocast).
template <bool RES, bool COPIES>
__global__ void __launch_bounds__(kThreads) bn_apply_kernel(const __nv_bfloat16* __restrict__ raw, int64_t raw_ld,
                                                            const float* __restrict__ scale,
                                                            const float* __restrict__ shift,
                                                            const __nv_bfloat16* __restrict__ res, int64_t res_ld,
                                                            __nv_bfloat16* __restrict__ out, int64_t out_ld,
                                                            int64_t npix, int C, int relu, int rev, const Copies X) {
  pdl_trigger();
  pdl_wait();
  Lanes L(C);
  if (!L.active) return;
  int64_t stride = static_cast<int64_t>(gridDim.x) * L.ppb;
  int64_t first = static_cast<int64_t>(blockIdx.x) * L.ppb + L.pl;
  if (first >= npix) return;
  const int g = L.g;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { sc[j] = __ldg(scale + g * 8 + j); sh[j] = __ldg(shift + g * 8 + j); }
  const bool r = relu != 0;
  int64_t left = (npix - first + stride - 1) / stride;
  // rev: walk from the END of the tensor — the tail of `raw` was written last by the convolution that precedes this
  // pass and is still in the 126 MB L2 (the head of a 0.5 GB tensor is not)
  if (rev) { first += (left - 1) * stride; stride = -stride; }
  const __nv_bfloat16* pr = raw + first * raw_ld + g * 8;
  const __nv_bfloat16* ps = RES ? res + first * res_ld + g * 8 : nullptr;
  __nv_bfloat16* po = out + first * out_ld + g * 8;
  const int64_t sr = stride * raw_ld, ss = stride * res_ld, so = stride * out_ld;
  auto one = [&](const uint4& u, const uint4& v, __nv_bfloat16* dst, int64_t pix) {
    float a[8];
    bn_relu8(u, sc, sh, a, r);
    if constexpr (RES) {
      float f[8];
      unpack8(v, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += f[j];
    }
    const uint4 o = pack8(a);
    stg16(dst, o);
    if constexpr (COPIES) {
#pragma unroll
      for (int e = 0; e < 3; ++e)
        if (e < X.n) stg16(X.p[e] + pix * X.ld[e] + g * 8, o);
    }
  };
  int64_t pix = first;
  for (; left >= 4; left -= 4) {
    uint4 u[4], v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      u[k] = ldg16(pr + k * sr);
      v[k] = RES ? ldg16(ps + k * ss) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) one(u[k], v[k], po + k * so, pix + k * stride);
    pr += 4 * sr; po += 4 * so; pix += 4 * stride;
    if constexpr (RES) ps += 4 * ss;
  }
  for (; left > 0; --left) {
    one(ldg16(pr), RES ? ldg16(ps) : make_uint4(0, 0, 0, 0), po, pix);
    pr += sr; po += so; pix += stride;
    if constexpr (RES) ps += ss;
  }
}

// Window order (0,0),(0,1),(1,0),(1,1); take `v > best || isnan(v)` -> first max wins, last NaN wins.
__device__ __forceinline__ void argmax4(const float (&a)[4][8], float* best, int* arg) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float b = a[0][j];
    int k = 0;
#pragma unroll
    for (int q = 1; q < 4; ++q) {
      const float v = a[q][j];
      if (v > b || v != v) { b = v; k = q; }
    }
    best[j] = b;
    arg[j] = k;
  }
}

template <bool COPIES>
__global__ void __launch_bounds__(kThreads)
bn_apply_pool_kernel(const __nv_bfloat16* __restrict__ raw, int64_t raw_ld, const float* __restrict__ scale,
                     const float* __restrict__ shift, __nv_bfloat16* __restrict__ out, int64_t out_ld,
                     __nv_bfloat16* __restrict__ pooled, int64_t pooled_ld, int N, int H, int W, int C,
                     int rev, const Copies X) {
  pdl_trigger();
  pdl_wait();
  const int cg = C >> 3, Ho = H >> 1, Wo = W >> 1;
  const int64_t total = static_cast<int64_t>(N) * Ho * Wo * cg;
  int64_t stride = static_cast<int64_t>(gridDim.x) * kThreads;  // multiple of cg
  int64_t i = blockIdx.x * static_cast<int64_t>(kThreads) + threadIdx.x;
  if (i >= total) return;
  const int g = static_cast<int>(i % cg);
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { sc[j] = __ldg(scale + g * 8 + j); sh[j] = __ldg(shift + g * 8 + j); }
  if (rev) { i += (total - 1 - i) / stride * stride; stride = -stride; }   // see bn_apply_kernel
  for (; i < total && i >= 0; i += stride) {
    int64_t t = i / cg;
    const int wo = static_cast<int>(t % Wo); t /= Wo;
    const int ho = static_cast<int>(t % Ho);
    const int n = static_cast<int>(t / Ho);
    const int64_t pix0 = (static_cast<int64_t>(n) * H + 2 * ho) * W + 2 * wo;
    uint4 u[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) u[q] = ldg16(raw + (pix0 + (q >> 1) * W + (q & 1)) * raw_ld + g * 8);
    float best[8];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float a[8];
      bn_relu8(u[q], sc, sh, a, true);
      const uint4 o = pack8(a);
      const int64_t pix = pix0 + (q >> 1) * W + (q & 1);
      stg16(out + pix * out_ld + g * 8, o);
      if constexpr (COPIES) {
#pragma unroll
        for (int e = 0; e < 3; ++e)
          if (e < X.n) stg16(X.p[e] + pix * X.ld[e] + g * 8, o);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (q == 0) best[j] = a[j];
        else if (a[j] > best[j] || a[j] != a[j]) best[j] = a[j];
      }
    }
    const int64_t opix = (static_cast<int64_t>(n) * Ho + ho) * Wo + wo;
    stg16(pooled + opix * pooled_ld + g * 8, pack8(best));
  }
}

// ------------------------------------------------------------------ MaxPool2d(2) standalone
// idx (optional): int64 NCHW-logical [N,C,Ho,Wo] holding h*W+w, exactly F.max_pool2d(return_indices=True).
__global__ void __launch_bounds__(kThreads)
maxpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, int64_t x_ld, __nv_bfloat16* __restrict__ y, int64_t y_ld,
                   long long* __restrict__ idx, int N, int H, int W, int C) {
  pdl_trigger();
  pdl_wait();
  const int cg = C >> 3, Ho = H >> 1, Wo = W >> 1;
  const int64_t total = static_cast<int64_t>(N) * Ho * Wo * cg;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(kThreads) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * kThreads) {
    const int g = static_cast<int>(i % cg);
    int64_t t = i / cg;
    const int wo = static_cast<int>(t % Wo); t /= Wo;
    const int ho = static_cast<int>(t % Ho);
    const int n = static_cast<int>(t / Ho);
    float a[4][8];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int64_t pix = (static_cast<int64_t>(n) * H + 2 * ho + (q >> 1)) * W + 2 * wo + (q & 1);
      unpack8(ldg16(x + pix * x_ld + g * 8), a[q]);
    }
    float best[8];
    int arg[8];
    argmax4(a, best, arg);
    const int64_t opix = (static_cast<int64_t>(n) * Ho + ho) * Wo + wo;
    stg16(y + opix * y_ld + g * 8, pack8(best));
    if (idx != nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = g * 8 + j;
        idx[((static_cast<int64_t>(n) * C + c) * Ho + ho) * Wo + wo] =
            static_cast<long long>(2 * ho + (arg[j] >> 1)) * W + 2 * wo + (arg[j] & 1);
      }
    }
  }
}

// dx[window pos] = (pos == argmax) ? dy : 0, argmax recomputed from x with the forward rule.
template <bool ACC>
__global__ void __launch_bounds__(kThreads)
maxpool_bwd_kernel(const __nv_bfloat16* __restrict__ x, int64_t x_ld, const __nv_bfloat16* __restrict__ dy,
                   int64_t dy_ld, __nv_bfloat16* __restrict__ dx, int64_t dx_ld, int N, int H, int W, int C) {
  pdl_trigger();
  pdl_wait();
  const int cg = C >> 3, Ho = H >> 1, Wo = W >> 1;
  const int64_t total = static_cast<int64_t>(N) * Ho * Wo * cg;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(kThreads) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * kThreads) {
    const int g = static_cast<int>(i % cg);
    int64_t t = i / cg;
    const int wo = static_cast<int>(t % Wo); t /= Wo;
    const int ho = static_cast<int>(t % Ho);
    const int n = static_cast<int>(t / Ho);
    float a[4][8];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int64_t pix = (static_cast<int64_t>(n) * H + 2 * ho + (q >> 1)) * W + 2 * wo + (q & 1);
      unpack8(ldg16(x + pix * x_ld + g * 8), a[q]);
    }
    float best[8], gy[8];
    int arg[8];
    argmax4(a, best, arg);
    const int64_t opix = (static_cast<int64_t>(n) * Ho + ho) * Wo + wo;
    unpack8(ldg16(dy + opix * dy_ld + g * 8), gy);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int64_t pix = (static_cast<int64_t>(n) * H + 2 * ho + (q >> 1)) * W + 2 * wo + (q & 1);
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (arg[j] == q) ? gy[j] : 0.f;
      if constexpr (ACC) {
        float old[8];
        unpack8(*reinterpret_cast<const uint4*>(dx + pix * dx_ld + g * 8), old);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += old[j];
      }
      stg16(dx + pix * dx_ld + g * 8, pack8(o));
    }
  }
  if constexpr (!ACC) {
    // odd H / W (MaxPool2d floors): the last row / column belongs to no window and gets a zero gradient
    const int64_t extra_w = (W & 1) ? static_cast<int64_t>(N) * H * cg : 0;             // pixels (n, h, W-1)
    const int64_t extra_h = (H & 1) ? static_cast<int64_t>(N) * (W - (W & 1)) * cg : 0;  // pixels (n, H-1, w < 2*Wo)
    for (int64_t i = blockIdx.x * static_cast<int64_t>(kThreads) + threadIdx.x; i < extra_w + extra_h;
         i += static_cast<int64_t>(gridDim.x) * kThreads) {
      int64_t pix;
      int g;
      if (i < extra_w) {
        g = static_cast<int>(i % cg);
        const int64_t t = i / cg;
        pix = (t / H * H + t % H) * W + (W - 1);
      } else {
        const int64_t j = i - extra_w;
        g = static_cast<int>(j % cg);
        const int64_t t = j / cg;
        const int Wv = W - (W & 1);
        pix = (t / Wv * H + (H - 1)) * W + t % Wv;
      }
      stg16(dx + pix * dx_ld + g * 8, make_uint4(0, 0, 0, 0));
    }
  }
}

// ------------------------------------------------------------------ backward of BN(+ReLU)(+pool,+skip add)
// Incoming gradient of the activation a = relu(bn(raw)):   g = g1 (same resolution, optional)
//                                                            + scatter(gp) through the 2x2 max-pool (optional).
// Unit of work: one 2x2 window x 8 channels when gp is given, else one pixel x 8 channels.
// Thread layout (Lanes): a thread keeps ONE channel group for the whole kernel and walks pixels with a fixed
// pointer step, so the loop body has no integer division (the first version spent ~10 of its ~25 instructions
// per element on 64-bit div/mod and was issue-bound at 39 % of HBM bandwidth, profiles/r01_ncu_bn_bwd_v1.txt).
struct BnBwdArgs {
  const __nv_bfloat16* raw; int64_t raw_ld;
  const __nv_bfloat16* g1; int64_t g1_ld;   // may be null
  const __nv_bfloat16* gp; int64_t gp_ld;   // may be null (pooled-resolution gradient)
  const float* scale; const float* shift; const float* mean; const float* invstd;
  int N, H, W, C, relu;
  FastDiv fd_wu, fd_hu;   // pooled variants: division by W/2 and H/2 without a hardware divide
  int rev;                // walk from the end of the tensors (their tails are the L2-resident part, see bn_apply_kernel)
};

// masked gradient of one pixel x 8 channels: gm = (relu && !(bf16(raw*sc+sh) > 0)) ? 0 : g
__device__ __forceinline__ void masked8(const uint4& ur, const uint4& ug, const float (&sc)[8], const float (&sh)[8],
                                        bool relu, float* r, float* gm) {
  unpack8(ur, r);
  unpack8(ug, gm);
  if (relu) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float z = bf16_round(fmaf(r[j], sc[j], sh[j]));
      gm[j] = (z > 0.f) ? gm[j] : 0.f;
    }
  }
}

// One 2x2 window x 8 channels: loads issued up front, argmax of the forward's pooled activation recomputed on the
// same bf16-rounded values (first max wins, NaN taken), per-pixel results streamed to emit(q, gm, r).
struct PoolWindow {
  uint4 ur[4], ug[4], ugp;
  __device__ __forceinline__ void load(const BnBwdArgs& A, int64_t pix0, int64_t opix, int g) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int64_t pix = pix0 + (q >> 1) * A.W + (q & 1);
      ur[q] = ldg16(A.raw + pix * A.raw_ld + g * 8);
      ug[q] = A.g1 ? ldg16(A.g1 + pix * A.g1_ld + g * 8) : make_uint4(0, 0, 0, 0);
    }
    ugp = ldg16(A.gp + opix * A.gp_ld + g * 8);
  }
  template <class Emit>
  __device__ __forceinline__ void visit(const BnBwdArgs& A, const float (&sc)[8], const float (&sh)[8],
                                        Emit&& emit) const {
    // Everything between the BatchNorm FMA and the per-channel sums runs on PACKED bf16 pairs (HMNMX2 / HSET2 /
    // HADD2 / LOP3): the first version evaluated mask, arg-max and gradient select per element in fp32 with bit
    // bookkeeping, ~23 instructions per element, which made the pooled passes issue-bound (55 % of the HBM peak
    // against 88 % for the un-pooled ones).
    // pass 1: activations of the four window pixels (bf16-rounded like the forward's) -> ReLU masks rm[q] and
    // "beats the running maximum" masks m[q] (strict >: the first maximum wins, as in ATen's max_pool2d)
    uint32_t rm[4][4], m[4][4], best[4];
    const __nv_bfloat162 zero2 = __float2bfloat162_rn(0.f);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float f[8];
      unpack8(ur[q], f);
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const __nv_bfloat162 z = __floats2bfloat162_rn(fmaf(f[2 * p], sc[2 * p], sh[2 * p]),
                                                       fmaf(f[2 * p + 1], sc[2 * p + 1], sh[2 * p + 1]));
        const __nv_bfloat162 v = __hmax2(z, zero2);   // the fused pool always sits behind a ReLU (checked on the host)
        rm[q][p] = __hgt2_mask(z, zero2);
        const uint32_t vb = *reinterpret_cast<const uint32_t*>(&v);
        if (q == 0) {
          best[p] = vb;
        } else {
          const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&best[p]);
          m[q][p] = __hgt2_mask(v, b);
          const __nv_bfloat162 nb = __hmax2(b, v);
          best[p] = *reinterpret_cast<const uint32_t*>(&nb);
        }
      }
    }
    // one-hot arg-max masks: the LAST q that beat the running maximum holds it
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const uint32_t m1 = m[1][p], m2 = m[2][p], m3 = m[3][p];
      m[2][p] = m2 & ~m3;
      m[1][p] = m1 & ~(m2 | m3);
      m[0][p] = ~(m1 | m2 | m3);
    }
    // pass 2: gm = relu-mask & (g1 + (arg-max ? gp : 0)); the sum of the two bf16 gradients is rounded to bf16, as
    // autograd's accumulation of the two consumers' gradients into one bf16 tensor is
    const uint32_t gy[4] = {ugp.x, ugp.y, ugp.z, ugp.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint32_t g[4] = {ug[q].x, ug[q].y, ug[q].z, ug[q].w};
      float r[8], gm[8];
      unpack8(ur[q], r);
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const uint32_t sel = gy[p] & m[q][p];
        const __nv_bfloat162 sum = __hadd2(*reinterpret_cast<const __nv_bfloat162*>(&g[p]),
                                           *reinterpret_cast<const __nv_bfloat162*>(&sel));
        const uint32_t sb = *reinterpret_cast<const uint32_t*>(&sum) & rm[q][p];
        gm[2 * p] = bf16_lo(sb);
        gm[2 * p + 1] = bf16_hi(sb);
      }
      emit(q, gm, r);
    }
  }
};

// Walks this thread's share of the tensor: pixels (POOL = false, kUnroll in flight) or 2x2 windows (POOL = true)
// of channel group L.g, calling emit(pixel index, gm[8], raw[8]).
template <bool POOL, class Emit>
__device__ __forceinline__ void bn_bwd_walk(const BnBwdArgs& A, const Lanes& L, const float (&sc)[8],
                                            const float (&sh)[8], Emit&& emit) {
  int64_t stride = static_cast<int64_t>(gridDim.x) * L.ppb;
  int64_t first = static_cast<int64_t>(blockIdx.x) * L.ppb + L.pl;
  const bool relu = A.relu != 0;
  if constexpr (!POOL) {
    const int64_t npix = static_cast<int64_t>(A.N) * A.H * A.W;
    if (first >= npix) return;
    int64_t left = (npix - first + stride - 1) / stride;   // pixels this thread visits (one division per thread)
    if (A.rev) { first += (left - 1) * stride; stride = -stride; }
    const __nv_bfloat16* pr = A.raw + first * A.raw_ld + L.g * 8;
    const __nv_bfloat16* pg = A.g1 + first * A.g1_ld + L.g * 8;
    const int64_t sr = stride * A.raw_ld, sg = stride * A.g1_ld;
    int64_t pix = first;
    constexpr int kUnroll = 4;
    for (; left >= kUnroll; left -= kUnroll) {
      uint4 ur[kUnroll], ug[kUnroll];
#pragma unroll
      for (int k = 0; k < kUnroll; ++k) { ur[k] = ldg16(pr + k * sr); ug[k] = ldg16(pg + k * sg); }
#pragma unroll
      for (int k = 0; k < kUnroll; ++k) {
        float r[8], gm[8];
        masked8(ur[k], ug[k], sc, sh, relu, r, gm);
        emit(pix + k * stride, gm, r);
      }
      pr += kUnroll * sr; pg += kUnroll * sg; pix += kUnroll * stride;
    }
    for (; left > 0; --left) {
      float r[8], gm[8];
      masked8(ldg16(pr), ldg16(pg), sc, sh, relu, r, gm);
      emit(pix, gm, r);
      pr += sr; pg += sg; pix += stride;
    }
  } else {
    const uint32_t Wu = static_cast<uint32_t>(A.W >> 1), Hu = static_cast<uint32_t>(A.H >> 1);
    const int64_t units = static_cast<int64_t>(A.N) * Hu * Wu;   // < 2^31 (checked on the host)
    auto window = [&](int64_t u) -> int64_t {   // first pixel of window u
      uint32_t wu, t, hu, n;
      A.fd_wu.divmod(static_cast<uint32_t>(u), t, wu);
      A.fd_hu.divmod(t, n, hu);
      return (static_cast<int64_t>(n) * A.H + 2 * hu) * A.W + 2 * wu;
    };
    // The nine sectors of this thread's NEXT window are pulled into L2 while the current window is evaluated (the
    // first version loaded, waited ~1 us for HBM and computed ~1 us, one window at a time: 45 % of the HBM peak on the
    // 64-channel layer; holding two windows in registers instead spills at the 128-register budget of 2 blocks/SM).
    if (first >= units) return;
    if (A.rev) { first += (units - 1 - first) / stride * stride; stride = -stride; }
    for (int64_t u = first; u < units && u >= 0; u += stride) {
      const int64_t pix0 = window(u);
      PoolWindow w;
      w.load(A, pix0, u, L.g);
      const int64_t un = u + stride;
      if (un < units && un >= 0) {
        const int64_t pn = window(un);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int64_t pix = pn + (q >> 1) * A.W + (q & 1);
          prefetch_l2(A.raw + pix * A.raw_ld + L.g * 8);
          if (A.g1) prefetch_l2(A.g1 + pix * A.g1_ld + L.g * 8);
        }
        prefetch_l2(A.gp + un * A.gp_ld + L.g * 8);
      }
      w.visit(A, sc, sh, [&](int q, const float* gm, const float* r) { emit(pix0 + (q >> 1) * A.W + (q & 1), gm, r); });
    }
  }
}

// partial[blk][0][c] = sum gm, partial[blk][1][c] = sum gm * (raw - mean)
template <bool POOL>
__global__ void __launch_bounds__(kThreads, 2) bn_bwd_reduce_kernel(const BnBwdArgs A, float* __restrict__ partial) {
  pdl_trigger();
  pdl_wait();
  Lanes L(A.C);
  float acc[2][8] = {};
  if (L.active) {
    float sc[8], sh[8], mu[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sc[j] = __ldg(A.scale + L.g * 8 + j); sh[j] = __ldg(A.shift + L.g * 8 + j); mu[j] = __ldg(A.mean + L.g * 8 + j);
    }
    bn_bwd_walk<POOL>(A, L, sc, sh, [&](int64_t, const float* gm, const float* r) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { acc[0][j] += gm[j]; acc[1][j] = fmaf(gm[j], r[j] - mu[j], acc[1][j]); }
    });
  }
  block_reduce_store<2>(acc, L, A.C, partial);
}

// sums[0][c] = S0 = sum gm, sums[1][c] = S1 = sum gm*(raw-mean).   dbeta = S0, dgamma = invstd * S1.
// The apply pass computes  draw = scale*(gm - dbeta/M - xhat*dgamma/M)  as  scale*gm + K1*raw + K0  with
//   K1 = -scale * invstd^2 * S1 / M,   K0 = -scale * S0 / M - K1 * mean          (coef[0][c] = K0, coef[1][c] = K1)
__global__ void bn_bwd_finalize_kernel(const double* __restrict__ sums, int C, double count,
                                       const float* __restrict__ scale, const float* __restrict__ mean,
                                       const float* __restrict__ invstd, float* dgamma, float* dbeta, int accumulate,
                                       float* __restrict__ coef, float* dconv_bias) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  // Bias of the convolution in FRONT of this BatchNorm (conv_block, up_conv, Recurrent_block, W_g/W_x/psi, ...):
  // its gradient is sum_pixels d(raw) = scale*S0 + K1*M*mean + K0*M, which is identically zero (batch statistics
  // remove any per-channel constant).  The reference's autograd sums the rounded d(raw) and gets rounding noise
  // around zero; we write the exact value instead of spending a full pass over d(raw) per layer.
  if (dconv_bias != nullptr && !accumulate) dconv_bias[c] = 0.f;
  const double s0 = sums[c], s1 = sums[C + c];
  const double is = invstd[c], sc = scale[c], mu = mean[c];
  const float db = static_cast<float>(s0), dg = static_cast<float>(is * s1);
  if (dbeta) dbeta[c] = accumulate ? dbeta[c] + db : db;
  if (dgamma) dgamma[c] = accumulate ? dgamma[c] + dg : dg;
  const double k1 = -sc * is * is * s1 / count;
  coef[c] = static_cast<float>(-sc * s0 / count - k1 * mu);
  coef[C + c] = static_cast<float>(k1);
}

// ACC: draw += (bf16 read-modify-write) instead of draw = ; used when `draw` is the gradient of a tensor that has
// other consumers (pre-activation BatchNorm of ResidualConv, unet_parts.py:458-459).
// RESG: the unit's output was relu(bn(raw)) + res (Recurrent_block's x + x1, RRCNN_block's residual, unet_parts.py:125-146):
// d(res) = the incoming gradient g1 itself, written (1) or accumulated (2) into `dres` from this pass instead of a separate
// add pass that reads g1 again (R2UNet: 45 such passes, 2.2 ms per step).
template <bool POOL, bool ACC, int RESG>
__global__ void __launch_bounds__(kThreads, 2)
bn_bwd_apply_kernel(const BnBwdArgs A, const float* __restrict__ coef, __nv_bfloat16* __restrict__ draw,
                    int64_t draw_ld, __nv_bfloat16* __restrict__ dres, int64_t dres_ld) {
  pdl_trigger();
  pdl_wait();
  Lanes L(A.C);
  if (!L.active) return;
  const int g = L.g;
  float sc[8], sh[8], k0[8], k1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = __ldg(A.scale + g * 8 + j); sh[j] = __ldg(A.shift + g * 8 + j);
    k0[j] = __ldg(coef + g * 8 + j); k1[j] = __ldg(coef + A.C + g * 8 + j);
  }
  bn_bwd_walk<POOL>(A, L, sc, sh, [&](int64_t pix, const float* gm, const float* r) {
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = fmaf(sc[j], gm[j], fmaf(k1[j], r[j], k0[j]));
    __nv_bfloat16* dst = draw + pix * draw_ld + g * 8;
    if constexpr (ACC) {
      float old[8];
      unpack8(*reinterpret_cast<const uint4*>(dst), old);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = old[j] + bf16_round(o[j]);
    }
    stg16(dst, pack8(o));
    if constexpr (RESG != 0) {
      const uint4 ug = ldg16(A.g1 + pix * A.g1_ld + g * 8);   // loaded by the walk a moment ago: an L1 hit
      __nv_bfloat16* rd = dres + pix * dres_ld + g * 8;
      if constexpr (RESG == 2) {
        float a[8], b[8];
        unpack8(*reinterpret_cast<const uint4*>(rd), a);
        unpack8(ug, b);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] += b[j];
        stg16(rd, pack8(a));
      } else {
        stg16(rd, ug);
      }
    }
  });
}

// grid for the flat (unit x channel-group) kernels whose threads keep their channel group: blocks*kThreads % cg == 0
int flat_grid_cg(int64_t total, int cg) {
  int64_t b = (total + kThreads - 1) / kThreads;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  // kThreads = 256 = 2^8; make b*256 a multiple of cg by rounding b up to a multiple of cg / gcd(cg, 256)
  int g = cg, a = kThreads;
  while (a) { int t = g % a; g = a; a = t; }
  const int m = cg / g;
  b = (b + m - 1) / m * m;
  return static_cast<int>(b);
}

int reduce_grid(int64_t units, int C) {
  const int cg = C / 8;
  int ppb = kThreads / cg;
  if (ppb < 1) ppb = 1;
  int64_t want = (units + ppb * 8 - 1) / (ppb * 8);  // >= 8 units per lane
  int64_t cap = static_cast<int64_t>(num_sms()) * 4;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return static_cast<int>(want);
}
// BN backward kernels: 2 resident blocks per SM (register budget), exactly one wave
int bwd_grid(int64_t units, int C) {
  const int g = reduce_grid(units, C), cap = num_sms() * 2;
  return g < cap ? g : cap;
}
size_t reduce_smem(int C, int K) {
  const int cg = C / 8;
  int ppb = kThreads / cg;
  if (ppb < 1) ppb = 1;
  return static_cast<size_t>(ppb) * K * C * sizeof(float);
}
int flat_grid(int64_t total) {
  int64_t b = (total + kThreads - 1) / kThreads;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

}  // namespace

// ---------------------------------------------------------------------------------------------- host side
// UNETK_REVERSE=0: every pass walks its tensors front to back (A/B of the L2-tail reuse)
static int reverse_walk() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("UNETK_REVERSE"); v = e ? atoi(e) : 1; }
  return v;
}

#define CHECK_C(C) UNETK_CHECK((C) % 8 == 0 && (C) >= 8 && (C) <= 2048, -1, "channel count %d must be a multiple of 8 in [8,2048]", (C))

size_t chan_partial_floats(int64_t units, int C) { return static_cast<size_t>(reduce_grid(units, C)) * 2 * C; }

static int launch_sums(const float* partial, int nblk, int C, double* sums, cudaStream_t s) {
  UNETK_CUDA(launch_pdl(chan_sums_kernel, dim3((2 * C + kSum2Lanes - 1) / kSum2Lanes), dim3(dim3(kSum2Lanes, kSum2Slices)), 0, s, partial, nblk, C, sums));
  UNETK_LAUNCHED();
  return 0;
}

// sums (double [2][C]): per-channel sum and sum of squares over npix pixels
int bn_stats_run(const void* x, int64_t ld, int64_t npix, int C, float* partial, double* sums, cudaStream_t s) {
  CHECK_C(C);
  const int grid = reduce_grid(npix, C);
  UNETK_CUDA(launch_pdl(stats_kernel, dim3(grid), dim3(kThreads), reduce_smem(C, 2), s, static_cast<const __nv_bfloat16*>(x), ld, npix, C, partial));
  UNETK_LAUNCHED();
  return launch_sums(partial, grid, C, sums, s);
}

int bn_finalize_run(const double* sums, int C, double count, const float* gamma, const float* beta, float eps,
                    float momentum, float* rm, float* rv, long long* nbt, float* scale, float* shift, float* mean,
                    float* invstd, cudaStream_t s) {
  UNETK_CUDA(launch_pdl(bn_finalize_kernel, dim3((C + 127) / 128), dim3(128), 0, s, sums, C, count, gamma, beta, eps, momentum, rm, rv, nbt, scale,
                                                    shift, mean, invstd));
  UNETK_LAUNCHED();
  return 0;
}

int bn_eval_fold_run(int C, const float* gamma, const float* beta, float eps, const float* rm, const float* rv,
                     float* scale, float* shift, float* mean, float* invstd, cudaStream_t s) {
  UNETK_CUDA(launch_pdl(bn_eval_fold_kernel, dim3((C + 127) / 128), dim3(128), 0, s, C, gamma, beta, eps, rm, rv, scale, shift, mean, invstd));
  UNETK_LAUNCHED();
  return 0;
}

int bn_eval_fold_bias_run(int C, const float* gamma, const float* beta, float eps, const float* rm, const float* rv,
                          const float* conv_bias, float* scale, float* shift, cudaStream_t s) {
  UNETK_CUDA(launch_pdl(bn_eval_fold_bias_kernel, dim3((C + 127) / 128), dim3(128), 0, s, C, gamma, beta, eps, rm, rv, conv_bias, scale, shift));
  UNETK_LAUNCHED();
  return 0;
}

int colsum_run(const void* x, int64_t ld, int64_t npix, int C, float* partial, float* out, int accumulate,
               cudaStream_t s) {
  CHECK_C(C);
  const int grid = reduce_grid(npix, C);
  UNETK_CUDA(launch_pdl(stats_kernel, dim3(grid), dim3(kThreads), reduce_smem(C, 2), s, static_cast<const __nv_bfloat16*>(x), ld, npix, C, partial));
  UNETK_LAUNCHED();
  UNETK_CUDA(launch_pdl(colsum_finalize_kernel, dim3((C + kSum2Lanes - 1) / kSum2Lanes), dim3(dim3(kSum2Lanes, kSum2Slices)), 0, s, partial, grid, C, out, accumulate));
  UNETK_LAUNCHED();
  return 0;
}

__global__ void sums_to_f32_kernel(const double* __restrict__ sums, int n, float* __restrict__ out, int accumulate) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = accumulate ? out[i] + static_cast<float>(sums[i]) : static_cast<float>(sums[i]);
}
int sums_to_f32_run(const double* sums, int n, float* out, int accumulate, cudaStream_t s) {
  UNETK_CUDA(launch_pdl(sums_to_f32_kernel, dim3((n + 127) / 128), dim3(128), 0, s, sums, n, out, accumulate));
  UNETK_LAUNCHED();
  return 0;
}

// grid for the Lanes-layout streaming kernels: one resident wave, >= 4 units per lane
static int lanes_grid(int64_t units, int C, int blocks_per_sm) {
  int ppb = kThreads / (C / 8);
  if (ppb < 1) ppb = 1;
  int64_t want = (units + ppb * 4 - 1) / (ppb * 4);
  const int64_t cap = static_cast<int64_t>(num_sms()) * blocks_per_sm;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return static_cast<int>(want);
}

int bn_apply_run(const void* raw, int64_t raw_ld, const float* scale, const float* shift, const void* res,
                 int64_t res_ld, void* out, int64_t out_ld, void* pooled, int64_t pooled_ld, int N, int H, int W, int C,
                 int relu, cudaStream_t s, int ncopies, void* const* copies, const int64_t* copies_ld) {
  CHECK_C(C);
  UNETK_CHECK(ncopies >= 0 && ncopies <= 3, -1, "bn_apply: at most 3 extra destinations (got %d)", ncopies);
  Copies X{};
  X.n = ncopies;
  for (int e = 0; e < ncopies; ++e) {
    UNETK_CHECK(copies[e] != nullptr && copies_ld[e] >= C && copies_ld[e] % 8 == 0, -1, "bn_apply: bad extra destination %d", e);
    X.p[e] = static_cast<__nv_bfloat16*>(copies[e]); X.ld[e] = copies_ld[e];
  }
  const __nv_bfloat16* rw = static_cast<const __nv_bfloat16*>(raw);
  __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out);
  if (pooled != nullptr) {
    UNETK_CHECK(H % 2 == 0 && W % 2 == 0 && relu, -1, "fused pool needs even H,W and relu");
    UNETK_CHECK(res == nullptr, -1, "bn_apply: fused pool and residual add cannot be combined");
    const int64_t total = static_cast<int64_t>(N) * (H / 2) * (W / 2) * (C / 8);
    const int grid = flat_grid_cg(total, C / 8);
    __nv_bfloat16* pl = static_cast<__nv_bfloat16*>(pooled);
    if (ncopies > 0)
      UNETK_CUDA(launch_pdl(bn_apply_pool_kernel<true>, dim3(grid), dim3(kThreads), 0, s, rw, raw_ld, scale, shift, o, out_ld, pl,
                            pooled_ld, N, H, W, C, reverse_walk(), X));
    else
      UNETK_CUDA(launch_pdl(bn_apply_pool_kernel<false>, dim3(grid), dim3(kThreads), 0, s, rw, raw_ld, scale, shift, o, out_ld, pl,
                            pooled_ld, N, H, W, C, reverse_walk(), X));
  } else {
    const int64_t npix = static_cast<int64_t>(N) * H * W;
    const int grid = lanes_grid(npix, C, 8);
    const __nv_bfloat16* rs = static_cast<const __nv_bfloat16*>(res);
    UNETK_CHECK(ncopies == 0 || res == nullptr, -1, "bn_apply: extra destinations and a residual input cannot be combined");
    if (res != nullptr)
      UNETK_CUDA(launch_pdl(bn_apply_kernel<true, false>, dim3(grid), dim3(kThreads), 0, s, rw, raw_ld, scale, shift, rs, res_ld, o,
                            out_ld, npix, C, relu, reverse_walk(), X));
    else if (ncopies > 0)
      UNETK_CUDA(launch_pdl(bn_apply_kernel<false, true>, dim3(grid), dim3(kThreads), 0, s, rw, raw_ld, scale, shift, nullptr, 0, o,
                            out_ld, npix, C, relu, reverse_walk(), X));
    else
      UNETK_CUDA(launch_pdl(bn_apply_kernel<false, false>, dim3(grid), dim3(kThreads), 0, s, rw, raw_ld, scale, shift, nullptr, 0, o,
                            out_ld, npix, C, relu, reverse_walk(), X));
  }
  UNETK_LAUNCHED();
  return 0;
}

int maxpool_fwd_run(const void* x, int64_t x_ld, void* y, int64_t y_ld, long long* idx, int N, int H, int W, int C,
                    cudaStream_t s) {
  CHECK_C(C);
  const int64_t total = static_cast<int64_t>(N) * (H / 2) * (W / 2) * (C / 8);
  if (total == 0) return 0;
  UNETK_CUDA(launch_pdl(maxpool_fwd_kernel, dim3(flat_grid(total)), dim3(kThreads), 0, s, static_cast<const __nv_bfloat16*>(x), x_ld,
                                                          static_cast<__nv_bfloat16*>(y), y_ld, idx, N, H, W, C));
  UNETK_LAUNCHED();
  return 0;
}

int maxpool_bwd_run(const void* x, int64_t x_ld, const void* dy, int64_t dy_ld, void* dx, int64_t dx_ld,
                    int accumulate, int N, int H, int W, int C, cudaStream_t s) {
  CHECK_C(C);
  const int64_t total = static_cast<int64_t>(N) * (H / 2) * (W / 2) * (C / 8) + 1;   // +1: the odd-edge loop runs even without windows
  if (accumulate)
    UNETK_CUDA(launch_pdl(maxpool_bwd_kernel<true>, dim3(flat_grid(total)), dim3(kThreads), 0, s, static_cast<const __nv_bfloat16*>(x), x_ld,
                                                                  static_cast<const __nv_bfloat16*>(dy), dy_ld,
                                                                  static_cast<__nv_bfloat16*>(dx), dx_ld, N, H, W, C));
  else
    UNETK_CUDA(launch_pdl(maxpool_bwd_kernel<false>, dim3(flat_grid(total)), dim3(kThreads), 0, s, static_cast<const __nv_bfloat16*>(x), x_ld,
                                                                   static_cast<const __nv_bfloat16*>(dy), dy_ld,
                                                                   static_cast<__nv_bfloat16*>(dx), dx_ld, N, H, W, C));
  UNETK_LAUNCHED();
  return 0;
}

static int bn_bwd_args(BnBwdArgs* A, const void* raw, int64_t raw_ld, const void* g1, int64_t g1_ld, const void* gp,
                       int64_t gp_ld, const float* scale, const float* shift, const float* mean, const float* invstd,
                       int N, int H, int W, int C, int relu) {
  CHECK_C(C);
  UNETK_CHECK(g1 != nullptr || gp != nullptr, -1, "bn_bwd: no incoming gradient");
  if (gp != nullptr) {
    UNETK_CHECK(H % 2 == 0 && W % 2 == 0, -1, "bn_bwd: pooled gradient needs even H,W");
    UNETK_CHECK(relu, -1, "bn_bwd: the fused max-pool sits behind a ReLU (unet_parts.py:28-29,43)");
    UNETK_CHECK(static_cast<int64_t>(N) * (H / 2) * (W / 2) < (1ll << 31), -1, "bn_bwd: too many pooling windows");
  }
  *A = BnBwdArgs{static_cast<const __nv_bfloat16*>(raw), raw_ld, static_cast<const __nv_bfloat16*>(g1), g1_ld,
                 static_cast<const __nv_bfloat16*>(gp), gp_ld, scale, shift, mean, invstd, N, H, W, C, relu,
                 FastDiv(static_cast<uint32_t>(W / 2 > 0 ? W / 2 : 1)), FastDiv(static_cast<uint32_t>(H / 2 > 0 ? H / 2 : 1)), 0};
  return 0;
}

// sums (double [2][C]): sum of masked gradient, sum of masked gradient * xhat
int bn_bwd_reduce_run(const void* raw, int64_t raw_ld, const void* g1, int64_t g1_ld, const void* gp, int64_t gp_ld,
                      const float* scale, const float* shift, const float* mean, const float* invstd, float* partial,
                      double* sums, int N, int H, int W, int C, int relu, cudaStream_t s) {
  BnBwdArgs A;
  if (int rc = bn_bwd_args(&A, raw, raw_ld, g1, g1_ld, gp, gp_ld, scale, shift, mean, invstd, N, H, W, C, relu)) return rc;
  const bool pool = gp != nullptr;
  // back to front: the tail of g1 was written last by the dgrad that produced it; the apply pass then runs front to
  // back and starts on what this pass touched last
  A.rev = reverse_walk();
  const int64_t units = pool ? static_cast<int64_t>(N) * (H / 2) * (W / 2) : static_cast<int64_t>(N) * H * W;
  const int grid = bwd_grid(units, C);
  const size_t smem = reduce_smem(C, 2);
  if (pool) UNETK_CUDA(launch_pdl(bn_bwd_reduce_kernel<true>, dim3(grid), dim3(kThreads), smem, s, A, partial));
  else UNETK_CUDA(launch_pdl(bn_bwd_reduce_kernel<false>, dim3(grid), dim3(kThreads), smem, s, A, partial));
  UNETK_LAUNCHED();
  return launch_sums(partial, grid, C, sums, s);
}

int bn_bwd_coef_run(const double* sums, int C, double count, const float* scale, const float* mean,
                    const float* invstd, float* dgamma, float* dbeta, int accumulate, float* coef, float* dconv_bias,
                    cudaStream_t s) {
  UNETK_CHECK(C >= 1, -1, "bn_bwd_coef: C=%d", C);
  UNETK_CUDA(launch_pdl(bn_bwd_finalize_kernel, dim3((C + 127) / 128), dim3(128), 0, s, sums, C, count, scale, mean, invstd, dgamma, dbeta, accumulate, coef,
                                                         dconv_bias));
  UNETK_LAUNCHED();
  return 0;
}

int bn_bwd_apply_run(const void* raw, int64_t raw_ld, const void* g1, int64_t g1_ld, const void* gp, int64_t gp_ld,
                     const float* scale, const float* shift, const float* mean, const float* invstd,
                     const double* sums, double count, float* dgamma, float* dbeta, int accumulate, float* coef,
                     float* dconv_bias, void* draw, int64_t draw_ld, int draw_accumulate, int N, int H, int W, int C,
                     int relu, cudaStream_t s, void* dres, int64_t dres_ld, int dres_accumulate) {
  UNETK_CHECK(dres == nullptr || (gp == nullptr && g1 != nullptr && dres_ld % 8 == 0), -1,
              "bn_bwd_apply: the residual gradient needs g1, no fused pool and a pixel stride that is a multiple of 8");
  BnBwdArgs A;
  if (int rc = bn_bwd_args(&A, raw, raw_ld, g1, g1_ld, gp, gp_ld, scale, shift, mean, invstd, N, H, W, C, relu)) return rc;
  const bool pool = gp != nullptr;
  const int64_t units = pool ? static_cast<int64_t>(N) * (H / 2) * (W / 2) : static_cast<int64_t>(N) * H * W;
  UNETK_CUDA(launch_pdl(bn_bwd_finalize_kernel, dim3((C + 127) / 128), dim3(128), 0, s, sums, C, count, scale, mean, invstd, dgamma, dbeta, accumulate, coef,
                                                         dconv_bias));
  UNETK_LAUNCHED();
  const int grid = bwd_grid(units, C);
  __nv_bfloat16* d = static_cast<__nv_bfloat16*>(draw);
  __nv_bfloat16* rg = static_cast<__nv_bfloat16*>(dres);
  const int64_t z = 0;
  if (pool) {
    if (draw_accumulate) UNETK_CUDA(launch_pdl(bn_bwd_apply_kernel<true, true, 0>, dim3(grid), dim3(kThreads), 0, s, A, coef, d, draw_ld, rg, z));
    else UNETK_CUDA(launch_pdl(bn_bwd_apply_kernel<true, false, 0>, dim3(grid), dim3(kThreads), 0, s, A, coef, d, draw_ld, rg, z));
  } else if (dres == nullptr) {
    if (draw_accumulate) UNETK_CUDA(launch_pdl(bn_bwd_apply_kernel<false, true, 0>, dim3(grid), dim3(kThreads), 0, s, A, coef, d, draw_ld, rg, z));
    else UNETK_CUDA(launch_pdl(bn_bwd_apply_kernel<false, false, 0>, dim3(grid), dim3(kThreads), 0, s, A, coef, d, draw_ld, rg, z));
  } else if (dres_accumulate) {
    if (draw_accumulate) UNETK_CUDA(launch_pdl(bn_bwd_apply_kernel<false, true, 2>, dim3(grid), dim3(kThreads), 0, s, A, coef, d, draw_ld, rg, dres_ld));
    else UNETK_CUDA(launch_pdl(bn_bwd_apply_kernel<false, false, 2>, dim3(grid), dim3(kThreads), 0, s, A, coef, d, draw_ld, rg, dres_ld));
  } else {
    if (draw_accumulate) UNETK_CUDA(launch_pdl(bn_bwd_apply_kernel<false, true, 1>, dim3(grid), dim3(kThreads), 0, s, A, coef, d, draw_ld, rg, dres_ld));
    else UNETK_CUDA(launch_pdl(bn_bwd_apply_kernel<false, false, 1>, dim3(grid), dim3(kThreads), 0, s, A, coef, d, draw_ld, rg, dres_ld));
  }
  UNETK_LAUNCHED();
  return 0;
}

}  // namespace unetk
