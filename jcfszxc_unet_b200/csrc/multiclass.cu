// n_classes > 1: the OutConv head with K output channels (UNet(n_channels, n_classes > 1), unet_parts.py:73-79) and
// the Dice coefficient as a stand-alone op (utils/dice_score.py:13-59: dice_coeff, multiclass_dice_coeff = dice_coeff
// of the (batch x class)-flattened tensors, dice_loss).  The fused head + BCE + dice kernel of loss.cu covers the
// n_classes == 1 training step; this file is the general forward / backward the nn.Module surface needs (the loss is
// then computed by the caller, e.g. nn.CrossEntropyLoss as at train.py:124, and autograd hands dL/dlogits back).
#include "host_common.cuh"
#include "kernels.cuh"
#include "ptx.cuh"
#include "reduce2.cuh"

namespace unetk {

namespace {

constexpr int kThreads = 256;
constexpr int kMaxK = 8;

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ float group_sum(float v, int lpp) {
  for (int o = lpp >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// logits[n][k][h][w] = bias[k] + sum_c x[pix][c] * w[k][c]     (lpp = C/8 lanes per pixel, K <= 8 classes)
template <int K>
__global__ void __launch_bounds__(kThreads)
head_multi_fwd_kernel(const __nv_bfloat16* __restrict__ x, int64_t ld, const float* __restrict__ w,
                      const float* __restrict__ bias, float* __restrict__ logits, int64_t npix, int64_t hw, int C) {
  const int lpp = C >> 3, gpb = kThreads / lpp;
  const int sub = threadIdx.x % lpp, grp = threadIdx.x / lpp;
  float wv[K][8];
#pragma unroll
  for (int k = 0; k < K; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) wv[k][j] = __ldg(w + k * C + sub * 8 + j);
  const int64_t step = static_cast<int64_t>(gridDim.x) * gpb;
  for (int64_t base = static_cast<int64_t>(blockIdx.x) * gpb; base < npix; base += step) {   // uniform trip count
    const int64_t pix = base + grp;
    const bool ok = pix < npix;
    float f[8];
    unpack8(ok ? __ldg(reinterpret_cast<const uint4*>(x + pix * ld + sub * 8)) : make_uint4(0, 0, 0, 0), f);
    const int64_t n = ok ? pix / hw : 0, r = ok ? pix - n * hw : 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      float dot = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) dot = fmaf(f[j], wv[k][j], dot);
      dot = group_sum(dot, lpp);
      if (ok && sub == 0) logits[(n * K + k) * hw + r] = dot + (bias ? __ldg(bias + k) : 0.f);
    }
  }
}

// dx[pix][c] = sum_k dl[n][k][hw] * w[k][c];  partial[blk][k][c] = sum_pix dl * x;  partial[blk][K*C + k] = sum_pix dl
template <int K>
__global__ void __launch_bounds__(kThreads)
head_multi_bwd_kernel(const __nv_bfloat16* __restrict__ x, int64_t ld, const float* __restrict__ w,
                      const float* __restrict__ dlogits, float gscale, __nv_bfloat16* __restrict__ dx, int64_t dx_ld,
                      int64_t npix, int64_t hw, int C, float* __restrict__ partial) {
  const int lpp = C >> 3, gpb = kThreads / lpp;
  const int sub = threadIdx.x % lpp, grp = threadIdx.x / lpp;
  float wv[K][8], acc[K][8], sdl[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    sdl[k] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { wv[k][j] = __ldg(w + k * C + sub * 8 + j); acc[k][j] = 0.f; }
  }
  const int64_t step = static_cast<int64_t>(gridDim.x) * gpb;
  for (int64_t pix = static_cast<int64_t>(blockIdx.x) * gpb + grp; pix < npix; pix += step) {
    float f[8], o[8] = {};
    unpack8(__ldg(reinterpret_cast<const uint4*>(x + pix * ld + sub * 8)), f);
    const int64_t n = pix / hw, r = pix - n * hw;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const float dl = gscale * __ldg(dlogits + (n * K + k) * hw + r);
#pragma unroll
      for (int j = 0; j < 8; ++j) { acc[k][j] = fmaf(dl, f[j], acc[k][j]); o[j] = fmaf(dl, wv[k][j], o[j]); }
      if (sub == 0) sdl[k] += dl;
    }
    uint4 w4;
    w4.x = pack_bf16x2(o[0], o[1]); w4.y = pack_bf16x2(o[2], o[3]);
    w4.z = pack_bf16x2(o[4], o[5]); w4.w = pack_bf16x2(o[6], o[7]);
    *reinterpret_cast<uint4*>(dx + pix * dx_ld + sub * 8) = w4;
  }
  extern __shared__ float red[];  // [gpb][K*C + K]
  const int row = K * C + K;
#pragma unroll
  for (int k = 0; k < K; ++k) {
#pragma unroll
    for (int j = 0; j < 8; ++j) red[grp * row + k * C + sub * 8 + j] = acc[k][j];
    if (sub == 0) red[grp * row + K * C + k] = sdl[k];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < row; i += kThreads) {
    float s = 0.f;
    for (int g = 0; g < gpb; ++g) s += red[g * row + i];
    partial[static_cast<size_t>(blockIdx.x) * row + i] = s;
  }
}

// dw[k][c] (+)= sum_blk partial[blk][k*C + c];  db[k] (+)= sum_blk partial[blk][K*C + k]   (ordered, deterministic)
__global__ void head_multi_finalize_kernel(const float* __restrict__ partial, int nblk, int K, int C, float* dw,
                                           float* db, int accumulate) {
  const int row = K * C + K;
  const int i = blockIdx.x * kSum2Lanes + threadIdx.x;
  const bool valid = i < row;
  const double s = sliced_ordered_sum(partial, nblk, valid, [&](int b) { return static_cast<size_t>(b) * row + i; });
  if (valid && threadIdx.y == 0) {
    float* o = (i < K * C) ? (dw ? dw + i : nullptr) : (db ? db + (i - K * C) : nullptr);
    if (o) *o = accumulate ? *o + static_cast<float>(s) : static_cast<float>(s);
  }
}

int mh_grid(int64_t npix, int C) {
  const int gpb = kThreads / (C / 8);
  int64_t b = (npix + gpb * 4 - 1) / (gpb * 4);
  const int64_t cap = static_cast<int64_t>(num_sms()) * 4;
  if (b > cap) b = cap;
  return static_cast<int>(b < 1 ? 1 : b);
}
bool mh_c_ok(int C) { return C == 8 || C == 16 || C == 32 || C == 64 || C == 128 || C == 256; }

// ---------------------------------------------------------------------------------------------- Dice
// group g = n elements; partial[g][blk][3] = (sum clamp(p)*t, sum clamp(p), sum t)
__global__ void __launch_bounds__(kThreads)
dice_sums_kernel(const float* __restrict__ p, const float* __restrict__ t, int64_t n, float lo, float hi,
                 float* __restrict__ partial) {
  const int64_t g = blockIdx.y;
  const float* pp = p + g * n;
  const float* tt = t + g * n;
  float s_pt = 0.f, s_p = 0.f, s_t = 0.f;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * kThreads) {
    const float a = fminf(fmaxf(__ldg(pp + i), lo), hi), b = __ldg(tt + i);
    s_pt = fmaf(a, b, s_pt); s_p += a; s_t += b;
  }
  __shared__ float red[3][kThreads / 32];
  s_pt = warp_sum(s_pt); s_p = warp_sum(s_p); s_t = warp_sum(s_t);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = s_pt; red[1][warp] = s_p; red[2][warp] = s_t; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float s = 0.f;
    for (int i = 0; i < kThreads / 32; ++i) s += red[threadIdx.x][i];
    partial[(g * gridDim.x + blockIdx.x) * 3 + threadIdx.x] = s;
  }
}
__global__ void dice_sums_final_kernel(const float* __restrict__ partial, int nblk, int64_t groups, double* __restrict__ sums) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= groups * 3) return;
  const int64_t g = i / 3;
  const int k = static_cast<int>(i % 3);
  double s = 0.0;
  for (int b = 0; b < nblk; ++b) s += partial[(g * nblk + b) * 3 + k];
  sums[i] = s;
}
// dp[i] = gout * (coef[g][0] * t[i] + coef[g][1]) where the clamp passes the gradient (lo <= p <= hi), else 0
__global__ void __launch_bounds__(kThreads)
dice_bwd_kernel(const float* __restrict__ p, const float* __restrict__ t, const float* __restrict__ coef,
                const float* __restrict__ gout, int64_t n, float lo, float hi, float* __restrict__ dp) {
  const int64_t g = blockIdx.y;
  const float ca = __ldg(coef + 2 * g) * __ldg(gout), cb = __ldg(coef + 2 * g + 1) * __ldg(gout);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * kThreads) {
    const float a = __ldg(p + g * n + i);
    dp[g * n + i] = (a >= lo && a <= hi) ? fmaf(ca, __ldg(t + g * n + i), cb) : 0.f;
  }
}
int dice_blocks(int64_t groups, int64_t n) {
  int64_t b = (n + kThreads * 8 - 1) / (kThreads * 8);
  int64_t cap = static_cast<int64_t>(num_sms()) * 8 / (groups > 0 ? groups : 1);
  if (cap < 1) cap = 1;
  if (b > cap) b = cap;
  return static_cast<int>(b < 1 ? 1 : b);
}

}  // namespace

size_t head_multi_partial_floats(int64_t npix, int C, int K) { return static_cast<size_t>(mh_grid(npix, C)) * (K * C + K); }

int head_multi_fwd_run(const void* x, int64_t ld, const float* w, const float* bias, float* logits, int N, int64_t hw,
                       int C, int K, cudaStream_t s) {
  UNETK_CHECK(mh_c_ok(C), -1, "head_multi: C=%d must be a power of two in [8,256]", C);
  UNETK_CHECK(K >= 1 && K <= kMaxK, -1, "head_multi: n_classes=%d (1..8)", K);
  const int64_t npix = static_cast<int64_t>(N) * hw;
  const int grid = mh_grid(npix, C);
  const __nv_bfloat16* xb = static_cast<const __nv_bfloat16*>(x);
#define UNETK_MH_FWD(KK) case KK: head_multi_fwd_kernel<KK><<<grid, kThreads, 0, s>>>(xb, ld, w, bias, logits, npix, hw, C); break;
  switch (K) { UNETK_MH_FWD(1) UNETK_MH_FWD(2) UNETK_MH_FWD(3) UNETK_MH_FWD(4) UNETK_MH_FWD(5) UNETK_MH_FWD(6) UNETK_MH_FWD(7) UNETK_MH_FWD(8) }
#undef UNETK_MH_FWD
  UNETK_LAUNCHED();
  return 0;
}

int head_multi_bwd_run(const void* x, int64_t ld, const float* w, const float* dlogits, float gscale, void* dx,
                       int64_t dx_ld, float* dw, float* db, int accumulate, int N, int64_t hw, int C, int K,
                       float* partial, cudaStream_t s) {
  UNETK_CHECK(mh_c_ok(C), -1, "head_multi: C=%d must be a power of two in [8,256]", C);
  UNETK_CHECK(K >= 1 && K <= kMaxK, -1, "head_multi: n_classes=%d (1..8)", K);
  const int64_t npix = static_cast<int64_t>(N) * hw;
  const int grid = mh_grid(npix, C);
  const int gpb = kThreads / (C / 8);
  const size_t smem = static_cast<size_t>(gpb) * (K * C + K) * sizeof(float);
  UNETK_CHECK(smem <= 200 * 1024, -1, "head_multi_bwd: C=%d x n_classes=%d needs %zu B of shared memory", C, K, smem);
  const __nv_bfloat16* xb = static_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* dxb = static_cast<__nv_bfloat16*>(dx);
#define UNETK_MH_BWD(KK)                                                                                              \
  case KK:                                                                                                            \
    if (smem > 48 * 1024)                                                                                             \
      UNETK_CUDA(cudaFuncSetAttribute(head_multi_bwd_kernel<KK>, cudaFuncAttributeMaxDynamicSharedMemorySize,         \
                                      static_cast<int>(smem)));                                                       \
    head_multi_bwd_kernel<KK><<<grid, kThreads, smem, s>>>(xb, ld, w, dlogits, gscale, dxb, dx_ld, npix, hw, C, partial); \
    break;
  switch (K) { UNETK_MH_BWD(1) UNETK_MH_BWD(2) UNETK_MH_BWD(3) UNETK_MH_BWD(4) UNETK_MH_BWD(5) UNETK_MH_BWD(6) UNETK_MH_BWD(7) UNETK_MH_BWD(8) }
#undef UNETK_MH_BWD
  UNETK_LAUNCHED();
  const int row = K * C + K;
  head_multi_finalize_kernel<<<(row + kSum2Lanes - 1) / kSum2Lanes, dim3(kSum2Lanes, kSum2Slices), 0, s>>>(partial, grid, K, C, dw, db, accumulate);
  UNETK_LAUNCHED();
  return 0;
}

size_t dice_partial_floats(int64_t groups, int64_t n) { return static_cast<size_t>(groups) * dice_blocks(groups, n) * 3; }

int dice_sums_run(const float* p, const float* t, int64_t groups, int64_t n, float lo, float hi, float* partial,
                  double* sums, cudaStream_t s) {
  UNETK_CHECK(groups >= 1 && groups <= 65535 && n >= 1, -1, "dice_sums: groups=%lld n=%lld", (long long)groups, (long long)n);
  const int nblk = dice_blocks(groups, n);
  dice_sums_kernel<<<dim3(nblk, static_cast<unsigned>(groups)), kThreads, 0, s>>>(p, t, n, lo, hi, partial);
  UNETK_LAUNCHED();
  dice_sums_final_kernel<<<static_cast<unsigned>((groups * 3 + 127) / 128), 128, 0, s>>>(partial, nblk, groups, sums);
  UNETK_LAUNCHED();
  return 0;
}

int dice_bwd_run(const float* p, const float* t, const float* coef, const float* gout, int64_t groups, int64_t n,
                 float lo, float hi, float* dp, cudaStream_t s) {
  UNETK_CHECK(groups >= 1 && groups <= 65535 && n >= 1, -1, "dice_bwd: groups=%lld n=%lld", (long long)groups, (long long)n);
  const int nblk = dice_blocks(groups, n);
  dice_bwd_kernel<<<dim3(nblk, static_cast<unsigned>(groups)), kThreads, 0, s>>>(p, t, coef, gout, n, lo, hi, dp);
  UNETK_LAUNCHED();
  return 0;
}

}  // namespace unetk
