// HBM-bound glue of the U-Net variants: n-ary residual adds / slice copies, nearest and bilinear 2x
// up-sampling with their backward passes, and a strided fp32 copy for derived weight caches.
// NHWC bf16, 16-byte vectors (8 channels per thread); a thread keeps one channel group and walks pixels
// with a constant pointer step (Lanes layout of elementwise.cu), so loops carry no integer division.
// Reference semantics replaced:
//   x + x1            Recurrent_block / RRCNN_block / ResidualConv / ResUNet   unet_parts.py:128,146,475, ResUNet.py:54
//   torch.cat copies  NestedUNet dense skips                                    UNetPP.py:75-99
//   nn.Upsample(scale_factor=2)  (nearest)  up_conv                             unet_parts.py:103
//   nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)            UNetPP.py:44
#include <cstdlib>
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

int lanes_grid(int64_t units, int C, int per_lane = 4, int blocks_per_sm = 8) {
  int ppb = kThreads / (C / 8);
  if (ppb < 1) ppb = 1;
  int64_t want = (units + static_cast<int64_t>(ppb) * per_lane - 1) / (static_cast<int64_t>(ppb) * per_lane);
  const int64_t cap = static_cast<int64_t>(num_sms()) * blocks_per_sm;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return static_cast<int>(want);
}

// ------------------------------------------------------------------ dst = [dst +] a [+ b [+ c [+ d]]]
struct AddArgs {
  __nv_bfloat16* dst; int64_t dst_ld;
  const __nv_bfloat16* src[4]; int64_t src_ld[4];
  int nsrc, accumulate;
  int64_t npix; int C;
};

// Every partial sum is rounded to bf16, like the chain of bf16 tensor adds it replaces.
template <int NSRC, bool ACC>
__global__ void __launch_bounds__(kThreads) add_n_kernel(const AddArgs A) {
  pdl_trigger();
  pdl_wait();
  Lanes L(A.C);
  if (!L.active) return;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * L.ppb;
  const int64_t first = static_cast<int64_t>(blockIdx.x) * L.ppb + L.pl;
  if (first >= A.npix) return;
  int64_t left = (A.npix - first + stride - 1) / stride;
  const __nv_bfloat16* ps[NSRC];
  int64_t ss[NSRC];
#pragma unroll
  for (int k = 0; k < NSRC; ++k) { ps[k] = A.src[k] + first * A.src_ld[k] + L.g * 8; ss[k] = stride * A.src_ld[k]; }
  __nv_bfloat16* pd = A.dst + first * A.dst_ld + L.g * 8;
  const int64_t sd = stride * A.dst_ld;
  constexpr int U = (NSRC + (ACC ? 1 : 0)) <= 2 ? 4 : 2;
  auto combine = [&](const uint4 (&v)[NSRC], const uint4& old, __nv_bfloat16* dst) {
    float acc[8], f[8];
    if constexpr (ACC) {
      unpack8(old, acc);
      unpack8(v[0], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = bf16_round(acc[j] + f[j]);
    } else {
      unpack8(v[0], acc);
    }
#pragma unroll
    for (int k = 1; k < NSRC; ++k) {
      unpack8(v[k], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = bf16_round(acc[j] + f[j]);
    }
    stg16(dst, pack8(acc));
  };
  for (; left >= U; left -= U) {
    uint4 v[U][NSRC], old[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int k = 0; k < NSRC; ++k) v[u][k] = ldg16(ps[k] + u * ss[k]);
      old[u] = ACC ? *reinterpret_cast<const uint4*>(pd + u * sd) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) combine(v[u], old[u], pd + u * sd);
#pragma unroll
    for (int k = 0; k < NSRC; ++k) ps[k] += U * ss[k];
    pd += U * sd;
  }
  for (; left > 0; --left) {
    uint4 v[NSRC];
#pragma unroll
    for (int k = 0; k < NSRC; ++k) v[k] = ldg16(ps[k]);
    const uint4 old = ACC ? *reinterpret_cast<const uint4*>(pd) : make_uint4(0, 0, 0, 0);
    combine(v, old, pd);
#pragma unroll
    for (int k = 0; k < NSRC; ++k) ps[k] += ss[k];
    pd += sd;
  }
}

template <int NSRC>
int launch_add(const AddArgs& A, int grid, cudaStream_t s) {
  if (A.accumulate) UNETK_CUDA(launch_pdl(add_n_kernel<NSRC, true>, dim3(grid), dim3(kThreads), 0, s, A));
  else UNETK_CUDA(launch_pdl(add_n_kernel<NSRC, false>, dim3(grid), dim3(kThreads), 0, s, A));
  return 0;
}

// ------------------------------------------------------------------ nearest 2x
// One unit = one INPUT pixel x 8 channels: fwd replicates it to the 2x2 output block; bwd sums the 2x2 block of
// dy in fp32 and rounds once (ATen's upsample_nearest2d_backward accumulates in float as well).
template <bool BWD, bool ACC>
__global__ void __launch_bounds__(kThreads)
nearest2x_kernel(const __nv_bfloat16* __restrict__ src, int64_t src_ld, __nv_bfloat16* __restrict__ dst,
                 int64_t dst_ld, int N, int H, int W, int C) {
  pdl_trigger();
  pdl_wait();
  // H, W: low-resolution size.  fwd: src low -> dst high.  bwd: src high (dy) -> dst low (dx).
  Lanes L(C);
  if (!L.active) return;
  const int64_t units = static_cast<int64_t>(N) * H * W;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * L.ppb;
  const uint32_t Wu = static_cast<uint32_t>(W), Hu = static_cast<uint32_t>(H);
  for (int64_t u = static_cast<int64_t>(blockIdx.x) * L.ppb + L.pl; u < units; u += stride) {
    const uint32_t uu = static_cast<uint32_t>(u);
    const uint32_t w = uu % Wu, t = uu / Wu;
    const uint32_t h = t % Hu, n = t / Hu;
    const int64_t hi0 = (static_cast<int64_t>(n) * 2 * H + 2 * h) * (2 * W) + 2 * w;
    if constexpr (!BWD) {
      const uint4 v = ldg16(src + u * src_ld + L.g * 8);
#pragma unroll
      for (int q = 0; q < 4; ++q) stg16(dst + (hi0 + (q >> 1) * 2 * W + (q & 1)) * dst_ld + L.g * 8, v);
    } else {
      uint4 v[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) v[q] = ldg16(src + (hi0 + (q >> 1) * 2 * W + (q & 1)) * src_ld + L.g * 8);
      float acc[8] = {}, f[8];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        unpack8(v[q], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += f[j];
      }
      __nv_bfloat16* d = dst + u * dst_ld + L.g * 8;
      if constexpr (ACC) {
        unpack8(*reinterpret_cast<const uint4*>(d), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = f[j] + bf16_round(acc[j]);
      }
      stg16(d, pack8(acc));
    }
  }
}

// ------------------------------------------------------------------ bilinear 2x, align_corners=True
// ATen upsample_bilinear2d (CUDA): scale = (in-1)/(out-1); src = scale*dst; i0 = (int)src; i1 = i0 + (i0 < in-1);
// l1 = src - i0; l0 = 1 - l1;  out = l0h*(l0w*v00 + l1w*v01) + l1h*(l0w*v10 + l1w*v11) in fp32, rounded once.
__device__ __forceinline__ void src_index(float scale, int o, int in, int* i0, int* i1, float* l0, float* l1) {
  const float s = scale * o;
  const int a = static_cast<int>(s);
  *i0 = a;
  *i1 = a + ((a < in - 1) ? 1 : 0);
  *l1 = s - a;
  *l0 = 1.f - *l1;
}

__global__ void __launch_bounds__(kThreads)
bilinear2x_fwd_kernel(const __nv_bfloat16* __restrict__ x, int64_t x_ld, __nv_bfloat16* __restrict__ y, int64_t y_ld,
                      int N, int H, int W, int C, float sh, float sw) {
  pdl_trigger();
  pdl_wait();
  Lanes L(C);
  if (!L.active) return;
  const int Ho = 2 * H, Wo = 2 * W;
  const int64_t units = static_cast<int64_t>(N) * Ho * Wo;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * L.ppb;
  for (int64_t u = static_cast<int64_t>(blockIdx.x) * L.ppb + L.pl; u < units; u += stride) {
    const uint32_t uu = static_cast<uint32_t>(u);
    const uint32_t wo = uu % Wo, t = uu / Wo;
    const uint32_t ho = t % Ho, n = t / Ho;
    int h0, h1, w0, w1;
    float lh0, lh1, lw0, lw1;
    src_index(sh, ho, H, &h0, &h1, &lh0, &lh1);
    src_index(sw, wo, W, &w0, &w1, &lw0, &lw1);
    const int64_t base = static_cast<int64_t>(n) * H * W;
    float a[8], b[8], c[8], d[8], o[8];
    unpack8(ldg16(x + (base + static_cast<int64_t>(h0) * W + w0) * x_ld + L.g * 8), a);
    unpack8(ldg16(x + (base + static_cast<int64_t>(h0) * W + w1) * x_ld + L.g * 8), b);
    unpack8(ldg16(x + (base + static_cast<int64_t>(h1) * W + w0) * x_ld + L.g * 8), c);
    unpack8(ldg16(x + (base + static_cast<int64_t>(h1) * W + w1) * x_ld + L.g * 8), d);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = lh0 * (lw0 * a[j] + lw1 * b[j]) + lh1 * (lw0 * c[j] + lw1 * d[j]);
    stg16(y + u * y_ld + L.g * 8, pack8(o));
  }
}

// ---- 2 x 2 output block per thread.  With scale = (in-1)/(2in-1) < 1/2 the source pair of output 2k is (k-1, k) and
// that of 2k+1 is (k, k+1) (clamped at the borders), so the four outputs (2k..2k+1, 2m..2m+1) read only the 3 x 3
// inputs around (k, m): 9 loads and 72 bf16->fp32 conversions for four results instead of 16 and 128, one index
// computation per block.  The per-pixel kernel above spent ~150 instructions per output and was issue-bound at
// 2.5 TB/s (UNet++: 1.7 ms per step for ten tensors).  The weights still come from src_index (ATen's arithmetic) and
// are attached to the loaded rows by comparing indices, so the borders (and H or W == 1) need no special case; in the
// interior the expression is the per-pixel kernel's, lh0*(lw0*a + lw1*b) + lh1*(lw0*c + lw1*d).
struct Pair3 {   // weights of the three loaded rows (or columns) k-1, k, k+1 for the even and the odd output
  float e0, e1, o1, o2;
};
__device__ __forceinline__ Pair3 pair_weights(float scale, int k, int in) {
  int i0, i1;
  float l0, l1;
  const int r0 = k > 0 ? k - 1 : 0, r1 = k, r2 = k + 1 < in ? k + 1 : in - 1;   // loaded (clamped) indices
  Pair3 w;
  src_index(scale, 2 * k, in, &i0, &i1, &l0, &l1);        // even output: rows r0, r1
  w.e0 = (r0 == i0 ? l0 : 0.f) + (r0 == i1 ? l1 : 0.f);
  w.e1 = (r1 != r0) ? (r1 == i0 ? l0 : 0.f) + (r1 == i1 ? l1 : 0.f) : 0.f;
  src_index(scale, 2 * k + 1, in, &i0, &i1, &l0, &l1);    // odd output: rows r1, r2
  w.o1 = (r1 == i0 ? l0 : 0.f) + (r1 == i1 ? l1 : 0.f);
  w.o2 = (r2 != r1) ? (r2 == i0 ? l0 : 0.f) + (r2 == i1 ? l1 : 0.f) : 0.f;
  return w;
}

__global__ void __launch_bounds__(kThreads)
bilinear2x_fwd_block_kernel(const __nv_bfloat16* __restrict__ x, int64_t x_ld, __nv_bfloat16* __restrict__ y,
                            int64_t y_ld, int N, int H, int W, int C, float sh, float sw, FastDiv fd_w, FastDiv fd_h) {
  pdl_trigger();
  pdl_wait();
  Lanes L(C);
  if (!L.active) return;
  const int64_t units = static_cast<int64_t>(N) * H * W;     // one per input pixel = per 2 x 2 output block
  const int64_t stride = static_cast<int64_t>(gridDim.x) * L.ppb;
  const int Wo = 2 * W;
  for (int64_t u = static_cast<int64_t>(blockIdx.x) * L.ppb + L.pl; u < units; u += stride) {
    uint32_t t, m, n, k;
    fd_w.divmod(static_cast<uint32_t>(u), t, m);
    fd_h.divmod(t, n, k);
    const Pair3 wr = pair_weights(sh, static_cast<int>(k), H), wc = pair_weights(sw, static_cast<int>(m), W);
    const int dk0 = k > 0 ? -1 : 0, dk2 = static_cast<int>(k) + 1 < H ? 1 : 0;
    const int dm0 = m > 0 ? -1 : 0, dm2 = static_cast<int>(m) + 1 < W ? 1 : 0;
    const __nv_bfloat16* p = x + u * x_ld + L.g * 8;          // input pixel (n, k, m)
    const int64_t rs = static_cast<int64_t>(W) * x_ld;
    uint4 v[3][3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const __nv_bfloat16* pr = p + (j == 0 ? dk0 : (j == 2 ? dk2 : 0)) * rs;
      v[j][0] = ldg16(pr + dm0 * x_ld);
      v[j][1] = ldg16(pr);
      v[j][2] = ldg16(pr + dm2 * x_ld);
    }
    float hx[3][2][8];                                        // row j, output column parity, channel
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float a[8], b[8], c[8];
      unpack8(v[j][0], a); unpack8(v[j][1], b); unpack8(v[j][2], c);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        hx[j][0][q] = wc.e0 * a[q] + wc.e1 * b[q];
        hx[j][1][q] = wc.o1 * b[q] + wc.o2 * c[q];
      }
    }
    __nv_bfloat16* o = y + ((static_cast<int64_t>(n) * 2 * H + 2 * k) * Wo + 2 * m) * y_ld + L.g * 8;
#pragma unroll
    for (int oc = 0; oc < 2; ++oc) {
      float e[8], od[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        e[q] = wr.e0 * hx[0][oc][q] + wr.e1 * hx[1][oc][q];
        od[q] = wr.o1 * hx[1][oc][q] + wr.o2 * hx[2][oc][q];
      }
      stg16(o + oc * y_ld, pack8(e));
      stg16(o + (static_cast<int64_t>(Wo) + oc) * y_ld, pack8(od));
    }
  }
}

// Gather form of the backward: input pixel i receives from every output o whose (i0, i1) pair contains i.
// With scale = (in-1)/(2in-1) < 1/2 those outputs lie in [2i-2, 2i+3]; at most four of them really touch i.
// The (output index, weight) lists of a row and of a column are built once per input pixel, so the inner
// loop is nothing but <= 16 vector loads and FMAs (the first version re-derived the source index of every
// candidate inside a 7 x 7 loop and ran at 4.5 ms per UNet++ step for ten tensors).
struct Taps {
  int idx[4];
  float w[4];
  int n;
};
__device__ __forceinline__ Taps contributors(float scale, int i, int in, int out) {
  Taps t;
  t.n = 0;
  int lo = 2 * i - 2, hi = 2 * i + 3;
  if (in == 1) { lo = 0; hi = out - 1; }   // every output reads input 0 with weight 1 (out == 2)
  if (lo < 0) lo = 0;
  if (hi > out - 1) hi = out - 1;
  for (int o = lo; o <= hi; ++o) {
    int i0, i1;
    float l0, l1;
    src_index(scale, o, in, &i0, &i1, &l0, &l1);
    float w = 0.f;
    if (i0 == i) w += l0;
    if (i1 == i) w += l1;
    if (w != 0.f && t.n < 4) { t.idx[t.n] = o; t.w[t.n] = w; ++t.n; }
  }
  return t;
}

template <bool ACC>
__global__ void __launch_bounds__(kThreads)
bilinear2x_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int64_t dy_ld, __nv_bfloat16* __restrict__ dx,
                      int64_t dx_ld, int N, int H, int W, int C, float sh, float sw) {
  pdl_trigger();
  pdl_wait();
  Lanes L(C);
  if (!L.active) return;
  const int Ho = 2 * H, Wo = 2 * W;
  const int64_t units = static_cast<int64_t>(N) * H * W;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * L.ppb;
  for (int64_t u = static_cast<int64_t>(blockIdx.x) * L.ppb + L.pl; u < units; u += stride) {
    const uint32_t uu = static_cast<uint32_t>(u);
    const int w = static_cast<int>(uu % W);
    const uint32_t t = uu / W;
    const int h = static_cast<int>(t % H);
    const int n = static_cast<int>(t / H);
    const Taps th = contributors(sh, h, H, Ho), tw = contributors(sw, w, W, Wo);
    float acc[8] = {};
    const __nv_bfloat16* base = dy + static_cast<int64_t>(n) * Ho * Wo * dy_ld + L.g * 8;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      if (a >= th.n) break;
      uint4 v[4];
#pragma unroll
      for (int b = 0; b < 4; ++b)
        v[b] = (b < tw.n) ? ldg16(base + (static_cast<int64_t>(th.idx[a]) * Wo + tw.idx[b]) * dy_ld) : make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        if (b >= tw.n) break;
        float f[8];
        unpack8(v[b], f);
        const float k = th.w[a] * tw.w[b];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(k, f[j], acc[j]);
      }
    }
    __nv_bfloat16* d = dx + u * dx_ld + L.g * 8;
    if constexpr (ACC) {
      float f[8];
      unpack8(*reinterpret_cast<const uint4*>(d), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = f[j] + bf16_round(acc[j]);
    }
    stg16(d, pack8(acc));
  }
}

// ---- 2 x 2 INPUT block per thread: inputs (k..k+1, m..m+1) receive from the 6 x 6 outputs (2k-1..2k+4, 2m-1..2m+4);
// input k takes rows 2k-1..2k+2 (positions 0..3 of the six), input k+1 rows 2k+1..2k+4 (positions 2..5).  Each output
// row is reduced along W first (two 4-term sums), then added to the one or two input rows it feeds: 36 loads and
// ~580 FMAs for four results against 64 loads and ~600 FMAs + four tap-list constructions in the per-pixel kernel
// (2.6 ms per UNet++ step, issue-bound at 1.6 TB/s).  Weights from src_index, attached by comparing indices.
struct Six {
  float lo[4], hi[4];   // weight of output position j (lo) / j + 2 (hi) for the block's first / second input
};
__device__ __forceinline__ Six six_weights(float scale, int k, int in) {
  Six w;
  const int out = 2 * in;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    const int o = 2 * k - 1 + j;
    float a = 0.f, b = 0.f;
    if (o >= 0 && o < out) {
      int i0, i1;
      float l0, l1;
      src_index(scale, o, in, &i0, &i1, &l0, &l1);
      a = (i0 == k ? l0 : 0.f) + (i1 == k ? l1 : 0.f);
      b = (i0 == k + 1 ? l0 : 0.f) + (i1 == k + 1 ? l1 : 0.f);
    }
    if (j < 4) w.lo[j] = a;
    if (j >= 2) w.hi[j - 2] = b;
  }
  return w;
}

template <bool ACC>
__global__ void __launch_bounds__(kThreads)
bilinear2x_bwd_block_kernel(const __nv_bfloat16* __restrict__ dy, int64_t dy_ld, __nv_bfloat16* __restrict__ dx,
                            int64_t dx_ld, int N, int H, int W, int C, float sh, float sw, FastDiv fd_wb, FastDiv fd_hb) {
  pdl_trigger();
  pdl_wait();
  Lanes L(C);
  if (!L.active) return;
  const int Hb = (H + 1) >> 1, Wb = (W + 1) >> 1, Ho = 2 * H, Wo = 2 * W;
  const int64_t units = static_cast<int64_t>(N) * Hb * Wb;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * L.ppb;
  for (int64_t u = static_cast<int64_t>(blockIdx.x) * L.ppb + L.pl; u < units; u += stride) {
    uint32_t t, mb, n, kb;
    fd_wb.divmod(static_cast<uint32_t>(u), t, mb);
    fd_hb.divmod(t, n, kb);
    const int k = 2 * static_cast<int>(kb), m = 2 * static_cast<int>(mb);
    const Six wr = six_weights(sh, k, H), wc = six_weights(sw, m, W);
    float acc[2][2][8] = {};
    const __nv_bfloat16* base = dy + static_cast<int64_t>(n) * Ho * Wo * dy_ld + L.g * 8;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const int r = 2 * k - 1 + j;
      if (r < 0 || r >= Ho) continue;
      const __nv_bfloat16* row = base + static_cast<int64_t>(r) * Wo * dy_ld;
      uint4 v[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const int c = 2 * m - 1 + i;
        v[i] = (c >= 0 && c < Wo) ? ldg16(row + static_cast<int64_t>(c) * dy_ld) : make_uint4(0, 0, 0, 0);
      }
      float h0[8] = {}, h1[8] = {};
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        float f[8];
        unpack8(v[i], f);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (i < 4) h0[q] = fmaf(wc.lo[i], f[q], h0[q]);
          if (i >= 2) h1[q] = fmaf(wc.hi[i - 2], f[q], h1[q]);
        }
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (j < 4) { acc[0][0][q] = fmaf(wr.lo[j], h0[q], acc[0][0][q]); acc[0][1][q] = fmaf(wr.lo[j], h1[q], acc[0][1][q]); }
        if (j >= 2) { acc[1][0][q] = fmaf(wr.hi[j - 2], h0[q], acc[1][0][q]); acc[1][1][q] = fmaf(wr.hi[j - 2], h1[q], acc[1][1][q]); }
      }
    }
#pragma unroll
    for (int a = 0; a < 2; ++a) {
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        if (k + a >= H || m + b >= W) continue;
        __nv_bfloat16* d = dx + ((static_cast<int64_t>(n) * H + k + a) * W + m + b) * dx_ld + L.g * 8;
        if constexpr (ACC) {
          float f[8];
          unpack8(*reinterpret_cast<const uint4*>(d), f);
#pragma unroll
          for (int q = 0; q < 8; ++q) acc[a][b][q] = f[q] + bf16_round(acc[a][b][q]);
        }
        stg16(d, pack8(acc[a][b]));
      }
    }
  }
}

__global__ void copy_f32_strided_kernel(float* __restrict__ dst, int64_t ds, const float* __restrict__ src, int64_t ss,
                                        int64_t n, int accumulate) {
  pdl_trigger();
  pdl_wait();
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float v = src[i * ss];
    dst[i * ds] = accumulate ? dst[i * ds] + v : v;
  }
}

}  // namespace

#define CHECK_C(C) UNETK_CHECK((C) % 8 == 0 && (C) >= 8 && (C) <= 2048, -1, "channel count %d must be a multiple of 8 in [8,2048]", (C))

int add_n_run(void* dst, int64_t dst_ld, int accumulate, const void* const* src, const int64_t* src_ld, int nsrc,
              int64_t npix, int C, cudaStream_t s) {
  CHECK_C(C);
  UNETK_CHECK(nsrc >= 1 && nsrc <= 4 && npix > 0, -1, "add_n: 1..4 sources");
  AddArgs A{};
  A.dst = static_cast<__nv_bfloat16*>(dst); A.dst_ld = dst_ld; A.nsrc = nsrc; A.accumulate = accumulate;
  A.npix = npix; A.C = C;
  for (int k = 0; k < nsrc; ++k) {
    UNETK_CHECK(src[k] != nullptr, -1, "add_n: null source %d", k);
    A.src[k] = static_cast<const __nv_bfloat16*>(src[k]); A.src_ld[k] = src_ld[k];
  }
  const int grid = lanes_grid(npix, C);
  switch (nsrc) {
    case 1: launch_add<1>(A, grid, s); break;
    case 2: launch_add<2>(A, grid, s); break;
    case 3: launch_add<3>(A, grid, s); break;
    default: launch_add<4>(A, grid, s); break;
  }
  UNETK_LAUNCHED();
  return 0;
}

int upsample_nearest2x_run(const void* src, int64_t src_ld, void* dst, int64_t dst_ld, int backward, int accumulate,
                           int N, int H, int W, int C, cudaStream_t s) {
  CHECK_C(C);
  const int64_t units = static_cast<int64_t>(N) * H * W;
  UNETK_CHECK(units > 0 && units < (1ll << 31), -1, "upsample_nearest2x: bad size");
  const int grid = lanes_grid(units, C, 2);
  const __nv_bfloat16* a = static_cast<const __nv_bfloat16*>(src);
  __nv_bfloat16* d = static_cast<__nv_bfloat16*>(dst);
  if (!backward) UNETK_CUDA(launch_pdl(nearest2x_kernel<false, false>, dim3(grid), dim3(kThreads), 0, s, a, src_ld, d, dst_ld, N, H, W, C));
  else if (accumulate) UNETK_CUDA(launch_pdl(nearest2x_kernel<true, true>, dim3(grid), dim3(kThreads), 0, s, a, src_ld, d, dst_ld, N, H, W, C));
  else UNETK_CUDA(launch_pdl(nearest2x_kernel<true, false>, dim3(grid), dim3(kThreads), 0, s, a, src_ld, d, dst_ld, N, H, W, C));
  UNETK_LAUNCHED();
  return 0;
}

int upsample_bilinear2x_run(const void* src, int64_t src_ld, void* dst, int64_t dst_ld, int backward, int accumulate,
                            int N, int H, int W, int C, cudaStream_t s) {
  CHECK_C(C);
  const int64_t units = static_cast<int64_t>(N) * H * W * (backward ? 1 : 4);
  UNETK_CHECK(units > 0 && units < (1ll << 31) && H >= 1 && W >= 1, -1, "upsample_bilinear2x: bad size");
  // area_pixel_compute_scale(align_corners=True): (in - 1) / (out - 1), 0 when out == 1 (cannot happen for 2x)
  const float sh = static_cast<float>(H - 1) / static_cast<float>(2 * H - 1);
  const float sw = static_cast<float>(W - 1) / static_cast<float>(2 * W - 1);
  const __nv_bfloat16* a = static_cast<const __nv_bfloat16*>(src);
  __nv_bfloat16* d = static_cast<__nv_bfloat16*>(dst);
  static int block = -1;   // UNETK_BILINEAR_BLOCK=0: the one-pixel-per-thread kernels
  if (block < 0) { const char* e = getenv("UNETK_BILINEAR_BLOCK"); block = e ? atoi(e) : 1; }
  if (block) {
    if (!backward) {
      const int g = lanes_grid(static_cast<int64_t>(N) * H * W, C, 2);
      UNETK_CUDA(launch_pdl(bilinear2x_fwd_block_kernel, dim3(g), dim3(kThreads), 0, s, a, src_ld, d, dst_ld, N, H, W, C, sh, sw,
                            FastDiv(static_cast<uint32_t>(W)), FastDiv(static_cast<uint32_t>(H))));
    } else {
      const int hb = (H + 1) / 2, wb = (W + 1) / 2;
      const int g = lanes_grid(static_cast<int64_t>(N) * hb * wb, C, 1);
      if (accumulate) UNETK_CUDA(launch_pdl(bilinear2x_bwd_block_kernel<true>, dim3(g), dim3(kThreads), 0, s, a, src_ld, d, dst_ld, N, H, W, C, sh, sw,
                                            FastDiv(static_cast<uint32_t>(wb)), FastDiv(static_cast<uint32_t>(hb))));
      else UNETK_CUDA(launch_pdl(bilinear2x_bwd_block_kernel<false>, dim3(g), dim3(kThreads), 0, s, a, src_ld, d, dst_ld, N, H, W, C, sh, sw,
                                 FastDiv(static_cast<uint32_t>(wb)), FastDiv(static_cast<uint32_t>(hb))));
    }
    UNETK_LAUNCHED();
    return 0;
  }
  const int grid = lanes_grid(units, C, 2);
  if (!backward) UNETK_CUDA(launch_pdl(bilinear2x_fwd_kernel, dim3(grid), dim3(kThreads), 0, s, a, src_ld, d, dst_ld, N, H, W, C, sh, sw));
  else if (accumulate) UNETK_CUDA(launch_pdl(bilinear2x_bwd_kernel<true>, dim3(grid), dim3(kThreads), 0, s, a, src_ld, d, dst_ld, N, H, W, C, sh, sw));
  else UNETK_CUDA(launch_pdl(bilinear2x_bwd_kernel<false>, dim3(grid), dim3(kThreads), 0, s, a, src_ld, d, dst_ld, N, H, W, C, sh, sw));
  UNETK_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------ patch gather (train.py:200-253 on the device)
// out_images[b][i][j][c] = images[img_b][c][x_b - P/2 + i][y_b - P/2 + j]  (fp32, channels_last batch);
// out_labels[b][i][j]    = labels[img_b][x_b - P/2 + i][y_b - P/2 + j].   centers = int32 [B][3] = (img, x, y).
struct GatherArgs {
  const float* images; long long si_n, si_c, si_h, si_w;
  const float* labels; long long sl_n, sl_h, sl_w;
  const int* centers;
  float* out_images; float* out_labels;
  int B, C, P, H, W;
};
__global__ void __launch_bounds__(kThreads) gather_patches_kernel(const GatherArgs a) {
  pdl_trigger();
  pdl_wait();
  const long long total = static_cast<long long>(a.B) * a.P * a.P;
  const int half = a.P / 2;
  for (long long u = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; u < total;
       u += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int j = static_cast<int>(u % a.P), i = static_cast<int>((u / a.P) % a.P), b = static_cast<int>(u / (static_cast<long long>(a.P) * a.P));
    const int img = __ldg(a.centers + 3 * b), x = __ldg(a.centers + 3 * b + 1) - half + i, y = __ldg(a.centers + 3 * b + 2) - half + j;
    const bool ok = x >= 0 && x < a.H && y >= 0 && y < a.W;   // the host validates the centres; never read out of bounds
    const float* src = a.images + img * a.si_n + x * a.si_h + y * a.si_w;
    float* dst = a.out_images + u * a.C;
    for (int c = 0; c < a.C; ++c) dst[c] = ok ? __ldg(src + c * a.si_c) : 0.f;
    if (a.out_labels != nullptr) a.out_labels[u] = ok ? __ldg(a.labels + img * a.sl_n + x * a.sl_h + y * a.sl_w) : 0.f;
  }
}
int gather_patches_run(const float* images, int64_t si_n, int64_t si_c, int64_t si_h, int64_t si_w, const float* labels,
                       int64_t sl_n, int64_t sl_h, int64_t sl_w, const int* centers, int B, int C, int P, int H, int W,
                       float* out_images, float* out_labels, cudaStream_t s) {
  GatherArgs a{images, si_n, si_c, si_h, si_w, labels, sl_n, sl_h, sl_w, centers, out_images, out_labels, B, C, P, H, W};
  const long long total = static_cast<long long>(B) * P * P;
  long long blocks = (total + kThreads - 1) / kThreads;
  if (blocks > 16LL * num_sms()) blocks = 16LL * num_sms();
  UNETK_CUDA(launch_pdl(gather_patches_kernel, dim3(static_cast<unsigned>(blocks)), dim3(kThreads), 0, s, a));
  UNETK_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------ sliding-window inference (evaluate.py:28-96)
// acc[h][w] += p_b[h - y_b][w - x_b] for every patch b of the batch covering (h, w), in batch order (the order of the
// reference's accumulation loop, evaluate.py:84-86); cnt likewise += 1.  One thread per output pixel: no atomics.
__global__ void __launch_bounds__(kThreads) tile_accumulate_kernel(const float* __restrict__ logits, const int* __restrict__ pos,
                                                                   int B, int P, int H, int W, int apply_sigmoid,
                                                                   double* __restrict__ acc, double* __restrict__ cnt) {
  pdl_trigger();
  pdl_wait();
  const long long total = static_cast<long long>(H) * W;
  for (long long u = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; u < total;
       u += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int w = static_cast<int>(u % W), h = static_cast<int>(u / W);
    double a = acc[u], c = cnt[u];
    for (int b = 0; b < B; ++b) {
      const int y = __ldg(pos + 2 * b), x = __ldg(pos + 2 * b + 1);
      if (h >= y && h < y + P && w >= x && w < x + P) {
        float v = __ldg(logits + (static_cast<long long>(b) * P + (h - y)) * P + (w - x));
        if (apply_sigmoid) v = 1.f / (1.f + expf(-v));
        a += static_cast<double>(v);
        c += 1.0;
      }
    }
    acc[u] = a;
    cnt[u] = c;
  }
}
__global__ void tile_finalize_kernel(const double* __restrict__ acc, const double* __restrict__ cnt, long long n,
                                     double* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  for (long long u = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; u < n;
       u += static_cast<long long>(gridDim.x) * blockDim.x)
    out[u] = cnt[u] != 0.0 ? acc[u] / cnt[u] : 0.0;   // np.divide(..., where=count != 0), evaluate.py:89-91
}
int tile_accumulate_run(const float* logits, const int* pos, int B, int P, int H, int W, int apply_sigmoid, double* acc,
                        double* cnt, cudaStream_t s) {
  const long long total = static_cast<long long>(H) * W;
  long long blocks = (total + kThreads - 1) / kThreads;
  if (blocks > 16LL * num_sms()) blocks = 16LL * num_sms();
  UNETK_CUDA(launch_pdl(tile_accumulate_kernel, dim3(static_cast<unsigned>(blocks)), dim3(kThreads), 0, s, logits, pos, B, P,
                        H, W, apply_sigmoid, acc, cnt));
  UNETK_LAUNCHED();
  return 0;
}
int tile_finalize_run(const double* acc, const double* cnt, long long n, double* out, cudaStream_t s) {
  long long blocks = (n + kThreads - 1) / kThreads;
  if (blocks > 16LL * num_sms()) blocks = 16LL * num_sms();
  UNETK_CUDA(launch_pdl(tile_finalize_kernel, dim3(static_cast<unsigned>(blocks)), dim3(kThreads), 0, s, acc, cnt, n, out));
  UNETK_LAUNCHED();
  return 0;
}

int copy_f32_strided_run(float* dst, int64_t ds, const float* src, int64_t ss, int64_t n, int accumulate,
                         cudaStream_t s) {
  UNETK_CHECK(dst && src && n > 0 && ds > 0 && ss > 0, -1, "copy_f32_strided: bad arguments");
  int64_t b = (n + 255) / 256;
  if (b > 1024) b = 1024;
  UNETK_CUDA(launch_pdl(copy_f32_strided_kernel, dim3(static_cast<int>(b)), dim3(256), 0, s, dst, ds, src, ss, n, accumulate));
  UNETK_LAUNCHED();
  return 0;
}

}  // namespace unetk
