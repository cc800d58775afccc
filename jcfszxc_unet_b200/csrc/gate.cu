// Attention gate (Attention_block, UNetFamily/utils/unet_parts.py:149-176) as fused HBM-bound kernels.
//
//   g1 = BN_g(conv1x1_g(g))            x1 = BN_x(conv1x1_x(x))         <- tensor-core 1x1 GEMMs with the BN
//   a  = relu(g1 + x1)                                                     statistics in their epilogue
//   s  = conv1x1_psi(a)  (F_int -> 1)  psi = sigmoid(BN_1(s))          out = x * psi
//
// The two 1x1 GEMMs write raw_g / raw_x (bf16, F_int channels).  Everything after them is four passes:
//   gate_fwd        : reads raw_g, raw_x        -> s[pixel] (+ sum s, sum s^2 for BN_1)         "add-relu-dot-w_psi"
//   gate_apply      : reads x, s                -> out = x * sigmoid(BN_1(s)) into the concat slice
//   gate_bwd_psi    : reads dout, x, s          -> dx (+)= dout*psi, dz = d/dBN_1-output, BN_1 backward sums
//   gate_bwd_reduce : reads raw_g, raw_x, s, dz -> per-channel sums for BN_g / BN_x backward, dw_psi, db_psi
//   gate_bwd_apply  : reads raw_g, raw_x, s, dz -> draw_g, draw_x (gradients of the two GEMM outputs)
// against ~7 elementwise/reduction launches plus their intermediates in the reference.  Intermediates are rounded
// to bf16 where the reference's autocast graph holds bf16 tensors (g1, x1, g1+x1, s, BN_1(s), psi).
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
__device__ __forceinline__ float sigmoidf(float z) { return 1.f / (1.f + __expf(-z)); }
__device__ __forceinline__ float group_sum(float v, int lpp) {
  for (int o = lpp >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// psi of one pixel, with the reference's bf16 roundings (BN_1 output, sigmoid output)
__device__ __forceinline__ float psi_of(float s, float sc1, float sh1) {
  return bf16_round(sigmoidf(bf16_round(fmaf(s, sc1, sh1))));
}

// a[j] = relu(bf16(bf16(bn_g) + bf16(bn_x))) for 8 channels
__device__ __forceinline__ void gate_act8(const uint4& ug, const uint4& ux, const float (&scg)[8], const float (&shg)[8],
                                          const float (&scx)[8], const float (&shx)[8], float* rg, float* rx, float* a) {
  unpack8(ug, rg);
  unpack8(ux, rx);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float v = bf16_round(bf16_round(fmaf(rg[j], scg[j], shg[j])) + bf16_round(fmaf(rx[j], scx[j], shx[j])));
    a[j] = fmaxf(v, 0.f);
  }
}

struct GateArgs {
  const __nv_bfloat16* rawg; int64_t rawg_ld;
  const __nv_bfloat16* rawx; int64_t rawx_ld;
  const float* scg; const float* shg; const float* scx; const float* shx;
  const float* wpsi; const float* bpsi;
  int64_t npix; int F;
};

// Pixel-group layout: lpp = F/8 lanes (power of two <= 32) share one pixel; gpb pixel groups per block.
// ------------------------------------------------------------------ s = w_psi . relu(g1 + x1) + b
__global__ void __launch_bounds__(kThreads)
gate_fwd_kernel(const GateArgs A, float* __restrict__ s_out, float* __restrict__ partial) {
  pdl_trigger();
  pdl_wait();
  const int lpp = A.F >> 3, gpb = kThreads / lpp;
  const int sub = threadIdx.x % lpp, grp = threadIdx.x / lpp;
  float scg[8], shg[8], scx[8], shx[8], wv[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = sub * 8 + j;
    scg[j] = __ldg(A.scg + c); shg[j] = __ldg(A.shg + c); scx[j] = __ldg(A.scx + c); shx[j] = __ldg(A.shx + c);
    wv[j] = bf16_round(__ldg(A.wpsi + c));   // autocast casts the conv weight to bf16
  }
  const float b = A.bpsi ? bf16_round(__ldg(A.bpsi)) : 0.f;
  float sum = 0.f, sumsq = 0.f;
  constexpr int U = 4;
  const int64_t step = static_cast<int64_t>(gridDim.x) * gpb;
  for (int64_t base = static_cast<int64_t>(blockIdx.x) * gpb; base < A.npix; base += U * step) {   // block-uniform trip count
    const int64_t p0 = base + grp;
    uint4 ug[U], ux[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t pix = p0 + u * step;
      const bool ok = pix < A.npix;
      ug[u] = ok ? ldg16(A.rawg + pix * A.rawg_ld + sub * 8) : make_uint4(0, 0, 0, 0);
      ux[u] = ok ? ldg16(A.rawx + pix * A.rawx_ld + sub * 8) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t pix = p0 + u * step;
      float rg[8], rx[8], a[8], dot = 0.f;
      gate_act8(ug[u], ux[u], scg, shg, scx, shx, rg, rx, a);
#pragma unroll
      for (int j = 0; j < 8; ++j) dot = fmaf(a[j], wv[j], dot);
      dot = group_sum(dot, lpp);
      if (sub == 0 && pix < A.npix) {
        const float s = bf16_round(dot + b);
        s_out[pix] = s;
        sum += s;
        sumsq = fmaf(s, s, sumsq);
      }
    }
  }
  __shared__ float red[2][kThreads / 32];
  sum = warp_sum(sum); sumsq = warp_sum(sumsq);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = sum; red[1][warp] = sumsq; }
  __syncthreads();
  if (threadIdx.x < 2) {
    float t = 0.f;
    for (int i = 0; i < kThreads / 32; ++i) t += red[threadIdx.x][i];
    partial[static_cast<size_t>(blockIdx.x) * 2 + threadIdx.x] = t;
  }
}

__global__ void pair_sums_kernel(const float* __restrict__ partial, int nblk, double* __restrict__ sums) {
  pdl_trigger();
  pdl_wait();
  const int k = threadIdx.x;
  if (k >= 2) return;
  double s = 0.0;
  for (int b = 0; b < nblk; ++b) s += partial[static_cast<size_t>(b) * 2 + k];
  sums[k] = s;
}

// ------------------------------------------------------------------ out = x * psi           (F_l channels)
// lpp = min(32, F/8) lanes per pixel, each lane walks channel groups sub, sub+lpp, ...
__global__ void __launch_bounds__(kThreads)
gate_apply_kernel(const __nv_bfloat16* __restrict__ x, int64_t x_ld, const float* __restrict__ s,
                  const float* __restrict__ sc1, const float* __restrict__ sh1, __nv_bfloat16* __restrict__ out,
                  int64_t out_ld, int64_t npix, int F) {
  pdl_trigger();
  pdl_wait();
  const int cg = F >> 3;
  const int lpp = cg < 32 ? cg : 32, gpb = kThreads / lpp;
  const int sub = threadIdx.x % lpp, grp = threadIdx.x / lpp;
  const float a1 = __ldg(sc1), b1 = __ldg(sh1);
  const int64_t step = static_cast<int64_t>(gridDim.x) * gpb;
  for (int64_t pix = static_cast<int64_t>(blockIdx.x) * gpb + grp; pix < npix; pix += step) {
    const float psi = psi_of(__ldg(s + pix), a1, b1);
    for (int g0 = sub; g0 < cg; g0 += 4 * lpp) {
      uint4 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        v[k] = (g0 + k * lpp < cg) ? ldg16(x + pix * x_ld + (g0 + k * lpp) * 8) : make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (g0 + k * lpp >= cg) break;
        float f[8];
        unpack8(v[k], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] *= psi;
        stg16(out + pix * out_ld + (g0 + k * lpp) * 8, pack8(f));
      }
    }
  }
}

// ------------------------------------------------------------------ backward through out = x * psi, sigmoid, into BN_1
// dx (+)= dout * psi;  dpsi = sum_c dout*x;  dz = dpsi * psi * (1 - psi);  partial: sum dz, sum dz * (s - mean1)
template <bool ACC>
__global__ void __launch_bounds__(kThreads)
gate_bwd_psi_kernel(const __nv_bfloat16* __restrict__ dout, int64_t dout_ld, const __nv_bfloat16* __restrict__ x,
                    int64_t x_ld, const float* __restrict__ s, const float* __restrict__ sc1,
                    const float* __restrict__ sh1, const float* __restrict__ mean1, __nv_bfloat16* __restrict__ dx,
                    int64_t dx_ld, float* __restrict__ dz, float* __restrict__ partial, int64_t npix, int F) {
  pdl_trigger();
  pdl_wait();
  const int cg = F >> 3;
  const int lpp = cg < 32 ? cg : 32, gpb = kThreads / lpp;
  const int sub = threadIdx.x % lpp, grp = threadIdx.x / lpp;
  const float a1 = __ldg(sc1), b1 = __ldg(sh1), mu = __ldg(mean1);
  float sum0 = 0.f, sum1 = 0.f;
  const int64_t step = static_cast<int64_t>(gridDim.x) * gpb;
  for (int64_t p0 = static_cast<int64_t>(blockIdx.x) * gpb; p0 < npix; p0 += step) {   // warp-uniform
    const int64_t pix = p0 + grp;
    const bool ok = pix < npix;
    const float sv = ok ? __ldg(s + pix) : 0.f;
    const float psi = psi_of(sv, a1, b1);
    float dot = 0.f;
    if (ok) {
      for (int g0 = sub; g0 < cg; g0 += 2 * lpp) {
        const bool two = g0 + lpp < cg;
        const uint4 d0 = ldg16(dout + pix * dout_ld + g0 * 8), x0 = ldg16(x + pix * x_ld + g0 * 8);
        const uint4 d1 = two ? ldg16(dout + pix * dout_ld + (g0 + lpp) * 8) : make_uint4(0, 0, 0, 0);
        const uint4 x1 = two ? ldg16(x + pix * x_ld + (g0 + lpp) * 8) : make_uint4(0, 0, 0, 0);
        float fd[8], fx[8], o[8];
        unpack8(d0, fd); unpack8(x0, fx);
#pragma unroll
        for (int j = 0; j < 8; ++j) { dot = fmaf(fd[j], fx[j], dot); o[j] = fd[j] * psi; }
        __nv_bfloat16* dst = dx + pix * dx_ld + g0 * 8;
        if constexpr (ACC) {
          float old[8];
          unpack8(*reinterpret_cast<const uint4*>(dst), old);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = old[j] + bf16_round(o[j]);
        }
        stg16(dst, pack8(o));
        if (two) {
          unpack8(d1, fd); unpack8(x1, fx);
#pragma unroll
          for (int j = 0; j < 8; ++j) { dot = fmaf(fd[j], fx[j], dot); o[j] = fd[j] * psi; }
          dst = dx + pix * dx_ld + (g0 + lpp) * 8;
          if constexpr (ACC) {
            float old[8];
            unpack8(*reinterpret_cast<const uint4*>(dst), old);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = old[j] + bf16_round(o[j]);
          }
          stg16(dst, pack8(o));
        }
      }
    }
    dot = group_sum(dot, lpp);
    if (sub == 0 && ok) {
      const float d = dot * psi * (1.f - psi);
      dz[pix] = d;
      sum0 += d;
      sum1 = fmaf(d, sv - mu, sum1);
    }
  }
  __shared__ float red[2][kThreads / 32];
  sum0 = warp_sum(sum0); sum1 = warp_sum(sum1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = sum0; red[1][warp] = sum1; }
  __syncthreads();
  if (threadIdx.x < 2) {
    float t = 0.f;
    for (int i = 0; i < kThreads / 32; ++i) t += red[threadIdx.x][i];
    partial[static_cast<size_t>(blockIdx.x) * 2 + threadIdx.x] = t;
  }
}

// ------------------------------------------------------------------ backward into BN_g / BN_x (F_int channels)
// Lanes layout by channel group: a thread keeps 8 channels and walks pixels.
struct GLanes {
  int cg, ppb, g, pl;
  bool active;
  __device__ GLanes(int C) {
    cg = C >> 3; ppb = kThreads / cg; if (ppb < 1) ppb = 1;
    g = threadIdx.x % cg; pl = threadIdx.x / cg; active = pl < ppb;
  }
};

struct GateBwdArgs {
  GateArgs G;
  const float* mug; const float* mux;
  const float* s; const float* dz;
  const float* sc1; const float* coef1;   // BN_1: scale, (K0, K1):  ds = sc1*dz + K1*s + K0
};

// partial[blk][k][F], k: 0 sum da, 1 sum da*(rawg-mug), 2 sum da*(rawx-mux), 3 sum ds*a (dw_psi), 4 sum ds (channel 0 only)
__global__ void __launch_bounds__(kThreads)
gate_bwd_reduce_kernel(const GateBwdArgs B, float* __restrict__ partial) {
  pdl_trigger();
  pdl_wait();
  const GateArgs& A = B.G;
  GLanes L(A.F);
  float acc[5][8] = {};
  if (L.active) {
    float scg[8], shg[8], scx[8], shx[8], wv[8], mg[8], mx[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = L.g * 8 + j;
      scg[j] = __ldg(A.scg + c); shg[j] = __ldg(A.shg + c); scx[j] = __ldg(A.scx + c); shx[j] = __ldg(A.shx + c);
      wv[j] = bf16_round(__ldg(A.wpsi + c)); mg[j] = __ldg(B.mug + c); mx[j] = __ldg(B.mux + c);
    }
    const float a1 = __ldg(B.sc1), k0 = __ldg(B.coef1), k1 = __ldg(B.coef1 + 1);
    const int64_t stride = static_cast<int64_t>(gridDim.x) * L.ppb;
    for (int64_t pix = static_cast<int64_t>(blockIdx.x) * L.ppb + L.pl; pix < A.npix; pix += 2 * stride) {
      const bool two = pix + stride < A.npix;
      const uint4 ug0 = ldg16(A.rawg + pix * A.rawg_ld + L.g * 8), ux0 = ldg16(A.rawx + pix * A.rawx_ld + L.g * 8);
      const uint4 ug1 = two ? ldg16(A.rawg + (pix + stride) * A.rawg_ld + L.g * 8) : make_uint4(0, 0, 0, 0);
      const uint4 ux1 = two ? ldg16(A.rawx + (pix + stride) * A.rawx_ld + L.g * 8) : make_uint4(0, 0, 0, 0);
      const float s0 = __ldg(B.s + pix), z0 = __ldg(B.dz + pix);
      const float s1 = two ? __ldg(B.s + pix + stride) : 0.f, z1 = two ? __ldg(B.dz + pix + stride) : 0.f;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (u == 1 && !two) break;
        const float ds = fmaf(a1, u ? z1 : z0, fmaf(k1, u ? s1 : s0, k0));
        float rg[8], rx[8], a[8];
        gate_act8(u ? ug1 : ug0, u ? ux1 : ux0, scg, shg, scx, shx, rg, rx, a);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float da = (a[j] > 0.f) ? ds * wv[j] : 0.f;
          acc[0][j] += da;
          acc[1][j] = fmaf(da, rg[j] - mg[j], acc[1][j]);
          acc[2][j] = fmaf(da, rx[j] - mx[j], acc[2][j]);
          acc[3][j] = fmaf(ds, a[j], acc[3][j]);
        }
        if (L.g == 0) acc[4][0] += ds;
      }
    }
  }
  extern __shared__ float red[];  // [ppb][5][F]
  const int F = A.F;
  if (L.active) {
#pragma unroll
    for (int k = 0; k < 5; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) red[(L.pl * 5 + k) * F + L.g * 8 + j] = acc[k][j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 5 * F; i += kThreads) {
    float t = 0.f;
    for (int pl = 0; pl < L.ppb; ++pl) t += red[pl * 5 * F + i];
    partial[static_cast<size_t>(blockIdx.x) * 5 * F + i] = t;
  }
}

// sums_g = double[2][F] (S0, S1g), sums_x = double[2][F] (S0, S1x); dwpsi[F], dbpsi[1] (accumulate optional)
__global__ void gate_bwd_sums_kernel(const float* __restrict__ partial, int nblk, int F, double* __restrict__ sums_g,
                                     double* __restrict__ sums_x, float* dwpsi, float* dbpsi, int accumulate) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 5 * F) return;
  const int k = i / F, c = i % F;
  if (k == 4 && c != 0) return;
  double t = 0.0;
  for (int b = 0; b < nblk; ++b) t += partial[static_cast<size_t>(b) * 5 * F + i];
  if (k == 0) { sums_g[c] = t; sums_x[c] = t; }
  else if (k == 1) sums_g[F + c] = t;
  else if (k == 2) sums_x[F + c] = t;
  else if (k == 3) { if (dwpsi) dwpsi[c] = accumulate ? dwpsi[c] + static_cast<float>(t) : static_cast<float>(t); }
  else { if (dbpsi) dbpsi[0] = accumulate ? dbpsi[0] + static_cast<float>(t) : static_cast<float>(t); }
}

// draw_g = scg*da + K1g*rawg + K0g,  draw_x = scx*da + K1x*rawx + K0x     (coef = [K0[F] | K1[F]])
__global__ void __launch_bounds__(kThreads)
gate_bwd_apply_kernel(const GateBwdArgs B, const float* __restrict__ coefg, const float* __restrict__ coefx,
                      __nv_bfloat16* __restrict__ drawg, int64_t drawg_ld, __nv_bfloat16* __restrict__ drawx,
                      int64_t drawx_ld) {
  pdl_trigger();
  pdl_wait();
  const GateArgs& A = B.G;
  GLanes L(A.F);
  if (!L.active) return;
  float scg[8], shg[8], scx[8], shx[8], wv[8], k0g[8], k1g[8], k0x[8], k1x[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = L.g * 8 + j;
    scg[j] = __ldg(A.scg + c); shg[j] = __ldg(A.shg + c); scx[j] = __ldg(A.scx + c); shx[j] = __ldg(A.shx + c);
    wv[j] = bf16_round(__ldg(A.wpsi + c));
    k0g[j] = __ldg(coefg + c); k1g[j] = __ldg(coefg + A.F + c);
    k0x[j] = __ldg(coefx + c); k1x[j] = __ldg(coefx + A.F + c);
  }
  const float a1 = __ldg(B.sc1), k0 = __ldg(B.coef1), k1 = __ldg(B.coef1 + 1);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * L.ppb;
  for (int64_t pix = static_cast<int64_t>(blockIdx.x) * L.ppb + L.pl; pix < A.npix; pix += 2 * stride) {
    const bool two = pix + stride < A.npix;
    const uint4 ug0 = ldg16(A.rawg + pix * A.rawg_ld + L.g * 8), ux0 = ldg16(A.rawx + pix * A.rawx_ld + L.g * 8);
    const uint4 ug1 = two ? ldg16(A.rawg + (pix + stride) * A.rawg_ld + L.g * 8) : make_uint4(0, 0, 0, 0);
    const uint4 ux1 = two ? ldg16(A.rawx + (pix + stride) * A.rawx_ld + L.g * 8) : make_uint4(0, 0, 0, 0);
    const float s0 = __ldg(B.s + pix), z0 = __ldg(B.dz + pix);
    const float s1 = two ? __ldg(B.s + pix + stride) : 0.f, z1 = two ? __ldg(B.dz + pix + stride) : 0.f;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (u == 1 && !two) break;
      const int64_t p = pix + u * stride;
      const float ds = fmaf(a1, u ? z1 : z0, fmaf(k1, u ? s1 : s0, k0));
      float rg[8], rx[8], a[8], og[8], ox[8];
      gate_act8(u ? ug1 : ug0, u ? ux1 : ux0, scg, shg, scx, shx, rg, rx, a);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float da = (a[j] > 0.f) ? ds * wv[j] : 0.f;
        og[j] = fmaf(scg[j], da, fmaf(k1g[j], rg[j], k0g[j]));
        ox[j] = fmaf(scx[j], da, fmaf(k1x[j], rx[j], k0x[j]));
      }
      stg16(drawg + p * drawg_ld + L.g * 8, pack8(og));
      stg16(drawx + p * drawx_ld + L.g * 8, pack8(ox));
    }
  }
}

bool pow2_f(int F) { return F == 8 || F == 16 || F == 32 || F == 64 || F == 128 || F == 256; }

int pix_grid(int64_t npix, int lpp, int per = 8) {
  const int gpb = kThreads / lpp;
  int64_t b = (npix + static_cast<int64_t>(gpb) * per - 1) / (static_cast<int64_t>(gpb) * per);
  const int64_t cap = static_cast<int64_t>(num_sms()) * 6;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}
int chan_grid(int64_t npix, int F) {
  int ppb = kThreads / (F / 8);
  if (ppb < 1) ppb = 1;
  int64_t b = (npix + static_cast<int64_t>(ppb) * 8 - 1) / (static_cast<int64_t>(ppb) * 8);
  const int64_t cap = static_cast<int64_t>(num_sms()) * 3;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

}  // namespace

size_t gate_partial_floats(int64_t npix, int F) {
  if (F < 8 || F % 8) return 0;
  const size_t a = static_cast<size_t>(num_sms()) * 6 * 2;
  const size_t b = static_cast<size_t>(chan_grid(npix, F)) * 5 * F;
  return a > b ? a : b;
}

static GateArgs gate_args(const void* rawg, int64_t rawg_ld, const void* rawx, int64_t rawx_ld, const float* scg,
                          const float* shg, const float* scx, const float* shx, const float* wpsi, const float* bpsi,
                          int64_t npix, int F) {
  return GateArgs{static_cast<const __nv_bfloat16*>(rawg), rawg_ld, static_cast<const __nv_bfloat16*>(rawx), rawx_ld,
                  scg, shg, scx, shx, wpsi, bpsi, npix, F};
}

int gate_fwd_run(const void* rawg, int64_t rawg_ld, const void* rawx, int64_t rawx_ld, const float* scg,
                 const float* shg, const float* scx, const float* shx, const float* wpsi, const float* bpsi, float* s,
                 float* partial, double* sums, int64_t npix, int F, cudaStream_t st) {
  UNETK_CHECK(pow2_f(F), -1, "gate: F_int=%d must be a power of two in [8,256]", F);
  const GateArgs A = gate_args(rawg, rawg_ld, rawx, rawx_ld, scg, shg, scx, shx, wpsi, bpsi, npix, F);
  const int grid = pix_grid(npix, F / 8);
  UNETK_CUDA(launch_pdl(gate_fwd_kernel, dim3(grid), dim3(kThreads), 0, st, A, s, partial));
  UNETK_LAUNCHED();
  UNETK_CUDA(launch_pdl(pair_sums_kernel, dim3(1), dim3(32), 0, st, partial, grid, sums));
  UNETK_LAUNCHED();
  return 0;
}

int gate_apply_run(const void* x, int64_t x_ld, const float* s, const float* sc1, const float* sh1, void* out,
                   int64_t out_ld, int64_t npix, int F, cudaStream_t st) {
  UNETK_CHECK(F % 8 == 0 && F >= 8 && (F <= 256 ? pow2_f(F) : F <= 4096), -1,
              "gate_apply: F_l=%d must be a power of two <= 256 or a multiple of 8 in (256, 4096]", F);
  const int lpp = F / 8 < 32 ? F / 8 : 32;
  UNETK_CUDA(launch_pdl(gate_apply_kernel, dim3(pix_grid(npix, lpp, 4)), dim3(kThreads), 0, st, static_cast<const __nv_bfloat16*>(x), x_ld, s, sc1,
                                                                sh1, static_cast<__nv_bfloat16*>(out), out_ld, npix, F));
  UNETK_LAUNCHED();
  return 0;
}

int gate_bwd_psi_run(const void* dout, int64_t dout_ld, const void* x, int64_t x_ld, const float* s, const float* sc1,
                     const float* sh1, const float* mean1, void* dx, int64_t dx_ld, int dx_accumulate, float* dz,
                     float* partial, double* sums, int64_t npix, int F, cudaStream_t st) {
  UNETK_CHECK(F % 8 == 0 && F >= 8 && (F <= 256 ? pow2_f(F) : F <= 4096), -1,
              "gate_bwd_psi: F_l=%d must be a power of two <= 256 or a multiple of 8 in (256, 4096]", F);
  const int lpp = F / 8 < 32 ? F / 8 : 32;
  const int grid = pix_grid(npix, lpp, 4);
  const __nv_bfloat16* d = static_cast<const __nv_bfloat16*>(dout);
  const __nv_bfloat16* xx = static_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* o = static_cast<__nv_bfloat16*>(dx);
  if (dx_accumulate)
    UNETK_CUDA(launch_pdl(gate_bwd_psi_kernel<true>, dim3(grid), dim3(kThreads), 0, st, d, dout_ld, xx, x_ld, s, sc1, sh1, mean1, o, dx_ld, dz, partial, npix, F));
  else
    UNETK_CUDA(launch_pdl(gate_bwd_psi_kernel<false>, dim3(grid), dim3(kThreads), 0, st, d, dout_ld, xx, x_ld, s, sc1, sh1, mean1, o, dx_ld, dz, partial, npix, F));
  UNETK_LAUNCHED();
  UNETK_CUDA(launch_pdl(pair_sums_kernel, dim3(1), dim3(32), 0, st, partial, grid, sums));
  UNETK_LAUNCHED();
  return 0;
}

int gate_bwd_reduce_run(const void* rawg, int64_t rawg_ld, const void* rawx, int64_t rawx_ld, const float* scg,
                        const float* shg, const float* mug, const float* scx, const float* shx, const float* mux,
                        const float* wpsi, const float* s, const float* dz, const float* sc1, const float* coef1,
                        float* partial, double* sums_g, double* sums_x, float* dwpsi, float* dbpsi, int accumulate,
                        int64_t npix, int F, cudaStream_t st) {
  UNETK_CHECK(pow2_f(F), -1, "gate: F_int=%d must be a power of two in [8,256]", F);
  GateBwdArgs B{gate_args(rawg, rawg_ld, rawx, rawx_ld, scg, shg, scx, shx, wpsi, nullptr, npix, F), mug, mux, s, dz,
                sc1, coef1};
  const int grid = chan_grid(npix, F);
  int ppb = kThreads / (F / 8);
  if (ppb < 1) ppb = 1;
  const size_t smem = static_cast<size_t>(ppb) * 5 * F * sizeof(float);
  UNETK_CUDA(launch_pdl(gate_bwd_reduce_kernel, dim3(grid), dim3(kThreads), smem, st, B, partial));
  UNETK_LAUNCHED();
  UNETK_CUDA(launch_pdl(gate_bwd_sums_kernel, dim3((5 * F + 127) / 128), dim3(128), 0, st, partial, grid, F, sums_g, sums_x, dwpsi, dbpsi, accumulate));
  UNETK_LAUNCHED();
  return 0;
}

int gate_bwd_apply_run(const void* rawg, int64_t rawg_ld, const void* rawx, int64_t rawx_ld, const float* scg,
                       const float* shg, const float* scx, const float* shx, const float* wpsi, const float* s,
                       const float* dz, const float* sc1, const float* coef1, const float* coefg, const float* coefx,
                       void* drawg, int64_t drawg_ld, void* drawx, int64_t drawx_ld, int64_t npix, int F,
                       cudaStream_t st) {
  UNETK_CHECK(pow2_f(F), -1, "gate: F_int=%d must be a power of two in [8,256]", F);
  GateBwdArgs B{gate_args(rawg, rawg_ld, rawx, rawx_ld, scg, shg, scx, shx, wpsi, nullptr, npix, F), nullptr, nullptr,
                s, dz, sc1, coef1};
  UNETK_CUDA(launch_pdl(gate_bwd_apply_kernel, dim3(chan_grid(npix, F)), dim3(kThreads), 0, st, B, coefg, coefx, static_cast<__nv_bfloat16*>(drawg),
                                                                drawg_ld, static_cast<__nv_bfloat16*>(drawx), drawx_ld));
  UNETK_LAUNCHED();
  return 0;
}

}  // namespace unetk
