// SegNet-style pool / unpool pair (reference UNetFamily/SegNet.py:89-138):
//   x_p, id = F.max_pool2d(x, 2, 2, return_indices=True)   ...   y = F.max_unpool2d(z, id, 2, 2)
// HBM-bound streaming kernels over NHWC bf16 (8 channels = one 16-byte vector per thread).  The arg-max travels as a
// compact CODE — one byte per pooled element holding the window position 0..3 (row-major), NHWC like the values, so
// a thread reads its 8 codes with one 8-byte load: 1 B per element instead of the 8 B of ATen's int64 indices.  The
// int64 form (NCHW-logical [N,C,Ho,Wo], h*W+w of the selected input element) is accepted as well: it is what
// unetk_maxpool2x2_fwd emits for the bit-exact index test and what a caller holding ATen indices passes.
// Windows of a 2x2 / stride-2 pool tile the plane, so the unpool WRITES every output element (the selected position
// gets the value, the other three get zero): no zero-fill pass, no scatter, no atomics.  Indices must come from a
// 2x2 / stride-2 max-pool (each inside its own window), which is the only way SegNet produces them.
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

struct Unit {   // one pooled pixel x one group of 8 channels
  int g, wo, ho, n;
  __device__ Unit(int64_t i, int cg, int Wo, int Ho) {
    g = static_cast<int>(i % cg);
    int64_t t = i / cg;
    wo = static_cast<int>(t % Wo); t /= Wo;
    ho = static_cast<int>(t % Ho);
    n = static_cast<int>(t / Ho);
  }
};

// window position of each of the 8 channels: from the byte codes (one 8-byte load) or from int64 indices
template <bool IDX>
__device__ __forceinline__ void load_codes(const void* __restrict__ where, const Unit& u, int64_t opix, int C, int Ho,
                                           int Wo, int* q) {
  if constexpr (IDX) {
    const long long* idx = static_cast<const long long*>(where);
    const int W = 2 * Wo;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const long long v = __ldg(idx + ((static_cast<int64_t>(u.n) * C + u.g * 8 + j) * Ho + u.ho) * Wo + u.wo);
      const int h = static_cast<int>(v / W), w = static_cast<int>(v - static_cast<long long>(h) * W);
      q[j] = ((h & 1) << 1) | (w & 1);
    }
  } else {
    const uint2 c = __ldg(reinterpret_cast<const uint2*>(static_cast<const uint8_t*>(where) + opix * C + u.g * 8));
#pragma unroll
    for (int j = 0; j < 4; ++j) { q[j] = (c.x >> (8 * j)) & 3; q[4 + j] = (c.y >> (8 * j)) & 3; }
  }
}

// MaxPool2d(2) forward that also emits the byte codes (first maximum in row-major window order, NaN always taken:
// the rule of ATen's max_pool2d_with_indices, same as unetk_maxpool2x2_fwd).
__global__ void __launch_bounds__(kThreads)
maxpool_codes_kernel(const __nv_bfloat16* __restrict__ x, int64_t x_ld, __nv_bfloat16* __restrict__ y, int64_t y_ld,
                     uint8_t* __restrict__ code, int N, int H, int W, int C) {
  pdl_trigger();
  pdl_wait();
  const int cg = C >> 3, Ho = H >> 1, Wo = W >> 1;
  const int64_t total = static_cast<int64_t>(N) * Ho * Wo * cg;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(kThreads) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * kThreads) {
    const Unit u(i, cg, Wo, Ho);
    uint4 v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int64_t pix = (static_cast<int64_t>(u.n) * H + 2 * u.ho + (q >> 1)) * W + 2 * u.wo + (q & 1);
      v[q] = __ldg(reinterpret_cast<const uint4*>(x + pix * x_ld + u.g * 8));
    }
    float a[4][8];
#pragma unroll
    for (int q = 0; q < 4; ++q) unpack8(v[q], a[q]);
    float best[8];
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float b = a[0][j];
      int k = 0;
#pragma unroll
      for (int q = 1; q < 4; ++q) {
        const float t = a[q][j];
        if (t > b || t != t) { b = t; k = q; }
      }
      best[j] = b;
      if (j < 4) lo |= static_cast<uint32_t>(k) << (8 * j); else hi |= static_cast<uint32_t>(k) << (8 * (j - 4));
    }
    const int64_t opix = (static_cast<int64_t>(u.n) * Ho + u.ho) * Wo + u.wo;
    uint4 o;
    o.x = pack_bf16x2(best[0], best[1]); o.y = pack_bf16x2(best[2], best[3]);
    o.z = pack_bf16x2(best[4], best[5]); o.w = pack_bf16x2(best[6], best[7]);
    *reinterpret_cast<uint4*>(y + opix * y_ld + u.g * 8) = o;
    *reinterpret_cast<uint2*>(code + opix * C + u.g * 8) = make_uint2(lo, hi);
  }
}

// out[window position q] = (q == code) ? x : 0   (F.max_unpool2d(x, idx, 2, 2), SegNet.py:115-138)
template <bool IDX>
__global__ void __launch_bounds__(kThreads)
max_unpool_kernel(const __nv_bfloat16* __restrict__ x, int64_t x_ld, const void* __restrict__ where,
                  __nv_bfloat16* __restrict__ out, int64_t out_ld, int N, int Ho, int Wo, int C) {
  pdl_trigger();
  pdl_wait();
  const int cg = C >> 3, W = 2 * Wo, H = 2 * Ho;
  const int64_t total = static_cast<int64_t>(N) * Ho * Wo * cg;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(kThreads) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * kThreads) {
    const Unit u(i, cg, Wo, Ho);
    const int64_t opix = (static_cast<int64_t>(u.n) * Ho + u.ho) * Wo + u.wo;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + opix * x_ld + u.g * 8));
    int q[8];
    load_codes<IDX>(where, u, opix, C, Ho, Wo, q);
    const uint32_t w32[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      uint32_t o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)   // two bf16 per word: keep a half where its channel selected this position (bit moves only)
        o[k] = (q[2 * k] == p ? (w32[k] & 0xffffu) : 0u) | (q[2 * k + 1] == p ? (w32[k] & 0xffff0000u) : 0u);
      const int64_t pix = (static_cast<int64_t>(u.n) * H + 2 * u.ho + (p >> 1)) * W + 2 * u.wo + (p & 1);
      *reinterpret_cast<uint4*>(out + pix * out_ld + u.g * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// backward of the unpool = gather: dx[pooled] (+)= dy[selected position]
template <bool IDX, bool ACC>
__global__ void __launch_bounds__(kThreads)
max_unpool_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int64_t dy_ld, const void* __restrict__ where,
                      __nv_bfloat16* __restrict__ dx, int64_t dx_ld, int N, int Ho, int Wo, int C) {
  pdl_trigger();
  pdl_wait();
  const int cg = C >> 3, W = 2 * Wo, H = 2 * Ho;
  const int64_t total = static_cast<int64_t>(N) * Ho * Wo * cg;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(kThreads) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * kThreads) {
    const Unit u(i, cg, Wo, Ho);
    const int64_t opix = (static_cast<int64_t>(u.n) * Ho + u.ho) * Wo + u.wo;
    int q[8];
    load_codes<IDX>(where, u, opix, C, Ho, Wo, q);
    uint32_t o[4] = {0, 0, 0, 0};
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int64_t pix = (static_cast<int64_t>(u.n) * H + 2 * u.ho + (p >> 1)) * W + 2 * u.wo + (p & 1);
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(dy + pix * dy_ld + u.g * 8));
      const uint32_t w32[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k)
        o[k] |= (q[2 * k] == p ? (w32[k] & 0xffffu) : 0u) | (q[2 * k + 1] == p ? (w32[k] & 0xffff0000u) : 0u);
    }
    __nv_bfloat16* dst = dx + opix * dx_ld + u.g * 8;
    if constexpr (ACC) {
      float a[8], b[8];
      unpack8(*reinterpret_cast<const uint4*>(dst), a);
      unpack8(make_uint4(o[0], o[1], o[2], o[3]), b);
      uint4 r;
      r.x = pack_bf16x2(a[0] + b[0], a[1] + b[1]); r.y = pack_bf16x2(a[2] + b[2], a[3] + b[3]);
      r.z = pack_bf16x2(a[4] + b[4], a[5] + b[5]); r.w = pack_bf16x2(a[6] + b[6], a[7] + b[7]);
      *reinterpret_cast<uint4*>(dst) = r;
    } else {
      *reinterpret_cast<uint4*>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

int grid_for(int64_t total) {
  int64_t b = (total + kThreads - 1) / kThreads;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 16;   // a multiple of the SM count, 16 resident CTAs' worth
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

}  // namespace

int maxpool_codes_run(const void* x, int64_t x_ld, void* y, int64_t y_ld, uint8_t* code, int N, int H, int W, int C,
                      cudaStream_t s) {
  UNETK_CHECK(C > 0 && C % 8 == 0, -1, "maxpool_codes: C=%d must be a multiple of 8", C);
  UNETK_CHECK(H % 2 == 0 && W % 2 == 0, -1, "maxpool_codes: odd spatial size %dx%d", H, W);
  UNETK_CHECK((reinterpret_cast<uintptr_t>(code) & 7) == 0, -1, "maxpool_codes: code buffer must be 8-byte aligned");
  const int64_t total = static_cast<int64_t>(N) * (H / 2) * (W / 2) * (C / 8);
  if (total == 0) return 0;
  UNETK_CUDA(launch_pdl(maxpool_codes_kernel, dim3(grid_for(total)), dim3(kThreads), 0, s, static_cast<const __nv_bfloat16*>(x), x_ld,
                        static_cast<__nv_bfloat16*>(y), y_ld, code, N, H, W, C));
  UNETK_LAUNCHED();
  return 0;
}

int max_unpool_run(const void* x, int64_t x_ld, const void* where, int is_idx, void* out, int64_t out_ld, int N, int Ho,
                   int Wo, int C, cudaStream_t s) {
  UNETK_CHECK(C > 0 && C % 8 == 0, -1, "max_unpool: C=%d must be a multiple of 8", C);
  UNETK_CHECK((reinterpret_cast<uintptr_t>(where) & 7) == 0, -1, "max_unpool: index buffer must be 8-byte aligned");
  const int64_t total = static_cast<int64_t>(N) * Ho * Wo * (C / 8);
  if (total == 0) return 0;
  const __nv_bfloat16* xb = static_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* ob = static_cast<__nv_bfloat16*>(out);
  if (is_idx) UNETK_CUDA(launch_pdl(max_unpool_kernel<true>, dim3(grid_for(total)), dim3(kThreads), 0, s, xb, x_ld, where, ob, out_ld, N, Ho, Wo, C));
  else UNETK_CUDA(launch_pdl(max_unpool_kernel<false>, dim3(grid_for(total)), dim3(kThreads), 0, s, xb, x_ld, where, ob, out_ld, N, Ho, Wo, C));
  UNETK_LAUNCHED();
  return 0;
}

int max_unpool_bwd_run(const void* dy, int64_t dy_ld, const void* where, int is_idx, void* dx, int64_t dx_ld, int accumulate,
                       int N, int Ho, int Wo, int C, cudaStream_t s) {
  UNETK_CHECK(C > 0 && C % 8 == 0, -1, "max_unpool_bwd: C=%d must be a multiple of 8", C);
  UNETK_CHECK((reinterpret_cast<uintptr_t>(where) & 7) == 0, -1, "max_unpool_bwd: index buffer must be 8-byte aligned");
  const int64_t total = static_cast<int64_t>(N) * Ho * Wo * (C / 8);
  if (total == 0) return 0;
  const __nv_bfloat16* g = static_cast<const __nv_bfloat16*>(dy);
  __nv_bfloat16* d = static_cast<__nv_bfloat16*>(dx);
  const dim3 grid(grid_for(total)), block(kThreads);
  if (is_idx) {
    if (accumulate) UNETK_CUDA(launch_pdl(max_unpool_bwd_kernel<true, true>, grid, block, 0, s, g, dy_ld, where, d, dx_ld, N, Ho, Wo, C));
    else UNETK_CUDA(launch_pdl(max_unpool_bwd_kernel<true, false>, grid, block, 0, s, g, dy_ld, where, d, dx_ld, N, Ho, Wo, C));
  } else {
    if (accumulate) UNETK_CUDA(launch_pdl(max_unpool_bwd_kernel<false, true>, grid, block, 0, s, g, dy_ld, where, d, dx_ld, N, Ho, Wo, C));
    else UNETK_CUDA(launch_pdl(max_unpool_bwd_kernel<false, false>, grid, block, 0, s, g, dy_ld, where, d, dx_ld, N, Ho, Wo, C));
  }
  UNETK_LAUNCHED();
  return 0;
}

}  // namespace unetk
