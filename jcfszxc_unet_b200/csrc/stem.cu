// Stem convolution: the network-input conv (Cin <= 4, e.g. 3 -> 64) is flatly HBM-bound
// (27 FLOP/B) and its K = 27 cannot feed a 128B TMA row, so it runs as a direct
// CUDA-core kernel that reads the fp32 image in whatever strides the caller has (NCHW or channels_last),
// applies the same bf16 rounding autocast would, and writes NHWC bf16.  Its weight gradient is the matching
// direct reduction.  No input gradient is needed (the image is a leaf).
// Reference semantics replaced: the first nn.Conv2d of DoubleConv `inc` (UNetFamily/UNet.py:21,
// unet_parts.py:24) and its weight gradient.
#include "host_common.cuh"

#include <cstdlib>
#include "kernels.cuh"
#include "ptx.cuh"

namespace unetk {

namespace {

constexpr int kTH = 4, kTW = 32;  // 128 pixels per tile

__device__ __forceinline__ float bf16r(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

struct StemGeom {
  const float* x; int64_t sn, sc, sh, sw;  // fp32 image, element strides
  int N, H, W, Cin, Cout;
};

// stage the (kTH+2) x (kTW+2) x Cin halo of tile (n, h0, w0) into smem as bf16-rounded floats
__device__ __forceinline__ void load_halo(const StemGeom& G, int n, int h0, int w0, float* sx) {
  const int per = (kTH + 2) * (kTW + 2);
  for (int i = threadIdx.x; i < per * G.Cin; i += blockDim.x) {
    const int ci = i / per, r = (i % per) / (kTW + 2), c = i % (kTW + 2);
    const int h = h0 + r - 1, w = w0 + c - 1;
    float v = 0.f;
    if (h >= 0 && h < G.H && w >= 0 && w < G.W) v = bf16r(__ldg(G.x + n * G.sn + ci * G.sc + h * G.sh + w * G.sw));
    sx[(r * (kTW + 2) + c) * 4 + ci] = v;
  }
}

template <int COUT>
__global__ void __launch_bounds__(128) stem_fwd_kernel(const StemGeom G, const float* __restrict__ w,
                                                       const float* __restrict__ bias, __nv_bfloat16* __restrict__ y,
                                                       int64_t y_ld, int tiles_h, int tiles_w) {
  __shared__ float sx[(kTH + 2) * (kTW + 2) * 4];
  __shared__ __align__(16) float sw[36 * COUT];  // [k = (ci*3+r)*3+s][co], bf16-rounded
  for (int i = threadIdx.x; i < 9 * G.Cin * COUT; i += blockDim.x) {
    const int k = i / COUT, co = i % COUT;  // k = ci*9 + r*3 + s
    sw[i] = co < G.Cout ? bf16r(__ldg(w + static_cast<int64_t>(co) * G.Cin * 9 + k)) : 0.f;
  }
  const int num_tiles = G.N * tiles_h * tiles_w;
  const int py = threadIdx.x / kTW, px = threadIdx.x % kTW;
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int tw = tile % tiles_w, th = (tile / tiles_w) % tiles_h, n = tile / (tiles_w * tiles_h);
    const int h0 = th * kTH, w0 = tw * kTW;
    __syncthreads();
    load_halo(G, n, h0, w0, sx);
    __syncthreads();
    float acc[COUT];
#pragma unroll
    for (int c = 0; c < COUT; ++c) acc[c] = 0.f;
    for (int ci = 0; ci < G.Cin; ++ci) {
#pragma unroll
      for (int rs = 0; rs < 9; ++rs) {
        const float xv = sx[((py + rs / 3) * (kTW + 2) + px + rs % 3) * 4 + ci];
        const float4* wr = reinterpret_cast<const float4*>(sw + (ci * 9 + rs) * COUT);
#pragma unroll
        for (int c4 = 0; c4 < COUT / 4; ++c4) {
          const float4 wv = wr[c4];
          acc[c4 * 4 + 0] = fmaf(xv, wv.x, acc[c4 * 4 + 0]);
          acc[c4 * 4 + 1] = fmaf(xv, wv.y, acc[c4 * 4 + 1]);
          acc[c4 * 4 + 2] = fmaf(xv, wv.z, acc[c4 * 4 + 2]);
          acc[c4 * 4 + 3] = fmaf(xv, wv.w, acc[c4 * 4 + 3]);
        }
      }
    }
    const int h = h0 + py, wq = w0 + px;
    if (h < G.H && wq < G.W) {
      __nv_bfloat16* o = y + ((static_cast<int64_t>(n) * G.H + h) * G.W + wq) * y_ld;
#pragma unroll
      for (int c8 = 0; c8 < COUT / 8; ++c8) {
        if (c8 * 8 < G.Cout) {
          float f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = acc[c8 * 8 + j] + (bias ? __ldg(bias + c8 * 8 + j) : 0.f);
          uint4 u;
          u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
          u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
          *reinterpret_cast<uint4*>(o + c8 * 8) = u;
        }
      }
    }
  }
}

// partial[blk][co][k] = sum over the block's tiles of dy[p][co] * x[p + off(k)][ci(k)],  k = ci*9 + r*3 + s
// One warp per pixel sub-lane: lane = (ci 0..3) x (8 groups of 8 output channels); every thread keeps an
// 8 (co) x 9 (tap) register tile, so a pixel costs 1 LDS.128 + 9 broadcast LDS for 72 FMAs.
__global__ void __launch_bounds__(256) stem_wgrad_kernel(const StemGeom G, const __nv_bfloat16* __restrict__ dy,
                                                         int64_t dy_ld, float* __restrict__ partial, int tiles_h,
                                                         int tiles_w) {
  __shared__ float sx[(kTH + 2) * (kTW + 2) * 4];
  __shared__ __align__(16) __nv_bfloat16 sg[kTH * kTW][64];
  __shared__ float sacc[64 * 36];
  const int K = 9 * G.Cin;
  const int warp = threadIdx.x >> 5, lid = threadIdx.x & 31;
  const int cog = lid & 7, ci = lid >> 3;
  const bool live = ci < G.Cin;
  float acc[8][9];
#pragma unroll
  for (int c = 0; c < 8; ++c)
#pragma unroll
    for (int i = 0; i < 9; ++i) acc[c][i] = 0.f;
  const int num_tiles = G.N * tiles_h * tiles_w;
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int tw = tile % tiles_w, th = (tile / tiles_w) % tiles_h, n = tile / (tiles_w * tiles_h);
    const int h0 = th * kTH, w0 = tw * kTW;
    __syncthreads();
    load_halo(G, n, h0, w0, sx);
    for (int i = threadIdx.x; i < kTH * kTW * 8; i += blockDim.x) {
      const int p = i >> 3, c8 = i & 7;
      const int h = h0 + p / kTW, w = w0 + p % kTW;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (c8 * 8 < G.Cout && h < G.H && w < G.W)
        v = __ldg(reinterpret_cast<const uint4*>(dy + ((static_cast<int64_t>(n) * G.H + h) * G.W + w) * dy_ld + c8 * 8));
      *reinterpret_cast<uint4*>(&sg[p][c8 * 8]) = v;
    }
    __syncthreads();
    if (live) {
#pragma unroll 2
      for (int p = warp; p < kTH * kTW; p += 8) {
        const uint4 gu = *reinterpret_cast<const uint4*>(&sg[p][cog * 8]);
        const float g[8] = {bf16_lo(gu.x), bf16_hi(gu.x), bf16_lo(gu.y), bf16_hi(gu.y),
                            bf16_lo(gu.z), bf16_hi(gu.z), bf16_lo(gu.w), bf16_hi(gu.w)};
        const float* xr = sx + ((p / kTW) * (kTW + 2) + (p % kTW)) * 4 + ci;
#pragma unroll
        for (int rs = 0; rs < 9; ++rs) {
          const float xv = xr[((rs / 3) * (kTW + 2) + rs % 3) * 4];
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[c][rs] = fmaf(g[c], xv, acc[c][rs]);
        }
      }
    }
  }
  // ordered (deterministic) reduction over the 8 warps of the block
  for (int i = threadIdx.x; i < 64 * 36; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  for (int w = 0; w < 8; ++w) {
    if (warp == w && live) {
#pragma unroll
      for (int c = 0; c < 8; ++c)
#pragma unroll
        for (int rs = 0; rs < 9; ++rs) sacc[(cog * 8 + c) * K + ci * 9 + rs] += acc[c][rs];
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < 64 * K; i += blockDim.x) partial[static_cast<size_t>(blockIdx.x) * 64 * K + i] = sacc[i];
}

__global__ void stem_wgrad_reduce_kernel(const float* __restrict__ partial, int nblk, int Cout, int K,
                                         float* __restrict__ dw, int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Cout * K) return;
  const int co = i / K, k = i % K;
  double s = 0.0;
  for (int b = 0; b < nblk; ++b) s += partial[(static_cast<size_t>(b) * 64 + co) * K + k];
  dw[i] = accumulate ? dw[i] + static_cast<float>(s) : static_cast<float>(s);
}

}  // namespace

int stem_grid(int N, int H, int W) {
  const int tiles = N * ((H + kTH - 1) / kTH) * ((W + kTW - 1) / kTW);
  const int cap = 2 * num_sms();
  return tiles < cap ? tiles : cap;
}

// stem_tc.cu: the tensor-core versions (default); UNETK_STEM_TC=0 selects the CUDA-core kernels of this file
bool stem_tc_ok(int Cin, int Cout);
int stem_tc_fwd_run(const float* x, int64_t sn, int64_t sc, int64_t sh, int64_t sw, const float* w, const float* bias,
                    void* y, int64_t y_ld, int N, int H, int W, int Cin, int Cout, cudaStream_t s,
                    float* stats_partial = nullptr, double* stats_sums = nullptr, const float* scale = nullptr, int relu = 0);
bool stem_tc_stats_ok(int Cin, int Cout);
size_t stem_tc_stats_partial_floats(int N, int H, int W, int Cout);
size_t stem_tc_wgrad_workspace(int N, int H, int W, int Cin, int Cout);
int stem_tc_wgrad_run(const float* x, int64_t sn, int64_t sc, int64_t sh, int64_t sw, const void* dy, int64_t dy_ld,
                      float* dw, int accumulate, int N, int H, int W, int Cin, int Cout, void* ws, size_t ws_bytes,
                      cudaStream_t s);
static bool use_tc(int Cin, int Cout, int max_cout) {
  static int env = -1;
  if (env < 0) { const char* e = getenv("UNETK_STEM_TC"); env = e ? atoi(e) : 1; }
  return env && stem_tc_ok(Cin, Cout) && Cout <= max_cout;
}

int stem_fwd_affine_run(const float* x, int64_t sn, int64_t sc, int64_t sh, int64_t sw, const float* w, const float* scale,
                        const float* shift, int relu, void* y, int64_t y_ld, int N, int H, int W, int Cin, int Cout,
                        cudaStream_t s) {
  UNETK_CHECK(stem_tc_ok(Cin, Cout) && Cout <= 256, -1, "stem (affine): Cin=%d Cout=%d not supported", Cin, Cout);
  return stem_tc_fwd_run(x, sn, sc, sh, sw, w, shift, y, y_ld, N, H, W, Cin, Cout, s, nullptr, nullptr, scale, relu);
}

int stem_fwd_run(const float* x, int64_t sn, int64_t sc, int64_t sh, int64_t sw, const float* w, const float* bias,
                 void* y, int64_t y_ld, int N, int H, int W, int Cin, int Cout, cudaStream_t s) {
  if (use_tc(Cin, Cout, 256)) return stem_tc_fwd_run(x, sn, sc, sh, sw, w, bias, y, y_ld, N, H, W, Cin, Cout, s);
  UNETK_CHECK(Cin >= 1 && Cin <= 4, -1, "stem: Cin=%d must be <= 4", Cin);
  UNETK_CHECK(Cout % 8 == 0 && Cout <= 64, -1, "stem: Cout=%d must be a multiple of 8 and <= 64", Cout);
  StemGeom G{x, sn, sc, sh, sw, N, H, W, Cin, Cout};
  const int th = (H + kTH - 1) / kTH, tw = (W + kTW - 1) / kTW;
  const int grid = N * th * tw < 8 * num_sms() ? N * th * tw : 8 * num_sms();
  if (Cout > 32)
    stem_fwd_kernel<64><<<grid, 128, 0, s>>>(G, w, bias, static_cast<__nv_bfloat16*>(y), y_ld, th, tw);
  else
    stem_fwd_kernel<32><<<grid, 128, 0, s>>>(G, w, bias, static_cast<__nv_bfloat16*>(y), y_ld, th, tw);
  UNETK_LAUNCHED();
  return 0;
}

// Stem conv + the BatchNorm statistics of its (bf16) output: in the tensor-core kernel's epilogue when it applies,
// else the conv followed by the statistics pass (same sums).
size_t stem_stats_partial_floats(int N, int H, int W, int Cout) {
  const size_t a = chan_partial_floats(static_cast<int64_t>(N) * H * W, Cout);
  const size_t b = (Cout % 64 == 0 && Cout <= 256) ? stem_tc_stats_partial_floats(N, H, W, Cout) : 0;
  return a > b ? a : b;
}
int stem_fwd_stats_run(const float* x, int64_t sn, int64_t sc, int64_t sh, int64_t sw, const float* w, const float* bias,
                       void* y, int64_t y_ld, float* partial, double* sums, int N, int H, int W, int Cin, int Cout,
                       cudaStream_t s) {
  static int fuse = -1;
  if (fuse < 0) { const char* e = getenv("UNETK_STEM_STATS"); fuse = e ? atoi(e) : 1; }
  if (fuse && use_tc(Cin, Cout, 256) && stem_tc_stats_ok(Cin, Cout))
    return stem_tc_fwd_run(x, sn, sc, sh, sw, w, bias, y, y_ld, N, H, W, Cin, Cout, s, partial, sums);
  if (int rc = stem_fwd_run(x, sn, sc, sh, sw, w, bias, y, y_ld, N, H, W, Cin, Cout, s)) return rc;
  return bn_stats_run(y, y_ld, static_cast<int64_t>(N) * H * W, Cout, partial, sums, s);
}

size_t stem_wgrad_workspace(int N, int H, int W, int Cin) {
  const size_t a = static_cast<size_t>(stem_grid(N, H, W)) * 64 * 9 * Cin * sizeof(float);
  const size_t b = stem_tc_wgrad_workspace(N, H, W, Cin, 128);   // the query does not know Cout: size for the largest
  return a > b ? a : b;
}

int stem_wgrad_run(const float* x, int64_t sn, int64_t sc, int64_t sh, int64_t sw, const void* dy, int64_t dy_ld,
                   float* dw, int accumulate, int N, int H, int W, int Cin, int Cout, void* ws, size_t ws_bytes,
                   cudaStream_t s) {
  if (use_tc(Cin, Cout, 128))
    return stem_tc_wgrad_run(x, sn, sc, sh, sw, dy, dy_ld, dw, accumulate, N, H, W, Cin, Cout, ws, ws_bytes, s);
  UNETK_CHECK(Cin >= 1 && Cin <= 4, -1, "stem: Cin=%d must be <= 4", Cin);
  UNETK_CHECK(Cout % 8 == 0 && Cout <= 64, -1, "stem: Cout=%d must be a multiple of 8 and <= 64", Cout);
  UNETK_CHECK(ws != nullptr && ws_bytes >= stem_wgrad_workspace(N, H, W, Cin), -1, "stem_wgrad: workspace too small");
  StemGeom G{x, sn, sc, sh, sw, N, H, W, Cin, Cout};
  const int th = (H + kTH - 1) / kTH, tw = (W + kTW - 1) / kTW;
  const int grid = stem_grid(N, H, W);
  stem_wgrad_kernel<<<grid, 256, 0, s>>>(G, static_cast<const __nv_bfloat16*>(dy), dy_ld, static_cast<float*>(ws), th,
                                         tw);
  UNETK_LAUNCHED();
  const int K = 9 * Cin;
  stem_wgrad_reduce_kernel<<<(Cout * K + 127) / 128, 128, 0, s>>>(static_cast<const float*>(ws), grid, Cout, K, dw,
                                                                accumulate);
  UNETK_LAUNCHED();
  return 0;
}

}  // namespace unetk
