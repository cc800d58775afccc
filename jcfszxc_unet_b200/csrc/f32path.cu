// fp32-accuracy forward ("fp32 mode", BASELINE.json configs[0]: vanilla UNet fp32 forward, logits within 1e-4 of the
// reference's fp32 CPU forward) on the SAME bf16 tcgen05 tap-GEMM.
//
// An fp32 value x is carried as three bf16 terms x = hi + mid + lo (8 + 8 + 8 mantissa bits); a product
// x*w is evaluated as the six terms  hi*hi + hi*mid + hi*lo + mid*hi + mid*mid + lo*hi  (error ~2^-24 |x w|), each an
// EXACT bf16 x bf16 product accumulated in fp32 in TMEM.  The six terms are laid out along the reduction axis:
//     activation "split tensor"  [N,H,W, 6C] = [ hi | hi | hi | mid | mid | lo ]      (bf16 NHWC, written here)
//     weight     "split pack"    [T][R][6K]  = [ hi | mid | lo | hi | mid | hi ]      (bf16, written here)
// so that one ordinary bf16 conv over 6*Cin channels IS the fp32-accurate conv; its epilogue writes fp32
// (conv_gemm.cu, F32OUT).  BatchNorm / ReLU / max-pool / the 1x1 head run in fp32 on CUDA cores in this file.
// Reference semantics: UNetFamily/UNet.py:39-55 executed by torch in fp32 (no autocast), unet_parts.py:17-79.
#include "host_common.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

namespace unetk {

namespace {

__device__ __forceinline__ void split3(float x, __nv_bfloat16& hi, __nv_bfloat16& mid, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  const float r1 = x - __bfloat162float(hi);       // exact
  mid = __float2bfloat16_rn(r1);
  const float r2 = r1 - __bfloat162float(mid);     // exact
  lo = __float2bfloat16_rn(r2);
}

// ---------------------------------------------------------------------------------------------- weights
// dst[t][r][6K] from the fp32 master src[r*sr + k*sk + t*st].  The K axis is a concat of `ns` slices (the channel
// slices of the consumer's input buffer); slice s of width Ks occupies dst columns [6*k0, 6*(k0+Ks)) as six planes.
struct SplitPackArgs {
  const float* src;
  __nv_bfloat16* dst;
  long long sr, sk, st;
  int R, K, T, ns;
  int k0[5];  // slice starts, k0[ns] = K
};
__global__ void pack_split3_kernel(const SplitPackArgs a) {
  const long long total = static_cast<long long>(a.T) * a.R * a.K;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(i % a.K);
    const int r = static_cast<int>((i / a.K) % a.R);
    const int t = static_cast<int>(i / (static_cast<long long>(a.K) * a.R));
    int s = 0;
    while (s + 1 < a.ns && k >= a.k0[s + 1]) ++s;
    const int ks = a.k0[s + 1] - a.k0[s], kk = k - a.k0[s];
    __nv_bfloat16 hi, mid, lo;
    split3(a.src[r * a.sr + k * a.sk + t * a.st], hi, mid, lo);
    __nv_bfloat16* row = a.dst + (static_cast<long long>(t) * a.R + r) * 6 * a.K + 6 * a.k0[s] + kk;
    row[0 * ks] = hi;   // pairs with the activation planes [hi, hi, hi, mid, mid, lo]
    row[1 * ks] = mid;
    row[2 * ks] = lo;
    row[3 * ks] = hi;
    row[4 * ks] = mid;
    row[5 * ks] = hi;
  }
}

// ---------------------------------------------------------------------------------------------- stem (Cin <= 4)
// y[n,h,w,co] = bias[co] + sum x[n,ci,h+r-1,w+s-1] * w[co,ci,r,s]   in fp32 FMAs; thread = (pixel, 8 output channels)
__global__ void stem_f32_kernel(const float* __restrict__ x, long long sn, long long sc, long long sh, long long sw,
                                const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ y,
                                long long y_ld, int N, int H, int W, int Cin, int Cout) {
  extern __shared__ float ws[];  // [Cin*9][Cout]
  for (int i = threadIdx.x; i < Cout * Cin * 9; i += blockDim.x) {
    const int co = i / (Cin * 9), rest = i % (Cin * 9);
    ws[rest * Cout + co] = w[i];
  }
  __syncthreads();
  const int groups = Cout / 8;
  const long long total = static_cast<long long>(N) * H * W * groups;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(i % groups);
    const long long pix = i / groups;
    const int pw = static_cast<int>(pix % W), ph = static_cast<int>((pix / W) % H), n = static_cast<int>(pix / (static_cast<long long>(W) * H));
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = bias ? bias[g * 8 + j] : 0.f;
    for (int ci = 0; ci < Cin; ++ci) {
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int hh = ph + r - 1;
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const int wwp = pw + s - 1;
          const bool ok = hh >= 0 && hh < H && wwp >= 0 && wwp < W;
          const float v = ok ? x[n * sn + ci * sc + hh * sh + wwp * sw] : 0.f;
          const float* wp = ws + ((ci * 3 + r) * 3 + s) * Cout + g * 8;
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(v, wp[j], acc[j]);
        }
      }
    }
    float* o = y + pix * y_ld + g * 8;
    *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    *reinterpret_cast<float4*>(o + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
}

// ---------------------------------------------------------------------------------------------- statistics
// per-channel (sum, sum of squares) of an fp32 NHWC tensor in double; block partials + ordered second stage
constexpr int kStatThreads = 256;
__global__ void f32_stats_kernel(const float* __restrict__ x, long long ld, long long npix, int C,
                                 double* __restrict__ partial) {
  const int lanes = C / 4;                   // threads across the channel axis (float4 each)
  const int rows = kStatThreads / lanes;     // pixels per block step
  const int lane = threadIdx.x % lanes, rw = threadIdx.x / lanes;
  double s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
  if (rw < rows) {
    for (long long p = static_cast<long long>(blockIdx.x) * rows + rw; p < npix; p += static_cast<long long>(gridDim.x) * rows) {
      const float4 v = *reinterpret_cast<const float4*>(x + p * ld + lane * 4);
      s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
      q[0] += static_cast<double>(v.x) * v.x; q[1] += static_cast<double>(v.y) * v.y;
      q[2] += static_cast<double>(v.z) * v.z; q[3] += static_cast<double>(v.w) * v.w;
    }
  }
  extern __shared__ double red[];  // [rows][2][C]
  if (rw < rows) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      red[(rw * 2 + 0) * C + lane * 4 + j] = s[j];
      red[(rw * 2 + 1) * C + lane * 4 + j] = q[j];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    double t = 0;
    for (int r = 0; r < rows; ++r) t += red[r * 2 * C + i];
    partial[static_cast<long long>(blockIdx.x) * 2 * C + i] = t;
  }
}
__global__ void f32_stats_final_kernel(const double* __restrict__ partial, int nblk, int C, double* __restrict__ sums) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * C) return;
  double t = 0;
  for (int b = 0; b < nblk; ++b) t += partial[static_cast<long long>(b) * 2 * C + i];
  sums[i] = t;
}

// ---------------------------------------------------------------------------------------------- BN + ReLU (+pool) + split
struct BnSplitArgs {
  const float* raw; long long raw_ld;
  const float* scale; const float* shift;   // null: identity (ConvTranspose output, bias already added)
  __nv_bfloat16* split; long long split_ld;  // 6C-wide slice or null
  float* out_f32; long long out_ld;          // C-wide fp32 copy or null
  __nv_bfloat16* pooled; long long pooled_ld;  // 6C-wide slice of the 2x2 max-pooled activation or null
  int N, H, W, C, relu;
};
__device__ __forceinline__ void store_split4(__nv_bfloat16* dst, int C, const float (&v)[4]) {
  __nv_bfloat16 hi[4], mid[4], lo[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) split3(v[j], hi[j], mid[j], lo[j]);
  const uint2 H2 = *reinterpret_cast<const uint2*>(hi), M2 = *reinterpret_cast<const uint2*>(mid),
              L2 = *reinterpret_cast<const uint2*>(lo);
  *reinterpret_cast<uint2*>(dst + 0 * C) = H2;
  *reinterpret_cast<uint2*>(dst + 1 * C) = H2;
  *reinterpret_cast<uint2*>(dst + 2 * C) = H2;
  *reinterpret_cast<uint2*>(dst + 3 * C) = M2;
  *reinterpret_cast<uint2*>(dst + 4 * C) = M2;
  *reinterpret_cast<uint2*>(dst + 5 * C) = L2;
}
__global__ void f32_bn_split_kernel(const BnSplitArgs a) {
  const int lanes = a.C / 4;
  const bool pool = a.pooled != nullptr;
  const int UH = pool ? a.H / 2 : a.H, UW = pool ? a.W / 2 : a.W;   // work units: pixels, or 2x2 windows
  const long long total = static_cast<long long>(a.N) * UH * UW * lanes;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int lane = static_cast<int>(i % lanes);
    const long long u = i / lanes;
    const int uw = static_cast<int>(u % UW), uh = static_cast<int>((u / UW) % UH), n = static_cast<int>(u / (static_cast<long long>(UW) * UH));
    float sc[4] = {1.f, 1.f, 1.f, 1.f}, sh[4] = {0.f, 0.f, 0.f, 0.f};
    if (a.scale != nullptr) {
      const float4 s4 = *reinterpret_cast<const float4*>(a.scale + lane * 4), h4 = *reinterpret_cast<const float4*>(a.shift + lane * 4);
      sc[0] = s4.x; sc[1] = s4.y; sc[2] = s4.z; sc[3] = s4.w;
      sh[0] = h4.x; sh[1] = h4.y; sh[2] = h4.z; sh[3] = h4.w;
    }
    const int np = pool ? 4 : 1;
    float best[4];
    for (int q = 0; q < np; ++q) {
      const int ph = pool ? 2 * uh + (q >> 1) : uh, pw = pool ? 2 * uw + (q & 1) : uw;
      const long long pix = (static_cast<long long>(n) * a.H + ph) * a.W + pw;
      const float4 r4 = *reinterpret_cast<const float4*>(a.raw + pix * a.raw_ld + lane * 4);
      float v[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (a.scale != nullptr) v[j] = fmaf(v[j], sc[j], sh[j]);
        if (a.relu) v[j] = (v[j] != v[j]) ? v[j] : fmaxf(v[j], 0.f);   // ReLU keeps NaN like torch
        if (q == 0 || v[j] > best[j] || v[j] != v[j]) best[j] = v[j];   // first max wins, NaN propagates (ATen)
      }
      if (a.split != nullptr) store_split4(a.split + pix * a.split_ld + lane * 4, a.C, v);
      if (a.out_f32 != nullptr) *reinterpret_cast<float4*>(a.out_f32 + pix * a.out_ld + lane * 4) = make_float4(v[0], v[1], v[2], v[3]);
    }
    if (pool) {
      const long long ppix = (static_cast<long long>(n) * UH + uh) * UW + uw;
      store_split4(a.pooled + ppix * a.pooled_ld + lane * 4, a.C, best);
    }
  }
}

// ---------------------------------------------------------------------------------------------- head
// logits[p] = bias + sum_c x[p,c] * w[c]  (OutConv, n_classes == 1): one warp per pixel, fixed shuffle tree
__global__ void head_f32_kernel(const float* __restrict__ x, long long ld, const float* __restrict__ w,
                                const float* __restrict__ bias, float* __restrict__ logits, long long npix, int C) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long p = warp; p < npix; p += nwarps) {
    float acc = 0.f;
    for (int c = lane; c < C; c += 32) acc = fmaf(x[p * ld + c], w[c], acc);
    acc = warp_sum(acc);
    if (lane == 0) logits[p] = acc + (bias ? bias[0] : 0.f);
  }
}

}  // namespace

int f32_pack_split3_run(const float* src, void* dst, long long sr, long long sk, long long st, int R, int K, int T,
                        const int* slices, int ns, cudaStream_t s) {
  UNETK_CHECK(ns >= 1 && ns <= 4, -1, "pack_split3: 1..4 input slices");
  SplitPackArgs a{};
  a.src = src; a.dst = static_cast<__nv_bfloat16*>(dst); a.sr = sr; a.sk = sk; a.st = st;
  a.R = R; a.K = K; a.T = T; a.ns = ns;
  int k = 0;
  for (int i = 0; i < ns; ++i) { a.k0[i] = k; k += slices[i]; }
  a.k0[ns] = k;
  UNETK_CHECK(k == K, -1, "pack_split3: slices sum to %d, K = %d", k, K);
  const long long total = static_cast<long long>(T) * R * K;
  const int grid = static_cast<int>((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  pack_split3_kernel<<<grid, 256, 0, s>>>(a);
  UNETK_LAUNCHED();
  return 0;
}

int f32_stem_run(const float* x, int64_t sn, int64_t sc, int64_t sh, int64_t sw, const float* w, const float* bias,
                 float* y, int64_t y_ld, int N, int H, int W, int Cin, int Cout, cudaStream_t s) {
  UNETK_CHECK(Cin >= 1 && Cin <= 4 && Cout % 8 == 0 && Cout <= 256 && y_ld % 4 == 0, -1, "stem_f32: Cin <= 4, Cout %% 8 == 0");
  const size_t smem = static_cast<size_t>(Cout) * Cin * 9 * sizeof(float);
  const long long total = static_cast<long long>(N) * H * W * (Cout / 8);
  const int grid = static_cast<int>((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  stem_f32_kernel<<<grid, 256, smem, s>>>(x, sn, sc, sh, sw, w, bias, y, y_ld, N, H, W, Cin, Cout);
  UNETK_LAUNCHED();
  return 0;
}

static int f32_stats_grid(long long npix, int C) {
  const int rows = kStatThreads / (C / 4);
  long long g = (npix + rows - 1) / rows;
  return static_cast<int>(g < 296 ? g : 296);
}
size_t f32_stats_partial_doubles(long long npix, int C) {
  if (C < 4 || C % 4 || C > 1024) return 0;
  return static_cast<size_t>(f32_stats_grid(npix, C)) * 2 * C;
}
int f32_stats_run(const float* x, int64_t ld, int64_t npix, int C, double* partial, double* sums, cudaStream_t s) {
  UNETK_CHECK(C >= 4 && C % 4 == 0 && C <= 1024 && ld % 4 == 0, -1, "f32_stats: C %% 4 == 0, C <= 1024");
  const int rows = kStatThreads / (C / 4);
  const int grid = f32_stats_grid(npix, C);
  f32_stats_kernel<<<grid, kStatThreads, static_cast<size_t>(rows) * 2 * C * sizeof(double), s>>>(x, ld, npix, C, partial);
  UNETK_LAUNCHED();
  f32_stats_final_kernel<<<(2 * C + 127) / 128, 128, 0, s>>>(partial, grid, C, sums);
  UNETK_LAUNCHED();
  return 0;
}

int f32_bn_split_run(const float* raw, int64_t raw_ld, const float* scale, const float* shift, void* split,
                     int64_t split_ld, float* out_f32, int64_t out_ld, void* pooled, int64_t pooled_ld, int N, int H,
                     int W, int C, int relu, cudaStream_t s) {
  UNETK_CHECK(C % 4 == 0 && raw_ld % 4 == 0 && split_ld % 4 == 0 && out_ld % 4 == 0 && pooled_ld % 4 == 0, -1,
              "f32_bn_split: channel counts and strides must be multiples of 4");
  UNETK_CHECK((scale == nullptr) == (shift == nullptr), -1, "f32_bn_split: scale and shift go together");
  UNETK_CHECK(pooled == nullptr || (H % 2 == 0 && W % 2 == 0), -1, "f32_bn_split: pooling needs even H, W");
  BnSplitArgs a{};
  a.raw = raw; a.raw_ld = raw_ld; a.scale = scale; a.shift = shift;
  a.split = static_cast<__nv_bfloat16*>(split); a.split_ld = split_ld;
  a.out_f32 = out_f32; a.out_ld = out_ld;
  a.pooled = static_cast<__nv_bfloat16*>(pooled); a.pooled_ld = pooled_ld;
  a.N = N; a.H = H; a.W = W; a.C = C; a.relu = relu;
  const long long units = static_cast<long long>(N) * (pooled ? H / 2 : H) * (pooled ? W / 2 : W) * (C / 4);
  const int grid = static_cast<int>((units + 255) / 256 < 148 * 16 ? (units + 255) / 256 : 148 * 16);
  f32_bn_split_kernel<<<grid, 256, 0, s>>>(a);
  UNETK_LAUNCHED();
  return 0;
}

int f32_head_run(const float* x, int64_t ld, const float* w, const float* bias, float* logits, int64_t npix, int C,
                 cudaStream_t s) {
  const long long warps = npix;
  const int grid = static_cast<int>((warps + 7) / 8 < 148 * 8 ? (warps + 7) / 8 : 148 * 8);
  head_f32_kernel<<<grid, 256, 0, s>>>(x, ld, w, bias, logits, npix, C);
  UNETK_LAUNCHED();
  return 0;
}

}  // namespace unetk
