// Weight cache: fp32 PyTorch-layout master weights -> bf16 tap-major packs read by the TMA.
//   src [A][B][T] fp32  ->  dst_ab [T][A][B] bf16 and/or dst_ba [T][B][A] bf16
// One CTA transposes a 32 x 32 (a, b) tile for all T taps through shared memory, so global reads are contiguous
// runs of 32*T floats and global writes are 64-byte runs in BOTH destinations (the first version wrote dst_ba with
// a stride of A elements per thread: 0.58 ms per U-Net step for 21 weights; profiles/r01_step_profile_v3.txt).
// The batched entry point packs every weight of a plan in ONE launch from a device-side table.
#include "host_common.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

namespace unetk {

namespace {
constexpr int kTile = 32;
constexpr int kMaxT = 9;

struct PackJob {   // one row of the device table (8 x int64)
  const float* src;
  __nv_bfloat16* dst_ab;
  __nv_bfloat16* dst_ba;
  long long A, B, T;
  long long first_tile;   // index of this weight's first tile in the batched grid
  long long mode;         // 0: plain tap-major packs; 1: sub-pixel packs of an up_conv weight (T = 9), see pack_tile_up
};

// nearest-2x up-sampling followed by conv3x3(pad 1) (up_conv, unet_parts.py:99-111 in the reference) is, per output phase
// q = (qy, qx) of the 2x grid, a 2x2-tap convolution of the LOW-resolution input whose window starts at (qy-1, qx-1):
// rows kh of the 3x3 filter that read the same low-resolution row collapse into one tap,
//   qy = 0: u = 0 <- {kh 0},    u = 1 <- {kh 1, 2};      qy = 1: u = 0 <- {kh 0, 1}, u = 1 <- {kh 2}   (same along kw).
// The sums are taken in fp32 from the masters and rounded to bf16 once.  Layouts (A = Cout, B = Cin, t4 = u*2 + v):
//   dst_ab = forward  B operand [t4][q][Cout][Cin]   (tap-major, rows q*Cout + co, K = Cin)
//   dst_ba = dgrad    B operand [q*4 + t4][Cin][Cout] (16 taps, rows ci, K = Cout)
__device__ __forceinline__ float up_tap(const float* w9, int qy, int qx, int u, int v) {
  const int h0 = (qy == 0) ? (u == 0 ? 0 : 1) : (u == 0 ? 0 : 2), h1 = (qy == 0) ? (u == 0 ? 0 : 2) : (u == 0 ? 1 : 2);
  const int w0 = (qx == 0) ? (v == 0 ? 0 : 1) : (v == 0 ? 0 : 2), w1 = (qx == 0) ? (v == 0 ? 0 : 2) : (v == 0 ? 1 : 2);
  float acc = 0.f;
  for (int kh = h0; kh <= h1; ++kh)
    for (int kw = w0; kw <= w1; ++kw) acc += w9[kh * 3 + kw];
  return acc;
}

__device__ __forceinline__ void pack_tile(const PackJob& j, int tile, float* s /* [32][32*T + 1] */) {
  const int A = static_cast<int>(j.A), B = static_cast<int>(j.B), T = static_cast<int>(j.T);
  const int tiles_b = (B + kTile - 1) / kTile;
  const int a0 = (tile / tiles_b) * kTile, b0 = (tile % tiles_b) * kTile;
  const int na = min(kTile, A - a0), nb = min(kTile, B - b0);
  const int row = kTile * T + 1;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8 threads, no integer division below
  // load: row a of the tile is the contiguous run src[(a*B + b0)*T ... + nb*T).  All (up to 36) loads of a thread are
  // issued before the first one is stored: the rolled loop had ONE load in flight per thread (6 KB per SM) and the
  // batched pack ran at 1.6 TB/s.
  float v[4][kMaxT];
#pragma unroll
  for (int ai = 0; ai < 4; ++ai) {
    const int a = ty + 8 * ai;
    const float* run = j.src + (static_cast<long long>(a0 + a) * B + b0) * T;
#pragma unroll
    for (int i = 0; i < kMaxT; ++i) {
      const int r = tx + 32 * i;
      v[ai][i] = (a < na && r < nb * T) ? __ldg(run + r) : 0.f;
    }
  }
#pragma unroll
  for (int ai = 0; ai < 4; ++ai) {
    const int a = ty + 8 * ai;
#pragma unroll
    for (int i = 0; i < kMaxT; ++i) {
      const int r = tx + 32 * i;
      if (a < na && r < nb * T) s[a * row + r] = v[ai][i];
    }
  }
  __syncthreads();
  if (j.mode == 1) {   // T == 9 (checked on the host)
    if (j.dst_ab != nullptr && tx < nb) {
      for (int a = ty; a < na; a += 8) {
        const float* w9 = s + a * row + tx * 9;
        for (int q = 0; q < 4; ++q)
          for (int t4 = 0; t4 < 4; ++t4)
            j.dst_ab[(static_cast<long long>(t4 * 4 + q) * A + a0 + a) * B + b0 + tx] =
                __float2bfloat16_rn(up_tap(w9, q >> 1, q & 1, t4 >> 1, t4 & 1));
      }
    }
    if (j.dst_ba != nullptr && tx < na) {
      for (int b = ty; b < nb; b += 8) {
        const float* w9 = s + tx * row + b * 9;
        for (int q = 0; q < 4; ++q)
          for (int t4 = 0; t4 < 4; ++t4)
            j.dst_ba[(static_cast<long long>(q * 4 + t4) * B + b0 + b) * A + a0 + tx] =
                __float2bfloat16_rn(up_tap(w9, q >> 1, q & 1, t4 >> 1, t4 & 1));
      }
    }
    return;
  }
  if (j.dst_ab != nullptr && tx < nb) {
    for (int t = 0; t < T; ++t)
      for (int a = ty; a < na; a += 8)
        j.dst_ab[(static_cast<long long>(t) * A + a0 + a) * B + b0 + tx] = __float2bfloat16_rn(s[a * row + tx * T + t]);
  }
  if (j.dst_ba != nullptr && tx < na) {
    for (int t = 0; t < T; ++t)
      for (int b = ty; b < nb; b += 8)
        j.dst_ba[(static_cast<long long>(t) * B + b0 + b) * A + a0 + tx] = __float2bfloat16_rn(s[tx * row + b * T + t]);
  }
}

__global__ void __launch_bounds__(256) pack_weight_kernel(const PackJob job) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float smem_pack[];
  pack_tile(job, blockIdx.x, smem_pack);
}

__global__ void __launch_bounds__(256) pack_weights_kernel(const PackJob* __restrict__ table, int n) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float smem_pack[];
  // binary search: last job whose first_tile <= blockIdx.x
  int lo = 0, hi = n - 1;
  const long long tile = blockIdx.x;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (table[mid].first_tile <= tile) lo = mid; else hi = mid - 1;
  }
  const PackJob j = table[lo];
  pack_tile(j, static_cast<int>(tile - j.first_tile), smem_pack);
}

size_t tile_smem(int T) { return static_cast<size_t>(kTile) * (kTile * T + 1) * sizeof(float); }
}  // namespace

long long pack_tiles(int A, int B) {
  return static_cast<long long>((A + kTile - 1) / kTile) * ((B + kTile - 1) / kTile);
}

int pack_weight_run(const float* src, void* dst_ab, void* dst_ba, int A, int B, int T, cudaStream_t stream) {
  UNETK_CHECK(T <= kMaxT, -1, "pack_weight: T=%d > %d taps", T, kMaxT);
  PackJob j{src, static_cast<__nv_bfloat16*>(dst_ab), static_cast<__nv_bfloat16*>(dst_ba), A, B, T, 0, 0};
  UNETK_CUDA(launch_pdl(pack_weight_kernel, dim3(static_cast<int>(pack_tiles(A, B))), dim3(256), tile_smem(T), stream, j));
  UNETK_LAUNCHED();
  return 0;
}

int pack_upconv_weight_run(const float* src, void* dst_fwd, void* dst_dgrad, int Cout, int Cin, cudaStream_t stream) {
  PackJob j{src, static_cast<__nv_bfloat16*>(dst_fwd), static_cast<__nv_bfloat16*>(dst_dgrad), Cout, Cin, 9, 0, 1};
  UNETK_CUDA(launch_pdl(pack_weight_kernel, dim3(static_cast<int>(pack_tiles(Cout, Cin))), dim3(256), tile_smem(9), stream, j));
  UNETK_LAUNCHED();
  return 0;
}

// table: device int64 [n][8] rows {src, dst_ab, dst_ba, A, B, T, first_tile, mode} with first_tile ascending from 0
int pack_weights_run(const long long* table, int n, long long total_tiles, cudaStream_t stream) {
  UNETK_CHECK(table != nullptr && n > 0 && total_tiles > 0 && total_tiles < (1ll << 31), -1, "pack_weights: bad arguments");
  static_assert(sizeof(PackJob) == 64, "PackJob must be 8 x int64");
  UNETK_CUDA(launch_pdl(pack_weights_kernel, dim3(static_cast<int>(total_tiles)), dim3(256), tile_smem(kMaxT), stream, reinterpret_cast<const PackJob*>(table), n));
  UNETK_LAUNCHED();
  return 0;
}

}  // namespace unetk
