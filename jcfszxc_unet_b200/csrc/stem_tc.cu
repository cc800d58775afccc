// Stem convolution on the tensor core: the network-input conv (Cin <= 4, 3x3, pad 1) as an im2col GEMM whose
// 128 x 48 A tile (128 pixels x the 9*Cin-wide patch, zero-padded to K = 48) is built IN SHARED MEMORY by the
// threads themselves — K = 27 cannot feed a 128-byte TMA row, but nothing forces the A operand to come from TMA:
// the un-swizzled K-major core-matrix layout ([k/8][row][8 elements], LBO = rows*16 B, SBO = 128 B, validated by
// the descriptor probe, profiles/r01_umma_descriptor_probe.txt mode 1) is trivial to write with 16-byte stores.
//
//   forward : D[pixel, co] = sum_k patch[pixel][k] * w[co][k]          3 tcgen05.mma (128 x Cout x 16) per tile
//   wgrad   : dW[co][k]    = sum_pixel dY[pixel][co] * patch[pixel][k] K = pixels; dY tile by TMA (MN-major, 128B
//             swizzle), patch tile written by the threads in the same swizzled MN-major layout; accumulators stay
//             in TMEM over all tiles of a CTA, fp32 partials + ordered reduce (deterministic)
//
// The CUDA-core kernels these replace (stem.cu) spent 1728 FFMA + 460 LDS per pixel and ran at 0.67 / 0.72 ms for
// a 16 x 3 x 512 x 512 batch; the output write alone (537 MB) takes 0.09 ms at HBM speed.
// Reference semantics replaced: the first nn.Conv2d of DoubleConv / conv_block / NestedUNet.DoubleConv / ResUNet
// input_layer + input_skip (UNet.py:21, unet_parts.py:24,85, UNetPP.py:18, ResUNet.py:23,30) and its weight gradient.
#include "fastdiv.cuh"
#include "host_common.cuh"
#include "kernels.cuh"
#include "ptx.cuh"
#include "reduce2.cuh"

namespace unetk {

int conv_stats_sums_launch(const float* partial, int grid, int num_n_tiles, int BN, int C, double* sums,
                           cudaStream_t stream);   // conv_gemm.cu

namespace {

constexpr int kTile = 128;     // pixels per tile (consecutive along W)
constexpr int kK = 48;         // padded patch width (>= 9 * Cin for Cin <= 5)
constexpr int kChunks = kK / 8;

struct StemTc {
  const float* x; int64_t sn, sc, sh, sw;  // fp32 image, element strides
  int N, H, W, Cin, Cout, tiles_w;
  FastDiv fd_tw, fd_h;                     // tile -> (row, column tile), row -> (image, h) without hardware division
  const float* scale = nullptr;            // forward only: eval-mode BatchNorm fold, y = relu?(acc * scale + bias)
  int relu = 0;
};

// tile -> pointer to x[n, 0, h, w] and (h, w) of this thread's pixel
struct StemPix {
  const float* p; int h, w, row, tw;
};
__device__ __forceinline__ StemPix stem_pix(const StemTc& G, int tile, int tid) {
  uint32_t row, tw, n, h;
  G.fd_tw.divmod(static_cast<uint32_t>(tile), row, tw);
  G.fd_h.divmod(row, n, h);
  StemPix q;
  q.h = static_cast<int>(h); q.w = static_cast<int>(tw) * kTile + tid; q.row = static_cast<int>(row); q.tw = static_cast<int>(tw);
  q.p = G.x + n * G.sn + h * G.sh + q.w * G.sw;
  return q;
}

// patch of pixel (n, h, w): k = ci*9 + r*3 + s  ->  bf16(x[n, ci, h+r-1, w+s-1]) (0 outside), k >= 9*Cin -> 0.
// Addresses are built from ONE pointer per pixel with +-stride steps: the first version multiplied four runtime
// 64-bit strides per tap and spent 452 instructions on 27 loads (the kernel was issue-bound at 74 % issue-active,
// profiles/r01_ncu_stem_head.txt).
__device__ __forceinline__ void load_patch(const StemTc& G, const StemPix& q, uint32_t (&pk)[kK / 2]) {
  float v[kK];
#pragma unroll
  for (int k = 0; k < kK; ++k) v[k] = 0.f;
  if (q.w < G.W) {
    const bool rok[3] = {q.h >= 1, true, q.h + 1 < G.H};
    const bool cok[3] = {q.w >= 1, true, q.w + 1 < G.W};
    const float* pc = q.p;
#pragma unroll
    for (int ci = 0; ci < 4; ++ci) {
      if (ci < G.Cin) {
        const float* pr = pc - G.sh;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          if (rok[r]) {
            if (cok[0]) v[ci * 9 + r * 3 + 0] = __ldg(pr - G.sw);
            v[ci * 9 + r * 3 + 1] = __ldg(pr);
            if (cok[2]) v[ci * 9 + r * 3 + 2] = __ldg(pr + G.sw);
          }
          pr += G.sh;
        }
      }
      pc += G.sc;
    }
  }
#pragma unroll
  for (int k = 0; k < kK / 2; ++k) pk[k] = pack_bf16x2(v[2 * k], v[2 * k + 1]);
}

// ---------------------------------------------------------------------------------------------- forward
template <int CMAX>   // TMEM columns allocated = max Cout handled (64 / 128 / 256): small CMAX => more CTAs per SM
__global__ void __launch_bounds__(kTile, (512 / CMAX) < 6 ? (512 / CMAX) : 6) stem_tc_fwd_kernel(const StemTc G, const float* __restrict__ wgt,
                                                            const float* __restrict__ bias,
                                                            __nv_bfloat16* __restrict__ y, int64_t y_ld, int num_tiles,
                                                            float* __restrict__ stats_partial) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~static_cast<uintptr_t>(127));
  uint8_t* sA = smem;                                  // [kChunks][128 rows][16 B] (12 KB); 16 KB: also output staging
  uint8_t* sB = sA + kTile * 128;                      // [kChunks][Cout rows][16 B]
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + kChunks * CMAX * 16);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  float* sStats = reinterpret_cast<float*>(bar + 8);   // [4 warps][2][CMAX]: BatchNorm sum / sum of squares (optional)
  float* sBias = sStats + 4 * 2 * CMAX;                // [CMAX]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Cout = G.Cout, K = 9 * G.Cin;
  if (stats_partial != nullptr)
    for (int i = tid; i < 4 * 2 * CMAX; i += kTile) sStats[i] = 0.f;
  const bool has_bias = bias != nullptr;
  if (has_bias)
    for (int i = tid; i < CMAX; i += kTile) sBias[i] = i < Cout ? __ldg(bias + i) : 0.f;

  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<CMAX>(slot);
  // B: row n = output channel, bf16-rounded weights (autocast), zero beyond K
  for (int n = tid; n < Cout; n += kTile) {
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = c * 8 + j;
        f[j] = k < K ? __ldg(wgt + static_cast<int64_t>(n) * K + k) : 0.f;
      }
      uint4 u;
      u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]); u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
      *reinterpret_cast<uint4*>(sB + (c * Cout + n) * 16) = u;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((static_cast<uint32_t>(Cout) >> 3) << 17) | ((128u >> 4) << 24);
  const uint32_t lbo_a = kTile * 16, lbo_b = static_cast<uint32_t>(Cout) * 16;
  uint32_t parity = 0;
  // loop-invariant pieces of the epilogue addressing (32-bit shared addresses, one 64-bit output pointer per tile)
  const uint32_t sA_u = smem_u32(sA), wbuf_u = sA_u + warp * 4096, sBias_u = smem_u32(sBias);
  const int piece = lane & 7, rsub = lane >> 3;                       // store phase: 16-byte piece / row within a group of 4
  const int64_t lane_off = rsub * y_ld + piece * 8, step4 = 4 * y_ld;
  uint32_t ld_off[8];                                                 // staged row i*4 + rsub, swizzled piece
#pragma unroll
  for (int i = 0; i < 8; ++i) ld_off[i] = (i * 4 + rsub) * 128 + ((piece ^ ((i * 4 + rsub) & 7)) << 4);
  uint32_t st_off[8];                                                 // statistics: row 8k' + k, this lane's channel pair
#pragma unroll
  for (int k = 0; k < 8; ++k) st_off[k] = k * 128 + (((lane >> 2) ^ k) << 4) + (lane & 3) * 4;

  // software pipeline: the patch of the NEXT tile is fetched (27 global loads per thread) while this tile's MMAs
  // and output stores are in flight
  uint32_t pk[kK / 2];
  StemPix cur = stem_pix(G, blockIdx.x < num_tiles ? blockIdx.x : 0, tid);
  if (blockIdx.x < num_tiles) load_patch(G, cur, pk);
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int tw = cur.tw, row = cur.row;
#pragma unroll
    for (int c = 0; c < kChunks; ++c)
      sts128(sA_u + (c * kTile + tid) * 16, make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]));
    fence_proxy_async_smem();          // generic-proxy writes -> visible to the tensor core's async-proxy reads
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < kK / 16; ++ks) {
        const uint64_t da = make_smem_desc(sA_u + ks * 2 * lbo_a, lbo_a, 128, kLayoutNone);
        const uint64_t db = make_smem_desc(smem_u32(sB) + ks * 2 * lbo_b, lbo_b, 128, kLayoutNone);
        umma_bf16(tmem, da, db, idesc, ks != 0);
      }
      umma_commit(bar);
    }
    {
      const int nt = tile + gridDim.x;
      if (nt < num_tiles) {
        cur = stem_pix(G, nt, tid);
        load_patch(G, cur, pk);
      }
    }
    mbar_wait(bar, parity);
    parity ^= 1u;
    tc_fence_after();
    // epilogue: thread = pixel = TMEM lane; Cout fp32 columns -> +bias -> bf16.  64-channel groups are transposed
    // through shared memory (the A tile's region is free once its MMAs retired) so that a warp store instruction
    // covers four whole 128-byte pixel rows instead of 16 bytes of 32 different rows (32x fewer LSU wavefronts).
    const int w0 = tw * kTile + warp * 32;                    // first pixel column of this warp's 32 rows
    const int rows = min(32, G.W - w0);                       // live rows (<= 0: the warp is past the image edge)
    __nv_bfloat16* const orow = y + (static_cast<int64_t>(row) * G.W + w0) * y_ld;
    const uint32_t taddr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    for (int c0 = 0; c0 < Cout; c0 += 64) {
      const bool full64 = c0 + 64 <= Cout;
      if (full64) __syncwarp();
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        if (c0 + half * 32 >= Cout) break;
        uint32_t r[32];
        tmem_ld32(taddr + c0 + half * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const int cb = c0 + half * 32 + v * 8;
          float f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(r[v * 8 + j]);
          if (G.scale != nullptr) {   // eval-mode BatchNorm fold (bias holds the folded shift)
            const float4 b0 = lds128_f(sBias_u + cb * 4), b1 = lds128_f(sBias_u + cb * 4 + 16);
            float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
            if (cb < Cout) { s0 = __ldg(reinterpret_cast<const float4*>(G.scale + cb)); s1 = __ldg(reinterpret_cast<const float4*>(G.scale + cb + 4)); }
            f[0] = fmaf(f[0], s0.x, b0.x); f[1] = fmaf(f[1], s0.y, b0.y); f[2] = fmaf(f[2], s0.z, b0.z); f[3] = fmaf(f[3], s0.w, b0.w);
            f[4] = fmaf(f[4], s1.x, b1.x); f[5] = fmaf(f[5], s1.y, b1.y); f[6] = fmaf(f[6], s1.z, b1.z); f[7] = fmaf(f[7], s1.w, b1.w);
            if (G.relu) {
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
            }
          } else if (has_bias) {
            const float4 b0 = lds128_f(sBias_u + cb * 4), b1 = lds128_f(sBias_u + cb * 4 + 16);
            f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
            f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
          }
          uint4 u;
          u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]); u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
          if (full64) sts128(wbuf_u + lane * 128 + (((half * 4 + v) ^ (lane & 7)) << 4), u);
          else if (lane < rows && cb < Cout) *reinterpret_cast<uint4*>(orow + lane * y_ld + cb) = u;
        }
      }
      if (full64) {
        __syncwarp();
        __nv_bfloat16* o = orow + lane_off + c0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (i * 4 + rsub < rows) *reinterpret_cast<uint4*>(o) = lds128(wbuf_u + ld_off[i]);
          o += step4;
        }
        if (stats_partial != nullptr) {
          // BatchNorm statistics of the bf16 values just stored: lane owns channels c0 + 2*lane, +1 and sums them
          // over the warp's staged rows (one 4-byte word per row: conflict-free), then adds into its own slot
          float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
          for (int r8 = 0; r8 < rows; r8 += 8) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              if (r8 + k < rows) {
                const uint32_t v = lds_u32(wbuf_u + r8 * 128 + st_off[k]);
                const float a = bf16_lo(v), b = bf16_hi(v);
                s0 += a; q0 = fmaf(a, a, q0);
                s1 += b; q1 = fmaf(b, b, q1);
              }
            }
          }
          float* st = sStats + warp * 2 * CMAX + c0 + 2 * lane;
          st[0] += s0; st[1] += s1; st[CMAX] += q0; st[CMAX + 1] += q1;
        }
      }
    }
    tc_fence_before();   // TMEM reads done before the next tile's MMAs overwrite the accumulator
    __syncthreads();     // ... and sA may be rewritten (its MMAs completed: the commit barrier was waited on)
  }
  if (stats_partial != nullptr) {   // partial[cta][2][Cout]; the loop's last __syncthreads ordered the slot updates
    for (int i = tid; i < 2 * Cout; i += kTile) {
      const int k = i / Cout, c = i % Cout;
      float t = 0.f;
#pragma unroll
      for (int wq = 0; wq < 4; ++wq) t += sStats[wq * 2 * CMAX + k * CMAX + c];
      stats_partial[static_cast<size_t>(blockIdx.x) * 2 * Cout + i] = t;
    }
  }
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<CMAX>(tmem);
  }
}

int stem_tc_fwd_grid(int tiles, int CMAX) {
  const int per_sm = (512 / CMAX) < 6 ? (512 / CMAX) : 6;   // 6 / 4 / 2
  return tiles < per_sm * num_sms() ? tiles : per_sm * num_sms();
}

template <int CMAX>
int stem_tc_fwd_launch(const StemTc& G, const float* w, const float* bias, void* y, int64_t y_ld, int tiles, cudaStream_t s,
                       float* stats_partial = nullptr, double* stats_sums = nullptr) {
  const int smem = kTile * 128 + kChunks * CMAX * 16 + 64 + 128 + 4 * 2 * CMAX * 4 + CMAX * 4;
  // CTAs per SM: bounded by TMEM (512 / CMAX columns); one CTA builds its patch tile while the others' MMAs /
  // stores run
  const int grid = stem_tc_fwd_grid(tiles, CMAX);
  static DeviceOnce once;
  UNETK_CUDA(once.run([smem] { return cudaFuncSetAttribute(stem_tc_fwd_kernel<CMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); }));
  UNETK_CUDA(launch_pdl(stem_tc_fwd_kernel<CMAX>, dim3(grid), dim3(kTile), smem, s, G, w, bias, static_cast<__nv_bfloat16*>(y), y_ld, tiles,
                        stats_partial));
  UNETK_LAUNCHED();
  if (stats_partial != nullptr) return conv_stats_sums_launch(stats_partial, grid, 1, G.Cout, G.Cout, stats_sums, s);
  return 0;
}

// ---------------------------------------------------------------------------------------------- weight gradient
// D[co][k] (TMEM, 128 lanes x 64 columns) += sum over the tile's 128 pixels of dY[pix][co] * patch[pix][k].
//   A = dY tile, MN-major (rows = pixels = K), two 64-channel boxes by TMA (the second is all out-of-bounds = zeros
//       when Cout <= 64), 128B swizzle;   B = patch tile [128 px][64 k-slots] written by the threads with the same
//       swizzle (16-byte chunk j of row r at position j ^ (r & 7)).  Same descriptors as wgrad.cu.
struct StemTcW {
  CUtensorMap tmDy;   // dims (Cout, W, N*H), box (64, 128, 1)
  StemTc G;
};

__global__ void __launch_bounds__(kTile) stem_tc_wgrad_kernel(const __grid_constant__ StemTcW P, float* __restrict__ partial,
                                                              int num_tiles) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  constexpr uint32_t kBox = kTile * 128;               // 16 KB: [128 px][64 elements] bf16
  uint8_t* sA = smem;                                  // 2 boxes (channels 0-63, 64-127)
  uint8_t* sB = smem + 2 * kBox;                       // patch tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + kBox);   // [0] dY landed, [1] MMAs retired
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 2);
  const StemTc& G = P.G;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    tma_prefetch_desc(&P.tmDy);
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<64>(slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  constexpr uint32_t idesc = make_idesc_bf16(128, 64, true, true);
  uint32_t parity = 0;
  bool first = true;
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const StemPix cur = stem_pix(G, tile, tid);
    const int tw = cur.tw, row = cur.row;
    if (tid == 0) {
      mbar_expect_tx(&bars[0], 2 * kBox);
      tma_load_3d(sA, &P.tmDy, &bars[0], 0, tw * kTile, row);
      tma_load_3d(sA + kBox, &P.tmDy, &bars[0], 64, tw * kTile, row);
    }
    uint32_t pk[kK / 2];
    load_patch(G, cur, pk);
    uint8_t* rowp = sB + tid * 128;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const uint4 v = c < kChunks ? make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]) : make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(rowp + ((c ^ (tid & 7)) << 4)) = v;
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      mbar_wait(&bars[0], parity);
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < kTile / 16; ++ks) {
        const uint64_t da = make_smem_desc(smem_u32(sA) + ks * 2048, kBox, 1024, kLayoutSW128);
        const uint64_t db = make_smem_desc(smem_u32(sB) + ks * 2048, kBox, 1024, kLayoutSW128);
        umma_bf16(tmem, da, db, idesc, (first && ks == 0) ? 0u : 1u);
      }
      umma_commit(&bars[1]);
    }
    first = false;
    mbar_wait(&bars[1], parity);    // smem may be refilled once these MMAs have read it
    parity ^= 1u;
  }
  tc_fence_after();
  // partial[blk][co][k]: lane = co (warp w holds channels 32w..32w+31), columns 0..K-1
  const int K = 9 * G.Cin;
  const uint32_t taddr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  uint32_t r0[32], r1[32];
  tmem_ld32(taddr, r0);
  tmem_ld32(taddr + 32, r1);
  tmem_ld_wait();
  const int co = tid;
  if (co < G.Cout) {
    float* dst = partial + (static_cast<size_t>(blockIdx.x) * G.Cout + co) * K;
    const bool any = blockIdx.x < num_tiles;   // a CTA without tiles never wrote its accumulator
#pragma unroll
    for (int k = 0; k < 32; ++k)
      if (k < K) dst[k] = any ? __uint_as_float(r0[k]) : 0.f;
#pragma unroll
    for (int k = 32; k < kK; ++k)
      if (k < K) dst[k] = any ? __uint_as_float(r1[k - 32]) : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<64>(tmem);
  }
}

// 32 x 32 threads: 32 outputs per block, the CTA partials summed in 32 slices and then in a fixed order (the first
// version walked up to 592 partials per thread: 47 us, profiles/r01_ncu_stem_head.txt)
__global__ void stem_tc_wgrad_reduce_kernel(const float* __restrict__ partial, int nblk, int n, float* __restrict__ dw,
                                            int accumulate) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * kSum2Lanes + threadIdx.x;
  const bool valid = i < n;
  const double s = sliced_ordered_sum(partial, nblk, valid, [&](int b) { return static_cast<size_t>(b) * n + i; });
  if (valid && threadIdx.y == 0) dw[i] = accumulate ? dw[i] + static_cast<float>(s) : static_cast<float>(s);
}

int stem_tc_wgrad_grid(int64_t tiles) {
  const int64_t cap = 4ll * num_sms();   // 49 KB of smem and 64 TMEM columns per CTA
  return static_cast<int>(tiles < cap ? tiles : cap);
}

}  // namespace

bool stem_tc_ok(int Cin, int Cout) { return Cin >= 1 && 9 * Cin <= kK && Cout % 16 == 0 && Cout >= 16 && Cout <= 256; }

// BatchNorm statistics in the epilogue: whole 64-channel groups only (they are taken from the staged store tiles)
bool stem_tc_stats_ok(int Cin, int Cout) { return stem_tc_ok(Cin, Cout) && Cout % 64 == 0; }
size_t stem_tc_stats_partial_floats(int N, int H, int W, int Cout) {
  const int64_t tiles = static_cast<int64_t>(N) * H * ((W + kTile - 1) / kTile);
  const int cmax = Cout <= 64 ? 64 : (Cout <= 128 ? 128 : 256);
  return static_cast<size_t>(stem_tc_fwd_grid(static_cast<int>(tiles < (1 << 30) ? tiles : (1 << 30)), cmax)) * 2 * Cout;
}

int stem_tc_fwd_run(const float* x, int64_t sn, int64_t sc, int64_t sh, int64_t sw, const float* w, const float* bias,
                    void* y, int64_t y_ld, int N, int H, int W, int Cin, int Cout, cudaStream_t s, float* stats_partial,
                    double* stats_sums, const float* scale, int relu) {
  UNETK_CHECK(stem_tc_ok(Cin, Cout), -1, "stem_tc: Cin=%d Cout=%d not supported", Cin, Cout);
  UNETK_CHECK(scale == nullptr || (bias != nullptr && stats_partial == nullptr && (reinterpret_cast<uintptr_t>(scale) & 15) == 0 && Cout % 8 == 0),
              -1, "stem_tc: the affine epilogue needs a 16-byte aligned scale, a shift vector, Cout %% 8 == 0 and no statistics");
  UNETK_CHECK(y_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0, -1, "stem_tc: output must be 16-byte aligned");
  StemTc G{x, sn, sc, sh, sw, N, H, W, Cin, Cout, (W + kTile - 1) / kTile,
           FastDiv(static_cast<uint32_t>((W + kTile - 1) / kTile)), FastDiv(static_cast<uint32_t>(H))};
  const int64_t tiles64 = static_cast<int64_t>(N) * H * G.tiles_w;
  UNETK_CHECK(tiles64 < (1ll << 31), -1, "stem_tc: too many tiles");
  const int tiles = static_cast<int>(tiles64);
  G.scale = scale;
  G.relu = relu;
  UNETK_CHECK(stats_partial == nullptr || (stem_tc_stats_ok(Cin, Cout) && stats_sums != nullptr), -1,
              "stem_tc: fused statistics need Cout %% 64 == 0");
  if (Cout <= 64) return stem_tc_fwd_launch<64>(G, w, bias, y, y_ld, tiles, s, stats_partial, stats_sums);
  if (Cout <= 128) return stem_tc_fwd_launch<128>(G, w, bias, y, y_ld, tiles, s, stats_partial, stats_sums);
  return stem_tc_fwd_launch<256>(G, w, bias, y, y_ld, tiles, s, stats_partial, stats_sums);
}

size_t stem_tc_wgrad_workspace(int N, int H, int W, int Cin, int Cout) {
  const int64_t tiles = static_cast<int64_t>(N) * H * ((W + kTile - 1) / kTile);
  return static_cast<size_t>(stem_tc_wgrad_grid(tiles)) * Cout * 9 * Cin * sizeof(float);
}

int stem_tc_wgrad_run(const float* x, int64_t sn, int64_t sc, int64_t sh, int64_t sw, const void* dy, int64_t dy_ld,
                      float* dw, int accumulate, int N, int H, int W, int Cin, int Cout, void* ws, size_t ws_bytes,
                      cudaStream_t s) {
  UNETK_CHECK(stem_tc_ok(Cin, Cout) && Cout <= 128, -1, "stem_tc_wgrad: Cin=%d Cout=%d not supported", Cin, Cout);
  UNETK_CHECK(ws != nullptr && ws_bytes >= stem_tc_wgrad_workspace(N, H, W, Cin, Cout), -1, "stem_tc_wgrad: workspace too small");
  UNETK_CHECK(dy_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(dy) & 15) == 0, -1, "stem_tc_wgrad: dy must be 16-byte aligned");
  StemTcW P{};
  P.G = StemTc{x, sn, sc, sh, sw, N, H, W, Cin, Cout, (W + kTile - 1) / kTile,
               FastDiv(static_cast<uint32_t>((W + kTile - 1) / kTile)), FastDiv(static_cast<uint32_t>(H))};
  const int64_t tiles64 = static_cast<int64_t>(N) * H * P.G.tiles_w;
  UNETK_CHECK(tiles64 < (1ll << 31), -1, "stem_tc_wgrad: too many tiles");
  {
    uint64_t dims[3] = {static_cast<uint64_t>(Cout), static_cast<uint64_t>(W), static_cast<uint64_t>(N) * H};
    uint64_t strides[2] = {static_cast<uint64_t>(dy_ld) * 2, static_cast<uint64_t>(dy_ld) * 2 * W};
    uint32_t box[3] = {64, kTile, 1};
    uint32_t es[3] = {1, 1, 1};
    if (int rc = make_tmap_bf16(&P.tmDy, dy, 3, dims, strides, box, es, true)) return rc;
  }
  const int grid = stem_tc_wgrad_grid(tiles64);
  const int smem = 3 * kTile * 128 + 64 + 1024;
  static DeviceOnce once;
  UNETK_CUDA(once.run([smem] { return cudaFuncSetAttribute(stem_tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); }));
  UNETK_CUDA(launch_pdl(stem_tc_wgrad_kernel, dim3(grid), dim3(kTile), smem, s, P, static_cast<float*>(ws), static_cast<int>(tiles64)));
  UNETK_LAUNCHED();
  const int n = Cout * 9 * Cin;
  UNETK_CUDA(launch_pdl(stem_tc_wgrad_reduce_kernel, dim3((n + kSum2Lanes - 1) / kSum2Lanes), dim3(kSum2Lanes, kSum2Slices), 0, s,
                        static_cast<const float*>(ws), grid, n, dw, accumulate));
  UNETK_LAUNCHED();
  return 0;
}

}  // namespace unetk
