// Weight-gradient GEMM on tcgen05: D_tap[m, n] = sum_pixels P[pix + offP(tap)][m] * Q[pix + offQ(tap)][n].
// The reduction dimension is the pixel index, which is the SLOW dimension of NHWC, so both operands
// are "MN-major" UMMA tiles: a TMA box of (64 channels, TW, TH, 1) lands as [64 pixels][64 ch] with
// 128B swizzle, which is exactly the canonical MN-major SW128 atom sequence (8 pixel rows per atom).
// Split over pixel ranges (deterministic: fp32 partials + ordered reduce in wgrad_reduce_kernel).
// Tap packing (tpn > 1): when the taps differ only in the Q operand's offset (ConvTranspose2d: the four output phases of
// dY against ONE X tile) and Nn <= 128, the Q boxes of tpn taps sit side by side in one N = tpn * Nn tile: X is read once
// instead of once per tap and the MMA runs at N = 256 instead of N = 64 (half rate, profiles/r01_mma_rate_probe.txt).
// Reference semantics replaced: autograd's weight gradient of nn.Conv2d(k=3,p=1 / k=1) and
// nn.ConvTranspose2d(k=2,s=2) (UNetFamily/utils/unet_parts.py:24-31,56-58 in the reference).
#include <cstdlib>
#include "fastdiv.cuh"
#include "host_common.cuh"
#include "ptx.cuh"
#include "wgrad.cuh"

namespace unetk {

int wgrad_reduce_launch(const float* partial, float* dw, int ksplit, int taps, int M, int Nn, int64_t sm, int64_t sn,
                        int64_t st, int accumulate, cudaStream_t stream);
int wgrad_upfold_launch(const float* partial, float* dw, int ksplit, int M, int Nn, int64_t sm, int64_t sn, int64_t st,
                        int accumulate, cudaStream_t stream);

namespace {

constexpr int kPix = 64;      // pixels per k-block
constexpr int kThreads = 192;
constexpr uint32_t kBoxBytes = kPix * 64 * 2;  // one [64 pix][64 ch] box = 8 KB

template <int BN>
struct WCfg {
  static constexpr int kBoxes = 2 + BN / 64;
  static constexpr uint32_t kStageBytes = kBoxes * kBoxBytes;
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 128 ? 6 : 7);
  static constexpr uint32_t kTmemCols = 2 * BN;
  static constexpr uint32_t kSmemBytes = kStages * kStageBytes + 1024 + 256;
};

template <int BN>
__global__ void __launch_bounds__(kThreads, 1) wgrad_kernel(const __grid_constant__ WgradParams p) {
  using C = WCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + C::kStages;
  uint64_t* tfull_bar = bars + 2 * C::kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmP);
    tma_prefetch_desc(&p.tmQ);
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<C::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // programmatic dependent launch: everything above overlapped the previous kernel's tail; from here on global memory
  pdl_trigger();
  pdl_wait();

  const int items_per_split = p.tgroups * p.m_tiles * p.n_tiles;
  const int num_items = items_per_split * p.ksplit;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int nt = item % p.n_tiles;
        const int mt = (item / p.n_tiles) % p.m_tiles;
        const int tap = ((item / (p.n_tiles * p.m_tiles)) % p.tgroups) * p.tpn;   // first tap of the item's group
        const int ks = item / items_per_split;
        const int kt0 = static_cast<int>(static_cast<int64_t>(p.pix_tiles) * ks / p.ksplit);
        const int kt1 = static_cast<int>(static_cast<int64_t>(p.pix_tiles) * (ks + 1) / p.ksplit);
        for (int kt = kt0; kt < kt1; ++kt) {
          const int tw = kt % p.tiles_w;
          const int th = (kt / p.tiles_w) % p.tiles_h;
          const int img = kt / (p.tiles_w * p.tiles_h);
          const int h0 = th * p.TH, w0 = tw * p.TW;
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sp = smem + stage * C::kStageBytes;
          uint8_t* sq = sp + 2 * kBoxBytes;
          mbar_expect_tx(&full_bar[stage], C::kStageBytes);
          const int ph = p.p_step * h0 + p.p_dh[tap], pw = p.p_step * w0 + p.p_dw[tap];
          const int qh = p.q_step * h0 + p.q_dh[tap], qw = p.q_step * w0 + p.q_dw[tap];
#pragma unroll
          for (int b = 0; b < 2; ++b)
            tma_load_4d(sp + b * kBoxBytes, &p.tmP, &full_bar[stage], mt * 128 + b * 64, pw, ph, img);
          if (p.tpn > 1) {
            // box b = 64 channels of tap (tap + b*64/Nn): the taps of the group side by side in N
#pragma unroll
            for (int b = 0; b < BN / 64; ++b) {
              const int tb = tap + (b * 64) / p.Nn;
              tma_load_4d(sq + b * kBoxBytes, &p.tmQ, &full_bar[stage], (b * 64) % p.Nn, p.q_step * w0 + p.q_dw[tb],
                          p.q_step * h0 + p.q_dh[tb], img);
            }
          } else {
#pragma unroll
            for (int b = 0; b < BN / 64; ++b)
              tma_load_4d(sq + b * kBoxBytes, &p.tmQ, &full_bar[stage], nt * BN + b * 64, qw, qh, img);
          }
          if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    {
      // MMA issuer: warp-convergent loop, the elected lane issues (see umma_bf16_p in ptx.cuh)
      const bool issue = elect_one();
      constexpr uint32_t idesc = make_idesc_bf16(128, BN, true, true);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        const int ks = item / items_per_split;
        const int kt0 = static_cast<int>(static_cast<int64_t>(p.pix_tiles) * ks / p.ksplit);
        const int kt1 = static_cast<int>(static_cast<int64_t>(p.pix_tiles) * (ks + 1) / p.ksplit);
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait_p(issue, &tempty_bar[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kt = kt0; kt < kt1; ++kt) {
          mbar_wait_p(issue, &full_bar[stage], phase);
          tc_fence_after();
          const uint32_t p_base = smem_u32(smem + stage * C::kStageBytes);
          const uint32_t q_base = p_base + 2 * kBoxBytes;
#pragma unroll
          for (int k = 0; k < kPix / 16; ++k) {
            // 16 pixel rows per MMA = 2 swizzle atoms of 8 rows (SBO 1024 B); 64-channel boxes are
            // kBoxBytes apart (LBO).
            const uint64_t da = make_smem_desc(p_base + k * 2048, kBoxBytes, 1024, kLayoutSW128);
            const uint64_t db = make_smem_desc(q_base + k * 2048, kBoxBytes, 1024, kLayoutSW128);
            umma_bf16_p(issue, d_tmem, da, db, idesc, (kt > kt0) || (k != 0));
          }
          umma_commit_p(issue, &empty_bar[stage]);
          if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit_p(issue, &tfull_bar[acc]);
      }
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    int it = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      const int nt = item % p.n_tiles;
      const int mt = (item / p.n_tiles) % p.m_tiles;
      const int tap = ((item / (p.n_tiles * p.m_tiles)) % p.tgroups) * p.tpn;
      const int ks = item / items_per_split;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int m = mt * 128 + row;
      const size_t tap_stride = static_cast<size_t>(p.M) * p.Nn;
      float* dst = p.partial + ((static_cast<size_t>(ks) * p.taps + tap) * p.M + m) * p.Nn + nt * BN;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(taddr + c * 32, r);
        tmem_ld_wait();
        if (m < p.M) {
          // packed taps: columns [j*Nn, (j+1)*Nn) of the tile are tap (tap + j); a 32-column chunk never straddles two
          float* dc = dst + c * 32;
          int col0 = nt * BN + c * 32;
          if (p.tpn > 1) {
            const int j = (c * 32) / p.Nn;
            col0 = c * 32 - j * p.Nn;
            dc = dst + j * tap_stride + col0;
          }
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            if (col0 + v * 4 < p.Nn) {
              float4 o = make_float4(__uint_as_float(r[v * 4]), __uint_as_float(r[v * 4 + 1]),
                                     __uint_as_float(r[v * 4 + 2]), __uint_as_float(r[v * 4 + 3]));
              *reinterpret_cast<float4*>(dc + v * 4) = o;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<C::kTmemCols>(tmem_base);
  }
}

// dw[m*sm + n*sn + tap*st] (+)= sum_ks partial[ks][tap][m][n]   (fixed order -> deterministic)
// One thread per (m, n): its reads are coalesced along n for every (split, tap) plane and it writes all taps of its
// filter (9 adjacent floats when st = +-1).  The first version ran one thread per (tap, m, n) with two 64-bit
// divisions and a lone 4-byte store 36 B from its neighbours' (0.94 ms over the UNet step's 23 launches under ncu).
template <int TAPS>
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, int ksplit, int M,
                                    int Nn, int64_t sm, int64_t sn, int64_t st, int accumulate, FastDiv fd_n) {
  pdl_trigger();
  pdl_wait();
  const uint32_t plane = static_cast<uint32_t>(M) * static_cast<uint32_t>(Nn);   // < 2^31 (checked on the host)
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < plane; j += gridDim.x * blockDim.x) {
    uint32_t m, n;
    fd_n.divmod(j, m, n);
    float s[TAPS];
#pragma unroll
    for (int t = 0; t < TAPS; ++t) s[t] = 0.f;
    for (int k = 0; k < ksplit; ++k) {
#pragma unroll
      for (int t = 0; t < TAPS; ++t) s[t] += __ldg(partial + (static_cast<size_t>(k) * TAPS + t) * plane + j);
    }
    float* o = dw + m * sm + n * sn;
#pragma unroll
    for (int t = 0; t < TAPS; ++t) o[t * st] = accumulate ? (o[t * st] + s[t]) : s[t];
  }
}

// Few splits over a large filter (ksplit < 8): a 32 x 8 block owns a 32 (m) x 32 (n) tile of the filter for ALL taps.
// Reads: 4 rows x TAPS independent 128-byte-per-warp loads per split and thread (the one-thread-per-(m, n) kernel above
// had at most TAPS four-byte loads in flight and wrote every tap 36 B from its neighbour's: 1.1 TB/s on the
// 1024 x 1024 x 9 filter, profiles/r01_ncu_launch_shares_v9.txt).  Writes: the tile goes through shared memory and
// leaves in the order of dw's own contiguity — runs of 32*TAPS floats along n (sn < sm: conv weights
// [Cout][Cin][kh][kw]) or along m (swapped operands / ConvTranspose2d [Cin][Cout][2][2]).  Same summation order over
// the splits as wgrad_reduce_kernel: bit-identical results.
template <int TAPS>
__global__ void __launch_bounds__(256) wgrad_reduce_tiled_kernel(const float* __restrict__ partial, float* __restrict__ dw,
                                                                 int ksplit, int M, int Nn, int64_t sm, int64_t sn,
                                                                 int64_t st, int accumulate) {
  __shared__ float tile[32][32 * TAPS + 1];
  pdl_trigger();
  pdl_wait();
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int n0 = blockIdx.x * 32, m0 = blockIdx.y * 32;
  const size_t plane = static_cast<size_t>(M) * Nn;
  float acc[4][TAPS];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int t = 0; t < TAPS; ++t) acc[i][t] = 0.f;
  const bool n_ok = n0 + tx < Nn;
  for (int k = 0; k < ksplit; ++k) {
    const float* pk = partial + static_cast<size_t>(k) * TAPS * plane + n0 + tx;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + ty + 8 * i;
      if (n_ok && m < M) {
#pragma unroll
        for (int t = 0; t < TAPS; ++t) acc[i][t] += __ldg(pk + t * plane + static_cast<size_t>(m) * Nn);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int t = 0; t < TAPS; ++t) tile[ty + 8 * i][tx * TAPS + t] = acc[i][t];
  __syncthreads();
  const bool along_n = (sn < 0 ? -sn : sn) <= (sm < 0 ? -sm : sm);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int outer = ty + 8 * i;                     // m (runs along n) or n (runs along m)
    for (int idx = tx; idx < 32 * TAPS; idx += 32) {
      const int inner = idx / TAPS, t = idx - inner * TAPS;
      const int m = along_n ? outer : inner, n = along_n ? inner : outer;
      if (m0 + m < M && n0 + n < Nn) {
        float* o = dw + (m0 + m) * sm + (n0 + n) * sn + t * st;
        const float v = tile[m][n * TAPS + t];
        *o = accumulate ? (*o + v) : v;
      }
    }
  }
}

// Many pixel splits over a small filter (the 64-channel layers: 98 splits of a 64 x 64 x 9 filter): a 32 x 32 block
// owns 32 (m, n) positions of one tap, slice y adds the splits y, y+32, ... and row 0 adds the 32 slice sums in order.
template <int S>   // S slices of the split axis per block (8, 16 or 32)
__global__ void wgrad_reduce_sliced_kernel(const float* __restrict__ partial, float* __restrict__ dw, int ksplit,
                                           int taps, int M, int Nn, int64_t sm, int64_t sn, int64_t st, int accumulate,
                                           FastDiv fd_n) {
  __shared__ float red[S][33];
  pdl_trigger();
  pdl_wait();
  const uint32_t plane = static_cast<uint32_t>(M) * static_cast<uint32_t>(Nn);
  const uint32_t j = blockIdx.x * 32 + threadIdx.x;
  const int tap = blockIdx.y;
  float s = 0.f;
  if (j < plane) {
    for (int k = threadIdx.y; k < ksplit; k += S) s += __ldg(partial + (static_cast<size_t>(k) * taps + tap) * plane + j);
  }
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && j < plane) {
    float t = 0.f;
#pragma unroll
    for (int y = 0; y < S; ++y) t += red[y][threadIdx.x];
    uint32_t m, n;
    fd_n.divmod(j, m, n);
    float* o = dw + m * sm + n * sn + tap * st;
    *o = accumulate ? (*o + t) : t;
  }
}

// Sub-pixel up-conv (nearest 2x + conv3x3 as four 2x2-tap convs, see pack.cu): partial[ks][q*4 + u*2 + v][M][Nn] holds the
// gradients of the sixteen (phase, window tap) filters; every 3x3 tap (kh, kw) was summed into window tap
// (u, v) = (U(qy, kh), U(qx, kw)) of EACH phase, so its gradient is the sum of those four.  U(0, k) = k > 0, U(1, k) = k > 1.
__global__ void wgrad_reduce_upfold_kernel(const float* __restrict__ partial, float* __restrict__ dw, int ksplit, int M,
                                           int Nn, int64_t sm, int64_t sn, int64_t st, int accumulate, FastDiv fd_n) {
  pdl_trigger();
  pdl_wait();
  const uint32_t plane = static_cast<uint32_t>(M) * static_cast<uint32_t>(Nn);
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < plane; j += gridDim.x * blockDim.x) {
    uint32_t m, n;
    fd_n.divmod(j, m, n);
    float g[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) g[t] = 0.f;
    for (int k = 0; k < ksplit; ++k) {
#pragma unroll
      for (int t = 0; t < 16; ++t) g[t] += __ldg(partial + (static_cast<size_t>(k) * 16 + t) * plane + j);
    }
    float* o = dw + m * sm + n * sn;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int u = (q >> 1) == 0 ? (kh > 0) : (kh > 1), v = (q & 1) == 0 ? (kw > 0) : (kw > 1);
          s += g[q * 4 + u * 2 + v];
        }
        const int t = kh * 3 + kw;
        o[t * st] = accumulate ? (o[t * st] + s) : s;
      }
    }
  }
}

// The same fold for MANY pixel splits over a small filter (the full-resolution up-conv: 148 splits of a 128 x 64 filter; the
// kernel above would walk them with 32 blocks of dependent loads: 0.3 ms): a 32 x S block owns 32 (m, n) positions of one
// 3x3 tap, slice y adds the four sub-filter planes of the splits y, y+S, ... and row 0 adds the S slice sums in order.
template <int S>
__global__ void wgrad_reduce_upfold_sliced_kernel(const float* __restrict__ partial, float* __restrict__ dw, int ksplit, int M,
                                                  int Nn, int64_t sm, int64_t sn, int64_t st, int accumulate, FastDiv fd_n) {
  __shared__ float red[S][33];
  pdl_trigger();
  pdl_wait();
  const uint32_t plane = static_cast<uint32_t>(M) * static_cast<uint32_t>(Nn);
  const uint32_t j = blockIdx.x * 32 + threadIdx.x;
  const int t9 = blockIdx.y, kh = t9 / 3, kw = t9 - kh * 3;
  int src[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int u = (q >> 1) == 0 ? (kh > 0) : (kh > 1), v = (q & 1) == 0 ? (kw > 0) : (kw > 1);
    src[q] = q * 4 + u * 2 + v;
  }
  float s = 0.f;
  if (j < plane) {
    for (int k = threadIdx.y; k < ksplit; k += S) {
      const float* pk = partial + static_cast<size_t>(k) * 16 * plane + j;
      // same association as the unsliced kernel is NOT required (each kernel is deterministic by itself)
      s += (__ldg(pk + src[0] * static_cast<size_t>(plane)) + __ldg(pk + src[1] * static_cast<size_t>(plane))) +
           (__ldg(pk + src[2] * static_cast<size_t>(plane)) + __ldg(pk + src[3] * static_cast<size_t>(plane)));
    }
  }
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && j < plane) {
    float t = 0.f;
#pragma unroll
    for (int y = 0; y < S; ++y) t += red[y][threadIdx.x];
    uint32_t m, n;
    fd_n.divmod(j, m, n);
    float* o = dw + m * sm + n * sn + t9 * st;
    *o = accumulate ? (*o + t) : t;
  }
}

template <int BN>
int launch(const WgradParams& p, cudaStream_t stream) {
  using C = WCfg<BN>;
  static DeviceOnce once;
  UNETK_CUDA(once.run([] { return cudaFuncSetAttribute(wgrad_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes); }));
  const int items = p.tgroups * p.m_tiles * p.n_tiles * p.ksplit;
  const int grid = items < num_sms() ? items : num_sms();
  UNETK_CUDA(launch_pdl(wgrad_kernel<BN>, dim3(grid), dim3(kThreads), C::kSmemBytes, stream, p));
  UNETK_LAUNCHED();
  return 0;
}

// taps that can share one N tile: same P offsets, Nn a whole number of 64-channel boxes, tpn * Nn <= 256
int taps_per_tile(const WgradDesc& d) {
  static int env = -1;
  if (env < 0) { const char* e = getenv("UNETK_WGRAD_PACK_TAPS"); env = e ? atoi(e) : 1; }
  if (!env || d.taps < 2 || (d.Nn != 64 && d.Nn != 128)) return 1;
  for (int t = 1; t < d.taps; ++t)
    if (d.p_dh[t] != d.p_dh[0] || d.p_dw[t] != d.p_dw[0]) return 1;
  int tpn = 256 / d.Nn;
  while (tpn > 1 && d.taps % tpn) tpn >>= 1;
  return tpn;
}

void plan(const WgradDesc& d, int* BN, int* ksplit, int* TH, int* TW, int* pix_tiles, int* tpn_out = nullptr) {
  const int tpn = taps_per_tile(d);
  if (tpn_out) *tpn_out = tpn;
  *BN = tpn > 1 ? tpn * d.Nn : (d.Nn >= 256 ? 256 : (d.Nn > 64 ? 128 : 64));
  int tw = 64;
  while (tw > d.W) tw >>= 1;
  if (tw < 1) tw = 1;
  *TW = tw;
  *TH = kPix / tw;
  const int tiles_h = (d.H + *TH - 1) / *TH, tiles_w = (d.W + tw - 1) / tw;
  *pix_tiles = d.N * tiles_h * tiles_w;
  const int m_tiles = (d.M + 127) / 128, n_tiles = tpn > 1 ? 1 : (d.Nn + *BN - 1) / *BN;
  const int base = (d.taps / tpn) * m_tiles * n_tiles;
  int ks = (2 * num_sms() + base - 1) / base;
  const int cap = *pix_tiles / 8 > 0 ? *pix_tiles / 8 : 1;
  if (ks > cap) ks = cap;
  if (ks < 1) ks = 1;
  *ksplit = ks;
}

}  // namespace

int wgrad_reduce_launch(const float* partial, float* dw, int ksplit, int taps, int M, int Nn, int64_t sm, int64_t sn,
                        int64_t st, int accumulate, cudaStream_t stream) {
  const int64_t plane = static_cast<int64_t>(M) * Nn;
  UNETK_CHECK(plane < (1ll << 31), -1, "wgrad_reduce: M*N too large");
  int blocks = static_cast<int>((plane + 255) / 256);
  if (blocks > 8 * num_sms()) blocks = 8 * num_sms();
  const FastDiv fd(static_cast<uint32_t>(Nn));
  if (ksplit >= 8) {
    const dim3 grid(static_cast<unsigned>((plane + 31) / 32), taps);
    if (ksplit >= 64)
      UNETK_CUDA(launch_pdl(wgrad_reduce_sliced_kernel<32>, grid, dim3(32, 32), 0, stream, partial, dw, ksplit, taps, M, Nn, sm, sn, st, accumulate, fd));
    else if (ksplit >= 24)
      UNETK_CUDA(launch_pdl(wgrad_reduce_sliced_kernel<16>, grid, dim3(32, 16), 0, stream, partial, dw, ksplit, taps, M, Nn, sm, sn, st, accumulate, fd));
    else
      UNETK_CUDA(launch_pdl(wgrad_reduce_sliced_kernel<8>, grid, dim3(32, 8), 0, stream, partial, dw, ksplit, taps, M, Nn, sm, sn, st, accumulate, fd));
    UNETK_LAUNCHED();
    return 0;
  }
  static int tiled = -1;
  if (tiled < 0) { const char* e = getenv("UNETK_WGRAD_REDUCE_TILED"); tiled = e ? atoi(e) : 1; }
  if (tiled && (taps == 9 || taps == 4 || taps == 1)) {
    const dim3 grid(static_cast<unsigned>((Nn + 31) / 32), static_cast<unsigned>((M + 31) / 32));
    if (taps == 9) UNETK_CUDA(launch_pdl(wgrad_reduce_tiled_kernel<9>, grid, dim3(32, 8), 0, stream, partial, dw, ksplit, M, Nn, sm, sn, st, accumulate));
    else if (taps == 4) UNETK_CUDA(launch_pdl(wgrad_reduce_tiled_kernel<4>, grid, dim3(32, 8), 0, stream, partial, dw, ksplit, M, Nn, sm, sn, st, accumulate));
    else UNETK_CUDA(launch_pdl(wgrad_reduce_tiled_kernel<1>, grid, dim3(32, 8), 0, stream, partial, dw, ksplit, M, Nn, sm, sn, st, accumulate));
    UNETK_LAUNCHED();
    return 0;
  }
  switch (taps) {
    case 9: UNETK_CUDA(launch_pdl(wgrad_reduce_kernel<9>, dim3(blocks), dim3(256), 0, stream, partial, dw, ksplit, M, Nn, sm, sn, st, accumulate, fd)); break;
    case 4: UNETK_CUDA(launch_pdl(wgrad_reduce_kernel<4>, dim3(blocks), dim3(256), 0, stream, partial, dw, ksplit, M, Nn, sm, sn, st, accumulate, fd)); break;
    case 1: UNETK_CUDA(launch_pdl(wgrad_reduce_kernel<1>, dim3(blocks), dim3(256), 0, stream, partial, dw, ksplit, M, Nn, sm, sn, st, accumulate, fd)); break;
    default: UNETK_CHECK(false, -1, "wgrad_reduce: taps=%d (1, 4 or 9)", taps);
  }
  UNETK_LAUNCHED();
  return 0;
}

int wgrad_upfold_launch(const float* partial, float* dw, int ksplit, int M, int Nn, int64_t sm, int64_t sn, int64_t st,
                        int accumulate, cudaStream_t stream) {
  const int64_t plane = static_cast<int64_t>(M) * Nn;
  UNETK_CHECK(plane < (1ll << 31), -1, "wgrad_upfold: M*N too large");
  const FastDiv fd(static_cast<uint32_t>(Nn));
  if (ksplit >= 8) {
    const dim3 grid(static_cast<unsigned>((plane + 31) / 32), 9);
    if (ksplit >= 64)
      UNETK_CUDA(launch_pdl(wgrad_reduce_upfold_sliced_kernel<32>, grid, dim3(32, 32), 0, stream, partial, dw, ksplit, M, Nn, sm, sn, st, accumulate, fd));
    else
      UNETK_CUDA(launch_pdl(wgrad_reduce_upfold_sliced_kernel<8>, grid, dim3(32, 8), 0, stream, partial, dw, ksplit, M, Nn, sm, sn, st, accumulate, fd));
    UNETK_LAUNCHED();
    return 0;
  }
  int blocks = static_cast<int>((plane + 255) / 256);
  if (blocks > 8 * num_sms()) blocks = 8 * num_sms();
  UNETK_CUDA(launch_pdl(wgrad_reduce_upfold_kernel, dim3(blocks), dim3(256), 0, stream, partial, dw, ksplit, M, Nn, sm, sn, st,
                        accumulate, FastDiv(static_cast<uint32_t>(Nn))));
  UNETK_LAUNCHED();
  return 0;
}

size_t wgrad_workspace_bytes(const WgradDesc& d) {
  int BN, ks, TH, TW, pt;
  plan(d, &BN, &ks, &TH, &TW, &pt);
  return static_cast<size_t>(ks) * d.taps * d.M * d.Nn * sizeof(float);
}

int wgrad_run(const WgradDesc& d, void* workspace, size_t ws_bytes, cudaStream_t stream) {
  UNETK_CHECK(d.M % 8 == 0 && d.Nn % 8 == 0, -1, "wgrad: channel counts must be multiples of 8 (M=%d N=%d)", d.M, d.Nn);
  UNETK_CHECK(d.p_ld % 8 == 0 && d.q_ld % 8 == 0, -1, "wgrad: pixel strides must be multiples of 8");
  UNETK_CHECK(d.taps >= 1 && d.taps <= 16, -1, "wgrad: taps=%d", d.taps);
  UNETK_CHECK(!d.fold_up || d.taps == 16, -1, "wgrad: the sub-pixel fold needs 16 taps");
  WgradParams p{};
  int BN;
  plan(d, &BN, &p.ksplit, &p.TH, &p.TW, &p.pix_tiles, &p.tpn);
  p.tgroups = d.taps / p.tpn;
  const size_t need = static_cast<size_t>(p.ksplit) * d.taps * d.M * d.Nn * sizeof(float);
  UNETK_CHECK(workspace != nullptr && ws_bytes >= need, -1, "wgrad: workspace too small (%zu < %zu)", ws_bytes, need);
  p.tiles_h = (d.H + p.TH - 1) / p.TH;
  p.tiles_w = (d.W + p.TW - 1) / p.TW;
  p.m_tiles = (d.M + 127) / 128;
  p.n_tiles = p.tpn > 1 ? 1 : (d.Nn + BN - 1) / BN;
  p.taps = d.taps;
  p.M = d.M; p.Nn = d.Nn;
  p.p_step = d.p_step; p.q_step = d.q_step;
  for (int t = 0; t < d.taps; ++t) {
    p.p_dh[t] = d.p_dh[t]; p.p_dw[t] = d.p_dw[t]; p.q_dh[t] = d.q_dh[t]; p.q_dw[t] = d.q_dw[t];
  }
  p.partial = static_cast<float*>(workspace);
  UNETK_CHECK(p.TW * d.p_step <= 256 && p.TH * d.p_step <= 256 && p.TW * d.q_step <= 256 && p.TH * d.q_step <= 256,
              -1, "wgrad: TMA box too large");
  auto mk = [&](CUtensorMap* tm, const void* base, int64_t ld, int C, int step) -> int {
    const int AH = d.H * step, AW = d.W * step;
    uint64_t dims[4] = {static_cast<uint64_t>(C), static_cast<uint64_t>(AW), static_cast<uint64_t>(AH),
                        static_cast<uint64_t>(d.N)};
    uint64_t strides[3] = {static_cast<uint64_t>(ld) * 2, static_cast<uint64_t>(ld) * 2 * AW,
                           static_cast<uint64_t>(ld) * 2 * AW * AH};
    uint32_t box[4] = {64, static_cast<uint32_t>(p.TW * step), static_cast<uint32_t>(p.TH * step), 1};
    uint32_t es[4] = {1, static_cast<uint32_t>(step), static_cast<uint32_t>(step), 1};
    return make_tmap_bf16(tm, base, 4, dims, strides, box, es, true);
  };
  if (int rc = mk(&p.tmP, d.p, d.p_ld, d.M, d.p_step)) return rc;
  if (int rc = mk(&p.tmQ, d.q, d.q_ld, d.Nn, d.q_step)) return rc;
  int rc;
  switch (BN) {
    case 256: rc = launch<256>(p, stream); break;
    case 128: rc = launch<128>(p, stream); break;
    default: rc = launch<64>(p, stream); break;
  }
  if (rc) return rc;
  if (d.fold_up) return wgrad_upfold_launch(p.partial, d.dw, p.ksplit, d.M, d.Nn, d.dw_sm, d.dw_sn, d.dw_st, d.accumulate, stream);
  return wgrad_reduce_launch(p.partial, d.dw, p.ksplit, d.taps, d.M, d.Nn, d.dw_sm, d.dw_sn, d.dw_st, d.accumulate, stream);
}

}  // namespace unetk
