// Weight gradient of the sub-pixel up-conv (nearest 2x + conv3x3, unet_parts.py:103-104 in the reference; forward in
// capi.cu unetk_upconv3x3_fwd): the sixteen (output phase q, window tap (u, v)) sub-filter gradients
//
//   G[q][u][v][ci][co] = sum_{n,i,j} X[n, i, j, ci] * dYq[n, i - (qy-1+u), j - (qx-1+v), co],   dYq[ii, jj] = dY[2 ii + qy, 2 jj + qx]
//
// with the pixel index as the reduction dimension (both operands MN-major, as in wgrad3x3.cu).  The per-tap kernel
// (wgrad.cu) loads a fresh X and dY tile for every tap and issues one MMA per pair of tiles: operand reads and TMA fills want
// far more than the ~128 B/clk of shared memory (230-760 TFLOP/s on AttentionUNet, profiles/r02_*).  Here one work item owns all FOUR taps of one
// phase: per 64-pixel k-block it loads
//   P: the X tile                       box (64 ch, TW, TH)           -> [64 px][64 ch]  x 2 (M = 128 input channels)
//   Q: ONE halo tile of the phase view  box (64 ch, TW+1, TH+1), element stride 2 over dY, starting at phase pixel
//      (h0 - qy, w0 - qx)                                             -> [(TH+1)(TW+1) px][64 ch] per 64 output channels
// and tap (u, v) is the MMA whose Q descriptor starts (1-u) halo rows and (1-v) pixels into that tile (tcgen05 applies the
// 128B-swizzle XOR on absolute shared-memory address bits, profiles/r01_umma_descriptor_probe.txt): 4 x 128 fp32
// accumulator columns = the whole TMEM, one X tile and one halo tile per four taps (ncu: 73-80 % tensor-pipe activity,
// bound by the shared-memory reads of its N = 128 MMAs like wgrad3x3_kernel<128>, profiles/r02_ncu_upconv.txt).  Cout <= 64 (the full-resolution up-conv, 128 -> 64): a work item
// owns the two phases (qy, 0), (qy, 1) instead and the two taps v of a (phase, u) pair are ONE MMA of N = 128 — two
// overlapping 64-wide N atoms one pixel (128 B) apart — so that the MMA rows are never half empty.
// Output: fp32 partials [ksplit][16][Cin][Cout], folded to the 3x3 master's gradient by wgrad_reduce_upfold (wgrad.cu).
#include "host_common.cuh"

#include <cstdlib>
#include "ptx.cuh"
#include "wgrad.cuh"

namespace unetk {

int wgrad_upfold_launch(const float* partial, float* dw, int ksplit, int M, int Nn, int64_t sm, int64_t sn, int64_t st,
                        int accumulate, cudaStream_t stream);

namespace {

constexpr int kPix = 64;
constexpr int kThreads = 192;
constexpr int kMaxStages = 5;
constexpr uint32_t kPBoxBytes = kPix * 128;  // [64 px][64 ch] bf16

struct WUParams {
  CUtensorMap tmP;  // X : dims (M, W, H, N), box (64, TW, TH, 1)
  CUtensorMap tmQ;  // dY: dims (Nn, 2W, 2H, N), box (64, 2(TW+1), 2(TH+1), 1), element strides (1, 2, 2, 1)
  float* partial;   // [ksplit][16][M][Nn]
  int TH, TW, tiles_h, tiles_w, pix_tiles;
  int m_tiles, n_tiles, ksplit, stages;
  int M, Nn;
  uint32_t q_box_bytes;  // (TH+1)*(TW+1)*128 rounded up to 1024
  uint32_t q_tx_bytes;   // (TH+1)*(TW+1)*128 (what the TMA really writes)
};

template <int NT>
__global__ void __launch_bounds__(kThreads, 1) wgrad_up_kernel(const __grid_constant__ WUParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const uint32_t stage_bytes = 2 * kPBoxBytes + 2 * p.q_box_bytes;
  const int nstages = p.stages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + nstages * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kMaxStages;
  uint64_t* tfull_bar = bars + 2 * kMaxStages;   // [1]
  uint64_t* tempty_bar = tfull_bar + 1;          // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmP);
    tma_prefetch_desc(&p.tmQ);
    for (int s = 0; s < kMaxStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, 4);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();

  // item -> (n tile, m tile, phase group g, pixel split ks).  NT = 128: g = phase q (4 groups); NT = 64: g = qy (2 groups,
  // both qx inside the item, one n tile)
  constexpr int kGroups = (NT == 128) ? 4 : 2;
  const int items_per_split = kGroups * p.m_tiles * p.n_tiles;
  const int num_items = items_per_split * p.ksplit;
  auto decode = [&](int item, int& nt, int& mt, int& g, int& ks) {
    nt = item % p.n_tiles;
    mt = (item / p.n_tiles) % p.m_tiles;
    g = (item / (p.n_tiles * p.m_tiles)) % kGroups;
    ks = item / items_per_split;
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx = 2 * kPBoxBytes + 2 * p.q_tx_bytes;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        int nt, mt, g, ks;
        decode(item, nt, mt, g, ks);
        const int qy = (NT == 128) ? (g >> 1) : g;
        const int kt0 = static_cast<int>(static_cast<int64_t>(p.pix_tiles) * ks / p.ksplit);
        const int kt1 = static_cast<int>(static_cast<int64_t>(p.pix_tiles) * (ks + 1) / p.ksplit);
        for (int kt = kt0; kt < kt1; ++kt) {
          const int tw = kt % p.tiles_w;
          const int th = (kt / p.tiles_w) % p.tiles_h;
          const int img = kt / (p.tiles_w * p.tiles_h);
          const int h0 = th * p.TH, w0 = tw * p.TW;
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sp = smem + stage * stage_bytes;
          uint8_t* sq = sp + 2 * kPBoxBytes;
          mbar_expect_tx(&full_bar[stage], tx);
#pragma unroll
          for (int b = 0; b < 2; ++b)
            tma_load_4d(sp + b * kPBoxBytes, &p.tmP, &full_bar[stage], mt * 128 + b * 64, w0, h0, img);
          // phase pixel (ii, jj) is dY[2 ii + qy, 2 jj + qx]; the halo tile starts at phase pixel (h0 - qy, w0 - qx)
          if constexpr (NT == 128) {
            const int qx = g & 1;
#pragma unroll
            for (int b = 0; b < 2; ++b)
              tma_load_4d(sq + b * p.q_box_bytes, &p.tmQ, &full_bar[stage], nt * 128 + b * 64, 2 * w0 - qx, 2 * h0 - qy, img);
          } else {
#pragma unroll
            for (int qx = 0; qx < 2; ++qx)
              tma_load_4d(sq + qx * p.q_box_bytes, &p.tmQ, &full_bar[stage], 0, 2 * w0 - qx, 2 * h0 - qy, img);
          }
          if (++stage == nstages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    {
      // MMA issuer: warp-convergent loop, the elected lane issues (see umma_bf16_p in ptx.cuh)
      const bool issue = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      const int row_pitch = p.TW + 1;                 // halo pixels per image row in the Q tile
      const int steps_per_row = p.TW >> 4;            // 16-pixel k-steps per image row
      uint32_t q_off[kPix / 16];
#pragma unroll
      for (int k = 0; k < kPix / 16; ++k)
        q_off[k] = static_cast<uint32_t>((k / steps_per_row) * row_pitch + (k % steps_per_row) * 16) * 128;
      const uint32_t row_bytes = static_cast<uint32_t>(row_pitch) * 128;
      // P = two 64-channel boxes kPBoxBytes apart; Q = two 64-channel halo boxes q_box_bytes apart (NT = 128) or two
      // overlapping 64-wide atoms one pixel (128 B) apart (NT = 64: columns 0..63 = tap v = 1, 64..127 = tap v = 0)
      const uint64_t p_desc0 = make_smem_desc(smem_u32(smem), kPBoxBytes, 1024, kLayoutSW128);
      const uint64_t q_desc0 = make_smem_desc(smem_u32(smem) + 2 * kPBoxBytes, (NT == 64) ? 128u : p.q_box_bytes, 1024,
                                              kLayoutSW128);
      constexpr uint32_t idesc = make_idesc_bf16(128, 128, true, true);
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        int nt_, mt_, g_, ks;
        decode(item, nt_, mt_, g_, ks);
        const int kt0 = static_cast<int>(static_cast<int64_t>(p.pix_tiles) * ks / p.ksplit);
        const int kt1 = static_cast<int>(static_cast<int64_t>(p.pix_tiles) * (ks + 1) / p.ksplit);
        mbar_wait_p(issue, tempty_bar, (it & 1) ^ 1u);
        tc_fence_after();
        for (int kt = kt0; kt < kt1; ++kt) {
          mbar_wait_p(issue, &full_bar[stage], phase);
          tc_fence_after();
          const uint64_t dp0 = desc_advance(p_desc0, static_cast<uint32_t>(stage) * stage_bytes);
          const uint64_t dq0 = desc_advance(q_desc0, static_cast<uint32_t>(stage) * stage_bytes);
          const bool first = (kt == kt0);
#pragma unroll
          for (int k = 0; k < kPix / 16; ++k) {
            const uint64_t da = desc_advance(dp0, k * 2048);
            const uint64_t dq = desc_advance(dq0, q_off[k]);
            const uint32_t acc = (first && k == 0) ? 0u : 1u;
            if constexpr (NT == 128) {
              // tap (u, v): X pixel (i, j) meets phase pixel (i - (qy-1+u), j - (qx-1+v)) = halo (row 1-u, column 1-v) + (i-h0, j-w0)
#pragma unroll
              for (int t4 = 0; t4 < 4; ++t4) {
                const int u = t4 >> 1, v = t4 & 1;
                umma_bf16_p(issue, tmem_base + t4 * 128, da, desc_advance(dq, (1 - u) * row_bytes + (1 - v) * 128), idesc, acc);
              }
            } else {
              // accumulator a4 = qx*2 + u: both taps v of (phase (qy, qx), u) in one N = 128 MMA
#pragma unroll
              for (int a4 = 0; a4 < 4; ++a4) {
                const int qx = a4 >> 1, u = a4 & 1;
                umma_bf16_p(issue, tmem_base + a4 * 128, da, desc_advance(dq, qx * p.q_box_bytes + (1 - u) * row_bytes), idesc, acc);
              }
            }
          }
          umma_commit_p(issue, &empty_bar[stage]);
          if (++stage == nstages) { stage = 0; phase ^= 1u; }
        }
        umma_commit_p(issue, tfull_bar);
      }
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    int it = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      int nt, mt, g, ks;
      decode(item, nt, mt, g, ks);
      mbar_wait(tfull_bar, it & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
      const int m = mt * 128 + row;
#pragma unroll 1
      for (int a4 = 0; a4 < 4; ++a4) {
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          // NT = 128: accumulator a4 = tap (u, v) of phase g, 128 output channels of n tile nt
          // NT = 64 : accumulator a4 = (qx, u) of row phase g; columns 0..63 = tap v = 1, 64..127 = tap v = 0
          int tap16, col0;
          if constexpr (NT == 128) {
            tap16 = g * 4 + a4;
            col0 = nt * 128 + c * 32;
          } else {
            const int qx = a4 >> 1, u = a4 & 1, v = (c < 2) ? 1 : 0;
            tap16 = (g * 2 + qx) * 4 + u * 2 + v;
            col0 = (c & 1) * 32;
          }
          uint32_t r[32];
          tmem_ld32(taddr + a4 * 128 + c * 32, r);
          tmem_ld_wait();
          if (m < p.M) {
            float* dst = p.partial + ((static_cast<size_t>(ks) * 16 + tap16) * p.M + m) * p.Nn + col0;
#pragma unroll
            for (int x = 0; x < 8; ++x) {
              if (col0 + x * 4 < p.Nn) {
                float4 o = make_float4(__uint_as_float(r[x * 4]), __uint_as_float(r[x * 4 + 1]),
                                       __uint_as_float(r[x * 4 + 2]), __uint_as_float(r[x * 4 + 3]));
                *reinterpret_cast<float4*>(dst + x * 4) = o;
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

struct WUPlan {
  int NT, TH, TW, tiles_h, tiles_w, pix_tiles, m_tiles, n_tiles, ksplit, stages;
  uint32_t q_box_bytes, q_tx_bytes, smem_bytes;
};

bool make_plan(int N, int H, int W, int M, int Nn, WUPlan* pl) {
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("UNETK_WGRAD_UP"); enabled = e ? atoi(e) : 1; }
  if (!enabled || W < 16 || M % 8 || Nn % 8) return false;
  int tw = 64;
  while (tw > W) tw >>= 1;
  pl->TW = tw;
  pl->TH = kPix / tw;
  pl->NT = Nn > 64 ? 128 : 64;
  pl->tiles_h = (H + pl->TH - 1) / pl->TH;
  pl->tiles_w = (W + tw - 1) / tw;
  pl->pix_tiles = N * pl->tiles_h * pl->tiles_w;
  pl->m_tiles = (M + 127) / 128;
  pl->n_tiles = pl->NT == 128 ? (Nn + 127) / 128 : 1;
  pl->q_tx_bytes = static_cast<uint32_t>((pl->TH + 1) * (tw + 1) * 128);
  pl->q_box_bytes = (pl->q_tx_bytes + 1023u) & ~1023u;
  const uint32_t stage = 2 * kPBoxBytes + 2 * pl->q_box_bytes;
  int stages = static_cast<int>((227u * 1024u - 1024u - 256u) / stage);
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 3) return false;
  pl->stages = stages;
  pl->smem_bytes = stages * stage + 1024 + 256;
  // pixel split: as in wgrad3x3.cu — few (m, n) tiles: ~2 items per SM; otherwise the split that minimises
  // waves * (k-blocks per item + ~12 k-blocks of epilogue) + 2 * ksplit
  const int base = (pl->NT == 128 ? 4 : 2) * pl->m_tiles * pl->n_tiles;
  const int cap = pl->pix_tiles / 8 > 0 ? pl->pix_tiles / 8 : 1;
  if (base < 12) {
    // one item per SM (the items are equally long): half the partials of the "~2 items per SM" rule of wgrad3x3.cu, and
    // the fold-reduce over them was 20 % of the call (profiles/r02_ncu_upconv.txt)
    int ks = num_sms() / base;
    if (ks > cap) ks = cap;
    pl->ksplit = ks < 1 ? 1 : ks;
    return true;
  }
  int best = 1;
  long best_cost = -1;
  for (int ks = 1; ks <= cap && ks <= 48; ++ks) {
    const long waves = (static_cast<long>(base) * ks + num_sms() - 1) / num_sms();
    const long cost = waves * ((pl->pix_tiles + ks - 1) / ks + 12) + 2L * ks;
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = ks; }
  }
  pl->ksplit = best;
  return true;
}

template <int NT>
int launch(const WUParams& p, const WUPlan& pl, cudaStream_t stream) {
  static DeviceOnce once;
  UNETK_CUDA(once.run([] { return cudaFuncSetAttribute(wgrad_up_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); }));
  const int items = (NT == 128 ? 4 : 2) * pl.m_tiles * pl.n_tiles * pl.ksplit;
  const int grid = items < num_sms() ? items : num_sms();
  UNETK_CUDA(launch_pdl(wgrad_up_kernel<NT>, dim3(grid), dim3(kThreads), pl.smem_bytes, stream, p));
  UNETK_LAUNCHED();
  return 0;
}

}  // namespace

// 0 when this kernel does not apply to the shape (the caller falls back to the per-tap kernel)
size_t wgrad_up_workspace_bytes(int N, int H, int W, int Cin, int Cout) {
  WUPlan pl;
  if (!make_plan(N, H, W, Cin, Cout, &pl)) return 0;
  return static_cast<size_t>(pl.ksplit) * 16 * Cin * Cout * sizeof(float);
}

// dw[co][ci][3][3] (+)= the folded sub-filter gradients; x = [N,H,W,Cin] (low resolution), dy = [N,2H,2W,Cout].
// Returns 1 if the shape is not eligible.
int wgrad_up_run(const void* x, int64_t x_ld, const void* dy, int64_t dy_ld, float* dw, int accumulate, int N, int H, int W,
                 int Cin, int Cout, void* workspace, size_t ws_bytes, cudaStream_t stream) {
  WUPlan pl;
  if (!make_plan(N, H, W, Cin, Cout, &pl)) return 1;
  UNETK_CHECK(x_ld % 8 == 0 && dy_ld % 8 == 0, -1, "wgrad_up: pixel strides must be multiples of 8");
  const size_t need = static_cast<size_t>(pl.ksplit) * 16 * Cin * Cout * sizeof(float);
  UNETK_CHECK(workspace != nullptr && ws_bytes >= need, -1, "wgrad_up: workspace too small (%zu < %zu)", ws_bytes, need);
  WUParams p{};
  p.partial = static_cast<float*>(workspace);
  p.TH = pl.TH; p.TW = pl.TW; p.tiles_h = pl.tiles_h; p.tiles_w = pl.tiles_w; p.pix_tiles = pl.pix_tiles;
  p.m_tiles = pl.m_tiles; p.n_tiles = pl.n_tiles; p.ksplit = pl.ksplit; p.stages = pl.stages;
  p.M = Cin; p.Nn = Cout;
  p.q_box_bytes = pl.q_box_bytes; p.q_tx_bytes = pl.q_tx_bytes;
  {
    uint64_t dims[4] = {static_cast<uint64_t>(Cin), static_cast<uint64_t>(W), static_cast<uint64_t>(H), static_cast<uint64_t>(N)};
    uint64_t strides[3] = {static_cast<uint64_t>(x_ld) * 2, static_cast<uint64_t>(x_ld) * 2 * W,
                           static_cast<uint64_t>(x_ld) * 2 * W * H};
    uint32_t box[4] = {64, static_cast<uint32_t>(pl.TW), static_cast<uint32_t>(pl.TH), 1};
    uint32_t es[4] = {1, 1, 1, 1};
    if (int rc = make_tmap_bf16(&p.tmP, x, 4, dims, strides, box, es, true)) return rc;
  }
  {
    const uint64_t W2 = 2ull * W, H2 = 2ull * H;
    uint64_t dims[4] = {static_cast<uint64_t>(Cout), W2, H2, static_cast<uint64_t>(N)};
    uint64_t strides[3] = {static_cast<uint64_t>(dy_ld) * 2, static_cast<uint64_t>(dy_ld) * 2 * W2,
                           static_cast<uint64_t>(dy_ld) * 2 * W2 * H2};
    uint32_t box[4] = {64, static_cast<uint32_t>(2 * (pl.TW + 1)), static_cast<uint32_t>(2 * (pl.TH + 1)), 1};
    uint32_t es[4] = {1, 2, 2, 1};
    if (int rc = make_tmap_bf16(&p.tmQ, dy, 4, dims, strides, box, es, true)) return rc;
  }
  int rc = (pl.NT == 128) ? launch<128>(p, pl, stream) : launch<64>(p, pl, stream);
  if (rc) return rc;
  // rows = input channels (stride 9), columns = output channels (stride Cin*9): the 3x3 master's gradient [Cout][Cin][3][3]
  return wgrad_upfold_launch(p.partial, dw, pl.ksplit, Cin, Cout, 9, static_cast<int64_t>(Cin) * 9, 1, accumulate, stream);
}

}  // namespace unetk
