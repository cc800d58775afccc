// 3x3 weight gradient, second generation: ONE halo load of the activations feeds the three taps of a
// filter row.
//
//   dW[co][ci][r][s] = sum_{n,h,w} dY[n,h,w,co] * X[n,h+r-1,w+s-1,ci]
//
// Work item = (pixel range, filter row r, 128 output channels, NT input channels).  Per 64-pixel k-block
// (TH x TW pixels, TW in {16,32,64}) the CTA loads
//   P: dY tile            box (64ch, TW,   TH)  -> [64 px][64 ch]   x 2 (M = 128)
//   Q: X halo of row r    box (64ch, TW+2, TH)  -> [TH*(TW+2) px][64 ch] per 64 input channels
// and issues, for every tap s = 0,1,2, MMAs whose B descriptor simply STARTS s pixel-rows (s*128 B)
// further into the same halo tile: tcgen05 applies the 128B-swizzle XOR on absolute smem address bits
// (probe: profiles/r01_umma_descriptor_probe.txt), so a row-shifted start is exact.  For Cin tiles of 64
// the three taps are even fused into ONE MMA of N = 192 by giving the descriptor a leading-dimension
// stride of 128 B (three overlapping 64-wide atoms).
// Versus the first kernel (one tap per work item, wgrad.cu) this cuts the operand traffic per FLOP by ~3x:
// that kernel re-read dY and X nine times and was HBM-bound on the 512^2 layers (138 TFLOP/s, 64->64).
// Accumulators: 3 taps x NT fp32 columns in TMEM; fp32 partials + ordered reduce (deterministic).
#include "host_common.cuh"

#include <cstdlib>
#include <mutex>
#include "ptx.cuh"
#include "wgrad.cuh"

namespace unetk {

int wgrad_reduce_launch(const float* partial, float* dw, int ksplit, int taps, int M, int Nn, int64_t sm, int64_t sn,
                        int64_t st, int accumulate, cudaStream_t stream);

namespace {

constexpr int kPix = 64;
constexpr int kThreads = 192;
constexpr uint32_t kPBoxBytes = kPix * 128;  // [64 px][64 ch] bf16

struct W3Params {
  CUtensorMap tmP;  // dY: dims (M, W, H, N), box (64, TW, TH, 1)
  CUtensorMap tmQ;  // X : dims (Nn, W, H, N), box (64, TW+2, TH, 1)
  CUtensorMap tmP2; // paired mode: dY box (64, 64, 2 or 3, 1) = image rows, one per filter row
  int paired;       // 1: M <= 64 -> the two halves of the 128 MMA rows carry two FILTER ROWS (see make_plan);
                    // 2: the same with BOTH row groups in one work item (three dY rows per X row, two accumulators)
  int stages;       // pipeline depth actually used (<= W3Cfg::kStages)
  int rows2;        // 1 (NT = 64, TH = 1, not paired): filter rows 0 and 1 share one pass (two halo rows of Q, two
                    // accumulators), filter row 2 runs alone: the first operand is read twice instead of three times
                    // 2: the nine taps split 5 + 4 over two EQUALLY long items that run side by side (g = 0: filter row 0
                    // + taps 0, 1 of row 1; g = 1: tap 2 of row 1 + filter row 2): the 128 x 576 fp32 gradient does not fit
                    // one SM's TMEM (512 columns), so two CTAs must sweep the pixels — but in step, on neighbouring SMs,
                    // so that the second one finds the first operand in L2 (the 1 + 1/2 split re-read it from HBM:
                    // 3.26 GB of DRAM traffic for 1.61 GB of operands, profiles/r02_ncu_tensor_pipe.txt)
  float* partial;   // [ksplit][9][M][Nn]
  int TH, TW, tiles_h, tiles_w, pix_tiles;
  int m_tiles, n_tiles, ksplit;
  int M, Nn;
  uint32_t q_box_bytes;  // TH*(TW+2)*128 rounded up to 1024
  uint32_t q_tx_bytes;   // TH*(TW+2)*128 (what the TMA really writes)
};

template <int NT>
struct W3Cfg {
  static constexpr int kQBoxes = NT / 64;
  static constexpr int kStages = (NT == 128) ? 5 : 7;
  static constexpr uint32_t kTmemCols = 512;   // NT = 64 merged mode: two 192-column accumulators
};

// PAIR (NT = 64, rows2 == 2 only): launched as clusters of two CTAs = the two halves (g = 0 / 1) of a pixel range.  What
// both halves read — the first operand's tile and the middle halo row — is fetched from L2 ONCE per pair and multicast
// into both CTAs' shared memory (each CTA loads one of the two 64-channel boxes; the halo row alternates), which takes
// the L2 reads per CTA from 32.5 to ~20.5 KB per k-block (measured: L2 traffic 5.5 -> 4.4 GB per launch, time unchanged,
// profiles/r02_ncu_wgrad_pair.txt: the limit is shared-memory bandwidth, see launch_pair).  A stage is free again when the
// MMAs of BOTH CTAs have retired (multicast commit onto both empty barriers).
template <int NT, bool PAIR>
__global__ void __launch_bounds__(kThreads, 1) wgrad3x3_kernel(const __grid_constant__ W3Params p) {
  using C = W3Cfg<NT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const bool merged = p.paired == 2;
  const uint32_t p_boxes = merged ? 3u : 2u;
  const bool rows2 = p.rows2 != 0;
  const bool balanced = p.rows2 == 2;
  const uint32_t q_boxes = rows2 ? 2u : static_cast<uint32_t>(C::kQBoxes);
  const uint32_t stage_bytes = p_boxes * kPBoxBytes + q_boxes * p.q_box_bytes;
  const int nstages = p.stages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + nstages * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + C::kStages;
  uint64_t* tfull_bar = bars + 2 * C::kStages;   // [1]
  uint64_t* tempty_bar = tfull_bar + 1;          // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmP);
    tma_prefetch_desc(&p.tmQ);
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], PAIR ? 2 : 1);
    }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, 4);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<C::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  uint32_t rank = 0;
  if constexpr (PAIR) {
    rank = cluster_ctarank();   // == blockIdx.x & 1 == g of every item of this CTA (even grid, balanced decode)
    cluster_sync_all();         // the peer's barriers exist before anything of ours can land on them
  }
  // programmatic dependent launch: everything above overlapped the previous kernel's tail; from here on global memory
  pdl_trigger();
  pdl_wait();

  const int items_per_split = merged ? p.n_tiles : (p.paired ? 2 * p.n_tiles : (rows2 ? 2 : 3) * p.m_tiles * p.n_tiles);
  // item -> (n tile, m tile, filter row or row group r, pixel split ks, g).  rows2: the long items (g = 0: filter rows
  // 0 + 1) come first, then the short ones (g = 1: row 2), so that a CTA striding over the items gets one of each
  auto decode = [&](int item, int& nt, int& mt, int& r, int& ks, int& g) {
    if (balanced) {   // the two halves of a pixel range on neighbouring CTAs
      g = item & 1;
      const int rem = item >> 1, mn = p.m_tiles * p.n_tiles;
      ks = rem / mn;
      const int t = rem - ks * mn;
      mt = t / p.n_tiles;
      nt = t - mt * p.n_tiles;
      r = g ? 2 : 0;
    } else if (rows2) {
      const int mn = p.m_tiles * p.n_tiles, per_g = mn * p.ksplit;
      g = item / per_g;
      const int rem = item - g * per_g;
      ks = rem / mn;
      const int t = rem - ks * mn;
      mt = t / p.n_tiles;
      nt = t - mt * p.n_tiles;
      r = g ? 2 : 0;
    } else {
      nt = item % p.n_tiles;
      mt = p.paired ? 0 : (item / p.n_tiles) % p.m_tiles;
      r = p.paired ? (item / p.n_tiles) % 2 : (item / (p.n_tiles * p.m_tiles)) % 3;
      ks = item / items_per_split;
      g = 0;
    }
  };
  const int num_items = items_per_split * p.ksplit;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        // paired: r is the filter-row GROUP (0: rows 1|0, 1: row 2) and the X row is the k-block's own row
        int nt, mt, r, ks, g;
        decode(item, nt, mt, r, ks, g);
        const uint32_t tx = p_boxes * kPBoxBytes + (rows2 ? ((g == 0 || balanced) ? 2u : 1u) : static_cast<uint32_t>(C::kQBoxes)) * p.q_tx_bytes;
        const int kt0 = static_cast<int>(static_cast<int64_t>(p.pix_tiles) * ks / p.ksplit);
        const int kt1 = static_cast<int>(static_cast<int64_t>(p.pix_tiles) * (ks + 1) / p.ksplit);
        for (int kt = kt0; kt < kt1; ++kt) {
          const int tw = kt % p.tiles_w;
          const int th = (kt / p.tiles_w) % p.tiles_h;
          const int img = kt / (p.tiles_w * p.tiles_h);
          const int h0 = th * p.TH, w0 = tw * p.TW;
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sp = smem + stage * stage_bytes;
          uint8_t* sq = sp + p_boxes * kPBoxBytes;
          mbar_expect_tx(&full_bar[stage], tx);
          if (p.paired) {
            // X row q = h0 meets dY row q-r+1 under filter row r: group 0 loads dY rows (q, q+1) = filter rows (1, 0),
            // group 1 loads rows (q-1, q) = filter rows (2, 1 [discarded]); rows outside the image read as zero
            // merged: dY rows (q-1, q, q+1) in one box; the two MMAs of a k-step read rows (q, q+1) and (q-1, q)
            tma_load_4d(sp, &p.tmP2, &full_bar[stage], 0, w0, merged ? h0 - 1 : h0 - r, img);
#pragma unroll
            for (int b = 0; b < C::kQBoxes; ++b)
              tma_load_4d(sq + b * p.q_box_bytes, &p.tmQ, &full_bar[stage], nt * NT + b * 64, w0 - 1, h0, img);
            if (++stage == nstages) { stage = 0; phase ^= 1u; }
            continue;
          }
          if constexpr (PAIR) {
            // tx (above) counts what LANDS in this CTA: both boxes of the first operand (one from the peer), the private
            // halo row and the shared one (from whichever CTA's turn it is)
            tma_load_4d_mc(sp + rank * kPBoxBytes, &p.tmP, &full_bar[stage], mt * 128 + static_cast<int>(rank) * 64, w0, h0, img, 0x3);
            tma_load_4d(sq, &p.tmQ, &full_bar[stage], nt * NT, w0 - 1, h0 + r - 1, img);
            if (((static_cast<uint32_t>(kt) ^ rank) & 1u) == 0)
              tma_load_4d_mc(sq + p.q_box_bytes, &p.tmQ, &full_bar[stage], nt * NT, w0 - 1, h0, img, 0x3);
            if (++stage == nstages) { stage = 0; phase ^= 1u; }
            continue;
          }
#pragma unroll
          for (int b = 0; b < 2; ++b)
            tma_load_4d(sp + b * kPBoxBytes, &p.tmP, &full_bar[stage], mt * 128 + b * 64, w0, h0, img);
          if (rows2) {   // NT == 64: one 64-channel box per halo row; g == 0: rows h0-1 (filter row 0) and h0 (row 1)
            tma_load_4d(sq, &p.tmQ, &full_bar[stage], nt * NT, w0 - 1, h0 + r - 1, img);
            if (g == 0 || balanced) tma_load_4d(sq + p.q_box_bytes, &p.tmQ, &full_bar[stage], nt * NT, w0 - 1, h0, img);
          } else {
#pragma unroll
            for (int b = 0; b < C::kQBoxes; ++b)
              tma_load_4d(sq + b * p.q_box_bytes, &p.tmQ, &full_bar[stage], nt * NT + b * 64, w0 - 1, h0 + r - 1, img);
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
      const int row_pitch = p.TW + 2;                 // halo pixels per image row in the Q tile
      const int steps_per_row = p.TW >> 4;            // 16-pixel k-steps per image row
      // stage-0 descriptors: P = two 64-channel boxes kPBoxBytes apart; Q = 64-channel halo boxes q_box_bytes apart,
      // or (NT == 64) three overlapping tap atoms one pixel row (128 B) apart
      uint32_t q_off[kPix / 16];
#pragma unroll
      for (int k = 0; k < kPix / 16; ++k)
        q_off[k] = static_cast<uint32_t>((k / steps_per_row) * row_pitch + (k % steps_per_row) * 16) * 128;
      const uint64_t p_desc0 = make_smem_desc(smem_u32(smem), kPBoxBytes, 1024, kLayoutSW128);
      const uint64_t q_desc0 = make_smem_desc(smem_u32(smem) + p_boxes * kPBoxBytes, (NT == 64) ? 128u : p.q_box_bytes, 1024,
                                              kLayoutSW128);
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        int nt_, mt_, r_, ks, g;
        decode(item, nt_, mt_, r_, ks, g);
        const bool two_rows = rows2 && g == 0;
        // balanced: the second MMA of a k-step takes taps 0, 1 (g = 0: N = 128) or tap 2 (g = 1: N = 64, two pixels in)
        // of filter row 1 from the second halo row
        const uint32_t idesc_b = (g == 0) ? make_idesc_bf16(128, 128, true, true) : make_idesc_bf16(128, 64, true, true);
        const uint32_t shift_b = p.q_box_bytes + ((g == 0) ? 0u : 256u);
        const int kt0 = static_cast<int>(static_cast<int64_t>(p.pix_tiles) * ks / p.ksplit);
        const int kt1 = static_cast<int>(static_cast<int64_t>(p.pix_tiles) * (ks + 1) / p.ksplit);
        mbar_wait_p(issue, tempty_bar, (it & 1) ^ 1u);
        tc_fence_after();
        for (int kt = kt0; kt < kt1; ++kt) {
          mbar_wait_p(issue, &full_bar[stage], phase);
          tc_fence_after();
          // descriptors are advanced from per-kernel bases (few instructions per MMA on the issuing thread)
          const uint64_t dp0 = desc_advance(p_desc0, static_cast<uint32_t>(stage) * stage_bytes);
          const uint64_t dq0 = desc_advance(q_desc0, static_cast<uint32_t>(stage) * stage_bytes);
          const bool first = (kt == kt0);
#pragma unroll
          for (int k = 0; k < kPix / 16; ++k) {
            // 16 consecutive pixels of one image row: P rows k*16.., Q rows i*(TW+2) + j0 (+ s per tap)
            const uint64_t da = desc_advance(dp0, k * 2048);
            const uint64_t dq = desc_advance(dq0, q_off[k]);
            if constexpr (NT == 64) {
              // taps s = 0,1,2 as three overlapping 64-wide N atoms, 128 B (one pixel row) apart (LBO = 128 B)
              constexpr uint32_t idesc = make_idesc_bf16(128, 192, true, true);
              if (merged) {   // uniform per launch
                // rows (q, q+1) -> filter rows (1, 0) in columns 0..191; rows (q-1, q) -> filter rows (2, [1]) in 192..383
                umma_bf16_p(issue, tmem_base, desc_advance(da, kPBoxBytes), dq, idesc, (first && k == 0) ? 0u : 1u);
                umma_bf16_p(issue, tmem_base + 192, da, dq, idesc, (first && k == 0) ? 0u : 1u);
              } else {
                umma_bf16_p(issue, tmem_base, da, dq, idesc, (first && k == 0) ? 0u : 1u);   // no control flow per MMA
                if (balanced)   // uniform per launch
                  umma_bf16_p(issue, tmem_base + 192, da, desc_advance(dq, shift_b), idesc_b, (first && k == 0) ? 0u : 1u);
                else if (two_rows)   // uniform per item: the second halo row (filter row 1) against the same first operand
                  umma_bf16_p(issue, tmem_base + 192, da, desc_advance(dq, p.q_box_bytes), idesc, (first && k == 0) ? 0u : 1u);
              }
            } else {
              constexpr uint32_t idesc = make_idesc_bf16(128, NT, true, true);
#pragma unroll
              for (int s = 0; s < 3; ++s) {
                umma_bf16_p(issue, tmem_base + s * NT, da, desc_advance(dq, s * 128), idesc, (first && k == 0) ? 0u : 1u);
              }
            }
          }
          if constexpr (PAIR) umma_commit_mc_p(issue, &empty_bar[stage], 0x3);   // frees the slot in BOTH CTAs
          else umma_commit_p(issue, &empty_bar[stage]);
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
      int nt, mt, r_item, ks, g;
      decode(item, nt, mt, r_item, ks, g);
      mbar_wait(tfull_bar, it & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll 1
      for (int a = 0; a < ((merged || balanced || (rows2 && g == 0)) ? 2 : 1); ++a) {
        // merged: accumulator 0 = row group 0, accumulator 1 = group 1; rows2 (g == 0): accumulator a = filter row a
        // balanced: accumulator 0 = filter row 0 (g = 0) or 2 (g = 1), accumulator 1 = taps [0, 2) / [2, 3) of filter row 1
        int r = merged ? a : (balanced ? (a ? 1 : r_item) : (rows2 && g == 0 ? a : r_item));
        const int s_lo = (balanced && a == 1 && g == 1) ? 2 : 0, s_hi = (balanced && a == 1 && g == 0) ? 2 : 3;
        int m = mt * 128 + row;
        if (p.paired) {
          // accumulator rows 0-63 = first dY row of the pair, 64-127 = second: group 0 -> filter rows (1, 0), group 1 -> (2, -)
          const int half = row >> 6;
          m = row & 63;
          if (r == 0) r = 1 - half;
          else { r = 2; if (half) m = p.M; }   // second half of group 1 repeats filter row 1: not stored
        }
#pragma unroll 1
        for (int s = s_lo; s < s_hi; ++s) {
          float* dst = p.partial + ((static_cast<size_t>(ks) * 9 + r * 3 + s) * p.M + m) * p.Nn + nt * NT;
#pragma unroll 1
          for (int c = 0; c < NT / 32; ++c) {
            uint32_t v[32];
            tmem_ld32(taddr + a * 192 + (s - s_lo) * NT + c * 32, v);
            tmem_ld_wait();
            if (m < p.M) {
#pragma unroll
              for (int x = 0; x < 8; ++x) {
                const int col = nt * NT + c * 32 + x * 4;
                if (col < p.Nn) {
                  float4 o = make_float4(__uint_as_float(v[x * 4]), __uint_as_float(v[x * 4 + 1]),
                                         __uint_as_float(v[x * 4 + 2]), __uint_as_float(v[x * 4 + 3]));
                  *reinterpret_cast<float4*>(dst + c * 32 + x * 4) = o;
                }
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
    tmem_dealloc<C::kTmemCols>(tmem_base);
  }
  // a CTA's shared memory (barriers the peer's commits arrive on, slots its loads write) must outlive the peer's work
  if constexpr (PAIR) cluster_sync_all();
}

struct W3Plan {
  int NT, TH, TW, tiles_h, tiles_w, pix_tiles, m_tiles, n_tiles, ksplit, paired, stages, rows2;
  uint32_t q_box_bytes, q_tx_bytes, smem_bytes;
};

bool make_plan(int N, int H, int W, int M, int Nn, W3Plan* pl) {
  if (W < 16 || M % 8 || Nn % 8) return false;
  int tw = 64;
  while (tw > W) tw >>= 1;
  pl->TW = tw;
  pl->TH = kPix / tw;
  pl->NT = Nn > 64 ? 128 : 64;
  pl->tiles_h = (H + pl->TH - 1) / pl->TH;
  pl->tiles_w = (W + tw - 1) / tw;
  pl->pix_tiles = N * pl->tiles_h * pl->tiles_w;
  pl->m_tiles = (M + 127) / 128;
  pl->n_tiles = (Nn + pl->NT - 1) / pl->NT;
  pl->q_tx_bytes = static_cast<uint32_t>(pl->TH * (tw + 2) * 128);
  pl->q_box_bytes = (pl->q_tx_bytes + 1023u) & ~1023u;
  static int pair_env = -1, merge_env = -1;
  if (pair_env < 0) { const char* e = getenv("UNETK_WGRAD3_PAIR"); pair_env = e ? atoi(e) : 1; }
  if (merge_env < 0) { const char* e = getenv("UNETK_WGRAD3_MERGE"); merge_env = e ? atoi(e) : 1; }
  pl->paired = (pair_env && M <= 64 && pl->NT == 64 && pl->TW == 64 && pl->TH == 1) ? (merge_env ? 2 : 1) : 0;
  // merged: both filter-row groups in one pass over the pixels.  The 64-channel layers are bound by the L2 -> SM
  // fabric (3.25 GB through the crossbar per 64 -> 64 launch = 9 TB/s at 60 % tensor-pipe activity,
  // profiles/r01_ncu_top_kernels_v6.txt): one X row + THREE dY rows per k-block (33.8 KB per 768 MMA clocks) instead of
  // one X row + two dY rows per group (2 x 24.6 KB)
  // rows2: more than 64 rows against <= 64 columns (swapped 128 -> 64, UNet++'s 96..192 -> 32): same fabric bound,
  // the first operand (16 KB per k-block) was read once per filter row; filter rows 0 and 1 now share it
  static int rows2_env = -1;
  if (rows2_env < 0) { const char* e = getenv("UNETK_WGRAD3_ROWS2"); rows2_env = e ? atoi(e) : 2; }
  pl->rows2 = (rows2_env && !pl->paired && pl->NT == 64 && pl->TW == 64 && pl->TH == 1) ? (rows2_env == 1 ? 1 : 2) : 0;
  const int stages = pl->NT == 128 ? 5 : ((pl->paired == 2 || pl->rows2) ? 6 : 7);
  pl->stages = stages;
  const uint32_t stage = (pl->paired == 2 ? 3 : 2) * kPBoxBytes + (pl->rows2 ? 2 : pl->NT / 64) * pl->q_box_bytes;
  pl->smem_bytes = stages * stage + 1024 + 256;
  if (pl->smem_bytes > 227 * 1024) return false;
  // items = 3 * m_tiles * n_tiles * ksplit run in waves of num_sms CTAs.  Pick the pixel split that minimises
  //   waves * (k-blocks per item + ~12 k-blocks of epilogue) + 2 * ksplit (partial write + reduce),
  // at least 8 k-blocks per item.  (1024 -> 1024 @32^2, 192 base items: the old "~2 items per SM" rule chose 1 split =
  // two waves at 65 % occupancy; 3 splits make 3.9 waves.)
  // M <= 64 (64-channel dY): half of the 128 MMA rows would be empty (59 % tensor-pipe activity for 30 % useful work,
  // profiles/r01_ncu_wgrad3x3.txt).  With one image row per k-block (TW = 64) the second half is given the NEXT dY row,
  // i.e. another filter row against the same X halo row: 2 work items per pixel range instead of 3.
  const int base = pl->paired == 2 ? pl->n_tiles : (pl->paired ? 2 * pl->n_tiles : (pl->rows2 ? 2 : 3) * pl->m_tiles * pl->n_tiles);
  const int cap = pl->pix_tiles / 8 > 0 ? pl->pix_tiles / 8 : 1;
  static int rule = -1;
  if (rule < 0) { const char* e = getenv("UNETK_WGRAD3_RULE"); rule = e ? atoi(e) : 0; }
  if (base < 12 && rule == 0) {
    // few (m, n) tiles = the wide-image thin-channel layers (L2 -> SM bound): ~2 items per SM.  The cost rule below
    // would pick one wave of items twice as long; measured on the same box it makes no difference there
    // (UNet step: 5.79-5.81 vs 5.61-5.80 ms over the 17 weight gradients), so these keep the simple rule.
    int ks = (2 * num_sms()) / (pl->paired == 2 ? 2 * base : base);   // merged items are twice as long: one wave
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
int launch(const W3Params& p, const W3Plan& pl, cudaStream_t stream) {
  static DeviceOnce once;
  UNETK_CUDA(once.run([] { return cudaFuncSetAttribute(wgrad3x3_kernel<NT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); }));
  const int items = (pl.paired == 2 ? pl.n_tiles : (pl.paired ? 2 * pl.n_tiles : (pl.rows2 ? 2 : 3) * pl.m_tiles * pl.n_tiles)) * pl.ksplit;
  const int grid = items < num_sms() ? items : num_sms();
  UNETK_CUDA(launch_pdl(wgrad3x3_kernel<NT, false>, dim3(grid), dim3(kThreads), pl.smem_bytes, stream, p));
  UNETK_LAUNCHED();
  return 0;
}

// CTA pairs the device can keep resident with this kernel's shared memory (GPCs with an odd number of free SMs leave one
// out): cached per device
int max_pairs(uint32_t smem_bytes) {
  static std::mutex mu;
  static int cached[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  std::lock_guard<std::mutex> lk(mu);
  if (cached[dev & 63] == 0) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * 148);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, wgrad3x3_kernel<64, true>, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); n = -1; }
    cached[dev & 63] = n;
  }
  return cached[dev & 63];
}

// rows2 == 2 as clusters of two CTAs (the g = 0 / g = 1 halves of a pixel range) with multicast loads; > 0: not launched
int launch_pair(const W3Params& p, const W3Plan& pl, cudaStream_t stream) {
  static int env = -1;
  // Off by default: measured neutral (0.571 -> 0.562 ms on 128 -> 64 @512^2, L2 traffic 5.5 -> 4.4 GB, tensor pipe 57 % either
  // way, profiles/r02_ncu_wgrad_pair.txt) — the kernel is bound by shared-memory bandwidth (MMA operand reads + TMA fills,
  // ~166 B/clk wanted of ~128), which a multicast does not lower: every CTA still receives all its bytes.
  if (env < 0) { const char* e = getenv("UNETK_WGRAD3_CLUSTER"); env = e ? atoi(e) : 0; }
  if (!env) return 1;
  static DeviceOnce once;
  UNETK_CUDA(once.run([] { return cudaFuncSetAttribute(wgrad3x3_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); }));
  const int pairs_fit = max_pairs(pl.smem_bytes);
  if (pairs_fit <= 0) return 1;
  const int pair_items = pl.m_tiles * pl.n_tiles * pl.ksplit;
  int pairs = num_sm_pairs();
  if (pairs > pairs_fit) pairs = pairs_fit;
  if (pairs > pair_items) pairs = pair_items;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = pl.smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  UNETK_CUDA(cudaLaunchKernelEx(&cfg, wgrad3x3_kernel<64, true>, p));
  UNETK_LAUNCHED();
  return 0;
}

}  // namespace

// 0 when this kernel does not apply to the shape (caller falls back to the per-tap kernel)
size_t wgrad3x3_workspace_bytes(int N, int H, int W, int M, int Nn) {
  W3Plan pl;
  if (!make_plan(N, H, W, M, Nn, &pl)) return 0;
  return static_cast<size_t>(pl.ksplit) * 9 * M * Nn * sizeof(float);
}

// dw[co][ci][3][3] (+)= dY (M = Cout channels) x X (Nn = Cin channels); returns 1 if the shape is not eligible.
// flip != 0: the operands are passed SWAPPED (first = X with M = Cin rows, second = dY with Nn = Cout columns), which
// computes tap (2-r, 2-s) of the transposed filter: sum_px X[px] dY[px + d] = sum_px' dY[px'] X[px' - d].  Used when
// Cout = 64 < Cin: the 128 MMA rows are then all real input channels instead of 64 output channels + 64 empty rows.
int wgrad3x3_run(const void* dy, int64_t dy_ld, const void* x, int64_t x_ld, float* dw, int accumulate, int N, int H,
                 int W, int M, int Nn, void* workspace, size_t ws_bytes, cudaStream_t stream, int flip) {
  W3Plan pl;
  if (!make_plan(N, H, W, M, Nn, &pl)) return 1;
  UNETK_CHECK(dy_ld % 8 == 0 && x_ld % 8 == 0, -1, "wgrad3x3: pixel strides must be multiples of 8");
  const size_t need = static_cast<size_t>(pl.ksplit) * 9 * M * Nn * sizeof(float);
  UNETK_CHECK(workspace != nullptr && ws_bytes >= need, -1, "wgrad3x3: workspace too small (%zu < %zu)", ws_bytes, need);
  W3Params p{};
  p.partial = static_cast<float*>(workspace);
  p.TH = pl.TH; p.TW = pl.TW; p.tiles_h = pl.tiles_h; p.tiles_w = pl.tiles_w; p.pix_tiles = pl.pix_tiles;
  p.m_tiles = pl.m_tiles; p.n_tiles = pl.n_tiles; p.ksplit = pl.ksplit;
  p.M = M; p.Nn = Nn;
  p.q_box_bytes = pl.q_box_bytes; p.q_tx_bytes = pl.q_tx_bytes;
  p.paired = pl.paired;
  p.stages = pl.stages;
  p.rows2 = pl.rows2;
  auto mk = [&](CUtensorMap* tm, const void* base, int64_t ld, int C, int halo) -> int {
    uint64_t dims[4] = {static_cast<uint64_t>(C), static_cast<uint64_t>(W), static_cast<uint64_t>(H),
                        static_cast<uint64_t>(N)};
    uint64_t strides[3] = {static_cast<uint64_t>(ld) * 2, static_cast<uint64_t>(ld) * 2 * W,
                           static_cast<uint64_t>(ld) * 2 * W * H};
    uint32_t box[4] = {64, static_cast<uint32_t>(pl.TW + halo), static_cast<uint32_t>(pl.TH), 1};
    uint32_t es[4] = {1, 1, 1, 1};
    return make_tmap_bf16(tm, base, 4, dims, strides, box, es, true);
  };
  if (int rc = mk(&p.tmP, dy, dy_ld, M, 0)) return rc;
  if (pl.paired) {
    uint64_t dims[4] = {static_cast<uint64_t>(M), static_cast<uint64_t>(W), static_cast<uint64_t>(H), static_cast<uint64_t>(N)};
    uint64_t strides[3] = {static_cast<uint64_t>(dy_ld) * 2, static_cast<uint64_t>(dy_ld) * 2 * W,
                           static_cast<uint64_t>(dy_ld) * 2 * W * H};
    uint32_t box[4] = {64, 64, pl.paired == 2 ? 3u : 2u, 1};
    uint32_t es[4] = {1, 1, 1, 1};
    if (int rc = make_tmap_bf16(&p.tmP2, dy, 4, dims, strides, box, es, true)) return rc;
  }
  if (int rc = mk(&p.tmQ, x, x_ld, Nn, 2)) return rc;
  int rc = 1;
  if (pl.rows2 == 2) rc = launch_pair(p, pl, stream);
  if (rc > 0) rc = (pl.NT == 128) ? launch<128>(p, pl, stream) : launch<64>(p, pl, stream);
  if (rc) return rc;
  if (flip)   // rows are input channels (stride 9), columns output channels (stride Cin*9), taps reversed
    return wgrad_reduce_launch(p.partial, dw + 8, pl.ksplit, 9, M, Nn, 9, static_cast<int64_t>(M) * 9, -1, accumulate, stream);
  return wgrad_reduce_launch(p.partial, dw, pl.ksplit, 9, M, Nn, static_cast<int64_t>(Nn) * 9, 9, 1, accumulate, stream);
}

}  // namespace unetk
