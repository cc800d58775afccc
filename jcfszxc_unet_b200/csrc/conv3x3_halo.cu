// conv3x3 (fwd / dgrad) for the wide-image, thin-channel layers: W >= 128 and at most 128 output channels.
//
// Why a second kernel.  The tap-GEMM in conv_gemm.cu loads a fresh 16 KB A tile for every (tap, 64-channel
// chunk); with N = 64 that is 24 KB of TMA traffic per 128-192 MMA cycles, and every counter in
// profiles/r01_ncu_conv_thin_layers.txt sits at ~30% (tensor 22%, L2 31%, DRAM 20%): the TMA/L2 path
// delivers ~40 B/clk/SM and the layer is bound by bytes per FLOP.  Here one CTA tile is TWO output rows x
// 128 columns and the four input rows it touches (with a one-pixel halo left and right) are loaded ONCE per
// 64-channel chunk; all nine taps of both rows read that halo through row-shifted UMMA descriptors
// (tcgen05 swizzles on absolute smem address bits, profiles/r01_umma_descriptor_probe.txt).  A-operand
// traffic drops 4.3x (66.5 KB instead of 18 x 16 KB) and each weight tile feeds two accumulators.
//
// Pipeline: warp0 = TMA producer (A halo ring of 2, B ring per tap), warp1 = MMA issuer, warps2-5 =
// epilogue (TMEM -> bf16 -> swizzled smem -> TMA store, optional fused BatchNorm statistics), accumulators
// double-buffered in TMEM (2 tiles x 2 rows x BN columns).
//
// TAPS = 4, phases: the sub-pixel up-conv (nearest 2x + conv3x3 as four 2x2-tap convs of the low-resolution tensor,
// capi.cu unetk_upconv3x3_fwd).  The four output phases are four groups of N tiles (q = nt / tiles_per_q) that share the
// SAME halo — phase (qy, qx) reads it one row / one pixel further in — use their own weight rows (q * ncols + co) and
// store through their own stride-2 view of the output.
#include "conv_gemm.cuh"
#include "host_common.cuh"
#include "ptx.cuh"

#include <cstdlib>
#include <mutex>

namespace unetk {

namespace {

constexpr int kThreads = 224;  // warp0 A producer, warp1 MMA, warps2-5 epilogue, warp6 B producer
constexpr int kEpiThreads = 128;
constexpr int kTW = 128;                              // output columns per tile
constexpr int kHaloW = kTW + 2;                       // 130 pixels per halo row
constexpr uint32_t kHaloBytes = 4 * kHaloW * 128;     // 66,560 B written by the TMA
constexpr uint32_t kHaloSlot = (kHaloBytes + 1023u) & ~1023u;  // 67,584 B per ring slot
constexpr uint32_t kStagingBytes = 128 * 64 * 2;

template <int BN>
struct HCfg {
  static constexpr uint32_t kBBytes = BN * 128;                // one (tap, chunk) weight tile
  static constexpr int kBStages = (BN == 64) ? 6 : 4;
  static constexpr int kStaging = (BN == 64) ? 2 : 1;          // staging buffers for the epilogue
  // Back-to-back MMAs that accumulate into the SAME TMEM tile retire one per ~141 clocks whatever N is
  // (profiles/r01_mma_rate_probe.txt): a 128x64 MMA needs 32, so two output rows (two chains) cap the tensor pipe
  // at 45 %.  With N = 64 there is TMEM to spare: each row gets kSplit accumulators that take alternate taps
  // (4 independent chains) and the epilogue adds the partial sums.
  static constexpr int kSplit = (BN == 64) ? 2 : 1;
  static constexpr uint32_t kTmemCols = 4 * BN * kSplit;       // 2 tiles x 2 rows x kSplit x BN
  static constexpr uint32_t kSmemBytes = 2 * kHaloSlot + kBStages * kBBytes + kStaging * kStagingBytes + 1024 + 256 + BN * 4 /*bias*/;
};

struct HaloParams {
  CUtensorMap tmA;    // dims (K, W, H, N), box (64, 130, 4, 1)
  CUtensorMap tmB;    // dims (K, ncols, 9), box (64, BN, 1)
  CUtensorMap tmOut[4];  // dims (ncols, W, H, N), box (64, 128, 1, 1); one per output phase ([0] only unless phased)
  const float* bias;
  const float* scale;   // non-null: eval-mode BatchNorm fold, out = relu?(acc * scale + bias) (the AFFINE instantiation)
  int relu;
  float* stats_partial;
  int accumulate;  // != 0: out += tile (TMA reduce-add)
  int H, W, tiles_h, tiles_w, num_m_tiles, num_n_tiles, ncols, kchunks;
  FastDiv fd_n_tiles, fd_tiles_w, fd_tiles_h;
  int resident;  // 1: all TAPS*kchunks weight tiles stay in smem for the whole kernel (they fit), no B ring
  int tiles_per_q;  // N tiles per output phase (num_n_tiles = tiles_per_q * phases)
  int q_shift;      // halo offset added per phase: (q >> 1, q & 1) * q_shift
  int l2_prefetch;  // > 0: L2-prefetch the halo of the tile `l2_prefetch` rounds ahead
  int8_t dh[9], dw[9], btap[9];
};

// PAIR (BN = 128): clusters of two CTAs own two neighbouring tiles (2 rows x 128 columns each) of the same N tile and run
// every tap as ONE M = 256 x N = 128 MMA over both SMs (tcgen05 cta_group::2).  An N = 128 MMA on one SM re-reads 4 KB of A
// and 4 KB of B per 64 clocks — all of an SM's shared-memory bandwidth, 55-69 % tensor-pipe activity
// (profiles/r01_ncu_halo_epilogue.txt); in the pair each CTA reads its own halo and its HALF of the weight tile (64 rows;
// the peer holds the other 64), 96 B/clk, and fetches half of the weights.  Barrier protocol as in wgrad3x3_2sm.cu.
template <int BN, bool AFFINE, int TAPS, bool PAIR = false>
__global__ void __launch_bounds__(kThreads, 1) conv3x3_halo_kernel(const __grid_constant__ HaloParams p) {
  using C = HCfg<BN>;
  static_assert(!PAIR || BN == 128, "the CTA-pair form exists for BN = 128");
  constexpr uint32_t kBBytes = PAIR ? C::kBBytes / 2 : C::kBBytes;   // this CTA's part of one (tap, chunk) weight tile
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;                                   // [2][kHaloSlot]
  uint8_t* sB = sA + 2 * kHaloSlot;                     // [kBStages][kBBytes]
  const uint32_t b_region = p.resident ? static_cast<uint32_t>(TAPS * p.kchunks) * kBBytes : C::kBStages * kBBytes;
  const int n_staging = p.resident ? 1 : C::kStaging;
  uint8_t* staging = sB + b_region;                     // [n_staging][16 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + n_staging * kStagingBytes);
  uint64_t* a_full = bars;                  // [2]
  uint64_t* a_empty = bars + 2;             // [2]
  uint64_t* b_full = bars + 4;              // [kBStages]
  uint64_t* b_empty = b_full + C::kBStages; // [kBStages]
  uint64_t* tfull = b_empty + C::kBStages;  // [2]
  uint64_t* tempty = tfull + 2;             // [2]
  uint64_t* w_full = tempty + 2;            // [1] resident weights landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    tma_prefetch_desc(&p.tmOut[0]);
    for (int i = 0; i < 2; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < C::kBStages; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], PAIR ? 8 : 4); }
    mbar_init(w_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (PAIR) tmem_alloc_2sm<C::kTmemCols>(tmem_slot);
    else tmem_alloc<C::kTmemCols>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  uint32_t rank = 0;
  if constexpr (PAIR) {
    rank = cluster_ctarank();
    cluster_sync_all();   // both CTAs' barriers and TMEM exist before anything crosses the pair
  }
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // programmatic dependent launch: everything above overlapped the previous kernel's tail; from here on global memory
  pdl_trigger();
  pdl_wait();

  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  // tile walk: single CTA: blockIdx.x, + gridDim.x, ...; PAIR: the cluster walks pair tiles (two neighbouring M tiles x one
  // N tile), rank r takes M tile 2*mp + r — past the end for the last pair of an odd count: all zero fill, nothing stored
  const int walk_first = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int walk_step = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int walk_end = PAIR ? ((p.num_m_tiles + 1) >> 1) * p.num_n_tiles : num_tiles;
  const int cta_nt = (PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x)) % p.num_n_tiles;   // this CTA's N tile
  auto tile_of = [&](int wt) -> int {
    if constexpr (!PAIR) return wt;
    uint32_t mp, nt;
    p.fd_n_tiles.divmod(static_cast<uint32_t>(wt), mp, nt);
    return static_cast<int>((2 * mp + rank) * static_cast<uint32_t>(p.num_n_tiles) + nt);
  };

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------ TMA producer
      int as = 0;
      uint32_t aph = 0;
      for (int wt = walk_first; wt < walk_end; wt += walk_step) {
        const int tile = tile_of(wt);
        uint32_t mt, tw, th, img, rest;
        mt = p.fd_n_tiles.div(tile);
        p.fd_tiles_w.divmod(mt, rest, tw);
        p.fd_tiles_h.divmod(rest, img, th);
        const int h0 = th * 2, w0 = tw * kTW;
        if (!PAIR && p.l2_prefetch > 0) {
          // optional: pull the halo of the tile this CTA reaches `l2_prefetch` rounds from now into L2
          const int ft = tile + p.l2_prefetch * gridDim.x;
          if (ft < num_tiles) {
            const int fmt = ft / p.num_n_tiles;
            const int fw0 = (fmt % p.tiles_w) * kTW, fh0 = ((fmt / p.tiles_w) % p.tiles_h) * 2;
            const int fimg = fmt / (p.tiles_w * p.tiles_h);
            for (int kc = 0; kc < p.kchunks; ++kc) tma_prefetch_l2_4d(&p.tmA, kc * 64, fw0 - 1, fh0 - 1, fimg);
          }
        }
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait(&a_empty[as], aph ^ 1u);
          if constexpr (PAIR) {   // both halos are counted on the leader's barrier
            if (rank == 0) mbar_expect_tx(&a_full[as], 2 * kHaloBytes);
            tma_load_4d_2sm(sA + as * kHaloSlot, &p.tmA, &a_full[as], kc * 64, w0 - 1, h0 - 1, img);
          } else {
            mbar_expect_tx(&a_full[as], kHaloBytes);
            tma_load_4d(sA + as * kHaloSlot, &p.tmA, &a_full[as], kc * 64, w0 - 1, h0 - 1, img);
          }
          if (++as == 2) { as = 0; aph ^= 1u; }
        }
      }
    }
  } else if (warp == 6) {
    if (lane == 0) {
      // ------------------------------------------------------------ weight producer (own warp: a full B ring must
      // never delay the next tile's halo load, and vice versa)
      if (p.resident) {
        // grid is a multiple of num_n_tiles => this CTA always works on N tile blockIdx % num_n_tiles
        const int nt = cta_nt;
        const uint32_t w_bytes = static_cast<uint32_t>(TAPS * p.kchunks) * kBBytes;
        if constexpr (PAIR) {
          if (rank == 0) mbar_expect_tx(w_full, 2 * w_bytes);
        } else {
          mbar_expect_tx(w_full, w_bytes);
        }
        for (int kc = 0; kc < p.kchunks; ++kc)
          for (int t = 0; t < TAPS; ++t) {
            if constexpr (PAIR) tma_load_3d_2sm(sB + (kc * TAPS + t) * kBBytes, &p.tmB, w_full, kc * 64, nt * BN + static_cast<int>(rank) * (BN / 2), p.btap[t]);
            else tma_load_3d(sB + (kc * TAPS + t) * kBBytes, &p.tmB, w_full, kc * 64, nt * BN, p.btap[t]);
          }
      } else {
        int bs = 0;
        uint32_t bph = 0;
        for (int wt = walk_first; wt < walk_end; wt += walk_step) {
          const int nt = cta_nt;
          for (int kc = 0; kc < p.kchunks; ++kc) {
            for (int t = 0; t < TAPS; ++t) {
              mbar_wait(&b_empty[bs], bph ^ 1u);
              if constexpr (PAIR) {
                if (rank == 0) mbar_expect_tx(&b_full[bs], 2 * kBBytes);
                tma_load_3d_2sm(sB + bs * kBBytes, &p.tmB, &b_full[bs], kc * 64, nt * BN + static_cast<int>(rank) * (BN / 2), p.btap[t]);
              } else {
                mbar_expect_tx(&b_full[bs], kBBytes);
                tma_load_3d(sB + bs * kBBytes, &p.tmB, &b_full[bs], kc * 64, nt * BN, p.btap[t]);
              }
              if (++bs == C::kBStages) { bs = 0; bph ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (!PAIR || rank == 0) {
      // ------------------------------------------------------------ MMA issuer (warp-convergent, elected lane issues;
      // PAIR: the leader issues for both SMs)
      const bool issue = elect_one();
      constexpr uint32_t idesc = make_idesc_bf16(PAIR ? 256 : 128, BN, false, false);
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      int it = 0;
      if (p.resident) { mbar_wait_p(issue, w_full, 0); tc_fence_after(); }
      for (int wt = walk_first; wt < walk_end; wt += walk_step, ++it) {
        const int acc = it & 1;
        mbar_wait_p(issue, &tempty[acc], ((it >> 1) & 1) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 2 * BN * C::kSplit;   // [split][row][BN]
        // output phase of this tile's N tile: its window starts (qy, qx) * q_shift further into the halo
        const int q = cta_nt / p.tiles_per_q;
        const int q_halo = ((q >> 1) * kHaloW + (q & 1)) * p.q_shift;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait_p(issue, &a_full[as], aph);
          tc_fence_after();
          // Descriptors are built once per chunk and only ADVANCED per MMA (one 64-bit add each): with N = 64 an
          // MMA lasts ~32-48 clocks, so the single issuing thread must spend only a few instructions per MMA.
          const uint64_t a_desc0 = make_smem_desc(smem_u32(sA + as * kHaloSlot), 16, 1024, kLayoutSW128);
          const uint64_t b_desc0 = make_smem_desc(smem_u32(sB), 16, 1024, kLayoutSW128);
          // Taps are issued kSplit at a time, one per accumulator set, with the k-steps of the sets interleaved:
          // consecutive MMAs then belong to 2*kSplit independent accumulation chains (row x set).
#pragma unroll
          for (int t0 = 0; t0 < TAPS; t0 += C::kSplit) {
            const int nset = (t0 + C::kSplit <= TAPS) ? C::kSplit : TAPS - t0;   // compile-time after unrolling
            uint64_t da0[C::kSplit], db0[C::kSplit];
#pragma unroll
            for (int s = 0; s < C::kSplit; ++s) {
              if (s < nset) {
                const int t = t0 + s;
                if (p.resident) {
                  db0[s] = desc_advance(b_desc0, static_cast<uint32_t>(kc * TAPS + t) * kBBytes);
                } else {
                  int st = bs + s;
                  uint32_t ph = bph;
                  if (st >= C::kBStages) { st -= C::kBStages; ph ^= 1u; }
                  mbar_wait_p(issue, &b_full[st], ph);
                  tc_fence_after();
                  db0[s] = desc_advance(b_desc0, static_cast<uint32_t>(st) * kBBytes);
                }
                // halo row of output row u and tap t: (u + dh + 1); halo column of output column 0: (dw + 1)
                da0[s] = desc_advance(a_desc0, static_cast<uint32_t>(((p.dh[t] + 1) * kHaloW + p.dw[t] + 1 + q_halo) * 128));
              }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
#pragma unroll
              for (int s = 0; s < C::kSplit; ++s) {
                if (s < nset) {
                  const uint32_t d_set = d_tmem + static_cast<uint32_t>(s) * 2 * BN;
                  const uint64_t da = desc_advance(da0[s], k * 32), db = desc_advance(db0[s], k * 32);
                  if constexpr (PAIR) {
                    const uint32_t accf = (t0 == 0 && k == 0 && kc == 0) ? 0u : 1u;
                    umma_bf16_2sm_p(issue, d_set, da, db, idesc, accf);
                    umma_bf16_2sm_p(issue, d_set + BN, desc_advance(da, kHaloW * 128), db, idesc, accf);
                  } else if (t0 == 0 && k == 0) {   // the first MMA of every accumulator overwrites (first chunk only)
                    umma_bf16_p(issue, d_set, da, db, idesc, kc != 0 ? 1u : 0u);
                    umma_bf16_p(issue, d_set + BN, desc_advance(da, kHaloW * 128), db, idesc, kc != 0 ? 1u : 0u);
                  } else {
                    umma_bf16_acc_p(issue, d_set, da, db, idesc);
                    umma_bf16_acc_p(issue, d_set + BN, desc_advance(da, kHaloW * 128), db, idesc);
                  }
                }
              }
            }
            if (!p.resident) {
#pragma unroll
              for (int s = 0; s < C::kSplit; ++s) {
                if (s < nset) {
                  if constexpr (PAIR) umma_commit_2sm_mc_p(issue, &b_empty[bs], 0x3);
                  else umma_commit_p(issue, &b_empty[bs]);
                  if (++bs == C::kBStages) { bs = 0; bph ^= 1u; }
                }
              }
            }
          }
          // all taps of this chunk have been issued
          if constexpr (PAIR) umma_commit_2sm_mc_p(issue, &a_empty[as], 0x3);
          else umma_commit_p(issue, &a_empty[as]);
          if (++as == 2) { as = 0; aph ^= 1u; }
        }
        if constexpr (PAIR) umma_commit_2sm_mc_p(issue, &tfull[acc], 0x3);
        else umma_commit_p(issue, &tfull[acc]);
      }
    }
  } else if (warp >= 2 && warp <= 5) {
    // -------------------------------------------------------------- epilogue (warps 2..5)
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;     // output column within the tile == TMEM lane
    const int et = threadIdx.x - 64;
    const bool leader = (et == 0);
    const int st_ch = et & 63, st_half = et >> 6;
    const uint32_t staging_a = smem_u32(staging);
    const uint32_t row_sw = static_cast<uint32_t>(row & 7) << 4;   // 128B-swizzle XOR of this thread's staging row
    uint32_t st_off[8];   // statistics: byte offset of channel st_ch in rows 8i+j of a staged chunk (swizzle resolved)
#pragma unroll
    for (int j = 0; j < 8; ++j)
      st_off[j] = static_cast<uint32_t>(j * 128 + ((((st_ch >> 3) ^ j) << 4) + (st_ch & 7) * 2));
    float ssum[BN / 64], ssq[BN / 64];
#pragma unroll
    for (int c = 0; c < BN / 64; ++c) { ssum[c] = 0.f; ssq[c] = 0.f; }
    uint32_t chunk_ctr = 0;
    int it = 0;
    // bias of this CTA's N tile -> shared memory once (see conv_gemm.cu)
    const uint32_t bias_a = smem_u32(bars) + 256;
    const float* scale_g = nullptr;   // AFFINE: read through the read-only cache (see conv_gemm.cu)
    if (p.bias != nullptr) {
      const int co_cta = (cta_nt % p.tiles_per_q) * BN;
      for (int i = et; i < BN; i += kEpiThreads) sts_f32(bias_a + i * 4, (co_cta + i < p.ncols) ? __ldg(p.bias + co_cta + i) : 0.f);
      if constexpr (AFFINE) scale_g = p.scale + co_cta;
      named_bar_sync(1, kEpiThreads);
    }
    for (int wt = walk_first; wt < walk_end; wt += walk_step, ++it) {
      const int tile = tile_of(wt);
      const bool tile_ok = !PAIR || tile < num_tiles;   // PAIR: the second CTA's tile past the end (odd tile count)
      const int acc = it & 1;
      uint32_t nt, mt, tw, th, img, rest;
      p.fd_n_tiles.divmod(tile, mt, nt);
      p.fd_tiles_w.divmod(mt, rest, tw);
      p.fd_tiles_h.divmod(rest, img, th);
      const int h0 = th * 2, w0 = tw * kTW;
      const int q = static_cast<int>(nt) / p.tiles_per_q;      // output phase (0 unless phased)
      const int co0 = (static_cast<int>(nt) - q * p.tiles_per_q) * BN;
      const int valid_w = (p.W - w0 < kTW) ? (p.W - w0) : kTW;   // columns of this tile inside the image

      mbar_wait(&tfull[acc], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * 2 * BN * C::kSplit;
      // NOT unrolled: the epilogue is instruction-fetch bound (one warp per scheduler, nothing hides an L0 I-cache miss;
      // stall_no_inst was 39 % of its samples with the 2-4x unrolled body), so the chunk body must stay resident
#pragma unroll 1
      for (int uc = 0; uc < 2 * (BN / 64); ++uc) {
        const int u = (BN == 64) ? uc : (uc >> 1);
        const bool row_ok = (h0 + u) < p.H && tile_ok;
        {
          const int c = (BN == 64) ? 0 : (uc & 1);
          const int colbase = co0 + c * 64;
          const bool live = row_ok && colbase < p.ncols;
          const bool last = (uc == 2 * (BN / 64) - 1);
          uint8_t* buf = staging + (n_staging == 2 ? (chunk_ctr & 1u) : 0u) * kStagingBytes;
          if (live) {
            if (leader) { if (n_staging == 2) bulk_wait_read<1>(); else bulk_wait_read<0>(); }
            named_bar_sync(1, kEpiThreads);
          }
          uint32_t r0[32], r1[32];
          if (live) {
            tmem_ld32(taddr + u * BN + c * 64, r0);
            tmem_ld32(taddr + u * BN + c * 64 + 32, r1);
            tmem_ld_wait();
            if constexpr (C::kSplit == 2) {   // add the second partial sum (odd taps)
              uint32_t q0[32], q1[32];
              tmem_ld32(taddr + 2 * BN + u * BN + c * 64, q0);
              tmem_ld32(taddr + 2 * BN + u * BN + c * 64 + 32, q1);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                r0[j] = __float_as_uint(__uint_as_float(r0[j]) + __uint_as_float(q0[j]));
                r1[j] = __float_as_uint(__uint_as_float(r1[j]) + __uint_as_float(q1[j]));
              }
            }
          }
          if (last) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if constexpr (PAIR) mbar_arrive_leader(&tempty[acc]);
              else mbar_arrive(&tempty[acc]);
            }
          }
          if (!live) continue;
          ++chunk_ctr;
          const uint32_t buf_a = staging_a + (buf - staging);
          const uint32_t row_a = buf_a + row * 128;
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            const uint32_t* src = (v < 4) ? &r0[v * 8] : &r1[(v - 4) * 8];
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(src[j]);
            if constexpr (AFFINE) {
              const float4 b0 = lds128_f(bias_a + (c * 64 + v * 8) * 4), b1 = lds128_f(bias_a + (c * 64 + v * 8 + 4) * 4);
              float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
              if (colbase + v * 8 < p.ncols) {
                s0 = __ldg(reinterpret_cast<const float4*>(scale_g + c * 64 + v * 8));
                s1 = __ldg(reinterpret_cast<const float4*>(scale_g + c * 64 + v * 8 + 4));
              }
              f[0] = fmaf(f[0], s0.x, b0.x); f[1] = fmaf(f[1], s0.y, b0.y); f[2] = fmaf(f[2], s0.z, b0.z); f[3] = fmaf(f[3], s0.w, b0.w);
              f[4] = fmaf(f[4], s1.x, b1.x); f[5] = fmaf(f[5], s1.y, b1.y); f[6] = fmaf(f[6], s1.z, b1.z); f[7] = fmaf(f[7], s1.w, b1.w);
              if (p.relu) {
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
              }
            } else if (p.bias != nullptr) {
              const float4 b0 = lds128_f(bias_a + (c * 64 + v * 8) * 4), b1 = lds128_f(bias_a + (c * 64 + v * 8 + 4) * 4);
              f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
              f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
            }
            uint4 o;
            o.x = pack_bf16x2(f[0], f[1]);
            o.y = pack_bf16x2(f[2], f[3]);
            o.z = pack_bf16x2(f[4], f[5]);
            o.w = pack_bf16x2(f[6], f[7]);
            sts128(row_a + ((v << 4) ^ row_sw), o);
          }
          fence_proxy_async_smem();
          named_bar_sync(1, kEpiThreads);
          if (leader) {
            if (p.accumulate) tma_reduce_add_4d(&p.tmOut[q], buf, colbase, w0, h0 + u, img);
            else tma_store_4d(&p.tmOut[q], buf, colbase, w0, h0 + u, img);
            bulk_commit();
          }
          if (p.stats_partial != nullptr && colbase + st_ch < p.ncols) {
            float s = 0.f, ss = 0.f, s2 = 0.f, ss2 = 0.f;   // two chains
            const int r_begin = st_half * 64;
            uint32_t base = buf_a + r_begin * 128;
#pragma unroll 1
            for (int r8 = 0; r8 < 8; ++r8, base += 1024) {
              uint32_t u[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) u[j] = lds_u16(base + st_off[j]);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float v = __uint_as_float(u[j] << 16);
                if (r_begin + r8 * 8 + j >= valid_w) v = 0.f;
                if (j & 1) { s2 += v; ss2 = fmaf(v, v, ss2); } else { s += v; ss = fmaf(v, v, ss); }
              }
            }
            s += s2;
            ss += ss2;
#pragma unroll
            for (int k = 0; k < BN / 64; ++k) {   // c is a run-time value: predicated adds keep the sums in registers
              if (k == c) { ssum[k] += s; ssq[k] += ss; }
            }
          }
        }
      }
    }
    if (leader) bulk_wait<0>();
    if (p.stats_partial != nullptr) {
      named_bar_sync(1, kEpiThreads);
      float* red = reinterpret_cast<float*>(staging);  // [2 halves][2][BN] floats <= 4 KB
#pragma unroll
      for (int c = 0; c < BN / 64; ++c) {
        red[(st_half * 2 + 0) * BN + c * 64 + st_ch] = ssum[c];
        red[(st_half * 2 + 1) * BN + c * 64 + st_ch] = ssq[c];
      }
      named_bar_sync(1, kEpiThreads);
      // PAIR: partial rows [0, pairs) = the leaders, [pairs, 2 pairs) = their peers (the pair count is a multiple of num_n_tiles)
      const size_t prow = PAIR ? (blockIdx.x >> 1) + static_cast<size_t>(rank) * (gridDim.x >> 1) : static_cast<size_t>(blockIdx.x);
      for (int i = et; i < 2 * BN; i += kEpiThreads)
        p.stats_partial[prow * 2 * BN + i] = red[i] + red[2 * BN + i];
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();   // nobody frees TMEM or leaves while the peer may still touch this CTA
  if (warp == 1) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_2sm<C::kTmemCols>(tmem_base);
    else tmem_dealloc<C::kTmemCols>(tmem_base);
  }
}

template <int BN, bool AFFINE, int TAPS>
int launch_t(HaloParams& p, int grid, cudaStream_t stream) {
  using C = HCfg<BN>;
  static DeviceOnce once;
  UNETK_CUDA(once.run([] { return cudaFuncSetAttribute(conv3x3_halo_kernel<BN, AFFINE, TAPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); }));
  // weights resident when all TAPS*kchunks tiles fit next to the two halo slots and one staging buffer
  const uint32_t w_bytes = static_cast<uint32_t>(TAPS * p.kchunks) * C::kBBytes;
  const uint32_t resident_smem = 2 * kHaloSlot + w_bytes + kStagingBytes + 1024 + 256 + BN * 4;
  p.resident = (resident_smem <= 227 * 1024) ? 1 : 0;
  const uint32_t smem_bytes = p.resident ? resident_smem : C::kSmemBytes;
  UNETK_CUDA(launch_pdl(conv3x3_halo_kernel<BN, AFFINE, TAPS>, dim3(grid), dim3(kThreads), smem_bytes, stream, p));
  UNETK_LAUNCHED();
  return 0;
}
// ---- CTA pairs (BN = 128)
template <bool AFFINE, int TAPS>
uint32_t pair_smem(HaloParams& p) {
  using C = HCfg<128>;
  const uint32_t w_bytes = static_cast<uint32_t>(TAPS * p.kchunks) * (C::kBBytes / 2);
  const uint32_t resident_smem = 2 * kHaloSlot + w_bytes + kStagingBytes + 1024 + 256 + 128 * 4;
  p.resident = (resident_smem <= 227 * 1024) ? 1 : 0;
  return p.resident ? resident_smem : 2 * kHaloSlot + C::kBStages * (C::kBBytes / 2) + C::kStaging * kStagingBytes + 1024 + 256 + 128 * 4;
}
int halo_max_pairs() {
  static std::mutex mu;
  static int cached[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  std::lock_guard<std::mutex> lk(mu);
  if (cached[dev & 63] == 0) {
    int n = -1;
    if (cudaFuncSetAttribute(conv3x3_halo_kernel<128, false, 9, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) == cudaSuccess) {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(2 * 148);
      cfg.blockDim = dim3(kThreads);
      cfg.dynamicSmemBytes = 227 * 1024;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      if (cudaOccupancyMaxActiveClusters(&n, conv3x3_halo_kernel<128, false, 9, true>, &cfg) != cudaSuccess || n <= 0) n = -1;
    }
    cudaGetLastError();
    cached[dev & 63] = n;
  }
  return cached[dev & 63];
}
template <bool AFFINE, int TAPS>
int launch_pair_t(HaloParams& p, int pairs, cudaStream_t stream) {
  static DeviceOnce once;
  UNETK_CUDA(once.run([] { return cudaFuncSetAttribute(conv3x3_halo_kernel<128, AFFINE, TAPS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); }));
  const uint32_t smem_bytes = pair_smem<AFFINE, TAPS>(p);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  UNETK_CUDA(cudaLaunchKernelEx(&cfg, conv3x3_halo_kernel<128, AFFINE, TAPS, true>, p));
  UNETK_LAUNCHED();
  return 0;
}
int launch_pair(HaloParams& p, int pairs, cudaStream_t stream, int taps) {
  if (taps == 4) return p.scale != nullptr ? launch_pair_t<true, 4>(p, pairs, stream) : launch_pair_t<false, 4>(p, pairs, stream);
  return p.scale != nullptr ? launch_pair_t<true, 9>(p, pairs, stream) : launch_pair_t<false, 9>(p, pairs, stream);
}

template <int BN>
int launch(HaloParams& p, int grid, cudaStream_t stream, int taps) {
  if (taps == 4) return p.scale != nullptr ? launch_t<BN, true, 4>(p, grid, stream) : launch_t<BN, false, 4>(p, grid, stream);
  return p.scale != nullptr ? launch_t<BN, true, 9>(p, grid, stream) : launch_t<BN, false, 9>(p, grid, stream);
}

}  // namespace

int conv_stats_sums_launch(const float* partial, int grid, int num_n_tiles, int BN, int C, double* sums,
                           cudaStream_t stream);
int conv_stats_sums_q_launch(const float* partial, int grid, int tiles_per_q, int q_groups, int BN, int C, double* sums,
                             cudaStream_t stream);

// the four-phase sub-pixel up-conv (capi.cu upconv_fwd_desc): 2x2 window taps at (-1..0, -1..0) + the phase
static bool halo_phased(const ConvGemmDesc& d) {
  if (!(d.taps == 4 && d.q_groups == 4 && d.q_shift == 1 && d.out_step == 2 && d.b_taps == 4 && !d.out_f32)) return false;
  for (int t = 0; t < 4; ++t)
    if (d.dh[t] < -1 || d.dh[t] > 0 || d.dw[t] < -1 || d.dw[t] > 0) return false;
  return d.ncols == 64 || d.ncols == 128;   // an N tile must not straddle two phases' weight rows
}

bool conv3x3_halo_eligible(const ConvGemmDesc& d) {
  static int enabled = -1, up = -1;
  if (enabled < 0) { const char* e = getenv("UNETK_HALO_CONV"); enabled = e ? atoi(e) : 1; }
  if (up < 0) { const char* e = getenv("UNETK_HALO_UPCONV"); up = e ? atoi(e) : 1; }
  if (!(enabled && d.a_step == 1 && d.W >= 128 && d.H >= 2 && d.ncols <= 128 && d.K >= 8)) return false;
  if (d.taps == 9 && d.out_step == 1 && d.q_groups == 1) return true;
  return up && halo_phased(d);
}

// Same contract as conv_gemm_run for the shapes conv3x3_halo_eligible() accepts.
int conv3x3_halo_run(const ConvGemmDesc& d, cudaStream_t stream) {
  HaloParams p{};
  const int BN = d.ncols > 64 ? 128 : 64;
  p.H = d.H; p.W = d.W;
  p.tiles_h = (d.H + 1) / 2;
  p.tiles_w = (d.W + kTW - 1) / kTW;
  p.num_m_tiles = d.N * p.tiles_h * p.tiles_w;
  const bool phased = d.q_groups > 1;
  p.tiles_per_q = (d.ncols + BN - 1) / BN;
  p.num_n_tiles = p.tiles_per_q * d.q_groups;
  p.q_shift = d.q_shift;
  p.ncols = d.ncols;
  p.fd_n_tiles = FastDiv(p.num_n_tiles);
  p.fd_tiles_w = FastDiv(p.tiles_w);
  p.fd_tiles_h = FastDiv(p.tiles_h);
  p.kchunks = (d.K + 63) / 64;
  for (int t = 0; t < d.taps; ++t) { p.dh[t] = d.dh[t]; p.dw[t] = d.dw[t]; p.btap[t] = d.btap[t]; }
  p.bias = d.bias;
  p.scale = d.scale;
  p.relu = d.relu;
  p.accumulate = d.accumulate;
  p.stats_partial = d.stats_sums ? d.stats_partial : nullptr;
  {
    // Measured on B200 (UNet B=16 512^2 step): prefetch 0 -> 26.98 ms, 2 -> 27.34 ms, 4 -> 27.61 ms: the extra L2
    // requests cost more than the latency they hide => off by default.  Also measured and dropped: replacing the
    // TMA store of the staged tile by coalesced st.global (64->64 dgrad 0.424 -> 0.478 ms, ConvTranspose fwd
    // 0.79 -> 1.15 ms) and one N = 256 tile over the four ConvTranspose phases (0.385 -> 0.402 ms).
    static int env = -1;
    if (env < 0) { const char* e = getenv("UNETK_HALO_PREFETCH"); env = e ? atoi(e) : 0; }
    p.l2_prefetch = env;
  }
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  int grid = tiles < num_sms() ? tiles : num_sms();
  grid = grid / p.num_n_tiles * p.num_n_tiles;
  if (grid < p.num_n_tiles) grid = p.num_n_tiles;
  // BN = 128: CTA pairs (cta_group::2): two neighbouring tiles of an N tile share one M = 256 MMA per tap and the weight tile
  int pairs = 0;
  {
    static int env = -1;
    if (env < 0) { const char* e = getenv("UNETK_HALO_2SM"); env = e ? atoi(e) : 1; }
    if (env && BN == 128 && p.num_m_tiles >= 2 && p.l2_prefetch == 0) {
      const int fit = halo_max_pairs();
      const int pair_tiles = ((p.num_m_tiles + 1) / 2) * p.num_n_tiles;
      int n = num_sm_pairs();
      if (n > fit) n = fit;
      if (n > pair_tiles) n = pair_tiles;
      n = n / p.num_n_tiles * p.num_n_tiles;
      if (n >= p.num_n_tiles && n > 0) pairs = n;
    }
    if (pairs) grid = 2 * pairs;
  }
  {
    uint64_t dims[4] = {static_cast<uint64_t>(d.K), static_cast<uint64_t>(d.W), static_cast<uint64_t>(d.H),
                        static_cast<uint64_t>(d.N)};
    uint64_t strides[3] = {static_cast<uint64_t>(d.a_ld) * 2, static_cast<uint64_t>(d.a_ld) * 2 * d.W,
                           static_cast<uint64_t>(d.a_ld) * 2 * d.W * d.H};
    uint32_t box[4] = {64, kHaloW, 4, 1};
    uint32_t es[4] = {1, 1, 1, 1};
    if (int rc = make_tmap_bf16(&p.tmA, d.a, 4, dims, strides, box, es, true)) return rc;
  }
  {
    const uint64_t rows = static_cast<uint64_t>(d.ncols) * d.q_groups;   // phased: weight rows q * ncols + co
    uint64_t dims[3] = {static_cast<uint64_t>(d.K), rows, static_cast<uint64_t>(d.b_taps)};
    uint64_t strides[2] = {static_cast<uint64_t>(d.K) * 2, static_cast<uint64_t>(d.K) * 2 * (d.b_rows ? static_cast<uint64_t>(d.b_rows) : rows)};
    uint32_t box[3] = {64, static_cast<uint32_t>(pairs ? BN / 2 : BN), 1};   // pair: each CTA loads half of the weight rows
    uint32_t es[3] = {1, 1, 1};
    if (int rc = make_tmap_bf16(&p.tmB, d.b, 3, dims, strides, box, es, true)) return rc;
  }
  {
    // output seen on the grid of GEMM rows: pixel (h, w) of phase q lives at out[(s*h + qy) * Wout + s*w + qx] (s = out_step)
    const int s = phased ? d.out_step : 1;
    const int64_t Wout = static_cast<int64_t>(d.W) * s, Hout = static_cast<int64_t>(d.H) * s;
    uint64_t dims[4] = {static_cast<uint64_t>(d.ncols), static_cast<uint64_t>(d.W), static_cast<uint64_t>(d.H),
                        static_cast<uint64_t>(d.N)};
    uint64_t strides[3] = {static_cast<uint64_t>(d.out_ld) * 2 * s, static_cast<uint64_t>(d.out_ld) * 2 * Wout * s,
                           static_cast<uint64_t>(d.out_ld) * 2 * Wout * Hout};
    uint32_t box[4] = {64, kTW, 1, 1};
    uint32_t es[4] = {1, 1, 1, 1};
    for (int q = 0; q < d.q_groups; ++q) {
      const uint8_t* base = static_cast<const uint8_t*>(d.out) + (static_cast<int64_t>(q >> 1) * Wout + (q & 1)) * d.out_ld * 2;
      if (int rc = make_tmap_bf16(&p.tmOut[q], base, 4, dims, strides, box, es, true)) return rc;
    }
  }
  const int rc = pairs ? launch_pair(p, pairs, stream, d.taps)
                       : ((BN == 128) ? launch<128>(p, grid, stream, d.taps) : launch<64>(p, grid, stream, d.taps));
  if (rc) return rc;
  if (d.stats_sums != nullptr) {
    if (phased) return conv_stats_sums_q_launch(d.stats_partial, grid, p.tiles_per_q, d.q_groups, BN, d.ncols, d.stats_sums, stream);
    return conv_stats_sums_launch(d.stats_partial, grid, p.num_n_tiles, BN, d.ncols, d.stats_sums, stream);
  }
  return 0;
}

}  // namespace unetk
