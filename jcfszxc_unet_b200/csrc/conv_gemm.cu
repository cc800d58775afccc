// Tap-GEMM on tcgen05: the one tensor-core kernel behind conv3x3 fwd/dgrad, conv1x1,
// ConvTranspose2d(2,2) fwd and dgrad.  NHWC bf16 activations, fp32 accumulation in TMEM.
//
//   D[pixel, n] = sum_{tap} sum_{k} A[pixel shifted by tap, k] * B[tap][n][k]
//
// * A tiles (128 output pixels x 64 channels) are fetched by TMA straight from the NHWC tensor with a
//   4-D box (64ch, TW, TH, 1) at signed coordinates; padding=1 is the TMA's out-of-bounds zero fill,
//   so there is no im2col buffer and no halo logic.  128B swizzle -> canonical K-major UMMA layout.
// * B tiles (BN x 64) come from the packed weight [tap][n][k] (K-major) the same way.
// * Warp-specialised persistent CTA: warp0 = TMA producer, warp1 = MMA issuer (one thread) + TMEM
//   owner, warps2-5 = epilogue.  Double-buffered accumulator: the epilogue of tile i overlaps the MMAs
//   of tile i+1.
// * Epilogue: TMEM -> registers -> (+bias) -> bf16 -> 128B-swizzled smem staging -> TMA store.  The TMA
//   writes full 128-byte rows per pixel (the first version stored 16 B per thread straight to global and
//   was LSU-bound on the thin layers: 64->64 @512^2 ran at 475 TFLOP/s, profiles/r01_*), clips ragged
//   tiles, and handles the ConvTranspose pixel shuffle through per-phase tensor maps.
// * Optional fused BatchNorm statistics: per-channel sum / sum-of-squares of the bf16 tile are taken
//   from the staging buffer while the TMA store drains it (deterministic: per-CTA partials, ordered
//   second stage), which removes one full read of every conv output.
// Reference semantics replaced: nn.Conv2d(k=3,p=1)/nn.Conv2d(k=1)/nn.ConvTranspose2d(k=2,s=2) as
// used by UNetFamily/utils/unet_parts.py:24-31,56-58,77 (reference), forward and input-gradient.
#include "conv_gemm.cuh"
#include "host_common.cuh"
#include "ptx.cuh"
#include "reduce2.cuh"

#include <cstdlib>
#include <mutex>

namespace unetk {

namespace {

constexpr int kTileM = 128;   // output pixels per tile (UMMA M)
constexpr int kTileK = 64;    // channels per k-block: 64 bf16 = one 128B swizzle row
constexpr int kUmmaK = 16;
constexpr int kThreads = 192;
constexpr int kEpiThreads = 128;
constexpr uint32_t kABytes = kTileM * kTileK * 2;
constexpr uint32_t kStagingBytes = kTileM * 64 * 2;  // one 128 x 64 bf16 chunk

template <int BN>
struct Cfg {
  static constexpr uint32_t kBBytes = BN * kTileK * 2;
  static constexpr uint32_t kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 192 ? 4 : (BN == 128 ? 6 : 7));
  static constexpr uint32_t kTmemCols = (BN == 192) ? 512 : 2 * BN;  // a power of two >= 32 (two BN-wide accumulators)
  static constexpr uint32_t kPipeBytes = kStages * kStageBytes;
  static constexpr uint32_t kSmemBytes = kPipeBytes + 2 * kStagingBytes + 1024 /*align slack*/ + 256 /*barriers*/ + BN * 4 /*bias*/;
  // PAIR (cta_group::2, BN = 256): a CTA stages its own A tile and HALF of the B tile (the peer holds the other 128 weight
  // rows): 32 KB per stage instead of 48, six stages in the same shared memory
  static constexpr uint32_t kPairStageBytes = kABytes + kBBytes / 2;
  static constexpr int kPairStages = 6;
  static constexpr uint32_t kPairSmemBytes = kPairStages * kPairStageBytes + 2 * kStagingBytes + 1024 + 256 + BN * 4;
};

// AFFINE: the eval-mode BatchNorm fold, out = relu?(acc * scale + shift) — its own instantiation, so the training kernels'
// epilogue (instruction-fetch bound, see below) does not carry the extra code.
// PAIR (BN = 256 only): launched as clusters of two CTAs that own two neighbouring M tiles of the same N tile and run ONE
// M = 256 x N = 256 MMA per k-step over both SMs (tcgen05 cta_group::2): every CTA reads its own A slab and its half of B
// from its own shared memory (64 B/clk instead of 96) and fetches half of the weights.  Protocol as in wgrad3x3_2sm.cu: all
// TMA bytes are counted on the leader's full barrier, the leader issues and commits onto both CTAs' empty / tfull barriers,
// both CTAs' epilogue warps arrive on the leader's tempty barrier.
template <int BN, bool F32OUT, bool AFFINE, bool PAIR = false>
__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_kernel(const __grid_constant__ ConvGemmParams p) {
  using C = Cfg<BN>;
  static_assert(!PAIR || (BN == 256 && !F32OUT), "the CTA-pair form exists for BN = 256 bf16 output");
  constexpr int kStages = PAIR ? C::kPairStages : C::kStages;
  constexpr uint32_t kStageBytes = PAIR ? C::kPairStageBytes : C::kStageBytes;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024B alignment.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* staging = smem + kStages * kStageBytes;  // 2 x 16 KB, 1024-aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + 2 * kStagingBytes);
  uint64_t* full_bar = bars;                     // [kStages]
  uint64_t* empty_bar = bars + kStages;          // [kStages]
  uint64_t* tfull_bar = bars + 2 * kStages;      // [2]
  uint64_t* tempty_bar = tfull_bar + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    if (!F32OUT) tma_prefetch_desc(&p.tmOut[0]);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], PAIR ? 8 : 4);  // one arrive per epilogue warp (of both CTAs)
    }
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
  const int kblocks = p.taps * p.kchunks;
  // Tile walk.  Single CTA: tiles blockIdx.x, blockIdx.x + gridDim.x, ...  PAIR: the cluster walks PAIR tiles (two
  // neighbouring M tiles x one N tile); rank r takes M tile 2*mp + r, which may lie past the end (odd M tile count):
  // such a tile is all TMA zero fill, its stores and statistics are skipped.
  const int walk_first = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int walk_step = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int walk_end = PAIR ? ((p.num_m_tiles + 1) >> 1) * p.num_n_tiles : num_tiles;
  auto tile_of = [&](int wt) -> int {
    if constexpr (!PAIR) return wt;
    uint32_t mp, nt;
    p.fd_n_tiles.divmod(static_cast<uint32_t>(wt), mp, nt);
    return static_cast<int>((2 * mp + rank) * static_cast<uint32_t>(p.num_n_tiles) + nt);
  };

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------ TMA producer
      int stage = 0;
      uint32_t phase = 0;
      for (int wt = walk_first; wt < walk_end; wt += walk_step) {
        const int tile = tile_of(wt);
        uint32_t nt, mt, tw, th, img, rest, q, ntq;
        p.fd_n_tiles.divmod(tile, mt, nt);
        p.fd_tiles_w.divmod(mt, rest, tw);
        p.fd_tiles_h.divmod(rest, img, th);
        p.fd_tiles_per_q.divmod(nt, q, ntq);
        const int h0 = th * p.TH, w0 = tw * p.TW;
        const int brow = q * p.rows_per_q + ntq * BN + (PAIR ? static_cast<int>(rank) * (BN / 2) : 0);
        if (!PAIR && p.l2_prefetch > 0) {
          // Streaming layers are latency-bound (the smem ring cannot hold enough bytes in flight to cover an HBM
          // miss): pull the A box of a tile this CTA will reach `l2_prefetch` rounds from now into L2.
          const int ftile = tile + p.l2_prefetch * gridDim.x;
          if (ftile < num_tiles && (ftile % p.num_n_tiles) == 0) {
            const int fmt = ftile / p.num_n_tiles;
            const int fw0 = (fmt % p.tiles_w) * p.TW, fh0 = ((fmt / p.tiles_w) % p.tiles_h) * p.TH;
            const int fimg = fmt / (p.tiles_w * p.tiles_h);
            for (int kc = 0; kc < p.kchunks; ++kc)
              tma_prefetch_l2_4d(&p.tmA, kc * kTileK, p.a_step * fw0, p.a_step * fh0, fimg);
          }
        }
        for (int t = 0; t < p.taps; ++t) {
          const int ah = p.a_step * h0 + p.dh[t] + p.q_shift * static_cast<int>(q >> 1);
          const int aw = p.a_step * w0 + p.dw[t] + p.q_shift * static_cast<int>(q & 1);
          const int bt = p.btap[t];
          for (int kc = 0; kc < p.kchunks; ++kc) {
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            uint8_t* sa = smem + stage * kStageBytes;
            uint8_t* sb = sa + kABytes;
            if constexpr (PAIR) {
              // the bytes of BOTH CTAs are counted on the leader's barrier (the only one the MMA issuer waits on)
              if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * kStageBytes);
              tma_load_4d_2sm(sa, &p.tmA, &full_bar[stage], kc * kTileK, aw, ah, img);
              tma_load_3d_2sm(sb, &p.tmB, &full_bar[stage], kc * kTileK, brow, bt);
            } else {
              mbar_expect_tx(&full_bar[stage], kStageBytes);
              tma_load_4d(sa, &p.tmA, &full_bar[stage], kc * kTileK, aw, ah, img);
              tma_load_3d(sb, &p.tmB, &full_bar[stage], kc * kTileK, brow, bt);
            }
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (!PAIR || rank == 0) {
      // ------------------------------------------------------------ MMA issuer (warp-convergent, elected lane issues;
      // PAIR: the leader CTA issues for both SMs)
      const bool issue = elect_one();
      constexpr uint32_t idesc = make_idesc_bf16(PAIR ? 2 * kTileM : kTileM, BN, false, false);
      const uint64_t a_desc0 = make_smem_desc(smem_u32(smem), 16, 1024, kLayoutSW128);  // stage 0, k = 0
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int wt = walk_first; wt < walk_end; wt += walk_step, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait_p(issue, &tempty_bar[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait_p(issue, &full_bar[stage], phase);
          tc_fence_after();
          // descriptors are advanced, not rebuilt: the issuing thread spends ~3 instructions per MMA
          const uint64_t da0 = desc_advance(a_desc0, static_cast<uint32_t>(stage) * kStageBytes);
          const uint64_t db0 = desc_advance(da0, kABytes);
          if constexpr (PAIR) umma_bf16_2sm_p(issue, d_tmem, da0, db0, idesc, kb == 0 ? 0u : 1u);
          else if (kb == 0) umma_bf16_p(issue, d_tmem, da0, db0, idesc, 0u);
          else umma_bf16_acc_p(issue, d_tmem, da0, db0, idesc);
          // K tail: the last 64-channel chunk of every tap holds ksteps_last real 16-channel steps (the rest of the
          // tile is TMA zero fill); K = 32 (UNet++'s 32-channel gradients) would otherwise issue 2x the MMAs
          const int ksteps = ((kb + 1) % p.kchunks == 0) ? p.ksteps_last : kTileK / kUmmaK;
#pragma unroll
          for (int k = 1; k < kTileK / kUmmaK; ++k)
            if (k < ksteps) {
              if constexpr (PAIR) umma_bf16_2sm_p(issue, d_tmem, desc_advance(da0, k * kUmmaK * 2), desc_advance(db0, k * kUmmaK * 2), idesc, 1u);
              else umma_bf16_acc_p(issue, d_tmem, desc_advance(da0, k * kUmmaK * 2), desc_advance(db0, k * kUmmaK * 2), idesc);
            }
          // frees the smem slot (PAIR: in both CTAs) once these MMAs retire
          if constexpr (PAIR) umma_commit_2sm_mc_p(issue, &empty_bar[stage], 0x3);
          else umma_commit_p(issue, &empty_bar[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        // accumulator complete -> epilogue (PAIR: of both CTAs)
        if constexpr (PAIR) umma_commit_2sm_mc_p(issue, &tfull_bar[acc], 0x3);
        else umma_commit_p(issue, &tfull_bar[acc]);
      }
    }
  } else {
    // -------------------------------------------------------------- epilogue (warps 2..5, 128 threads)
    const int quarter = warp & 3;            // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;     // tile row == TMEM lane
    const int et = threadIdx.x - 64;         // 0..127
    const bool leader = (et == 0);
    const int st_ch = et & 63, st_half = et >> 6;  // statistics: this thread owns channel st_ch of every 64-chunk
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
    // bias of this CTA's N tile (the grid is a multiple of num_n_tiles: a CTA keeps its N tile) -> shared memory once;
    // the per-element predicated __ldg it replaces was ~450 instructions of the chunk body
    const uint32_t bias_a = smem_u32(bars) + 256;
    // AFFINE: the scale vector is read through the read-only cache (every thread of the CTA reads the same 32 bytes at a
    // time: one broadcast transaction); the 227 KB of shared memory are full (BN = 256: 4 stages + 2 staging buffers)
    const float* scale_g = nullptr;
    if (!F32OUT && p.bias != nullptr) {
      const int co_cta = static_cast<int>((PAIR ? (blockIdx.x >> 1) : blockIdx.x) % p.num_n_tiles % p.tiles_per_q) * BN;
      for (int i = et; i < BN; i += kEpiThreads) sts_f32(bias_a + i * 4, (co_cta + i < p.ncols) ? __ldg(p.bias + co_cta + i) : 0.f);
      if constexpr (AFFINE) scale_g = p.scale + co_cta;
      named_bar_sync(1, kEpiThreads);
    }
    for (int wt = walk_first; wt < walk_end; wt += walk_step, ++it) {
      const int tile = tile_of(wt);
      const bool tile_ok = !PAIR || tile < num_tiles;   // PAIR: the second CTA's M tile past the end (odd tile count)
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      uint32_t nt, mt, tw, th, img, rest, q, ntq;
      p.fd_n_tiles.divmod(tile, mt, nt);
      p.fd_tiles_w.divmod(mt, rest, tw);
      p.fd_tiles_h.divmod(rest, img, th);
      p.fd_tiles_per_q.divmod(nt, q, ntq);
      const int h0 = th * p.TH, w0 = tw * p.TW;
      const int co0 = ntq * BN;
      const bool ragged = (h0 + p.TH > p.H) || (w0 + p.TW > p.W);

      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
      if constexpr (F32OUT) {
        // fp32 output (f32path.cu): each thread owns one output pixel and writes its row of the accumulator as is
        const int ph = h0 + (row >> p.tw_shift), pw = w0 + (row & (p.TW - 1));
        const bool ok = ph < p.H && pw < p.W;
        const long long Wout = static_cast<long long>(p.W) * p.out_step, Hout = static_cast<long long>(p.H) * p.out_step;
        float* orow = p.out_f32 + ((static_cast<long long>(img) * Hout + static_cast<long long>(ph) * p.out_step + (q >> 1)) * Wout +
                                   static_cast<long long>(pw) * p.out_step + (q & 1)) * p.out_ld;
#pragma unroll
        for (int c = 0; c < BN / 32; ++c) {
          const int colbase = co0 + c * 32;
          uint32_t r[32];
          if (colbase < p.ncols) {   // warp-uniform
            tmem_ld32(taddr + c * 32, r);
            tmem_ld_wait();
          }
          if (c == BN / 32 - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
          }
          if (colbase >= p.ncols || !ok) continue;
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            const int col = colbase + v * 4;
            if (col < p.ncols) {   // ncols is a multiple of 8
              float4 o;
              o.x = __uint_as_float(r[v * 4 + 0]);
              o.y = __uint_as_float(r[v * 4 + 1]);
              o.z = __uint_as_float(r[v * 4 + 2]);
              o.w = __uint_as_float(r[v * 4 + 3]);
              if (p.bias != nullptr) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col));
                o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
              }
              *reinterpret_cast<float4*>(orow + col) = o;
            }
          }
        }
        continue;
      }
      // NOT unrolled: the epilogue is instruction-fetch bound (one warp per scheduler, nothing hides an L0 I-cache
      // miss; stall_no_inst was 39 % of its samples with the 2-4x unrolled body), so the loop body must stay resident
#pragma unroll 1
      for (int c = 0; c < BN / 64; ++c) {
        const int colbase = co0 + c * 64;
        const bool live = colbase < p.ncols && tile_ok;  // false: whole chunk beyond the real channels (ragged N tile)
        uint8_t* buf = staging + (chunk_ctr & 1u) * kStagingBytes;
        if (live) {
          // the buffer is free once the TMA store issued two chunks ago has read it (and everybody has passed
          // the statistics pass over it, which the barrier below also guarantees)
          if (leader) bulk_wait_read<1>();
          named_bar_sync(1, kEpiThreads);
        }
        uint32_t r0[32], r1[32];
        if (live) {
          tmem_ld32(taddr + c * 64, r0);
          tmem_ld32(taddr + c * 64 + 32, r1);
          tmem_ld_wait();
        }
        if (c == BN / 64 - 1) {  // accumulator fully drained into registers: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (PAIR) mbar_arrive_leader(&tempty_bar[acc]);
            else mbar_arrive(&tempty_bar[acc]);
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
            // (live chunk: colbase + 64 <= round_up(ncols, 8); ncols % 8 == 0 and the group starts at a multiple of 8)
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
          sts128(row_a + ((v << 4) ^ row_sw), o);  // 128B swizzle, conflict-free
        }
        fence_proxy_async_smem();
        named_bar_sync(1, kEpiThreads);
        if (leader) {
          if (p.accumulate) tma_reduce_add_4d(&p.tmOut[q], buf, colbase, w0, h0, img);
          else tma_store_4d(&p.tmOut[q], buf, colbase, w0, h0, img);
          bulk_commit();
        }
        if (p.stats_partial != nullptr && colbase + st_ch < p.ncols) {
          // per-channel sum / sum of squares of the bf16 values just staged (rows outside the image excluded).
          // Measured and dropped: one 4-byte LDS per channel PAIR over 32 rows, fully unrolled with two independent
          // chains per channel (64->64 @512^2 fwd 0.504 -> 0.596 ms): thin layers are bound by shared-memory
          // bandwidth (the MMA re-reads A for every 64 output channels), not by this loop's latency.
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
              if (ragged) {
                const int r = r_begin + r8 * 8 + j;
                if (!((h0 + (r >> p.tw_shift) < p.H) && (w0 + (r & (p.TW - 1)) < p.W))) v = 0.f;
              }
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
    // all bulk stores of this thread must be complete before the CTA exits
    if (leader && !F32OUT) bulk_wait<0>();
    if (p.stats_partial != nullptr) {
      // combine the two row-halves through the (now idle) staging buffer, then one partial row per CTA
      named_bar_sync(1, kEpiThreads);
      float* red = reinterpret_cast<float*>(staging);  // [2 halves][2][BN]
#pragma unroll
      for (int c = 0; c < BN / 64; ++c) {
        red[(st_half * 2 + 0) * BN + c * 64 + st_ch] = ssum[c];
        red[(st_half * 2 + 1) * BN + c * 64 + st_ch] = ssq[c];
      }
      named_bar_sync(1, kEpiThreads);
      // PAIR: partial rows [0, pairs) = the leaders, [pairs, 2 pairs) = their peers; the pair count is a multiple of
      // num_n_tiles, so that row % num_n_tiles is the row's N tile in both halves (conv_stats_sums_kernel)
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

// sums[k][c] = sum over the CTAs that own channel c's N tile (CTA b owns tile b % num_n_tiles), fixed order
__global__ void conv_stats_sums_kernel(const float* __restrict__ partial, int grid, int num_n_tiles, int BN, int C,
                                       double* __restrict__ sums) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * kSum2Lanes + threadIdx.x;
  const bool valid = i < 2 * C;
  const int k = valid ? i / C : 0, c = valid ? i % C : 0;
  const int nt = c / BN, cc = c % BN;
  const int count = (grid - nt + num_n_tiles - 1) / num_n_tiles;   // CTAs nt, nt + num_n_tiles, ...
  const double s = sliced_ordered_sum(partial, count, valid, [&](int j) {
    return (static_cast<size_t>(nt + j * num_n_tiles) * 2 + k) * BN + cc;
  });
  if (valid && threadIdx.y == 0) sums[i] = s;
}

// Output phases (q_groups > 1, the sub-pixel up-conv): channel c's N tile is owned by the CTAs b with
// (b % num_n_tiles) % tiles_per_q == c / BN — one group of grid / num_n_tiles CTAs per phase (the grid is a multiple of
// num_n_tiles); summed phase by phase in a fixed order.
__global__ void conv_stats_sums_q_kernel(const float* __restrict__ partial, int grid, int tiles_per_q, int q_groups, int BN,
                                         int C, double* __restrict__ sums) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * kSum2Lanes + threadIdx.x;
  const bool valid = i < 2 * C;
  const int k = valid ? i / C : 0, c = valid ? i % C : 0;
  const int nt = c / BN, cc = c % BN;
  const int num_n_tiles = tiles_per_q * q_groups;
  const int rounds = grid / num_n_tiles;
  const double s = sliced_ordered_sum(partial, rounds * q_groups, valid, [&](int j) {
    const int q = j % q_groups, r = j / q_groups;
    return (static_cast<size_t>(q * tiles_per_q + nt + r * num_n_tiles) * 2 + k) * BN + cc;
  });
  if (valid && threadIdx.y == 0) sums[i] = s;
}

template <int BN, bool F32OUT, bool AFFINE>
int launch_t(const ConvGemmParams& p, int grid, cudaStream_t stream) {
  using C = Cfg<BN>;
  static DeviceOnce once;
  UNETK_CUDA(once.run([] {
    return cudaFuncSetAttribute(conv_gemm_kernel<BN, F32OUT, AFFINE>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
  }));
  UNETK_CUDA(launch_pdl(conv_gemm_kernel<BN, F32OUT, AFFINE>, dim3(grid), dim3(kThreads), C::kSmemBytes, stream, p));
  UNETK_LAUNCHED();
  return 0;
}
template <int BN>
int launch(const ConvGemmParams& p, int grid, cudaStream_t stream) {
  if (p.out_f32 != nullptr) return launch_t<BN, true, false>(p, grid, stream);
  return p.scale != nullptr ? launch_t<BN, false, true>(p, grid, stream) : launch_t<BN, false, false>(p, grid, stream);
}

// CTA pairs the device can keep resident with the pair kernel's shared memory (cached per device; <= 0: unavailable)
int conv_max_pairs() {
  static std::mutex mu;
  static int cached[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  std::lock_guard<std::mutex> lk(mu);
  if (cached[dev & 63] == 0) {
    int n = -1;
    if (cudaFuncSetAttribute(conv_gemm_kernel<256, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             Cfg<256>::kPairSmemBytes) == cudaSuccess) {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(2 * 148);
      cfg.blockDim = dim3(kThreads);
      cfg.dynamicSmemBytes = Cfg<256>::kPairSmemBytes;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      if (cudaOccupancyMaxActiveClusters(&n, conv_gemm_kernel<256, false, false, true>, &cfg) != cudaSuccess || n <= 0) n = -1;
    }
    cudaGetLastError();
    cached[dev & 63] = n;
  }
  return cached[dev & 63];
}

template <bool AFFINE>
int launch_pair_t(const ConvGemmParams& p, int pairs, cudaStream_t stream) {
  static DeviceOnce once;
  UNETK_CUDA(once.run([] {
    return cudaFuncSetAttribute(conv_gemm_kernel<256, false, AFFINE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<256>::kPairSmemBytes);
  }));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = Cfg<256>::kPairSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  UNETK_CUDA(cudaLaunchKernelEx(&cfg, conv_gemm_kernel<256, false, AFFINE, true>, p));
  UNETK_LAUNCHED();
  return 0;
}
int launch_pair(const ConvGemmParams& p, int pairs, cudaStream_t stream) {
  return p.scale != nullptr ? launch_pair_t<true>(p, pairs, stream) : launch_pair_t<false>(p, pairs, stream);
}

int pick_bn(int ncols, int q_groups) {
  if (q_groups > 1) {
    // ConvTranspose fwd: an N tile must not straddle two output phases q (a ragged last tile reads the next
    // phase's weight rows; those columns are never stored)
    return (ncols % 256 == 0) ? 256 : (ncols % 128 == 0 ? 128 : 64);
  }
  // 129..192 columns (UNet++ / ResUNet concat widths 160, 192): one 192-wide tile instead of two 128-wide ones
  return ncols >= 256 ? 256 : (ncols > 128 && ncols <= 192 ? 192 : (ncols > 64 ? 128 : 64));
}

}  // namespace

int conv_stats_sums_launch(const float* partial, int grid, int num_n_tiles, int BN, int C, double* sums,
                           cudaStream_t stream) {
  UNETK_CUDA(launch_pdl(conv_stats_sums_kernel, dim3((2 * C + kSum2Lanes - 1) / kSum2Lanes), dim3(kSum2Lanes, kSum2Slices), 0, stream, partial, grid, num_n_tiles, BN, C, sums));
  UNETK_LAUNCHED();
  return 0;
}

int conv_stats_sums_q_launch(const float* partial, int grid, int tiles_per_q, int q_groups, int BN, int C, double* sums,
                             cudaStream_t stream) {
  UNETK_CUDA(launch_pdl(conv_stats_sums_q_kernel, dim3((2 * C + kSum2Lanes - 1) / kSum2Lanes), dim3(kSum2Lanes, kSum2Slices), 0, stream,
                        partial, grid, tiles_per_q, q_groups, BN, C, sums));
  UNETK_LAUNCHED();
  return 0;
}

bool conv3x3_halo_eligible(const ConvGemmDesc& d);
int conv3x3_halo_run(const ConvGemmDesc& d, cudaStream_t stream);
bool conv3x3_rows_eligible(const ConvGemmDesc& d);
int conv3x3_rows_run(const ConvGemmDesc& d, cudaStream_t stream);

// see conv_gemm_run; UNETK_SPLIT_WIDE=0 disables, the value is the largest K that is split (default 64)
static bool split_wide_thin(const ConvGemmDesc& d) {
  static int kmax = -1;
  if (kmax < 0) { const char* e = getenv("UNETK_SPLIT_WIDE"); kmax = e ? atoi(e) : 64; }
  return kmax > 0 && d.taps == 9 && d.a_step == 1 && d.out_step == 1 && d.q_groups == 1 && !d.out_f32 &&
         d.stats_partial == nullptr && d.W >= 128 && d.H >= 2 && d.ncols > 128 && d.K <= kmax;
}

size_t conv_gemm_stats_partial_floats(int ncols) {
  return static_cast<size_t>(num_sms()) * 2 * pick_bn(ncols, 1);
}

int conv_gemm_run(const ConvGemmDesc& d, cudaStream_t stream) {
  UNETK_CHECK(d.K % 8 == 0 && d.K >= 8, -1, "conv_gemm: K=%d must be a multiple of 8", d.K);
  UNETK_CHECK(d.ncols % 8 == 0, -1, "conv_gemm: ncols=%d must be a multiple of 8", d.ncols);
  UNETK_CHECK(d.a_ld % 8 == 0 && d.out_ld % 8 == 0, -1, "conv_gemm: pixel strides must be multiples of 8");
  UNETK_CHECK((reinterpret_cast<uintptr_t>(d.a) & 15) == 0 && (reinterpret_cast<uintptr_t>(d.out) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(d.b) & 15) == 0,
              -1, "conv_gemm: pointers must be 16-byte aligned");
  UNETK_CHECK(d.taps >= 1 && d.taps <= 16, -1, "conv_gemm: taps=%d", d.taps);
  UNETK_CHECK(d.stats_sums == nullptr || d.stats_partial != nullptr, -1, "conv_gemm: fused statistics need a partial buffer");
  UNETK_CHECK(d.q_groups == 1 || d.q_groups == 4, -1, "conv_gemm: q_groups=%d (1 or 4)", d.q_groups);

  UNETK_CHECK(d.scale == nullptr || (reinterpret_cast<uintptr_t>(d.scale) & 15) == 0, -1, "conv_gemm: scale must be 16-byte aligned");
  UNETK_CHECK(d.scale == nullptr || (d.bias != nullptr && !d.accumulate && !d.out_f32 && d.stats_sums == nullptr), -1,
              "conv_gemm: the affine (folded BatchNorm) epilogue needs a shift vector and excludes accumulate / fp32 output / statistics");
  if (d.out_f32) {
    UNETK_CHECK(!d.accumulate && d.stats_sums == nullptr, -1, "conv_gemm: the fp32 output path neither accumulates nor takes statistics");
    UNETK_CHECK(d.out_ld % 4 == 0 && (d.bias == nullptr || (reinterpret_cast<uintptr_t>(d.bias) & 15) == 0), -1,
                "conv_gemm: fp32 output needs out_ld %% 4 == 0 and a 16-byte aligned bias");
  } else if (split_wide_thin(d)) {
    // Wide output, thin reduction (the dgrad of UNet++'s concat-fed convs: 32 -> 160 / 192 channels at full
    // resolution): the generic kernel re-streams 9 x BN x 64 weights through the pipeline for every 128-pixel tile
    // and ran at 310-460 TFLOP/s; column slices of <= 128 channels take the halo / row-stacked kernels with their
    // weights resident in shared memory.  The slices re-read the (thin) input, nothing else changes.
    const int total = d.b_rows ? d.b_rows : d.ncols;
    for (int c0 = 0; c0 < d.ncols; c0 += 128) {
      ConvGemmDesc s = d;
      s.ncols = d.ncols - c0 < 128 ? d.ncols - c0 : 128;
      s.b_rows = total;
      s.b = static_cast<const __nv_bfloat16*>(d.b) + static_cast<size_t>(c0) * d.K;
      s.out = static_cast<__nv_bfloat16*>(d.out) + c0;
      s.bias = d.bias ? d.bias + c0 : nullptr;
      s.scale = d.scale ? d.scale + c0 : nullptr;
      if (int rc = conv_gemm_run(s, stream)) return rc;
    }
    return 0;
  } else if (conv3x3_rows_eligible(d)) {
    return conv3x3_rows_run(d, stream);  // wide images, <= 64 output channels: filter rows stacked in N
  } else if (conv3x3_halo_eligible(d)) {
    return conv3x3_halo_run(d, stream);  // wide images, <= 128 output channels
  }

  ConvGemmParams p{};
  const int BN = pick_bn(d.ncols, d.q_groups);
  p.tiles_per_q = (d.ncols + BN - 1) / BN;
  p.rows_per_q = d.ncols;
  p.num_n_tiles = p.tiles_per_q * d.q_groups;
  p.ncols = d.ncols;

  // ---- M tiling: TH x TW = 128 output positions, TW a power of two
  int TW = 128, shift = 7;
  while (TW > d.W) { TW >>= 1; --shift; }
  if (TW < 1) { TW = 1; shift = 0; }
  const int TH = kTileM / TW;
  p.TH = TH; p.TW = TW; p.tw_shift = shift;
  p.H = d.H; p.W = d.W;
  p.tiles_h = (d.H + TH - 1) / TH;
  p.tiles_w = (d.W + TW - 1) / TW;
  p.num_m_tiles = d.N * p.tiles_h * p.tiles_w;
  p.fd_n_tiles = FastDiv(p.num_n_tiles);
  p.fd_tiles_w = FastDiv(p.tiles_w);
  p.fd_tiles_h = FastDiv(p.tiles_h);
  p.fd_tiles_per_q = FastDiv(p.tiles_per_q);
  p.taps = d.taps;
  p.kchunks = (d.K + kTileK - 1) / kTileK;
  p.ksteps_last = ((d.K - 1) % kTileK) / kUmmaK + 1;
  p.a_step = d.a_step;
  p.q_shift = d.q_shift;
  for (int t = 0; t < d.taps; ++t) { p.dh[t] = d.dh[t]; p.dw[t] = d.dw[t]; p.btap[t] = d.btap[t]; }
  p.bias = d.bias;
  p.scale = d.scale;
  p.relu = d.relu;
  p.accumulate = d.accumulate;
  p.stats_partial = d.stats_sums ? d.stats_partial : nullptr;
  {
    // L2 prefetch only pays when A streams from HBM (tensor much larger than what L2 keeps between taps)
    static int env = -1;
    // measured on B200: no gain (0.640 -> 0.667 ms on 64->64 @512^2, profiles/r01_notes.md) => off by default
    if (env < 0) { const char* e = getenv("UNETK_L2_PREFETCH"); env = e ? atoi(e) : 0; }
    const double a_bytes = 2.0 * d.N * d.H * d.W * d.a_step * d.a_step * d.K;
    p.l2_prefetch = (a_bytes > 48e6) ? env : 0;
  }
  UNETK_CHECK(TW * d.a_step <= 256 && TH * d.a_step <= 256, -1, "conv_gemm: TMA box too large");

  // persistent grid; a multiple of num_n_tiles so that every CTA keeps one N tile (fused statistics rely on it)
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  int grid = tiles < num_sms() ? tiles : num_sms();
  grid = grid / p.num_n_tiles * p.num_n_tiles;
  if (grid < p.num_n_tiles) grid = p.num_n_tiles;
  // BN = 256: CTA pairs (cta_group::2) — two neighbouring M tiles of an N tile share one M = 256 MMA and the weight tile
  int pairs = 0;
  {
    static int env = -1;
    if (env < 0) { const char* e = getenv("UNETK_CONV_2SM"); env = e ? atoi(e) : 1; }
    if (env && BN == 256 && !d.out_f32 && p.num_m_tiles >= 2 && p.l2_prefetch == 0) {
      const int fit = conv_max_pairs();
      const int pair_tiles = ((p.num_m_tiles + 1) / 2) * p.num_n_tiles;
      int n = num_sm_pairs();
      if (n > fit) n = fit;
      if (n > pair_tiles) n = pair_tiles;
      n = n / p.num_n_tiles * p.num_n_tiles;   // every pair keeps one N tile
      if (n >= p.num_n_tiles && n > 0) pairs = n;
    }
    if (pairs) grid = 2 * pairs;
  }

  // ---- tensor maps
  {
    const int AH = d.H * d.a_step, AW = d.W * d.a_step;  // spatial extent of the A tensor
    uint64_t dims[4] = {static_cast<uint64_t>(d.K), static_cast<uint64_t>(AW), static_cast<uint64_t>(AH),
                        static_cast<uint64_t>(d.N)};
    uint64_t strides[3] = {static_cast<uint64_t>(d.a_ld) * 2, static_cast<uint64_t>(d.a_ld) * 2 * AW,
                           static_cast<uint64_t>(d.a_ld) * 2 * AW * AH};
    uint32_t box[4] = {kTileK, static_cast<uint32_t>(TW * d.a_step), static_cast<uint32_t>(TH * d.a_step), 1};
    uint32_t es[4] = {1, static_cast<uint32_t>(d.a_step), static_cast<uint32_t>(d.a_step), 1};
    if (int rc = make_tmap_bf16(&p.tmA, d.a, 4, dims, strides, box, es, true)) return rc;
  }
  {
    const uint64_t rows = static_cast<uint64_t>(d.ncols) * d.q_groups;
    uint64_t dims[3] = {static_cast<uint64_t>(d.K), rows, static_cast<uint64_t>(d.b_taps)};
    uint64_t strides[2] = {static_cast<uint64_t>(d.K) * 2, static_cast<uint64_t>(d.K) * 2 * (d.b_rows ? static_cast<uint64_t>(d.b_rows) : rows)};
    uint32_t box[3] = {kTileK, static_cast<uint32_t>(pairs ? BN / 2 : BN), 1};   // pair: each CTA loads half of the rows
    uint32_t es[3] = {1, 1, 1};
    if (int rc = make_tmap_bf16(&p.tmB, d.b, 3, dims, strides, box, es, true)) return rc;
  }
  if (d.out_f32) {
    p.out_f32 = static_cast<float*>(d.out);
    p.out_ld = d.out_ld;
    p.out_step = d.out_step;
  } else {
    // Output seen on the grid of GEMM rows: pixel (h, w) of phase q lives at out[(s*h + qy) * Wout + s*w + qx].
    const int s = d.out_step;
    const int64_t Wout = static_cast<int64_t>(d.W) * s, Hout = static_cast<int64_t>(d.H) * s;
    uint64_t dims[4] = {static_cast<uint64_t>(d.ncols), static_cast<uint64_t>(d.W), static_cast<uint64_t>(d.H),
                        static_cast<uint64_t>(d.N)};
    uint64_t strides[3] = {static_cast<uint64_t>(d.out_ld) * 2 * s, static_cast<uint64_t>(d.out_ld) * 2 * Wout * s,
                           static_cast<uint64_t>(d.out_ld) * 2 * Wout * Hout};
    uint32_t box[4] = {64, static_cast<uint32_t>(TW), static_cast<uint32_t>(TH), 1};
    uint32_t es[4] = {1, 1, 1, 1};
    for (int q = 0; q < d.q_groups; ++q) {
      const uint8_t* base = static_cast<const uint8_t*>(d.out) + (static_cast<int64_t>(q >> 1) * Wout + (q & 1)) * d.out_ld * 2;
      if (int rc = make_tmap_bf16(&p.tmOut[q], base, 4, dims, strides, box, es, true)) return rc;
    }
  }
  int rc;
  switch (BN) {
    case 256: rc = pairs ? launch_pair(p, pairs, stream) : launch<256>(p, grid, stream); break;
    case 192: rc = launch<192>(p, grid, stream); break;
    case 128: rc = launch<128>(p, grid, stream); break;
    default: rc = launch<64>(p, grid, stream); break;
  }
  if (rc) return rc;
  if (d.stats_sums != nullptr) {
    if (d.q_groups > 1)
      return conv_stats_sums_q_launch(d.stats_partial, grid, p.tiles_per_q, d.q_groups, BN, d.ncols, d.stats_sums, stream);
    return conv_stats_sums_launch(d.stats_partial, grid, p.num_n_tiles, BN, d.ncols, d.stats_sums, stream);
  }
  return 0;
}

}  // namespace unetk
