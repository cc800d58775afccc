// conv3x3 (fwd / dgrad) for the thinnest layers: at most 64 output channels on wide images (W >= 128), K <= 128.
//
// Why a third kernel.  With N = 64 an M = 128 tcgen05.mma still costs ~64 clocks (profiles/r01_mma_rate_probe.txt:
// max(64, N/2)): the tensor core re-reads the 4 KB A operand from shared memory for every 64 output channels, so the
// 64-channel layers at 512^2 are capped at half the MMA rate whatever the pipeline does.  This kernel raises N
// instead: ONE input row feeds up to THREE output rows.  Input row i contributes to output rows i-1, i, i+1 through
// the filter rows kh = 2, 1, 0, so with the weights of one filter column kw laid out in shared memory as
// [kh=2 | kh=1 | kh=0] (64 rows each) a single MMA of N = 192 = 3 x 64 multiplies the input row by all three filter
// rows and accumulates into three ADJACENT 64-column accumulator blocks = three consecutive output rows:
//
//     tile = 4 output rows x 128 columns, TMEM accumulator [4 rows][64 ch] = 256 columns (double-buffered: 512)
//     input row i = 0..5 of the tile's halo:   blocks touched        MMA N     filter rows (smem order)
//         0                                    [0]                    64        kh=0
//         1                                    [0,1]                  128       kh=1,0
//         2                                    [0,1,2]                192       kh=2,1,0
//         3                                    [1,2,3]                192       kh=2,1,0
//         4                                    [2,3]                  128       kh=2,1
//         5                                    [3]                    64        kh=2
//
// 12 tap-rows x (3 kw x 4 k-steps) in 460 MMA clocks instead of 12 x 64 = 768: 1.7x the N = 64 rate, and every input
// row is loaded once per tile (1.5x halo over-fetch instead of 2x in conv3x3_halo.cu).  An accumulator block that an
// MMA touches for the first time must be overwritten while its neighbours accumulate; tcgen05.mma has one
// accumulate flag per instruction, so the very first MMA of input rows 1..3 is issued as two (N-64 | 64).
//
// Pipeline: warp0 = TMA producer (ring of input rows, 130 px x 64 ch each), warp1 = MMA issuer, warps2-5 = epilogue
// (TMEM -> bf16 -> swizzled smem -> TMA store / reduce-add, optional fused BatchNorm statistics), warp6 loads the
// resident weights once.  The kw shift is a row-shifted UMMA descriptor start (tcgen05 swizzles on absolute smem
// address bits, profiles/r01_umma_descriptor_probe.txt).
// Reference semantics replaced: nn.Conv2d(k=3,p=1) forward and input gradient of the 64-channel DoubleConv layers,
// UNetFamily/utils/unet_parts.py:24-31.
#include "conv_gemm.cuh"
#include "host_common.cuh"
#include "ptx.cuh"

#include <cstdlib>

namespace unetk {

namespace {

constexpr int kThreads = 224;   // warp0 A producer, warp1 MMA, warps2-5 epilogue, warp6 weight loader
constexpr int kEpiThreads = 128;
constexpr int kTW = 128;                          // output columns per tile
constexpr int kRows = 4;                          // output rows per tile
constexpr int kHaloW = kTW + 2;
constexpr uint32_t kRowBytes = kHaloW * 128;      // 16,640 B written by the TMA per (input row, 64-channel chunk)
constexpr uint32_t kRowSlot = (kRowBytes + 1023u) & ~1023u;   // 17,408 B per ring slot (1024-aligned for SW128)
constexpr int kMaxSlots = 8;

// BW = accumulator block width = output channels per row block: 64, or 32 for the 32-channel layers of UNet++
// (N = 96 for three filter rows; the staging tile then has 64-byte rows and the 64B swizzle).
template <int BW>
struct RCfg {
  static constexpr uint32_t kWBlock = BW * 128;             // one (kc, kw, kh) weight block: BW rows x 64 k
  static constexpr uint32_t kStagingBytes = 128 * BW * 2;   // 128 pixels x BW channels
  static constexpr uint32_t kTmemCols = 2 * kRows * BW;     // 2 tiles x 4 rows x BW channels (512 / 256)
  static constexpr int kGroups = kEpiThreads / BW;          // statistics: row groups (2 / 4) of BW rows each
};

struct RowsParams {
  CUtensorMap tmA;    // dims (K, W, H, N), box (64, 130, 1, 1)
  CUtensorMap tmB;    // dims (K, ncols, 9), box (64, 64, 1)
  CUtensorMap tmOut;  // dims (ncols, W, H, N), box (64, 128, 1, 1)
  const float* bias;
  const float* scale;   // non-null: eval-mode BatchNorm fold, out = relu?(acc * scale + bias) (the AFFINE instantiation)
  int relu;
  float* stats_partial;
  int accumulate;
  int H, W, tiles_h, tiles_w, num_tiles, ncols, kchunks;
  int slots, n_staging;
  int ksteps_last;   // 16-channel MMA steps of the last 64-channel chunk (K tail), 1..4
  FastDiv fd_tiles_w, fd_tiles_h;
  int8_t dh[9], dw[9], btap[9];
};

template <int BW, bool AFFINE>
__global__ void __launch_bounds__(kThreads, 1) conv3x3_rows_kernel(const __grid_constant__ RowsParams p) {
  using C = RCfg<BW>;
  constexpr uint32_t kWBlock = C::kWBlock, kStagingBytes = C::kStagingBytes, kTmemCols = C::kTmemCols;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;                                                     // [slots][kRowSlot]
  uint8_t* sW = sA + p.slots * kRowSlot;                                  // [kc][kw][kh: 2,1,0][64 rows][128 B]
  uint8_t* staging = sW + static_cast<uint32_t>(9 * p.kchunks) * kWBlock; // [n_staging][16 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + p.n_staging * kStagingBytes);
  uint64_t* a_full = bars;                   // [kMaxSlots]
  uint64_t* a_empty = bars + kMaxSlots;      // [kMaxSlots]
  uint64_t* tfull = a_empty + kMaxSlots;     // [2]
  uint64_t* tempty = tfull + 2;              // [2]
  uint64_t* w_full = tempty + 2;             // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    tma_prefetch_desc(&p.tmOut);
    for (int i = 0; i < kMaxSlots; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
    mbar_init(w_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // programmatic dependent launch: everything above overlapped the previous kernel's tail; from here on global memory
  pdl_trigger();
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------ input-row producer
      int as = 0;
      uint32_t aph = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        uint32_t tw, th, img, rest;
        p.fd_tiles_w.divmod(tile, rest, tw);
        p.fd_tiles_h.divmod(rest, img, th);
        const int h0 = th * kRows, w0 = tw * kTW;
        for (int i = 0; i < kRows + 2; ++i) {
          for (int kc = 0; kc < p.kchunks; ++kc) {
            mbar_wait(&a_empty[as], aph ^ 1u);
            mbar_expect_tx(&a_full[as], kRowBytes);
            tma_load_4d(sA + as * kRowSlot, &p.tmA, &a_full[as], kc * 64, w0 - 1, h0 - 1 + i, img);   // OOB rows/cols = 0
            if (++as == p.slots) { as = 0; aph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 6) {
    if (lane == 0) {
      // ------------------------------------------------------------ resident weights: tap t = (dh, dw) -> block [kc][dw+1][1-dh]
      mbar_expect_tx(w_full, static_cast<uint32_t>(9 * p.kchunks) * kWBlock);
      for (int kc = 0; kc < p.kchunks; ++kc)
        for (int t = 0; t < 9; ++t) {
          const int kh = p.dh[t] + 1, kw = p.dw[t] + 1;
          tma_load_3d(sW + ((kc * 3 + kw) * 3 + (2 - kh)) * kWBlock, &p.tmB, w_full, kc * 64, 0, p.btap[t]);
        }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer (warp-convergent, elected lane issues)
    const bool issue = elect_one();
    constexpr uint32_t idesc64 = make_idesc_bf16(128, BW, false, false);        // one / two / three row blocks
    constexpr uint32_t idesc128 = make_idesc_bf16(128, 2 * BW, false, false);
    constexpr uint32_t idesc192 = make_idesc_bf16(128, 3 * BW, false, false);
    const uint64_t a_desc0 = make_smem_desc(smem_u32(sA), 16, 1024, kLayoutSW128);
    const uint64_t w_desc0 = make_smem_desc(smem_u32(sW), 16, 1024, kLayoutSW128);
    int as = 0;
    uint32_t aph = 0;
    int it = 0;
    mbar_wait_p(issue, w_full, 0);
    tc_fence_after();
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      mbar_wait_p(issue, &tempty[acc], ((it >> 1) & 1) ^ 1u);
      tc_fence_after();
      const uint32_t d_tile = tmem_base + acc * (kRows * BW);
#pragma unroll
      for (int i = 0; i < kRows + 2; ++i) {
        // compile-time geometry of input row i (see the table in the header)
        const int jmin = (i - 2 > 0) ? i - 2 : 0, jmax = (i < kRows - 1) ? i : kRows - 1;
        const int nblk = jmax - jmin + 1;
        const int wslot0 = 2 - (i - jmin);              // first filter-row block of the stacked B operand
        const bool fresh = (i <= kRows - 1);            // accumulator block j = i is touched for the first time
        const uint32_t d_row = d_tile + jmin * BW;
        const uint32_t idesc = (nblk == 3) ? idesc192 : (nblk == 2 ? idesc128 : idesc64);
        const uint32_t idesc_rest = (nblk == 3) ? idesc128 : idesc64;   // the nblk-1 older blocks of a fresh row
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait_p(issue, &a_full[as], aph);
          tc_fence_after();
          const uint64_t da_row = desc_advance(a_desc0, static_cast<uint32_t>(as) * kRowSlot);
          const uint64_t dw_kc = desc_advance(w_desc0, static_cast<uint32_t>(kc * 9 + wslot0) * kWBlock);
          const int ksteps = (kc == p.kchunks - 1) ? p.ksteps_last : 4;   // K tail: skip the zero-filled steps
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const uint64_t da_kw = desc_advance(da_row, kw * 128);          // halo column kw = output column 0 shifted
            const uint64_t db_kw = desc_advance(dw_kc, kw * 3 * kWBlock);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t da = desc_advance(da_kw, k * 32), db = desc_advance(db_kw, k * 32);
              if (k >= ksteps) continue;
              if (fresh && kw == 0 && k == 0) {
                // first MMA onto a new accumulator block: older blocks accumulate, the new one is overwritten
                // (first chunk only; later chunks accumulate everywhere)
                if (nblk > 1) umma_bf16_acc_p(issue, d_row, da, db, idesc_rest);
                umma_bf16_p(issue, d_row + (nblk - 1) * BW, da, desc_advance(db, (nblk - 1) * kWBlock), idesc64,
                            kc != 0 ? 1u : 0u);
              } else {
                umma_bf16_acc_p(issue, d_row, da, db, idesc);
              }
            }
          }
          umma_commit_p(issue, &a_empty[as]);
          if (++as == p.slots) { as = 0; aph ^= 1u; }
        }
      }
      umma_commit_p(issue, &tfull[acc]);
    }
  } else if (warp >= 2 && warp <= 5) {
    // -------------------------------------------------------------- epilogue (warps 2..5)
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;     // output column within the tile == TMEM lane
    const int et = threadIdx.x - 64;
    const bool leader = (et == 0);
    const int st_ch = et % BW, st_half = et / BW;     // statistics: channel x row group (BW rows per group)
    const uint32_t staging_a = smem_u32(staging);
    // staging rows are 2*BW bytes: 128B swizzle (16-byte chunk ^= row & 7) or 64B swizzle (chunk ^= (row >> 1) & 3)
    const uint32_t row_sw = (BW == 64) ? (static_cast<uint32_t>(row & 7) << 4) : (static_cast<uint32_t>((row >> 1) & 3) << 4);
    uint32_t st_off[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int x = (BW == 64) ? j : ((j >> 1) & 3);
      st_off[j] = static_cast<uint32_t>(j * (2 * BW) + ((((st_ch >> 3) ^ x) << 4) + (st_ch & 7) * 2));
    }
    float ssum = 0.f, ssq = 0.f;
    uint32_t chunk_ctr = 0;
    int it = 0;
    const uint32_t bias_a = smem_u32(bars) + 256;
    if (p.bias != nullptr) {
      if (et < BW) sts_f32(bias_a + et * 4, (et < p.ncols) ? __ldg(p.bias + et) : 0.f);
      named_bar_sync(1, kEpiThreads);
    }
    const int n_staging = p.n_staging;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      uint32_t tw, th, img, rest;
      p.fd_tiles_w.divmod(tile, rest, tw);
      p.fd_tiles_h.divmod(rest, img, th);
      const int h0 = th * kRows, w0 = tw * kTW;
      const int valid_w = (p.W - w0 < kTW) ? (p.W - w0) : kTW;

      mbar_wait(&tfull[acc], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * (kRows * BW);
#pragma unroll 1
      for (int u = 0; u < kRows; ++u) {   // rolled on purpose: the body must stay resident in the L0 I-cache
        const bool live = (h0 + u) < p.H;
        const bool last = (u == kRows - 1);
        const uint32_t buf_off = (n_staging == 2 ? (chunk_ctr & 1u) : 0u) * kStagingBytes;
        if (live) {
          if (leader) { if (n_staging == 2) bulk_wait_read<1>(); else bulk_wait_read<0>(); }
          named_bar_sync(1, kEpiThreads);
        }
        uint32_t r0[32], r1[32];
        if (live) {
          tmem_ld32(taddr + u * BW, r0);
          if constexpr (BW == 64) tmem_ld32(taddr + u * BW + 32, r1);
          tmem_ld_wait();
        }
        if (last) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[acc]);
        }
        if (!live) continue;
        ++chunk_ctr;
        const uint32_t buf_a = staging_a + buf_off;
        const uint32_t row_a = buf_a + row * (2 * BW);
#pragma unroll
        for (int v = 0; v < BW / 8; ++v) {
          const uint32_t* src = (v < 4) ? &r0[v * 8] : &r1[(v - 4) * 8];
          float f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(src[j]);
          if constexpr (AFFINE) {
            const float4 b0 = lds128_f(bias_a + (v * 8) * 4), b1 = lds128_f(bias_a + (v * 8 + 4) * 4);
            float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;   // scale through the read-only cache (see conv_gemm.cu)
            if (v * 8 < p.ncols) {
              s0 = __ldg(reinterpret_cast<const float4*>(p.scale + v * 8));
              s1 = __ldg(reinterpret_cast<const float4*>(p.scale + v * 8 + 4));
            }
            f[0] = fmaf(f[0], s0.x, b0.x); f[1] = fmaf(f[1], s0.y, b0.y); f[2] = fmaf(f[2], s0.z, b0.z); f[3] = fmaf(f[3], s0.w, b0.w);
            f[4] = fmaf(f[4], s1.x, b1.x); f[5] = fmaf(f[5], s1.y, b1.y); f[6] = fmaf(f[6], s1.z, b1.z); f[7] = fmaf(f[7], s1.w, b1.w);
            if (p.relu) {
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
            }
          } else if (p.bias != nullptr) {
            const float4 b0 = lds128_f(bias_a + (v * 8) * 4), b1 = lds128_f(bias_a + (v * 8 + 4) * 4);
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
          if (p.accumulate) tma_reduce_add_4d(&p.tmOut, staging + buf_off, 0, w0, h0 + u, img);
          else tma_store_4d(&p.tmOut, staging + buf_off, 0, w0, h0 + u, img);
          bulk_commit();
        }
        if (p.stats_partial != nullptr && st_ch < p.ncols) {
          float s = 0.f, ss = 0.f, s2 = 0.f, ss2 = 0.f;
          const int r_begin = st_half * BW;
          uint32_t base = buf_a + r_begin * (2 * BW);
#pragma unroll 1
          for (int r8 = 0; r8 < BW / 8; ++r8, base += 8 * 2 * BW) {
            uint32_t x[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = lds_u16(base + st_off[j]);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float v = __uint_as_float(x[j] << 16);
              if (r_begin + r8 * 8 + j >= valid_w) v = 0.f;
              if (j & 1) { s2 += v; ss2 = fmaf(v, v, ss2); } else { s += v; ss = fmaf(v, v, ss); }
            }
          }
          ssum += s + s2;
          ssq += ss + ss2;
        }
      }
    }
    if (leader) bulk_wait<0>();
    if (p.stats_partial != nullptr) {
      named_bar_sync(1, kEpiThreads);
      float* red = reinterpret_cast<float*>(staging);  // [groups][2][BW]
      red[(st_half * 2 + 0) * BW + st_ch] = ssum;
      red[(st_half * 2 + 1) * BW + st_ch] = ssq;
      named_bar_sync(1, kEpiThreads);
      if (et < 2 * BW) {
        float t = 0.f;
#pragma unroll
        for (int g = 0; g < C::kGroups; ++g) t += red[g * 2 * BW + et];
        p.stats_partial[static_cast<size_t>(blockIdx.x) * 2 * BW + et] = t;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// smem plan: resident weights + staging + as many row slots as fit (>= 3)
bool rows_plan(int kchunks, int BW, int* slots, int* n_staging, uint32_t* smem_bytes) {
  const uint32_t kWBlock = BW * 128, kStagingBytes = 128 * BW * 2;
  const uint32_t fixed = static_cast<uint32_t>(9 * kchunks) * kWBlock + 1024 /*align*/ + 256 /*barriers*/ + 256 /*bias*/;
  const uint32_t budget = 227 * 1024;
  for (int ns = 2; ns >= 1; --ns) {
    const uint32_t rest = fixed + ns * kStagingBytes;
    if (rest >= budget) continue;
    int s = static_cast<int>((budget - rest) / kRowSlot);
    if (s > kMaxSlots) s = kMaxSlots;
    if (s >= (ns == 2 ? 5 : 3)) {
      *slots = s; *n_staging = ns; *smem_bytes = rest + s * kRowSlot;
      return true;
    }
  }
  return false;
}

}  // namespace

int conv_stats_sums_launch(const float* partial, int grid, int num_n_tiles, int BN, int C, double* sums,
                           cudaStream_t stream);

bool conv3x3_rows_eligible(const ConvGemmDesc& d) {
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("UNETK_ROWS_CONV"); enabled = e ? atoi(e) : 1; }
  if (!enabled) return false;
  if (!(d.taps == 9 && d.a_step == 1 && d.out_step == 1 && d.q_groups == 1 && !d.out_f32 && d.W >= 128 && d.H >= 1 &&
        d.ncols <= 64 && d.K >= 8 && d.K <= (d.ncols <= 32 ? 256 : 128)))
    return false;
  for (int t = 0; t < 9; ++t)
    if (d.dh[t] < -1 || d.dh[t] > 1 || d.dw[t] < -1 || d.dw[t] > 1) return false;
  int s, ns;
  uint32_t bytes;
  return rows_plan((d.K + 63) / 64, d.ncols <= 32 ? 32 : 64, &s, &ns, &bytes);
}

// Same contract as conv_gemm_run for the shapes conv3x3_rows_eligible() accepts.
int conv3x3_rows_run(const ConvGemmDesc& d, cudaStream_t stream) {
  RowsParams p{};
  p.H = d.H; p.W = d.W;
  p.tiles_h = (d.H + kRows - 1) / kRows;
  p.tiles_w = (d.W + kTW - 1) / kTW;
  p.num_tiles = d.N * p.tiles_h * p.tiles_w;
  p.ncols = d.ncols;
  p.kchunks = (d.K + 63) / 64;
  p.ksteps_last = ((d.K - 1) % 64) / 16 + 1;
  const int BW = d.ncols <= 32 ? 32 : 64;
  p.fd_tiles_w = FastDiv(p.tiles_w);
  p.fd_tiles_h = FastDiv(p.tiles_h);
  for (int t = 0; t < 9; ++t) { p.dh[t] = d.dh[t]; p.dw[t] = d.dw[t]; p.btap[t] = d.btap[t]; }
  p.bias = d.bias;
  p.scale = d.scale;
  p.relu = d.relu;
  p.accumulate = d.accumulate;
  p.stats_partial = d.stats_sums ? d.stats_partial : nullptr;
  uint32_t smem_bytes = 0;
  UNETK_CHECK(rows_plan(p.kchunks, BW, &p.slots, &p.n_staging, &smem_bytes), -1, "conv3x3_rows: weights do not fit");
  const int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
  {
    uint64_t dims[4] = {static_cast<uint64_t>(d.K), static_cast<uint64_t>(d.W), static_cast<uint64_t>(d.H),
                        static_cast<uint64_t>(d.N)};
    uint64_t strides[3] = {static_cast<uint64_t>(d.a_ld) * 2, static_cast<uint64_t>(d.a_ld) * 2 * d.W,
                           static_cast<uint64_t>(d.a_ld) * 2 * d.W * d.H};
    uint32_t box[4] = {64, kHaloW, 1, 1};
    uint32_t es[4] = {1, 1, 1, 1};
    if (int rc = make_tmap_bf16(&p.tmA, d.a, 4, dims, strides, box, es, true)) return rc;
  }
  {
    uint64_t dims[3] = {static_cast<uint64_t>(d.K), static_cast<uint64_t>(d.ncols), static_cast<uint64_t>(d.b_taps)};
    uint64_t strides[2] = {static_cast<uint64_t>(d.K) * 2, static_cast<uint64_t>(d.K) * 2 * (d.b_rows ? d.b_rows : d.ncols)};
    uint32_t box[3] = {64, static_cast<uint32_t>(BW), 1};
    uint32_t es[3] = {1, 1, 1};
    if (int rc = make_tmap_bf16(&p.tmB, d.b, 3, dims, strides, box, es, true)) return rc;
  }
  {
    uint64_t dims[4] = {static_cast<uint64_t>(d.ncols), static_cast<uint64_t>(d.W), static_cast<uint64_t>(d.H),
                        static_cast<uint64_t>(d.N)};
    uint64_t strides[3] = {static_cast<uint64_t>(d.out_ld) * 2, static_cast<uint64_t>(d.out_ld) * 2 * d.W,
                           static_cast<uint64_t>(d.out_ld) * 2 * d.W * d.H};
    uint32_t box[4] = {static_cast<uint32_t>(BW), kTW, 1, 1};
    uint32_t es[4] = {1, 1, 1, 1};
    if (int rc = make_tmap_bf16(&p.tmOut, d.out, 4, dims, strides, box, es, true, BW == 32 ? 64 : 128)) return rc;
  }
  static DeviceOnce once;
  UNETK_CUDA(once.run([] {
    cudaError_t e = cudaSuccess;
    auto set = [&e](auto kernel) { if (e == cudaSuccess) e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); };
    set(conv3x3_rows_kernel<64, false>); set(conv3x3_rows_kernel<32, false>);
    set(conv3x3_rows_kernel<64, true>); set(conv3x3_rows_kernel<32, true>);
    return e;
  }));
  if (p.scale != nullptr) {
    if (BW == 32) UNETK_CUDA(launch_pdl(conv3x3_rows_kernel<32, true>, dim3(grid), dim3(kThreads), smem_bytes, stream, p));
    else UNETK_CUDA(launch_pdl(conv3x3_rows_kernel<64, true>, dim3(grid), dim3(kThreads), smem_bytes, stream, p));
  } else if (BW == 32) UNETK_CUDA(launch_pdl(conv3x3_rows_kernel<32, false>, dim3(grid), dim3(kThreads), smem_bytes, stream, p));
  else UNETK_CUDA(launch_pdl(conv3x3_rows_kernel<64, false>, dim3(grid), dim3(kThreads), smem_bytes, stream, p));
  UNETK_LAUNCHED();
  if (d.stats_sums != nullptr) return conv_stats_sums_launch(d.stats_partial, grid, 1, BW, d.ncols, d.stats_sums, stream);
  return 0;
}

}  // namespace unetk
