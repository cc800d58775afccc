// 3x3 weight gradient on CTA PAIRS (tcgen05 cta_group::2), for layers with >= 256 output channels.
//
//   dW[co][ci][r][s] = sum_{n,h,w} dY[n,h,w,co] * X[n,h+r-1,w+s-1,ci]
//
// wgrad3x3_kernel<128> (wgrad3x3.cu) issues, per 16-pixel k-step, three M = 128 x N = 128 MMAs (one per tap of a filter row):
// each re-reads its 4 KB A slab and 4 KB of B from shared memory in 64 clocks = 128 B/clk, the whole shared-memory
// bandwidth of an SM, and the kernel sits at 75-80 % tensor-pipe activity (profiles/r02_ncu_wgrad_pair.txt).  Here the two
// SMs of a pair run ONE M = 256 x N = 128 MMA: each CTA holds 128 of the 256 dY channels (its own A rows, its own 128
// accumulator rows in its own TMEM) and 64 of the 128 X channels (its half of B, read once for both tensor cores):
// 6 KB per 64 clocks = 96 B/clk per SM, and each CTA fills 25 KB per k-block by TMA instead of 33.
//
// Protocol (cluster of 2, rank 0 = leader):
//   * both CTAs run the TMA producer for their own boxes; all bytes are counted on the LEADER's full barrier
//     (cp.async.bulk.tensor.cta_group::2 with the mbarrier address of CTA 0), which only the leader's MMA warp waits on;
//   * the leader issues the MMAs and commits each stage onto BOTH CTAs' empty barriers, and the finished accumulator onto
//     both CTAs' tfull barriers (tcgen05.commit.cta_group::2 ... multicast);
//   * each CTA's four epilogue warps drain their own TMEM and arrive on the LEADER's tempty barrier (count 8).
// Work item = (filter row r, 256 output channels, 128 input channels, pixel split), same k-block order as wgrad3x3.cu; fp32
// partials [ksplit][9][M][Nn] + the same ordered reduce.
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
constexpr int kMaxStages = 8;
constexpr uint32_t kPBoxBytes = kPix * 128;  // [64 px][64 ch] bf16
constexpr uint32_t kTmemCols = 512;          // 3 taps x 128 columns (a power of two)

struct W2Params {
  CUtensorMap tmP;  // dY: dims (M, W, H, N), box (64, TW, TH, 1)
  CUtensorMap tmQ;  // X : dims (Nn, W, H, N), box (64, TW+2, TH, 1)
  float* partial;   // [ksplit][9][M][Nn]
  int TH, TW, tiles_h, tiles_w, pix_tiles;
  int m_pairs, n_tiles, ksplit, stages;
  int M, Nn;
  uint32_t q_box_bytes;  // TH*(TW+2)*128 rounded up to 1024
  uint32_t q_tx_bytes;   // TH*(TW+2)*128
};

__global__ void __launch_bounds__(kThreads, 1) wgrad3x3_2sm_kernel(const __grid_constant__ W2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const uint32_t stage_bytes = 2 * kPBoxBytes + p.q_box_bytes;   // this CTA's half of a stage
  const int nstages = p.stages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + nstages * stage_bytes);
  uint64_t* full_bar = bars;                      // [kMaxStages] (used in the leader only)
  uint64_t* empty_bar = bars + kMaxStages;        // [kMaxStages] (each CTA its own; the leader's commit arrives on both)
  uint64_t* tfull_bar = bars + 2 * kMaxStages;    // [1] each CTA its own
  uint64_t* tempty_bar = tfull_bar + 1;           // [1] leader's: 8 arrivals (4 epilogue warps x 2 CTAs)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmP);
    tma_prefetch_desc(&p.tmQ);
    for (int s = 0; s < kMaxStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, 8);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2sm<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // both CTAs' barriers and TMEM exist before anything crosses the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();

  // item -> (n tile, m pair, filter row r, pixel split ks); the pair walks the items together
  const int items_per_split = 3 * p.m_pairs * p.n_tiles;
  const int num_items = items_per_split * p.ksplit;
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  auto decode = [&](int item, int& nt, int& mp, int& r, int& ks) {
    nt = item % p.n_tiles;
    mp = (item / p.n_tiles) % p.m_pairs;
    r = (item / (p.n_tiles * p.m_pairs)) % 3;
    ks = item / items_per_split;
  };

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------ TMA producer (both CTAs, each for its own boxes)
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_both = 2 * (2 * kPBoxBytes + p.q_tx_bytes);
      for (int item = pair; item < num_items; item += num_pairs) {
        int nt, mp, r, ks;
        decode(item, nt, mp, r, ks);
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
          if (leader) mbar_expect_tx(&full_bar[stage], tx_both);   // counts the peer's bytes as well
          const int m0 = mp * 256 + static_cast<int>(rank) * 128;
#pragma unroll
          for (int b = 0; b < 2; ++b)
            tma_load_4d_2sm(sp + b * kPBoxBytes, &p.tmP, &full_bar[stage], m0 + b * 64, w0, h0, img);
          tma_load_4d_2sm(sq, &p.tmQ, &full_bar[stage], nt * 128 + static_cast<int>(rank) * 64, w0 - 1, h0 + r - 1, img);
          if (++stage == nstages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ------------------------------------------------------------ MMA issuer (leader only; warp-convergent, elected lane)
      const bool issue = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      const int row_pitch = p.TW + 2;
      const int steps_per_row = p.TW >> 4;
      uint32_t q_off[kPix / 16];
#pragma unroll
      for (int k = 0; k < kPix / 16; ++k)
        q_off[k] = static_cast<uint32_t>((k / steps_per_row) * row_pitch + (k % steps_per_row) * 16) * 128;
      // A: this CTA's 128 rows = two 64-channel boxes kPBoxBytes apart (the peer's 128 rows sit at the same offsets in the
      // peer's shared memory); B: this CTA's 64 of the 128 columns = one 64-channel halo box
      const uint64_t p_desc0 = make_smem_desc(smem_u32(smem), kPBoxBytes, 1024, kLayoutSW128);
      const uint64_t q_desc0 = make_smem_desc(smem_u32(smem) + 2 * kPBoxBytes, p.q_box_bytes, 1024, kLayoutSW128);
      constexpr uint32_t idesc = make_idesc_bf16(256, 128, true, true);
      for (int item = pair; item < num_items; item += num_pairs, ++it) {
        int nt_, mp_, r_, ks;
        decode(item, nt_, mp_, r_, ks);
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
#pragma unroll
            for (int s = 0; s < 3; ++s)
              umma_bf16_2sm_p(issue, tmem_base + s * 128, da, desc_advance(dq, s * 128), idesc, (first && k == 0) ? 0u : 1u);
          }
          umma_commit_2sm_mc_p(issue, &empty_bar[stage], 0x3);   // the slot is free in BOTH CTAs
          if (++stage == nstages) { stage = 0; phase ^= 1u; }
        }
        umma_commit_2sm_mc_p(issue, tfull_bar, 0x3);             // both CTAs' epilogues may drain their halves
      }
    }
  } else {
    // -------------------------------------------------------------- epilogue (both CTAs: their own 128 accumulator rows)
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    int it = 0;
    for (int item = pair; item < num_items; item += num_pairs, ++it) {
      int nt, mp, r, ks;
      decode(item, nt, mp, r, ks);
      mbar_wait(tfull_bar, it & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
      const int m = mp * 256 + static_cast<int>(rank) * 128 + row;
#pragma unroll 1
      for (int s = 0; s < 3; ++s) {
        float* dst = p.partial + ((static_cast<size_t>(ks) * 9 + r * 3 + s) * p.M + m) * p.Nn + nt * 128;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld32(taddr + s * 128 + c * 32, v);
          tmem_ld_wait();
          if (m < p.M) {
#pragma unroll
            for (int x = 0; x < 8; ++x) {
              const int col = nt * 128 + c * 32 + x * 4;
              if (col < p.Nn) {
                float4 o = make_float4(__uint_as_float(v[x * 4]), __uint_as_float(v[x * 4 + 1]),
                                       __uint_as_float(v[x * 4 + 2]), __uint_as_float(v[x * 4 + 3]));
                *reinterpret_cast<float4*>(dst + c * 32 + x * 4) = o;
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(tempty_bar);
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // nobody frees TMEM or leaves while the peer may still touch this CTA's barriers / accumulators
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm<kTmemCols>(tmem_base);
  }
}

struct W2Plan {
  int TH, TW, tiles_h, tiles_w, pix_tiles, m_pairs, n_tiles, ksplit, stages, pairs;
  uint32_t q_box_bytes, q_tx_bytes, smem_bytes;
};

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
    if (cudaOccupancyMaxActiveClusters(&n, wgrad3x3_2sm_kernel, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); n = -1; }
    cached[dev & 63] = n;
  }
  return cached[dev & 63];
}

bool make_plan(int N, int H, int W, int M, int Nn, W2Plan* pl) {
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("UNETK_WGRAD3_2SM"); enabled = e ? atoi(e) : 1; }
  // >= 256 rows (a pair owns 256 of them), more than 64 columns (N = 128 MMAs; thinner layers have their own modes)
  if (!enabled || W < 16 || M < 256 || Nn <= 64 || M % 8 || Nn % 8) return false;
  int tw = 64;
  while (tw > W) tw >>= 1;
  pl->TW = tw;
  pl->TH = kPix / tw;
  pl->tiles_h = (H + pl->TH - 1) / pl->TH;
  pl->tiles_w = (W + tw - 1) / tw;
  pl->pix_tiles = N * pl->tiles_h * pl->tiles_w;
  pl->m_pairs = (M + 255) / 256;
  pl->n_tiles = (Nn + 127) / 128;
  pl->q_tx_bytes = static_cast<uint32_t>(pl->TH * (tw + 2) * 128);
  pl->q_box_bytes = (pl->q_tx_bytes + 1023u) & ~1023u;
  const uint32_t stage = 2 * kPBoxBytes + pl->q_box_bytes;
  int stages = static_cast<int>((227u * 1024u - 1024u - 256u) / stage);
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 3) return false;
  pl->stages = stages;
  pl->smem_bytes = stages * stage + 1024 + 256;
  static DeviceOnce once;
  if (once.run([] { return cudaFuncSetAttribute(wgrad3x3_2sm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); }) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  int pairs = max_pairs(227 * 1024);
  if (pairs <= 0) return false;
  if (pairs > num_sm_pairs()) pairs = num_sm_pairs();
  if (pairs < 1) return false;
  // pixel split: the cost rule of wgrad3x3.cu with the pairs in the place of the SMs
  const int base = 3 * pl->m_pairs * pl->n_tiles;
  const int cap = pl->pix_tiles / 8 > 0 ? pl->pix_tiles / 8 : 1;
  int best = 1;
  long best_cost = -1;
  for (int ks = 1; ks <= cap && ks <= 48; ++ks) {
    const long waves = (static_cast<long>(base) * ks + pairs - 1) / pairs;
    const long cost = waves * ((pl->pix_tiles + ks - 1) / ks + 12) + 2L * ks;
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = ks; }
  }
  pl->ksplit = best;
  const int items = base * best;
  pl->pairs = items < pairs ? items : pairs;
  return true;
}

}  // namespace

// 0 when this kernel does not apply to the shape
size_t wgrad3x3_2sm_workspace_bytes(int N, int H, int W, int M, int Nn) {
  W2Plan pl;
  if (!make_plan(N, H, W, M, Nn, &pl)) return 0;
  return static_cast<size_t>(pl.ksplit) * 9 * M * Nn * sizeof(float);
}

// dw[co][ci][3][3] (+)= dY (M = Cout >= 256 channels) x X (Nn = Cin channels); returns 1 if the shape is not eligible
int wgrad3x3_2sm_run(const void* dy, int64_t dy_ld, const void* x, int64_t x_ld, float* dw, int accumulate, int N, int H,
                     int W, int M, int Nn, void* workspace, size_t ws_bytes, cudaStream_t stream) {
  W2Plan pl;
  if (!make_plan(N, H, W, M, Nn, &pl)) return 1;
  UNETK_CHECK(dy_ld % 8 == 0 && x_ld % 8 == 0, -1, "wgrad3x3_2sm: pixel strides must be multiples of 8");
  const size_t need = static_cast<size_t>(pl.ksplit) * 9 * M * Nn * sizeof(float);
  UNETK_CHECK(workspace != nullptr && ws_bytes >= need, -1, "wgrad3x3_2sm: workspace too small (%zu < %zu)", ws_bytes, need);
  W2Params p{};
  p.partial = static_cast<float*>(workspace);
  p.TH = pl.TH; p.TW = pl.TW; p.tiles_h = pl.tiles_h; p.tiles_w = pl.tiles_w; p.pix_tiles = pl.pix_tiles;
  p.m_pairs = pl.m_pairs; p.n_tiles = pl.n_tiles; p.ksplit = pl.ksplit; p.stages = pl.stages;
  p.M = M; p.Nn = Nn;
  p.q_box_bytes = pl.q_box_bytes; p.q_tx_bytes = pl.q_tx_bytes;
  auto mk = [&](CUtensorMap* tm, const void* base, int64_t ld, int C, int halo) -> int {
    uint64_t dims[4] = {static_cast<uint64_t>(C), static_cast<uint64_t>(W), static_cast<uint64_t>(H), static_cast<uint64_t>(N)};
    uint64_t strides[3] = {static_cast<uint64_t>(ld) * 2, static_cast<uint64_t>(ld) * 2 * W, static_cast<uint64_t>(ld) * 2 * W * H};
    uint32_t box[4] = {64, static_cast<uint32_t>(pl.TW + halo), static_cast<uint32_t>(pl.TH), 1};
    uint32_t es[4] = {1, 1, 1, 1};
    return make_tmap_bf16(tm, base, 4, dims, strides, box, es, true);
  };
  if (int rc = mk(&p.tmP, dy, dy_ld, M, 0)) return rc;
  if (int rc = mk(&p.tmQ, x, x_ld, Nn, 2)) return rc;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pl.pairs);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = pl.smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  UNETK_CUDA(cudaLaunchKernelEx(&cfg, wgrad3x3_2sm_kernel, p));
  UNETK_LAUNCHED();
  return wgrad_reduce_launch(p.partial, dw, pl.ksplit, 9, M, Nn, static_cast<int64_t>(Nn) * 9, 9, 1, accumulate, stream);
}

}  // namespace unetk
