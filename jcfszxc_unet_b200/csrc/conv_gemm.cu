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
//   owner, warps2-5 = epilogue (TMEM -> registers -> bf16 -> global).  Double-buffered accumulator so
//   the epilogue of tile i overlaps the MMAs of tile i+1.
// Reference semantics replaced: nn.Conv2d(k=3,p=1)/nn.Conv2d(k=1)/nn.ConvTranspose2d(k=2,s=2) as
// used by UNetFamily/utils/unet_parts.py:24-31,56-58,77 (reference), forward and input-gradient.
#include "host_common.cuh"
#include "ptx.cuh"
#include "conv_gemm.cuh"

namespace unetk {

namespace {

constexpr int kTileM = 128;   // output pixels per tile (UMMA M)
constexpr int kTileK = 64;    // channels per k-block: 64 bf16 = one 128B swizzle row
constexpr int kUmmaK = 16;
constexpr int kThreads = 192;
constexpr uint32_t kABytes = kTileM * kTileK * 2;

template <int BN>
struct Cfg {
  static constexpr uint32_t kBBytes = BN * kTileK * 2;
  static constexpr uint32_t kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr uint32_t kTmemCols = 2 * BN;  // 128 / 256 / 512: all powers of two >= 32
  static constexpr uint32_t kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_kernel(const __grid_constant__ ConvGemmParams p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024B alignment.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
  uint64_t* full_bar = bars;                     // [kStages]
  uint64_t* empty_bar = bars + C::kStages;       // [kStages]
  uint64_t* tfull_bar = bars + 2 * C::kStages;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 4);  // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<C::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const int kblocks = p.taps * p.kchunks;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------ TMA producer
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int nt = tile % p.num_n_tiles;
        const int mt = tile / p.num_n_tiles;
        const int tw = mt % p.tiles_w;
        const int th = (mt / p.tiles_w) % p.tiles_h;
        const int img = mt / (p.tiles_w * p.tiles_h);
        const int h0 = th * p.TH, w0 = tw * p.TW;
        const int q = nt / p.tiles_per_q;
        const int brow = q * p.rows_per_q + (nt % p.tiles_per_q) * BN;
        for (int t = 0; t < p.taps; ++t) {
          const int ah = p.a_step * h0 + p.dh[t];
          const int aw = p.a_step * w0 + p.dw[t];
          const int bt = p.btap[t];
          for (int kc = 0; kc < p.kchunks; ++kc) {
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            uint8_t* sa = smem + stage * C::kStageBytes;
            uint8_t* sb = sa + kABytes;
            mbar_expect_tx(&full_bar[stage], C::kStageBytes);
            tma_load_4d(sa, &p.tmA, &full_bar[stage], kc * kTileK, aw, ah, img);
            tma_load_3d(sb, &p.tmB, &full_bar[stage], kc * kTileK, brow, bt);
            if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------------------------------------ MMA issuer
      constexpr uint32_t idesc = make_idesc_bf16(kTileM, BN, false, false);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_base = smem_u32(smem + stage * C::kStageBytes);
          const uint32_t b_base = a_base + kABytes;
#pragma unroll
          for (int k = 0; k < kTileK / kUmmaK; ++k) {
            const uint64_t da = make_smem_desc(a_base + k * kUmmaK * 2, 16, 1024, kLayoutSW128);
            const uint64_t db = make_smem_desc(b_base + k * kUmmaK * 2, 16, 1024, kLayoutSW128);
            umma_bf16(d_tmem, da, db, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
          if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
      }
    }
  } else {
    // -------------------------------------------------------------- epilogue (warps 2..5)
    const int quarter = warp & 3;            // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;     // tile row == TMEM lane
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int nt = tile % p.num_n_tiles;
      const int mt = tile / p.num_n_tiles;
      const int tw = mt % p.tiles_w;
      const int th = (mt / p.tiles_w) % p.tiles_h;
      const int img = mt / (p.tiles_w * p.tiles_h);
      const int hh = th * p.TH + row / p.TW;
      const int ww = tw * p.TW + row % p.TW;
      const int q = nt / p.tiles_per_q;
      const int co0 = (nt % p.tiles_per_q) * BN;
      const bool valid = (hh < p.H) && (ww < p.W);
      const int oh = p.out_step * hh + (q >> 1);
      const int ow = p.out_step * ww + (q & 1);
      __nv_bfloat16* optr =
          p.out + (static_cast<size_t>(img) * p.Hout * p.Wout + static_cast<size_t>(oh) * p.Wout + ow) *
                      p.out_ld + co0;

      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(taddr + c * 32, r);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const int col = co0 + c * 32 + v * 8;
            if (col < p.ncols) {
              float f[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(r[v * 8 + j]);
              if (p.bias != nullptr) {
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] += __ldg(p.bias + col + j);
              }
              uint4 o;
              o.x = pack_bf16x2(f[0], f[1]);
              o.y = pack_bf16x2(f[2], f[3]);
              o.z = pack_bf16x2(f[4], f[5]);
              o.w = pack_bf16x2(f[6], f[7]);
              *reinterpret_cast<uint4*>(optr + c * 32 + v * 8) = o;
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

template <int BN>
int launch(const ConvGemmParams& p, cudaStream_t stream) {
  using C = Cfg<BN>;
  static bool configured = false;
  if (!configured) {
    UNETK_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    C::kSmemBytes));
    configured = true;
  }
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  conv_gemm_kernel<BN><<<grid, kThreads, C::kSmemBytes, stream>>>(p);
  UNETK_LAUNCHED();
  return 0;
}

}  // namespace

int conv_gemm_run(const ConvGemmDesc& d, cudaStream_t stream) {
  UNETK_CHECK(d.K % 8 == 0 && d.K >= 8, -1, "conv_gemm: K=%d must be a multiple of 8", d.K);
  UNETK_CHECK(d.ncols % 8 == 0, -1, "conv_gemm: ncols=%d must be a multiple of 8", d.ncols);
  UNETK_CHECK(d.a_ld % 8 == 0 && d.out_ld % 8 == 0, -1, "conv_gemm: pixel strides must be multiples of 8");
  UNETK_CHECK((reinterpret_cast<uintptr_t>(d.a) & 15) == 0 && (reinterpret_cast<uintptr_t>(d.out) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(d.b) & 15) == 0,
              -1, "conv_gemm: pointers must be 16-byte aligned");
  UNETK_CHECK(d.taps >= 1 && d.taps <= 9, -1, "conv_gemm: taps=%d", d.taps);

  ConvGemmParams p{};
  // ---- N tiling
  int BN;
  if (d.q_groups > 1) {
    // ConvTranspose fwd: an N tile must not straddle two output phases q.
    // (a ragged last tile reads the next phase's weight rows; those columns are masked in the epilogue)
    BN = (d.ncols % 256 == 0) ? 256 : (d.ncols % 128 == 0 ? 128 : 64);
  } else {
    BN = d.ncols >= 256 ? 256 : (d.ncols > 64 ? 128 : 64);
  }
  p.tiles_per_q = (d.ncols + BN - 1) / BN;
  p.rows_per_q = d.ncols;
  p.num_n_tiles = p.tiles_per_q * d.q_groups;
  p.ncols = d.ncols;

  // ---- M tiling: TH x TW = 128 output positions, TW a power of two
  int TW = 128;
  while (TW > d.W) TW >>= 1;
  if (TW < 1) TW = 1;
  const int TH = kTileM / TW;
  p.TH = TH; p.TW = TW;
  p.H = d.H; p.W = d.W;
  p.tiles_h = (d.H + TH - 1) / TH;
  p.tiles_w = (d.W + TW - 1) / TW;
  p.num_m_tiles = d.N * p.tiles_h * p.tiles_w;
  p.taps = d.taps;
  p.kchunks = (d.K + kTileK - 1) / kTileK;
  p.a_step = d.a_step;
  for (int t = 0; t < d.taps; ++t) { p.dh[t] = d.dh[t]; p.dw[t] = d.dw[t]; p.btap[t] = d.btap[t]; }
  p.out = reinterpret_cast<__nv_bfloat16*>(d.out);
  p.out_ld = d.out_ld;
  p.out_step = d.out_step;
  p.Hout = d.H * d.out_step;
  p.Wout = d.W * d.out_step;
  p.bias = d.bias;
  UNETK_CHECK(TW * d.a_step <= 256 && TH * d.a_step <= 256, -1, "conv_gemm: TMA box too large");

  // ---- tensor maps
  {
    const int AH = d.H * d.a_step, AW = d.W * d.a_step;  // spatial extent of the A tensor
    uint64_t dims[4] = {static_cast<uint64_t>(d.K), static_cast<uint64_t>(AW), static_cast<uint64_t>(AH),
                        static_cast<uint64_t>(d.N)};
    uint64_t strides[3] = {static_cast<uint64_t>(d.a_ld) * 2, static_cast<uint64_t>(d.a_ld) * 2 * AW,
                           static_cast<uint64_t>(d.a_ld) * 2 * AW * AH};
    uint32_t box[4] = {kTileK, static_cast<uint32_t>(TW * d.a_step), static_cast<uint32_t>(TH * d.a_step), 1};
    uint32_t es[4] = {1, static_cast<uint32_t>(d.a_step), static_cast<uint32_t>(d.a_step), 1};
    int rc = make_tmap_bf16(&p.tmA, d.a, 4, dims, strides, box, es, true);
    if (rc) return rc;
  }
  {
    const uint64_t rows = static_cast<uint64_t>(d.ncols) * d.q_groups;
    uint64_t dims[3] = {static_cast<uint64_t>(d.K), rows, static_cast<uint64_t>(d.b_taps)};
    uint64_t strides[2] = {static_cast<uint64_t>(d.K) * 2, static_cast<uint64_t>(d.K) * 2 * rows};
    uint32_t box[3] = {kTileK, static_cast<uint32_t>(BN), 1};
    uint32_t es[3] = {1, 1, 1};
    int rc = make_tmap_bf16(&p.tmB, d.b, 3, dims, strides, box, es, true);
    if (rc) return rc;
  }
  switch (BN) {
    case 256: return launch<256>(p, stream);
    case 128: return launch<128>(p, stream);
    default: return launch<64>(p, stream);
  }
}

}  // namespace unetk
