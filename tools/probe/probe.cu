// Hardware-semantics probe (test infrastructure, not on the product path): checks how tcgen05.mma
// interprets shared-memory descriptors whose start address is NOT aligned to the swizzle repeat.
// The answer decides whether a 3x3 conv can reuse one halo tile in smem for all nine taps.
//
// mode 0: A K-major, SWIZZLE_128B, TMA-written, start = base + shift*128 B, base_offset = bo
// mode 1: A K-major, no swizzle, chunk-major [k/8][row][8], SBO = 128 B, LBO = rows*16 B, start += shift*16 B
// mode 2: A MN-major, SWIZZLE_128B (rows are K), M = 128 as two 64-wide boxes, start = base + shift*128 B
// B is always the plain aligned K-major SWIZZLE_128B [64 n][64 k] tile.  D = 128 x 64 fp32.
#include "host_common.cuh"
#include "ptx.cuh"

namespace unetk {

namespace {
constexpr int kRows = 144;  // rows staged for A (128 + room for shifts)

struct ProbeParams {
  CUtensorMap tmA;
  CUtensorMap tmA2;  // mode 2: second 64-column half
  CUtensorMap tmB;
  float* d;
  int mode, shift, bo;
};

__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ ProbeParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;                       // up to 2 * 144 * 128 = 36864 B
  uint8_t* sB = smem + 40960;               // 64 * 128 = 8192 B (1024-aligned)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 40960 + 8192);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<64>(slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;

  if (threadIdx.x == 0) {
    uint32_t bytes = 8192;
    if (p.mode == 0) bytes += kRows * 128;
    if (p.mode == 1) bytes += kRows * 128;
    if (p.mode == 2) bytes += 2 * kRows * 128;
    mbar_expect_tx(&bars[0], bytes);
    if (p.mode == 0) {
      tma_load_2d(sA, &p.tmA, &bars[0], 0, 0);
    } else if (p.mode == 1) {
      tma_load_3d(sA, &p.tmA, &bars[0], 0, 0, 0);
    } else {
      tma_load_2d(sA, &p.tmA, &bars[0], 0, 0);
      tma_load_2d(sA + kRows * 128, &p.tmA2, &bars[0], 0, 0);
    }
    tma_load_2d(sB, &p.tmB, &bars[0], 0, 0);
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    if (p.mode == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, 64, false, false);
      for (int k = 0; k < 4; ++k) {
        uint64_t da = make_smem_desc(a0 + p.shift * 128 + k * 32, 16, 1024, kLayoutSW128, p.bo);
        uint64_t db = make_smem_desc(b0 + k * 32, 16, 1024, kLayoutSW128);
        umma_bf16(tmem, da, db, idesc, k != 0);
      }
    } else if (p.mode == 1) {
      constexpr uint32_t idesc = make_idesc_bf16(128, 64, false, false);
      for (int k = 0; k < 4; ++k) {
        // K step of 16 elements = 2 chunks of 8; chunk stride = kRows*16 B
        uint64_t da = make_smem_desc(a0 + p.shift * 16 + k * 2 * kRows * 16, kRows * 16, 128, kLayoutNone);
        uint64_t db = make_smem_desc(b0 + k * 32, 16, 1024, kLayoutSW128);
        umma_bf16(tmem, da, db, idesc, k != 0);
      }
    } else {
      constexpr uint32_t idesc = make_idesc_bf16(128, 64, true, false);
      for (int k = 0; k < 4; ++k) {
        // K step of 16 rows = 2048 B; the two 64-wide MN atoms are kRows*128 B apart (LBO)
        uint64_t da = make_smem_desc(a0 + p.shift * 128 + k * 2048, kRows * 128, 1024, kLayoutSW128, p.bo);
        uint64_t db = make_smem_desc(b0 + k * 32, 16, 1024, kLayoutSW128);
        umma_bf16(tmem, da, db, idesc, k != 0);
      }
    }
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c = 0; c < 2; ++c) {
    uint32_t r[32];
    tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c * 32, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) p.d[row * 64 + c * 32 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<64>(tmem);
  }
}
}  // namespace

// ---------------------------------------------------------------------------------------------------------
// MMA issue-rate probe: one thread issues `iters` x 4 tcgen05.mma (K = 64 per group) from operands that already
// sit in shared memory and reports clocks per MMA.  Answers "what does an M=128 x N MMA really cost when A is
// (a) a 1024-byte-aligned tile, (b) a row-shifted view of a halo tile, (c) alternating between two accumulators"
// — the numbers the thin-layer conv kernels are designed against (profiles/r01_mma_rate_probe.txt).
template <int N, int EX>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int a_shift_rows, int iters, int b_tiles,
                                                         long long* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  constexpr uint32_t kABytes = 4 * 130 * 128 + 1024;                 // a halo-sized A region
  constexpr uint32_t kAReg = (kABytes + 1023u) & ~1023u;
  uint8_t* sA = smem;
  uint8_t* sB = smem + kAReg;                                         // b_tiles x (N x 128 B)
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + b_tiles * N * 128);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5;
  // small finite bf16 values everywhere (0x3c00 = 2^-7): content does not matter, NaN/denormal garbage might
  for (uint32_t i = threadIdx.x; i < (kAReg + static_cast<uint32_t>(b_tiles) * N * 128) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  fence_proxy_async_smem();
  if (warp == 0) tmem_alloc<512>(slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (warp == 1) {
    // warp-convergent issue loop (elected lane issues): measures the tensor pipe, not the issuing thread
    const bool issue = elect_one();
    constexpr uint32_t idesc = make_idesc_bf16(128, N, false, false);
    const uint64_t a0 = make_smem_desc(smem_u32(sA) + a_shift_rows * 128, 16, 1024, kLayoutSW128);
    const uint64_t b0 = make_smem_desc(smem_u32(sB), 16, 1024, kLayoutSW128);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint64_t db0 = desc_advance(b0, static_cast<uint32_t>(it % b_tiles) * (N * 128));
      const uint64_t da0 = desc_advance(a0, static_cast<uint32_t>((it % 3) * 128));   // tap-like column shift
#pragma unroll
      for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int e = 0; e <= EX; ++e)   // EX = number of EXTRA independent accumulators the MMAs rotate over
          umma_bf16_acc_p(issue, tmem + e * N, desc_advance(da0, e * 130 * 128 + k * 32), desc_advance(db0, k * 32), idesc);
      }
    }
    umma_commit_p(issue, bar);
    mbar_wait_p(issue, bar, 0);
    const long long t1 = clock64();
    if (issue) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

template <int N, int EX>
static int mma_rate_launch(int grid, int shift, int iters, int b_tiles, long long* out, cudaStream_t s) {
  const int smem = ((4 * 130 * 128 + 1024 + 1023) & ~1023) + b_tiles * N * 128 + 64 + 1024;
  UNETK_CHECK(smem <= 227 * 1024, -1, "mma_rate: smem %d", smem);
  UNETK_CUDA((cudaFuncSetAttribute(mma_rate_kernel<N, EX>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)));
  mma_rate_kernel<N, EX><<<grid, 128, smem, s>>>(shift, iters, b_tiles, out);
  UNETK_LAUNCHED();
  return 0;
}

// out: int64 [grid] clocks for iters*4*(1+two_acc) MMAs of shape 128 x N x 16
int probe_mma_rate_run(int N, int grid, int a_shift_rows, int two_acc, int iters, int b_tiles, long long* out,
                       cudaStream_t stream) {
  UNETK_CHECK(grid >= 1 && iters >= 1 && b_tiles >= 1 && b_tiles <= 9 && a_shift_rows >= 0 && a_shift_rows < 8, -1,
              "probe_mma_rate: bad arguments");
  UNETK_CHECK((two_acc == 0 || two_acc == 1 || two_acc == 3) && (two_acc + 1) * N <= 512, -1,
              "probe_mma_rate: extra accumulators 0/1/3 within 512 TMEM columns");
#define UNETK_RATE(NN, EE) if (N == NN && two_acc == EE) return mma_rate_launch<NN, EE>(grid, a_shift_rows, iters, b_tiles, out, stream)
  UNETK_RATE(64, 0); UNETK_RATE(64, 1); UNETK_RATE(64, 3);
  UNETK_RATE(128, 0); UNETK_RATE(128, 1); UNETK_RATE(128, 3);
  UNETK_RATE(192, 0); UNETK_RATE(192, 1);
  UNETK_RATE(256, 0); UNETK_RATE(256, 1);
#undef UNETK_RATE
  UNETK_CHECK(false, -1, "probe_mma_rate: unsupported (N, extra accumulators) = (%d, %d)", N, two_acc);
  return 0;
}

// a: bf16 [144][64] (modes 0,1) or [144][128] (mode 2); b: bf16 [64][64]; d: fp32 [128][64]
int probe_run(const void* a, const void* b, float* d, int mode, int shift, int bo, cudaStream_t stream) {
  ProbeParams p{};
  p.d = d; p.mode = mode; p.shift = shift; p.bo = bo;
  uint32_t es[3] = {1, 1, 1};
  if (mode == 0) {
    uint64_t dims[2] = {64, kRows}; uint64_t st[1] = {128}; uint32_t box[2] = {64, kRows};
    if (int rc = make_tmap_bf16(&p.tmA, a, 2, dims, st, box, es, true)) return rc;
  } else if (mode == 1) {
    uint64_t dims[3] = {8, kRows, 8}; uint64_t st[2] = {128, 16}; uint32_t box[3] = {8, kRows, 8};
    if (int rc = make_tmap_bf16(&p.tmA, a, 3, dims, st, box, es, false)) return rc;
  } else {
    uint64_t dims[2] = {64, kRows}; uint64_t st[1] = {256}; uint32_t box[2] = {64, kRows};
    if (int rc = make_tmap_bf16(&p.tmA, a, 2, dims, st, box, es, true)) return rc;
    if (int rc = make_tmap_bf16(&p.tmA2, static_cast<const uint8_t*>(a) + 128, 2, dims, st, box, es, true)) return rc;
  }
  {
    uint64_t dims[2] = {64, 64}; uint64_t st[1] = {128}; uint32_t box[2] = {64, 64};
    if (int rc = make_tmap_bf16(&p.tmB, b, 2, dims, st, box, es, true)) return rc;
  }
  const int smem = 40960 + 8192 + 64 + 1024;
  UNETK_CUDA(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe_kernel<<<1, 128, smem, stream>>>(p);
  UNETK_LAUNCHED();
  return 0;
}

}  // namespace unetk
