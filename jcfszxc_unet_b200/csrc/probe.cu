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
