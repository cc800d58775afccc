// Weight cache: fp32 PyTorch-layout master weights -> bf16 tap-major packs read by the TMA.
#include "host_common.cuh"
#include "ptx.cuh"

namespace unetk {

namespace {
// src [A][B][T] fp32 -> dst_ab [T][A][B] bf16 and/or dst_ba [T][B][A] bf16
__global__ void pack_weight_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst_ab,
                                   __nv_bfloat16* __restrict__ dst_ba, int A, int B, int T) {
  const int64_t total = static_cast<int64_t>(A) * B * T;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    // iterate in destination-ab order so writes to dst_ab coalesce (weights are tiny either way)
    const int b = static_cast<int>(i % B);
    const int a = static_cast<int>((i / B) % A);
    const int t = static_cast<int>(i / (static_cast<int64_t>(A) * B));
    const __nv_bfloat16 v = __float2bfloat16_rn(src[(static_cast<int64_t>(a) * B + b) * T + t]);
    if (dst_ab) dst_ab[i] = v;
    if (dst_ba) dst_ba[(static_cast<int64_t>(t) * B + b) * A + a] = v;
  }
}
}  // namespace

int pack_weight_run(const float* src, void* dst_ab, void* dst_ba, int A, int B, int T, cudaStream_t stream) {
  const int64_t total = static_cast<int64_t>(A) * B * T;
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > 8 * num_sms()) blocks = 8 * num_sms();
  pack_weight_kernel<<<blocks, 256, 0, stream>>>(src, static_cast<__nv_bfloat16*>(dst_ab),
                                                 static_cast<__nv_bfloat16*>(dst_ba), A, B, T);
  UNETK_LAUNCHED();
  return 0;
}

}  // namespace unetk
