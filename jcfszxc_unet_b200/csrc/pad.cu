// F.pad of the Up block (unet_parts.py:64-67): when the skip connection is larger than the up-sampled tensor (input
// sizes that are not multiples of 16: MaxPool2d floors), the reference zero-pads the ConvTranspose output to the
// skip's size before the concat.  One streaming kernel serves forward and backward:
//   dst[n, y, x, :] = (0 <= y - oy < Hs && 0 <= x - ox < Ws) ? src[n, y - oy, x - ox, :] : 0
// forward : dst = channel slice of the concat buffer (Hd x Wd = skip size), src = dense ConvTranspose output, (oy, ox) =
//           (diffY / 2, diffX / 2) >= 0: copy + zero border in one pass;
// backward: dst = dense gradient of the ConvTranspose output, src = the concat-gradient slice, (oy, ox) negated: a crop.
// NHWC bf16, 16 bytes (8 channels) per thread; HBM-bound, 4 B per element.
#include "host_common.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

namespace unetk {

namespace {
constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads)
shift_copy_kernel(__nv_bfloat16* __restrict__ dst, int64_t dst_ld, int Hd, int Wd, const __nv_bfloat16* __restrict__ src,
                  int64_t src_ld, int Hs, int Ws, int oy, int ox, int N, int C) {
  pdl_trigger();
  pdl_wait();
  const int cg = C >> 3;
  const int64_t total = static_cast<int64_t>(N) * Hd * Wd * cg;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(kThreads) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * kThreads) {
    const int g = static_cast<int>(i % cg);
    int64_t t = i / cg;
    const int x = static_cast<int>(t % Wd); t /= Wd;
    const int y = static_cast<int>(t % Hd);
    const int n = static_cast<int>(t / Hd);
    const int sy = y - oy, sx = x - ox;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (sy >= 0 && sy < Hs && sx >= 0 && sx < Ws)
      v = __ldg(reinterpret_cast<const uint4*>(src + ((static_cast<int64_t>(n) * Hs + sy) * Ws + sx) * src_ld + g * 8));
    *reinterpret_cast<uint4*>(dst + ((static_cast<int64_t>(n) * Hd + y) * Wd + x) * dst_ld + g * 8) = v;
  }
}
}  // namespace

int shift_copy_run(void* dst, int64_t dst_ld, int Hd, int Wd, const void* src, int64_t src_ld, int Hs, int Ws, int oy, int ox,
                   int N, int C, cudaStream_t s) {
  UNETK_CHECK(C > 0 && C % 8 == 0, -1, "shift_copy: C=%d must be a multiple of 8", C);
  UNETK_CHECK(dst_ld % 8 == 0 && src_ld % 8 == 0, -1, "shift_copy: pixel strides must be multiples of 8");
  const int64_t total = static_cast<int64_t>(N) * Hd * Wd * (C / 8);
  if (total == 0) return 0;
  int64_t b = (total + kThreads - 1) / kThreads;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 16;
  if (b > cap) b = cap;
  UNETK_CUDA(launch_pdl(shift_copy_kernel, dim3(static_cast<int>(b)), dim3(kThreads), 0, s, static_cast<__nv_bfloat16*>(dst), dst_ld, Hd,
                        Wd, static_cast<const __nv_bfloat16*>(src), src_ld, Hs, Ws, oy, ox, N, C));
  UNETK_LAUNCHED();
  return 0;
}

}  // namespace unetk
