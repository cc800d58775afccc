// C entry points of the tcgen05 probes (tools only: built into tools/probe/libunetk_probe.so by tools/probe/build.py,
// never linked into libunetk.so).
#include "host_common.cuh"

namespace unetk {
int probe_run(const void* a, const void* b, float* d, int mode, int shift, int bo, cudaStream_t stream);
int probe_mma_rate_run(int N, int grid, int a_shift_rows, int two_acc, int iters, int b_tiles, long long* out,
                       cudaStream_t stream);
}  // namespace unetk
using namespace unetk;
static inline cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }

extern "C" {
int unetk_probe_umma(const void* a, const void* b, float* d, int mode, int shift, int base_offset, void* stream) {
  return probe_run(a, b, d, mode, shift, base_offset, S(stream));
}
int unetk_probe_mma_rate(int N, int grid, int a_shift_rows, int two_acc, int iters, int b_tiles, int64_t* out,
                         void* stream) {
  UNETK_CHECK(out != nullptr, -1, "probe_mma_rate: null output");
  return probe_mma_rate_run(N, grid, a_shift_rows, two_acc, iters, b_tiles, reinterpret_cast<long long*>(out), S(stream));
}
const char* unetk_probe_last_error(void) { return last_error(); }
}
