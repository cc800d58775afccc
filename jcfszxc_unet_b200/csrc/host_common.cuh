// Host-side helpers shared by the C-ABI translation units: error reporting and TMA tensor maps.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <mutex>

namespace unetk {

// Error slot returned by unetk_last_error(); written only on failure.
void set_error(const char* fmt, ...);

#define UNETK_CHECK(cond, code, ...)  \
  do {                                \
    if (!(cond)) {                    \
      ::unetk::set_error(__VA_ARGS__); \
      return (code);                  \
    }                                 \
  } while (0)

#define UNETK_CUDA(call)                                                              \
  do {                                                                                \
    cudaError_t e__ = (call);                                                         \
    if (e__ != cudaSuccess) {                                                         \
      ::unetk::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, \
                         __LINE__);                                                   \
      return -3;                                                                      \
    }                                                                                 \
  } while (0)

// Every kernel launch site ends with this: counts the launch (unetk_launch_count) and checks for errors.
void count_launch();
#define UNETK_LAUNCHED()              \
  do {                                \
    ::unetk::count_launch();          \
    UNETK_CUDA(cudaGetLastError());   \
  } while (0)

// Programmatic dependent launch: the kernel may be scheduled while its predecessor in the stream is still draining
// (its CTAs run their prologue — barrier init, TMEM allocation, descriptor prefetch — and then block in
// griddepcontrol.wait until the predecessor has completed and flushed).  EVERY kernel launched through this helper
// must execute pdl_wait() (ptx.cuh) before its first access to global memory.  Measured slower on the UNet step
// (host_common.cu): a plain launch unless UNETK_PDL=1.
bool pdl_enabled();
int pdl_small_grid();   // UNETK_PDL_SMALL=n: grids of <= n blocks (the one-block second stages / finalize kernels) use PDL
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl_enabled() || static_cast<long long>(grid.x) * grid.y * grid.z <= pdl_small_grid()) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

long long launch_count();
int device_sms();          // SM count of the CURRENT device (cached per device)
int num_sms();             // what grids are sized for: device_sms(), or the calling thread's limit (set_sm_limit)
int num_sm_pairs();        // CTA pairs (clusters of two) a persistent pair kernel may use under that limit
int set_sm_limit(int n);   // n > 0: persistent kernels launched by THIS thread use at most n SMs; 0: all.  Returns the old limit
const char* last_error();  // the calling thread's last error text (thread-local, errno-style)

// One-time, per-device, thread-safe set-up of a kernel (cudaFuncSetAttribute for > 48 KB of dynamic shared memory is a
// per-device property of the function): `static DeviceOnce once; UNETK_CUDA(once.run([&] { return cudaFunc...; }));`
// A process that drives several devices, or calls from autograd worker threads, configures every (kernel, device) pair
// exactly once (re-entrant, no hidden process-wide state).
struct DeviceOnce {
  std::atomic<unsigned long long> done{0};
  std::mutex mu;
  template <class Fn>
  cudaError_t run(Fn&& fn) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const unsigned long long bit = 1ull << (dev & 63);
    if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
    std::lock_guard<std::mutex> lk(mu);
    if (done.load(std::memory_order_relaxed) & bit) return cudaSuccess;
    e = fn();
    if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_release);
    return e;
  }
};

// A bf16 tensor map with up to 5 dims. dims[0] is the contiguous one; strides_bytes[i] is the stride
// of dims[i+1]. OOB elements read as zero. Returns 0 or a negative code (error text set).
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides,
                   bool swizzle, int swizzle_bytes = 128);   // swizzle_bytes: 128 or 64 (box rows of 64 bytes)

}  // namespace unetk
