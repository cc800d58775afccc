#include "host_common.cuh"

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace unetk {

namespace {
// errno-style: each thread sees the text of ITS last failing call (the autograd worker thread that runs a backward
// must not read, or overwrite, the message of a forward failing on the main thread)
thread_local char g_err[1024] = "";
}  // namespace

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

const char* last_error() { return g_err; }

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

int pdl_small_grid() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("UNETK_PDL_SMALL"); v = e ? atoi(e) : 0; }
  return v;
}

bool pdl_enabled() {
  static int v = -1;
  // Measured on B200 (UNet B=16 512^2, CUDA graphs, same box, 2 rounds): PDL off 24.12 / 23.85 ms per step, on 24.79 /
  // 24.76 ms.  The kernels are persistent 1-CTA-per-SM grids that fill shared memory: a dependent CTA cannot become
  // resident before its predecessor's CTA on that SM exits, so only ~1 us of prologue could overlap, and the early
  // scheduling costs more than that.  Off by default; UNETK_PDL=1 enables it (every kernel honours griddepcontrol.wait).
  if (v < 0) { const char* e = getenv("UNETK_PDL"); v = e ? atoi(e) : 0; }
  return v != 0;
}

namespace {
thread_local int g_sm_limit = 0;   // > 0: persistent grids of this thread use at most this many SMs (unetk_set_sm_limit)
}
int set_sm_limit(int n) {
  const int prev = g_sm_limit;
  g_sm_limit = n > 0 ? n : 0;
  return prev;
}

int num_sms() {
  const int real = device_sms();
  return (g_sm_limit > 0 && g_sm_limit < real) ? g_sm_limit : real;
}
// CTA pairs (clusters of two = the two SMs of a TPC) a persistent pair kernel may use.  Under an SM limit the SMs left to
// the other kernel (NCCL's CTAs while gradient buckets are in flight) may sit in as many different TPCs, each of which then
// cannot host a pair: every reserved SM takes a whole TPC out, so that each launched pair finds a free TPC at once (a pair
// that has to wait for one runs after all the others and doubles the kernel's time).
int num_sm_pairs() {
  const int real = device_sms(), lim = num_sms();
  const int pairs = real / 2 - (real - lim);
  return pairs > 0 ? pairs : 0;
}

int device_sms() {
  static std::atomic<int> cached[64];   // per device ordinal; 0 = not queried yet
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  std::atomic<int>& slot = cached[dev & 63];
  int n = slot.load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    slot.store(n, std::memory_order_relaxed);
  }
  return n;
}

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                   CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                   CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    // Resolved through the runtime so the library has no link-time dependency on libcuda.
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides,
                   bool swizzle128, int swizzle_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  UNETK_CHECK(fn != nullptr, -2, "cuTensorMapEncodeTiled not available from the driver");
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base),
                  dims, strides_bytes, box, elem_strides, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle128 ? (swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B)
                             : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims [%llu %llu %llu %llu %llu] box [%u %u %u %u %u]",
              static_cast<int>(r), rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
              (unsigned long long)(rank > 4 ? dims[4] : 0), box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0,
              rank > 3 ? box[3] : 0, rank > 4 ? box[4] : 0);
    return -2;
  }
  return 0;
}

}  // namespace unetk
