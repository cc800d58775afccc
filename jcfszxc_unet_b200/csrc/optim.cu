// Optimizer tail over FLAT fp32 buffers (all parameters / gradients / RMSprop state are views into
// four contiguous arrays): global-norm clip + RMSprop with momentum in two passes.
// Reference semantics replaced: torch.nn.utils.clip_grad_norm_(params, 1.0) and optim.RMSprop.step
// (train.py:107-112,299-300; alpha = 0.99, eps = 1e-8 are the torch defaults the reference relies on).
#include "host_common.cuh"
#include "ptx.cuh"

namespace unetk {

namespace {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads) sqnorm_kernel(const float* __restrict__ g, int64_t n,
                                                          float* __restrict__ partial) {
  pdl_trigger();
  pdl_wait();
  float s = 0.f;
  const int64_t n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (int64_t i = blockIdx.x * static_cast<int64_t>(kThreads) + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * kThreads) {
    const float4 v = __ldg(g4 + i);
    s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) { const float v = g[(n4 << 2) + threadIdx.x]; s = fmaf(v, v, s); }
  __shared__ float red[kThreads / 32];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < kThreads / 32; ++i) t += red[i];
    partial[blockIdx.x] = t;
  }
}

// out[0] = total L2 norm, out[1] = clip coefficient min(1, max_norm / (norm + 1e-6)); gscale pre-multiplies
// the gradient (1/world_size after a sum all-reduce).
__global__ void clip_finalize_kernel(const float* __restrict__ partial, int nblk, float gscale, float max_norm,
                                     float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  // one warp: lane l adds partials l, l+32, ... (independent loads), lane 0 adds the 32 lane sums in order: a fixed
  // order, bit-identical run to run (one thread walking the ~600 partials took 20 us)
  const int lane = threadIdx.x & 31;
  double s = 0.0;
  for (int b = lane; b < nblk; b += 32) s += partial[b];
  double tot = 0.0;
#pragma unroll
  for (int l = 0; l < 32; ++l) tot += __shfl_sync(0xffffffffu, s, l);
  if (threadIdx.x != 0) return;
  const float norm = static_cast<float>(sqrt(tot)) * gscale;
  float coef = max_norm / (norm + 1e-6f);
  if (coef > 1.f) coef = 1.f;
  out[0] = norm;
  out[1] = coef * gscale;
}

__device__ __forceinline__ void rmsprop_one(float& pv, float gv0, float& sqv, float& bv, float coef, float lr, float alpha,
                                            float eps, float wd, float momentum) {
  const float gv = fmaf(wd, pv, gv0 * coef);
  const float s = fmaf(1.f - alpha, gv * gv, alpha * sqv);
  sqv = s;
  const float avg = sqrtf(s) + eps;
  float step = gv / avg;
  if (momentum > 0.f) {
    const float b = fmaf(momentum, bv, step);
    bv = b;
    step = b;
  }
  pv = fmaf(-lr, step, pv);
}

// 16-byte vectors, two per thread and array in flight (the scalar version below ran at 72 % of the HBM peak)
__global__ void __launch_bounds__(kThreads)
rmsprop4_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ sq, float4* __restrict__ buf,
                int64_t n4, float lr, float alpha, float eps, float wd, float momentum, const float* __restrict__ clip,
                const float* __restrict__ hyper) {
  pdl_trigger();
  pdl_wait();
  // hyper != NULL: {lr, alpha, eps, weight_decay, momentum} live in device memory, so a CUDA graph that captured this
  // launch follows later changes of the learning rate (ReduceLROnPlateau in the reference loop, train.py:114-122,355)
  if (hyper) { lr = __ldg(hyper); alpha = __ldg(hyper + 1); eps = __ldg(hyper + 2); wd = __ldg(hyper + 3); momentum = __ldg(hyper + 4); }
  const float coef = clip ? __ldg(clip + 1) : 1.f;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kThreads;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(kThreads) + threadIdx.x; i < n4; i += 2 * stride) {
    const int64_t i1 = i + stride;
    const bool two = i1 < n4;
    float4 pv[2], gv[2], sv[2], bv[2];
    pv[0] = p[i]; gv[0] = __ldg(g + i); sv[0] = sq[i]; bv[0] = momentum > 0.f ? buf[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    if (two) { pv[1] = p[i1]; gv[1] = __ldg(g + i1); sv[1] = sq[i1]; bv[1] = momentum > 0.f ? buf[i1] : make_float4(0.f, 0.f, 0.f, 0.f); }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      if (k == 1 && !two) break;
      rmsprop_one(pv[k].x, gv[k].x, sv[k].x, bv[k].x, coef, lr, alpha, eps, wd, momentum);
      rmsprop_one(pv[k].y, gv[k].y, sv[k].y, bv[k].y, coef, lr, alpha, eps, wd, momentum);
      rmsprop_one(pv[k].z, gv[k].z, sv[k].z, bv[k].z, coef, lr, alpha, eps, wd, momentum);
      rmsprop_one(pv[k].w, gv[k].w, sv[k].w, bv[k].w, coef, lr, alpha, eps, wd, momentum);
      const int64_t j = k ? i1 : i;
      p[j] = pv[k]; sq[j] = sv[k];
      if (momentum > 0.f) buf[j] = bv[k];
    }
  }
}

__global__ void __launch_bounds__(kThreads)
rmsprop_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ sq, float* __restrict__ buf,
               int64_t n, float lr, float alpha, float eps, float wd, float momentum,
               const float* __restrict__ clip, const float* __restrict__ hyper) {
  pdl_trigger();
  pdl_wait();
  if (hyper) { lr = __ldg(hyper); alpha = __ldg(hyper + 1); eps = __ldg(hyper + 2); wd = __ldg(hyper + 3); momentum = __ldg(hyper + 4); }
  const float coef = clip ? __ldg(clip + 1) : 1.f;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(kThreads) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * kThreads) {
    float pv = p[i];
    const float gv = fmaf(wd, pv, g[i] * coef);
    const float s = fmaf(1.f - alpha, gv * gv, alpha * sq[i]);
    sq[i] = s;
    const float avg = sqrtf(s) + eps;
    float step = gv / avg;
    if (momentum > 0.f) {
      const float b = fmaf(momentum, buf[i], step);
      buf[i] = b;
      step = b;
    }
    p[i] = fmaf(-lr, step, pv);
  }
}

}  // namespace

int sqnorm_blocks(int64_t n) {
  int64_t b = (n / 4 + kThreads * 4 - 1) / (kThreads * 4);
  const int64_t cap = static_cast<int64_t>(num_sms()) * 4;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

int grad_clip_coef_run(const float* g, int64_t n, float gscale, float max_norm, float* partial, float* out,
                       cudaStream_t s) {
  UNETK_CHECK((reinterpret_cast<uintptr_t>(g) & 15) == 0, -1, "grad buffer must be 16-byte aligned");
  const int nb = sqnorm_blocks(n);
  UNETK_CUDA(launch_pdl(sqnorm_kernel, dim3(nb), dim3(kThreads), 0, s, g, n, partial));
  UNETK_LAUNCHED();
  UNETK_CUDA(launch_pdl(clip_finalize_kernel, dim3(1), dim3(32), 0, s, partial, nb, gscale, max_norm, out));
  UNETK_LAUNCHED();
  return 0;
}

int rmsprop_run(float* p, const float* g, float* sq, float* buf, int64_t n, float lr, float alpha, float eps, float wd,
                float momentum, const float* clip, const float* hyper, cudaStream_t s) {
  const int64_t cap = static_cast<int64_t>(num_sms()) * 8;
  const bool aligned = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(sq) |
                         reinterpret_cast<uintptr_t>(buf)) & 15) == 0;
  const int64_t n4 = aligned ? n / 4 : 0;
  if (n4 > 0) {
    int64_t b = (n4 + kThreads * 2 - 1) / (kThreads * 2);
    if (b > cap) b = cap;
    UNETK_CUDA(launch_pdl(rmsprop4_kernel, dim3(static_cast<int>(b)), dim3(kThreads), 0, s, reinterpret_cast<float4*>(p),
                          reinterpret_cast<const float4*>(g), reinterpret_cast<float4*>(sq), reinterpret_cast<float4*>(buf), n4,
                          lr, alpha, eps, wd, momentum, clip, hyper));
    UNETK_LAUNCHED();
  }
  const int64_t done = n4 * 4, rest = n - done;   // tail (or everything, for unaligned buffers): scalar kernel
  if (rest > 0) {
    int64_t b = (rest + kThreads * 4 - 1) / (kThreads * 4);
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    UNETK_CUDA(launch_pdl(rmsprop_kernel, dim3(static_cast<int>(b)), dim3(kThreads), 0, s, p + done, g + done, sq + done,
                          buf + done, rest, lr, alpha, eps, wd, momentum, clip, hyper));
    UNETK_LAUNCHED();
  }
  return 0;
}

}  // namespace unetk
