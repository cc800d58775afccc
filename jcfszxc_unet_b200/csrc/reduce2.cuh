// Ordered second stage of the two-stage reductions: out[i] = sum_j partial[addr(i, j)], j < count, in double and in
// a FIXED order (bit-identical run to run, no atomics).
//
// The first version ran one thread per output walking all `count` block partials (up to 592 dependent
// load->add steps): 20-46 us per launch, ~1 ms of the UNet step over its ~45 launches
// (profiles/r01_ncu_launch_shares_v3.txt: chan_sums / conv_stats_sums / colsum_finalize).  Here a 32x32 block owns
// 32 outputs: slice y adds the partials j = y, y+32, ... and thread row 0 adds the 32 slice sums in order.
#pragma once
#include <cuda_runtime.h>

namespace unetk {

constexpr int kSum2Lanes = 32, kSum2Slices = 32;

// Call from ALL threads of a dim3(32, 32) block; `valid` = this lane's output index is in range.  The total is
// returned to the threads with threadIdx.y == 0 (others get 0).
template <class Addr>
__device__ __forceinline__ double sliced_ordered_sum(const float* __restrict__ partial, int count, bool valid, Addr addr) {
  __shared__ double sm[kSum2Slices][kSum2Lanes + 1];
  double s = 0.0;
  if (valid) {
    for (int j = threadIdx.y; j < count; j += kSum2Slices) s += static_cast<double>(partial[addr(j)]);
  }
  sm[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.y == 0) {
#pragma unroll 8
    for (int y = 0; y < kSum2Slices; ++y) t += sm[y][threadIdx.x];
  }
  return t;
}

}  // namespace unetk
