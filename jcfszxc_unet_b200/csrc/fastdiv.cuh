#pragma once
#include <cstdint>

namespace unetk {

// Division by a runtime constant as multiply-high + shift (valid for dividends < 2^31): the tile -> (image, row,
// column) decomposition runs once per tile in every epilogue thread; five hardware-less integer divisions cost
// ~120 instructions there (10 % of the epilogue's samples, profiles/r01_ncu_halo_epilogue.txt).
struct FastDiv {
  uint32_t d, mul, shr;
  FastDiv() : d(1), mul(0), shr(0) {}
  explicit FastDiv(uint32_t div) : d(div ? div : 1), mul(0), shr(0) {
    if (d > 1) {
      uint32_t lg = 0;
      while ((1ull << lg) < d) ++lg;   // ceil(log2(d))
      const uint32_t pw = 31 + lg;
      mul = static_cast<uint32_t>(((1ull << pw) + d - 1) / d);
      shr = pw - 32;
    }
  }
#ifdef __CUDACC__
  __device__ __forceinline__ uint32_t div(uint32_t x) const { return d == 1 ? x : (__umulhi(x, mul) >> shr); }
  __device__ __forceinline__ void divmod(uint32_t x, uint32_t& q, uint32_t& r) const {
    q = div(x);
    r = x - q * d;
  }
#endif
};

}  // namespace unetk
