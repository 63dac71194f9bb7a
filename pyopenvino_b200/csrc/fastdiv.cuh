// Division by a runtime constant without the ~40-instruction integer divide: host precomputes a magic
// multiplier, the device does one multiply-high and a shift.  Valid for dividends < 2^31.
#pragma once
#include <stdint.h>

namespace b200ov {

struct FastDiv {
  uint32_t d, mul, shr;
  FastDiv() : d(1), mul(0), shr(0) {}
  explicit FastDiv(uint32_t divisor) : d(divisor), mul(0), shr(0) {
    if (divisor > 1) {
      uint32_t l = 0;
      while ((1ull << l) < divisor) ++l;               // l = ceil(log2(d))
      shr = l - 1;
      mul = (uint32_t)(((1ull << (32 + shr)) + divisor - 1) / divisor);   // ceil(2^(32+shr) / d), fits: d > 2^(l-1)
    }
  }
  __host__ __device__ __forceinline__ uint32_t div(uint32_t n) const {
#ifdef __CUDA_ARCH__
    return d == 1 ? n : (__umulhi(n, mul) >> shr);
#else
    return n / d;
#endif
  }
  // q = n / d, r = n % d
  __host__ __device__ __forceinline__ void divmod(uint32_t n, uint32_t& q, uint32_t& r) const {
    q = div(n);
    r = n - q * d;
  }
};

}  // namespace b200ov
