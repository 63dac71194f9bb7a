// FP32 -> (hi, scaled lo) FP16 pair of the f16x2 contraction:  c = hi + 2^-11 * lo  up to 2^-22 |c|  (conv_f16x2.cu).
#pragma once
#include <cuda_fp16.h>

namespace b200ov {

// scalar form (layout / packing kernels; the contraction's producers use the packed 5-instruction form)
__device__ __forceinline__ void split_f16x2(float c, __half& hi, __half& lo) {
  hi = __float2half_rn(c);
  lo = __float2half_rn((c - __half2float(hi)) * 2048.f);       // (c - hi) and its 2^11 scaling are exact in FP32
}

}  // namespace b200ov
