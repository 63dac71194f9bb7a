// FP32 -> (hi, scaled lo) FP16 pair of the f16x2 contraction:  c = hi + 2^-11 * lo  up to 2^-22 |c|  (conv_f16x2.cu).
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace b200ov {

// scalar form (layout / packing kernels; the contraction's producers use the packed 5-instruction form)
__device__ __forceinline__ void split_f16x2(float c, __half& hi, __half& lo) {
  hi = __float2half_rn(c);
  lo = __float2half_rn((c - __half2float(hi)) * 2048.f);       // (c - hi) and its 2^11 scaling are exact in FP32
}

// packed form: (c0, c1) -> FP16 hi pair and scaled-residual lo pair.  (c - hi) is exact in FP32 and so is its 2^11 scaling,
// hence fma(hi, -2^11, 2^11 c) is exact too; the FMA is the mixed-precision one (FHFMA: f16 x f16 + f32 -> f32), which reads hi
// straight from the packed pair: five instructions per pair (F2FP, FMUL2, 2 FHFMA, F2FP).  Same bits as split_f16x2.
__device__ __forceinline__ void split_pair_f16x2(float c0, float c1, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(c0, c1);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  unsigned long long c2, k2, s2;
  asm("mov.b64 %0, {%1, %2};" : "=l"(c2) : "f"(c0), "f"(c1));
  asm("mov.b64 %0, {%1, %1};" : "=l"(k2) : "f"(2048.f));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(s2) : "l"(c2), "l"(k2));
  float sx, sy, r0, r1;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(sx), "=f"(sy) : "l"(s2));
  asm("{\n\t.reg .b16 h0, h1, m;\n\t"
      "mov.b32 {h0, h1}, %2;\n\t"
      "mov.b16 m, 0xE800;\n\t"                       // -2048 as an FP16 number
      "fma.rn.f32.f16 %0, h0, m, %3;\n\t"
      "fma.rn.f32.f16 %1, h1, m, %4;\n\t}"
      : "=f"(r0), "=f"(r1)
      : "r"(hi), "f"(sx), "f"(sy));
  const __half2 l = __floats2half2_rn(r0, r1);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// four consecutive channels -> the 16-byte group of a B200OV_DT_HL tensor: [hi(c0,c1) hi(c2,c3) lo(c0,c1) lo(c2,c3)]
__device__ __forceinline__ float4 encode_hl4(float c0, float c1, float c2, float c3) {
  uint32_t h01, l01, h23, l23;
  split_pair_f16x2(c0, c1, h01, l01);
  split_pair_f16x2(c2, c3, h23, l23);
  return make_float4(__uint_as_float(h01), __uint_as_float(h23), __uint_as_float(l01), __uint_as_float(l23));
}

}  // namespace b200ov
