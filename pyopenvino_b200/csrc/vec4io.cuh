// 4-channel vector access to a feature map stored as float32 or float16 (opt-in FP16 storage mode): the arithmetic of
// every bandwidth kernel is FP32; only the bytes in HBM / shared memory change.  float16 stores round to nearest even.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace b200ov {

template <typename T> struct Vec4IO;

template <> struct Vec4IO<float> {
  static constexpr int DT = 0;     // B200OV_DT_F32
  __device__ __forceinline__ static float4 ldg(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
  __device__ __forceinline__ static float4 ld(const float* p) { return *reinterpret_cast<const float4*>(p); }
  __device__ __forceinline__ static void st(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
};

template <> struct Vec4IO<__half> {
  static constexpr int DT = 1;     // B200OV_DT_F16
  __device__ __forceinline__ static float4 unpack(uint2 u) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
  }
  __device__ __forceinline__ static float4 ldg(const __half* p) { return unpack(__ldg(reinterpret_cast<const uint2*>(p))); }
  __device__ __forceinline__ static float4 ld(const __half* p) { return unpack(*reinterpret_cast<const uint2*>(p)); }
  __device__ __forceinline__ static void st(__half* p, float4 v) {
    const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<const uint32_t*>(&a);
    u.y = *reinterpret_cast<const uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = u;
  }
};

// alignment of a 4-channel access: 16 bytes (float32) / 8 bytes (float16)
template <typename T>
inline bool aligned_vec4(const void* p) { return (reinterpret_cast<uintptr_t>(p) & (4 * sizeof(T) - 1)) == 0; }

}  // namespace b200ov
