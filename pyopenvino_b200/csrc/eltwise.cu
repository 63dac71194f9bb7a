// Elementwise tail, SoftMax and LRN (all HBM-bandwidth bound).
//
//  affine_act : Add.py:9-14 / Multiply.py:9-17 with a broadcast (scalar or per-channel) operand,
//               ReLU.py:9-12, Clamp.py:9-12, Sigmoid.py:10-13 as standalone nodes.  Multiply and
//               Add are rounded separately (no FMA) so a fused Multiply->Add chain stays
//               bit-identical to the two numpy ops.
//  binary     : Add / Multiply of two same-shape tensors.
//  softmax    : SoftMax.py:10-14, one row per image, max-shifted, warp-shuffle reductions.
//  lrn        : LRN.py:10-22 across channels (contiguous in NHWC), alpha not divided by size.
#include "common.cuh"
#include "fastdiv.cuh"
#include "vec4io.cuh"

namespace b200ov {

template <int V>
__global__ void __launch_bounds__(256) affine_act_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                         long long rows, int c, int x_ld, int y_ld, int has_scale,
                                                         const float* __restrict__ scale_vec, float scale_s,
                                                         int has_shift, const float* __restrict__ shift_vec,
                                                         float shift_s, int act, float lo, float hi) {
  B200OV_PDL_SYNC();
  const int cg = c / V;
  const long long total = rows * cg;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long row = idx / cg;
    const int c0 = (int)(idx - row * cg) * V;
    float v[V];
    if constexpr (V == 4) {
      float4 t = __ldg(reinterpret_cast<const float4*>(x + row * x_ld + c0));
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
      v[0] = __ldg(x + row * x_ld + c0);
    }
    if (has_scale) {
#pragma unroll
      for (int j = 0; j < V; ++j) v[j] = __fmul_rn(v[j], scale_vec ? __ldg(scale_vec + c0 + j) : scale_s);
    }
    if (has_shift) {
#pragma unroll
      for (int j = 0; j < V; ++j) v[j] = __fadd_rn(v[j], shift_vec ? __ldg(shift_vec + c0 + j) : shift_s);
    }
#pragma unroll
    for (int j = 0; j < V; ++j) v[j] = apply_act(v[j], act, lo, hi);
    if constexpr (V == 4) *reinterpret_cast<float4*>(y + row * y_ld + c0) = make_float4(v[0], v[1], v[2], v[3]);
    else y[row * y_ld + c0] = v[0];
  }
}

template <int V>
__global__ void __launch_bounds__(256) binary_kernel(int op, const float* __restrict__ a, const float* __restrict__ b,
                                                     float* __restrict__ y, long long count) {
  B200OV_PDL_SYNC();
  const long long total = count / V;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    if constexpr (V == 4) {
      float4 p = __ldg(reinterpret_cast<const float4*>(a) + idx), q = __ldg(reinterpret_cast<const float4*>(b) + idx);
      float4 r = op == 0 ? make_float4(__fadd_rn(p.x, q.x), __fadd_rn(p.y, q.y), __fadd_rn(p.z, q.z), __fadd_rn(p.w, q.w))
                         : make_float4(__fmul_rn(p.x, q.x), __fmul_rn(p.y, q.y), __fmul_rn(p.z, q.z), __fmul_rn(p.w, q.w));
      reinterpret_cast<float4*>(y)[idx] = r;
    } else {
      y[idx] = op == 0 ? __fadd_rn(a[idx], b[idx]) : __fmul_rn(a[idx], b[idx]);
    }
  }
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// one CTA (128 threads) per row
__global__ void __launch_bounds__(128) softmax_kernel(const float* __restrict__ x, float* __restrict__ y, int cols) {
  B200OV_PDL_SYNC();
  __shared__ float red[4];
  __shared__ float bcast;
  const float* xr = x + (long long)blockIdx.x * cols;
  float* yr = y + (long long)blockIdx.x * cols;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  float m = -INFINITY;
  for (int i = tid; i < cols; i += 128) m = fmaxf(m, xr[i]);
  m = warp_max(m);
  if (lane == 0) red[wid] = m;
  __syncthreads();
  if (tid == 0) bcast = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  __syncthreads();
  m = bcast;
  float s = 0.f;
  for (int i = tid; i < cols; i += 128) s += expf(xr[i] - m);
  s = warp_sum(s);
  __syncthreads();
  if (lane == 0) red[wid] = s;
  __syncthreads();
  if (tid == 0) bcast = (red[0] + red[1]) + (red[2] + red[3]);
  __syncthreads();
  s = bcast;
  for (int i = tid; i < cols; i += 128) yr[i] = __fdiv_rn(expf(xr[i] - m), s);
}

__device__ __forceinline__ float lrn_pow(float v, float beta);

__global__ void __launch_bounds__(256) lrn_kernel(const float* __restrict__ x, float* __restrict__ y, long long pixels,
                                                  int c, int x_ld, int y_ld, int half, float alpha, float beta,
                                                  float bias) {
  B200OV_PDL_SYNC();
  const long long total = pixels * c;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long pix = idx / c;
    const int ch = (int)(idx - pix * c);
    const float* xp = x + pix * x_ld;
    const int c_lo = max(0, ch - half), c_hi = min(c, ch + half + 1);
    float s = 0.f;
    for (int k = c_lo; k < c_hi; ++k) {
      float v = __ldg(xp + k);
      float sq = __fmul_rn(v, v);
      s = (k == c_lo) ? sq : __fadd_rn(s, sq);
    }
    float den = lrn_pow(__fadd_rn(bias, __fmul_rn(alpha, s)), beta);   // scalar fallback: full-precision powf + divide
    y[pix * y_ld + ch] = __fdiv_rn(__ldg(xp + ch), den);
  }
}

// x / v^beta for the LRN output with v = bias + alpha * sum(x^2) > 0:  x * 2^(-beta * log2 v) on the
// SFU (two MUFU ops, ~3e-7 relative error, far inside the FP32 tolerance class) instead of powf + an IEEE
// divide (~100 instructions per element, which made the kernel issue-bound at 40% of HBM bandwidth).
__device__ __forceinline__ float lrn_scale(float x, float v, float beta) {
  return __fmul_rn(x, exp2f(-beta * __log2f(v)));
}

__device__ __forceinline__ float lrn_pow(float v, float beta) { return powf(v, beta); }

// Vector form: a thread owns 4 consecutive channels of one pixel and reads the neighbouring float4s
// for the window (half <= 4), i.e. 3 128-bit loads per 4 outputs instead of 4 * (2*half + 2) scalar loads.
// HALF > 0: compile-time half-window; HALF = 0: runtime `half`.
template <int HALF, typename T = float>
__global__ void __launch_bounds__(256) lrn_vec4_kernel(const T* __restrict__ x, T* __restrict__ y, uint32_t total,
                                                       FastDiv d_cg, int x_ld, int y_ld, int half_rt, float alpha,
                                                       float beta, float bias) {
  B200OV_PDL_SYNC();
  const int half = HALF > 0 ? HALF : half_rt;
  const int cg = (int)d_cg.d;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    uint32_t pix, gu;
    d_cg.divmod(idx, pix, gu);
    const int g = (int)gu;
    const T* xp = x + (size_t)pix * x_ld + 4 * g;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 cur = Vec4IO<T>::ldg(xp);
    const float4 prv = g > 0 ? Vec4IO<T>::ldg(xp - 4) : zero;
    const float4 nxt = g + 1 < cg ? Vec4IO<T>::ldg(xp + 4) : zero;
    const float v[12] = {prv.x, prv.y, prv.z, prv.w, cur.x, cur.y, cur.z, cur.w, nxt.x, nxt.y, nxt.z, nxt.w};
    float sq[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) sq[i] = __fmul_rn(v[i], v[i]);   // out-of-range neighbours are 0: adding them is exact
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float s = 0.f;
#pragma unroll
      for (int d = -4; d <= 4; ++d)
        if (d >= -half && d <= half) s = __fadd_rn(s, sq[4 + j + d]);
      o[j] = lrn_scale(v[4 + j], __fadd_rn(bias, __fmul_rn(alpha, s)), beta);
    }
    Vec4IO<T>::st(y + (size_t)pix * y_ld + 4 * g, make_float4(o[0], o[1], o[2], o[3]));
  }
}

}  // namespace b200ov

using namespace b200ov;

extern "C" {

int b200ov_affine_act(const float* x, float* y, int64_t rows, int c, int x_ld, int y_ld, int has_scale,
                      const float* scale_vec, float scale_s, int has_shift, const float* shift_vec, float shift_s,
                      int act, float act_lo, float act_hi, void* stream) {
  B200OV_REQUIRE(x && y && rows >= 0 && c > 0 && x_ld >= c && y_ld >= c, "affine_act: bad argument");
  B200OV_REQUIRE(act >= B200OV_ACT_NONE && act <= B200OV_ACT_SIGMOID, "affine_act: bad activation");
  if (rows == 0) return B200OV_OK;
  cudaStream_t s = as_stream(stream);
  // no per-channel operand and dense rows: the channel structure does not matter, treat the tensor as one flat row
  // (e.g. Sigmoid over the SSD class scores, 91 channels: 128-bit accesses instead of scalar ones)
  if (scale_vec == nullptr && shift_vec == nullptr && x_ld == c && y_ld == c && (rows * c) % 4 == 0 && rows * c < 0x7fffffffLL &&
      aligned16(x) && aligned16(y) && c % 4 != 0) {
    const int flat = (int)(rows * c);
    launch_k(affine_act_kernel<4>, bw_grid(flat / 4, 256), 256, 0, s, x, y, 1, flat, flat, flat, has_scale, nullptr, scale_s, has_shift,
                                                             nullptr, shift_s, act, act_lo, act_hi);
    B200OV_LAUNCH_CHECK("affine_act_kernel");
    return B200OV_OK;
  }
  const bool vec = (c % 4 == 0) && (x_ld % 4 == 0) && (y_ld % 4 == 0) && aligned16(x) && aligned16(y);
  if (vec)
    launch_k(affine_act_kernel<4>, bw_grid(rows * (c / 4), 256), 256, 0, s, x, y, rows, c, x_ld, y_ld, has_scale, scale_vec, scale_s,
                                                                      has_shift, shift_vec, shift_s, act, act_lo, act_hi);
  else
    launch_k(affine_act_kernel<1>, bw_grid(rows * c, 256), 256, 0, s, x, y, rows, c, x_ld, y_ld, has_scale, scale_vec, scale_s,
                                                                has_shift, shift_vec, shift_s, act, act_lo, act_hi);
  B200OV_LAUNCH_CHECK("affine_act_kernel");
  return B200OV_OK;
}

int b200ov_binary(int op, const float* a, const float* b, float* y, int64_t count, void* stream) {
  B200OV_REQUIRE(a && b && y && count >= 0 && (op == 0 || op == 1), "binary: bad argument");
  if (count == 0) return B200OV_OK;
  cudaStream_t s = as_stream(stream);
  if (count % 4 == 0 && aligned16(a) && aligned16(b) && aligned16(y))
    launch_k(binary_kernel<4>, bw_grid(count / 4, 256), 256, 0, s, op, a, b, y, count);
  else
    launch_k(binary_kernel<1>, bw_grid(count, 256), 256, 0, s, op, a, b, y, count);
  B200OV_LAUNCH_CHECK("binary_kernel");
  return B200OV_OK;
}

int b200ov_softmax(const float* x, float* y, int rows, int cols, void* stream) {
  B200OV_REQUIRE(x && y && rows >= 0 && cols > 0, "softmax: bad argument");
  if (rows == 0) return B200OV_OK;
  launch_k(softmax_kernel, rows, 128, 0, as_stream(stream), x, y, cols);
  B200OV_LAUNCH_CHECK("softmax_kernel");
  return B200OV_OK;
}

int b200ov_lrn_st(const void* x, void* y, int dtype, int64_t pixels, int c, int x_ld, int y_ld, int size, float alpha, float beta,
                  float bias, void* stream) {
  if (dtype == B200OV_DT_F32)
    return b200ov_lrn(static_cast<const float*>(x), static_cast<float*>(y), pixels, c, x_ld, y_ld, size, alpha, beta, bias, stream);
  B200OV_REQUIRE(dtype == B200OV_DT_F16, "lrn: bad storage type");
  B200OV_REQUIRE(x && y && pixels >= 0 && c > 0 && x_ld >= c && y_ld >= c && size > 0, "lrn: bad argument");
  if (pixels == 0) return B200OV_OK;
  const long long items = pixels * (c / 4);
  if (!((c % 4 == 0) && (x_ld % 4 == 0) && (y_ld % 4 == 0) && aligned_vec4<__half>(x) && aligned_vec4<__half>(y) && size / 2 <= 4 &&
        items < 0x7fffffffLL))
    return set_error(B200OV_ERR_UNSUPPORTED, "lrn: this shape has no FP16-storage kernel");
  const int g = bw_grid(items, 256);
  const __half* xh = static_cast<const __half*>(x);
  __half* yh = static_cast<__half*>(y);
  if (size / 2 == 2)
    launch_k(lrn_vec4_kernel<2, __half>, g, 256, 0, as_stream(stream), xh, yh, (uint32_t)items, FastDiv(c / 4), x_ld, y_ld, 2, alpha, beta, bias);
  else
    launch_k(lrn_vec4_kernel<0, __half>, g, 256, 0, as_stream(stream), xh, yh, (uint32_t)items, FastDiv(c / 4), x_ld, y_ld, size / 2, alpha, beta, bias);
  B200OV_LAUNCH_CHECK("lrn_kernel");
  return B200OV_OK;
}

int b200ov_lrn(const float* x, float* y, int64_t pixels, int c, int x_ld, int y_ld, int size, float alpha, float beta,
               float bias, void* stream) {
  B200OV_REQUIRE(x && y && pixels >= 0 && c > 0 && x_ld >= c && y_ld >= c && size > 0, "lrn: bad argument");
  if (pixels == 0) return B200OV_OK;
  const long long items = pixels * (c / 4);
  const bool vec = (c % 4 == 0) && (x_ld % 4 == 0) && (y_ld % 4 == 0) && aligned16(x) && aligned16(y) && size / 2 <= 4 &&
                   items < 0x7fffffffLL;
  if (vec) {
    const int g = bw_grid(items, 256);
    if (size / 2 == 2)
      launch_k(lrn_vec4_kernel<2>, g, 256, 0, as_stream(stream), x, y, (uint32_t)items, FastDiv(c / 4), x_ld, y_ld, 2, alpha, beta, bias);
    else
      launch_k(lrn_vec4_kernel<0>, g, 256, 0, as_stream(stream), x, y, (uint32_t)items, FastDiv(c / 4), x_ld, y_ld, size / 2, alpha, beta, bias);
  } else {
    launch_k(lrn_kernel, bw_grid(pixels * c, 256), 256, 0, as_stream(stream), x, y, pixels, c, x_ld, y_ld, size / 2, alpha, beta, bias);
  }
  B200OV_LAUNCH_CHECK("lrn_kernel");
  return B200OV_OK;
}

}  // extern "C"
