// Layout glue: batched 2-D transpose (NCHW <-> NHWC, MatMul transpose flags), the fused
// Parameter pre-processing (NCHW -> NHWC with per-channel scale / shift) and strided row copies
// (Concat when a producer could not write in place).  All bandwidth-bound; the transposes go
// through a padded 32x32 shared-memory tile so both the read and the write are coalesced.
#include <cuda_fp16.h>

#include "common.cuh"
#include "f16split.cuh"

namespace b200ov {

// Host inputs arrive in their native width (uint8 camera frames, FP16, FP32; Parameter.py:13 casts with
// `.astype(precision)`) and are widened here, on the device, instead of on the host before the PCIe copy.
// Every widening is exact, so the result is bit-identical to the reference's host-side cast.
__device__ __forceinline__ float widen(float v) { return v; }
__device__ __forceinline__ float widen(__half v) { return __half2float(v); }
__device__ __forceinline__ float widen(uint8_t v) { return (float)v; }
__device__ __forceinline__ float widen(int8_t v) { return (float)v; }
template <typename T>
__device__ __forceinline__ float load_widen(const T* p) { return widen(__ldg(p)); }
__device__ __forceinline__ void store_narrow(float* p, float v) { *p = v; }
__device__ __forceinline__ void store_narrow(__half* p, float v) { *p = __float2half_rn(v); }

// x: [batch][rows][x_ld] (cols valid)  ->  y: [batch][cols][y_ld] (rows valid)
// AFFINE: v = v*scale[row] + shift[row] (row = source row = channel of an NCHW tensor)
template <bool AFFINE, typename TIN = float, typename TOUT = float>
__global__ void __launch_bounds__(256) transpose_kernel(const TIN* __restrict__ x, TOUT* __restrict__ y, int rows,
                                                        int cols, int x_ld, int y_ld, int tiles_r, int tiles_c,
                                                        int has_scale, const float* __restrict__ scale_vec,
                                                        float scale_s, int has_shift,
                                                        const float* __restrict__ shift_vec, float shift_s) {
  B200OV_PDL_SYNC();
  __shared__ float tile[32][33];
  long long bid = blockIdx.x;
  const int tc = (int)(bid % tiles_c);
  bid /= tiles_c;
  const int tr = (int)(bid % tiles_r);
  const int b = (int)(bid / tiles_r);
  const TIN* xb = x + (long long)b * rows * x_ld;
  TOUT* yb = y + (long long)b * cols * y_ld;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    int r = tr * 32 + ty + i, c = tc * 32 + tx;
    float v = 0.f;
    if (r < rows && c < cols) {
      v = load_widen(xb + (long long)r * x_ld + c);
      if (AFFINE) {
        if (has_scale) v = __fmul_rn(v, scale_vec ? __ldg(scale_vec + r) : scale_s);
        if (has_shift) v = __fadd_rn(v, shift_vec ? __ldg(shift_vec + r) : shift_s);
      }
    }
    tile[ty + i][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    int c = tc * 32 + ty + i, r = tr * 32 + tx;
    if (r < rows && c < cols) store_narrow(yb + (long long)c * y_ld + r, tile[tx][ty + i]);
  }
}

// NCHW -> NHWC for few channels (the network input: C = 1 or 3).  The tiled transpose above wastes most
// of its 32-row tile here; instead a thread owns one pixel, reads its C planes (coalesced across the warp)
// and writes the pixel's channel run.  PAD = 4 / 8: y_ld == PAD >= C, 128-bit stores with the unused lanes
// zeroed (the tcgen05 stem convolution gathers whole 8-channel runs and needs finite padding).
template <int MAXC, int PAD, typename TIN = float>
__global__ void __launch_bounds__(256) nchw_to_nhwc_smallc_kernel(const TIN* __restrict__ x, float* __restrict__ y,
                                                                  long long pixels, int c, int hw, int y_ld, int has_scale,
                                                                  const float* __restrict__ scale_vec, float scale_s,
                                                                  int has_shift, const float* __restrict__ shift_vec,
                                                                  float shift_s) {
  B200OV_PDL_SYNC();
  float sc[MAXC], sf[MAXC];
#pragma unroll
  for (int ch = 0; ch < MAXC; ++ch) {
    sc[ch] = (has_scale && ch < c) ? (scale_vec ? __ldg(scale_vec + ch) : scale_s) : 1.f;
    sf[ch] = (has_shift && ch < c) ? (shift_vec ? __ldg(shift_vec + ch) : shift_s) : 0.f;
  }
  for (long long pix = blockIdx.x * (long long)blockDim.x + threadIdx.x; pix < pixels;
       pix += (long long)gridDim.x * blockDim.x) {
    const long long img = pix / hw;
    const int r = (int)(pix - img * hw);
    const TIN* xp = x + img * c * hw + r;
    float v[MAXC];
#pragma unroll
    for (int ch = 0; ch < MAXC; ++ch) {
      v[ch] = 0.f;
      if (ch < c) {
        v[ch] = load_widen(xp + (long long)ch * hw);
        if (has_scale) v[ch] = __fmul_rn(v[ch], sc[ch]);
        if (has_shift) v[ch] = __fadd_rn(v[ch], sf[ch]);
      }
    }
    if constexpr (PAD == 44) {
      // pre-split for the stem contraction: [hi(c0,c1) hi(c2,c3) lo(c0,c1) lo(c2,c3)], 16 bytes per pixel like 4 floats
      __half h[4], l[4];
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) split_f16x2(v[ch], h[ch], l[ch]);
      const __half2 h01 = __halves2half2(h[0], h[1]), h23 = __halves2half2(h[2], h[3]);
      const __half2 l01 = __halves2half2(l[0], l[1]), l23 = __halves2half2(l[2], l[3]);
      uint4 o;
      o.x = *reinterpret_cast<const uint32_t*>(&h01); o.y = *reinterpret_cast<const uint32_t*>(&h23);
      o.z = *reinterpret_cast<const uint32_t*>(&l01); o.w = *reinterpret_cast<const uint32_t*>(&l23);
      *reinterpret_cast<uint4*>(y + pix * 4) = o;
    } else if constexpr (PAD == 4) {
      *reinterpret_cast<float4*>(y + pix * 4) = make_float4(v[0], v[1], v[2], v[3]);
    } else if constexpr (PAD == 8) {
      *reinterpret_cast<float4*>(y + pix * 8) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(y + pix * 8 + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
#pragma unroll
      for (int ch = 0; ch < MAXC; ++ch)
        if (ch < c) y[pix * y_ld + ch] = v[ch];
    }
  }
}

// The same with four adjacent pixels per thread: one vector load per plane (16 bytes of FP32, 8 of FP16, 4 of uint8 / int8:
// 512 / 256 / 128 bytes per warp and plane instead of 128 / 64 / 32) and four 128-bit stores (2 KB contiguous per warp).
// SPLIT: the stem contraction's pre-split (hi, lo) pair form instead of FP32 (see PAD == 44 above; same bits).
// Needs hw % 4 == 0, a source aligned to four elements and y_ld == 4.
template <typename TIN>
__device__ __forceinline__ void load4_widen(const TIN* p, float (&f)[4]) {
  if constexpr (sizeof(TIN) == 4) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    f[0] = t.x; f[1] = t.y; f[2] = t.z; f[3] = t.w;
  } else if constexpr (sizeof(TIN) == 2) {
    const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
    const __half2 a = *reinterpret_cast<const __half2*>(&t.x), b = *reinterpret_cast<const __half2*>(&t.y);
    f[0] = __low2float(a); f[1] = __high2float(a); f[2] = __low2float(b); f[3] = __high2float(b);
  } else {
    const uint32_t word = __ldg(reinterpret_cast<const uint32_t*>(p));
#pragma unroll
    for (int px = 0; px < 4; ++px) {
      const uint32_t b = (word >> (8 * px)) & 0xffu;
      f[px] = TIN(-1) < TIN(0) ? (float)(int8_t)b : (float)b;
    }
  }
}

template <typename TIN, bool SPLIT>
__global__ void __launch_bounds__(256) nchw_to_nhwc4_x4_kernel(const TIN* __restrict__ x, float* __restrict__ y,
                                                               long long quads, int c, int hw, int has_scale,
                                                               const float* __restrict__ scale_vec, float scale_s,
                                                               int has_shift, const float* __restrict__ shift_vec,
                                                               float shift_s) {
  B200OV_PDL_SYNC();
  float sc[4], sf[4];
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    sc[ch] = (has_scale && ch < c) ? (scale_vec ? __ldg(scale_vec + ch) : scale_s) : 1.f;
    sf[ch] = (has_shift && ch < c) ? (shift_vec ? __ldg(shift_vec + ch) : shift_s) : 0.f;
  }
  const int qhw = hw >> 2;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < quads; q += (long long)gridDim.x * blockDim.x) {
    const long long img = q / qhw;
    const int r = (int)(q - img * qhw) * 4;
    const TIN* xp = x + img * c * hw + r;
    float v[4][4];                                   // [pixel][channel]
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      float f[4] = {0.f, 0.f, 0.f, 0.f};
      if (ch < c) {
        load4_widen(xp + (long long)ch * hw, f);
#pragma unroll
        for (int px = 0; px < 4; ++px) {
          if (has_scale) f[px] = __fmul_rn(f[px], sc[ch]);
          if (has_shift) f[px] = __fadd_rn(f[px], sf[ch]);
        }
      }
#pragma unroll
      for (int px = 0; px < 4; ++px) v[px][ch] = f[px];
    }
    uint4* yp = reinterpret_cast<uint4*>(y + (img * hw + r) * 4);
#pragma unroll
    for (int px = 0; px < 4; ++px) {
      uint4 o;
      if constexpr (SPLIT) {
        __half h[4], l[4];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) split_f16x2(v[px][ch], h[ch], l[ch]);
        const __half2 h01 = __halves2half2(h[0], h[1]), h23 = __halves2half2(h[2], h[3]);
        const __half2 l01 = __halves2half2(l[0], l[1]), l23 = __halves2half2(l[2], l[3]);
        o.x = *reinterpret_cast<const uint32_t*>(&h01); o.y = *reinterpret_cast<const uint32_t*>(&h23);
        o.z = *reinterpret_cast<const uint32_t*>(&l01); o.w = *reinterpret_cast<const uint32_t*>(&l23);
      } else {
        o.x = __float_as_uint(v[px][0]); o.y = __float_as_uint(v[px][1]); o.z = __float_as_uint(v[px][2]); o.w = __float_as_uint(v[px][3]);
      }
      yp[px] = o;
    }
  }
}

template <typename TIN>
static bool x4_ok(const TIN* x, const float* y, int hw, int y_ld) {
  return y_ld == 4 && aligned16(y) && hw % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & (4 * sizeof(TIN) - 1)) == 0;
}

template <int V>
__global__ void __launch_bounds__(256) copy2d_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                     long long rows, int cols, int src_ld, int dst_ld) {
  B200OV_PDL_SYNC();
  const int cg = cols / V;
  const long long total = rows * cg;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / cg;
    const int c0 = (int)(idx - r * cg) * V;
    if constexpr (V == 4)
      *reinterpret_cast<float4*>(dst + r * dst_ld + c0) = __ldg(reinterpret_cast<const float4*>(src + r * src_ld + c0));
    else
      dst[r * dst_ld + c0] = __ldg(src + r * src_ld + c0);
  }
}

// Concat of up to B200OV_CONCAT_MAX_PARTS row blocks in ONE launch: dst[r][off_p + c] = src_p[r][c].  The 2-D Concats of the SSD heads
// (six [n][h*w*273] class-score blocks -> [n][1917*91]) were six launches of a few microseconds of work each.
struct ConcatP {
  const float* src[B200OV_CONCAT_MAX_PARTS];
  int cols[B200OV_CONCAT_MAX_PARTS], off[B200OV_CONCAT_MAX_PARTS + 1];
  int nparts, total;
};
template <int V>
__global__ void __launch_bounds__(256) concat_rows_kernel(const ConcatP p, float* __restrict__ dst, long long rows) {
  B200OV_PDL_SYNC();
  // blockIdx.y walks the rows, the threads of a row walk its columns: no division per element
  const int tg = p.total / V;
  for (long long r = blockIdx.y; r < rows; r += gridDim.y) {
    float* drow = dst + r * p.total;
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < tg; g += gridDim.x * blockDim.x) {
      const int c0 = g * V;
      int part = 0;
#pragma unroll
      for (int i = 1; i < B200OV_CONCAT_MAX_PARTS; ++i)
        if (i < p.nparts && c0 >= p.off[i]) part = i;
      const float* sp = p.src[part] + r * p.cols[part] + (c0 - p.off[part]);
      if constexpr (V == 4) *reinterpret_cast<float4*>(drow + c0) = __ldg(reinterpret_cast<const float4*>(sp));
      else drow[c0] = __ldg(sp);
    }
  }
}

template <typename TIN, typename TOUT>
__global__ void __launch_bounds__(256) copy2d_cast_kernel(const TIN* __restrict__ src, TOUT* __restrict__ dst, long long rows, int cols,
                                                          int src_ld, int dst_ld) {
  B200OV_PDL_SYNC();
  // two elements per thread where the row pitch allows (cols, src_ld, dst_ld even): 32-bit / 64-bit accesses
  const long long total = rows * cols;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / cols;
    const int c = (int)(idx - r * cols);
    store_narrow(dst + r * dst_ld + c, load_widen(src + r * src_ld + c));
  }
}

template <typename TIN, typename TOUT>
static int launch_transpose_st(const TIN* x, TOUT* y, int batch, int rows, int cols, int x_ld, int y_ld, cudaStream_t s) {
  const int tiles_r = ceil_div(rows, 32), tiles_c = ceil_div(cols, 32);
  const long long blocks = (long long)batch * tiles_r * tiles_c;
  if (blocks > 0x7fffffffLL) return set_error(B200OV_ERR_INVALID, "transpose: grid too large");
  if (blocks == 0) return B200OV_OK;
  launch_k(transpose_kernel<false, TIN, TOUT>, (unsigned)blocks, 256, 0, s, x, y, rows, cols, x_ld, y_ld, tiles_r, tiles_c, 0, nullptr, 0.f, 0,
                                                                     nullptr, 0.f);
  B200OV_LAUNCH_CHECK("transpose_kernel");
  return B200OV_OK;
}

static int launch_transpose(bool affine, const float* x, float* y, int batch, int rows, int cols, int x_ld, int y_ld,
                            int has_scale, const float* scale_vec, float scale_s, int has_shift,
                            const float* shift_vec, float shift_s, cudaStream_t s) {
  int tiles_r = ceil_div(rows, 32), tiles_c = ceil_div(cols, 32);
  long long blocks = (long long)batch * tiles_r * tiles_c;
  if (blocks > 0x7fffffffLL) return set_error(B200OV_ERR_INVALID, "transpose: grid too large");
  if (blocks == 0) return B200OV_OK;
  if (affine)
    launch_k(transpose_kernel<true>, (unsigned)blocks, 256, 0, s, x, y, rows, cols, x_ld, y_ld, tiles_r, tiles_c, has_scale,
                                                           scale_vec, scale_s, has_shift, shift_vec, shift_s);
  else
    launch_k(transpose_kernel<false>, (unsigned)blocks, 256, 0, s, x, y, rows, cols, x_ld, y_ld, tiles_r, tiles_c, 0, nullptr,
                                                            0.f, 0, nullptr, 0.f);
  B200OV_LAUNCH_CHECK("transpose_kernel");
  return B200OV_OK;
}

template <typename TIN>
static int input_to_nhwc_typed(const TIN* x, float* y, int n, int c, int hw, int y_ld, int has_scale, const float* scale_vec,
                               float scale_s, int has_shift, const float* shift_vec, float shift_s, cudaStream_t s) {
  if (c <= 4) {
    const long long pixels = (long long)n * hw;
    if (x4_ok(x, y, hw, y_ld)) {
      launch_k(nchw_to_nhwc4_x4_kernel<TIN, false>, bw_grid(pixels / 4, 256), 256, 0, s, x, y, pixels / 4, c, hw, has_scale, scale_vec, scale_s,
                                                                                  has_shift, shift_vec, shift_s);
      B200OV_LAUNCH_CHECK("nchw_to_nhwc4_x4_kernel");
      return B200OV_OK;
    }
    if (y_ld == 8 && aligned16(y))
      launch_k(nchw_to_nhwc_smallc_kernel<4, 8, TIN>, bw_grid(pixels, 256), 256, 0, s, x, y, pixels, c, hw, y_ld, has_scale, scale_vec,
                                                                                scale_s, has_shift, shift_vec, shift_s);
    else if (y_ld == 4 && aligned16(y))
      launch_k(nchw_to_nhwc_smallc_kernel<4, 4, TIN>, bw_grid(pixels, 256), 256, 0, s, x, y, pixels, c, hw, y_ld, has_scale, scale_vec,
                                                                                scale_s, has_shift, shift_vec, shift_s);
    else
      launch_k(nchw_to_nhwc_smallc_kernel<4, 0, TIN>, bw_grid(pixels, 256), 256, 0, s, x, y, pixels, c, hw, y_ld, has_scale, scale_vec,
                                                                                scale_s, has_shift, shift_vec, shift_s);
    B200OV_LAUNCH_CHECK("nchw_to_nhwc_smallc_kernel");
    return B200OV_OK;
  }
  const int tiles_r = ceil_div(c, 32), tiles_c = ceil_div(hw, 32);
  const long long blocks = (long long)n * tiles_r * tiles_c;
  if (blocks > 0x7fffffffLL) return set_error(B200OV_ERR_INVALID, "input_to_nhwc: grid too large");
  launch_k(transpose_kernel<true, TIN>, (unsigned)blocks, 256, 0, s, x, y, c, hw, hw, y_ld, tiles_r, tiles_c, has_scale, scale_vec,
                                                              scale_s, has_shift, shift_vec, shift_s);
  B200OV_LAUNCH_CHECK("transpose_kernel");
  return B200OV_OK;
}

template <typename TIN>
static int input_to_nhwc_split_typed(const TIN* x, float* y, int n, int c, int hw, int has_scale, const float* scale_vec, float scale_s,
                                     int has_shift, const float* shift_vec, float shift_s, cudaStream_t s) {
  const long long pixels = (long long)n * hw;
  if (x4_ok(x, y, hw, 4)) {
    launch_k(nchw_to_nhwc4_x4_kernel<TIN, true>, bw_grid(pixels / 4, 256), 256, 0, s, x, y, pixels / 4, c, hw, has_scale, scale_vec, scale_s,
                                                                               has_shift, shift_vec, shift_s);
    B200OV_LAUNCH_CHECK("nchw_to_nhwc4_x4_kernel");
    return B200OV_OK;
  }
  launch_k(nchw_to_nhwc_smallc_kernel<4, 44, TIN>, bw_grid(pixels, 256), 256, 0, s, x, y, pixels, c, hw, 4, has_scale, scale_vec, scale_s,
                                                                            has_shift, shift_vec, shift_s);
  B200OV_LAUNCH_CHECK("nchw_to_nhwc_smallc_kernel");
  return B200OV_OK;
}

// plain (non 4-D) inputs: widen only
template <typename TIN>
__global__ void __launch_bounds__(256) widen_kernel(const TIN* __restrict__ x, float* __restrict__ y, long long count) {
  B200OV_PDL_SYNC();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
    y[i] = load_widen(x + i);
}

}  // namespace b200ov

using namespace b200ov;

extern "C" {

int b200ov_transpose(const float* x, float* y, int batch, int rows, int cols, int x_ld, int y_ld, void* stream) {
  B200OV_REQUIRE(x && y && batch >= 0 && rows > 0 && cols > 0 && x_ld >= cols && y_ld >= rows, "transpose: bad argument");
  return launch_transpose(false, x, y, batch, rows, cols, x_ld, y_ld, 0, nullptr, 0.f, 0, nullptr, 0.f, as_stream(stream));
}

int b200ov_nchw_to_nhwc_affine(const float* x, float* y, int n, int c, int hw, int y_ld, int has_scale,
                               const float* scale_vec, float scale_s, int has_shift, const float* shift_vec,
                               float shift_s, void* stream) {
  return b200ov_input_to_nhwc(x, B200OV_DT_F32, y, n, c, hw, y_ld, has_scale, scale_vec, scale_s, has_shift, shift_vec,
                              shift_s, stream);
}

int b200ov_input_to_nhwc_split(const void* x, int dtype, void* y, int n, int c, int hw, int has_scale, const float* scale_vec,
                               float scale_s, int has_shift, const float* shift_vec, float shift_s, void* stream) {
  B200OV_REQUIRE(x && y && n >= 0 && c > 0 && c <= 4 && hw > 0 && aligned16(y), "input_to_nhwc_split: bad argument (needs C <= 4)");
  if (n == 0) return B200OV_OK;
  float* yf = static_cast<float*>(y);
  cudaStream_t s = as_stream(stream);
  switch (dtype) {
    case B200OV_DT_F32: return input_to_nhwc_split_typed(static_cast<const float*>(x), yf, n, c, hw, has_scale, scale_vec, scale_s, has_shift, shift_vec, shift_s, s);
    case B200OV_DT_F16: return input_to_nhwc_split_typed(static_cast<const __half*>(x), yf, n, c, hw, has_scale, scale_vec, scale_s, has_shift, shift_vec, shift_s, s);
    case B200OV_DT_U8: return input_to_nhwc_split_typed(static_cast<const uint8_t*>(x), yf, n, c, hw, has_scale, scale_vec, scale_s, has_shift, shift_vec, shift_s, s);
    case B200OV_DT_I8: return input_to_nhwc_split_typed(static_cast<const int8_t*>(x), yf, n, c, hw, has_scale, scale_vec, scale_s, has_shift, shift_vec, shift_s, s);
    default: return set_error(B200OV_ERR_INVALID, "input_to_nhwc_split: unknown element type %d", dtype);
  }
}

int b200ov_input_to_nhwc(const void* x, int dtype, float* y, int n, int c, int hw, int y_ld, int has_scale,
                         const float* scale_vec, float scale_s, int has_shift, const float* shift_vec, float shift_s,
                         void* stream) {
  B200OV_REQUIRE(x && y && n >= 0 && c > 0 && hw > 0 && y_ld >= c, "input_to_nhwc: bad argument");
  B200OV_REQUIRE(dtype == B200OV_DT_F32 || dtype == B200OV_DT_F16 || dtype == B200OV_DT_U8 || dtype == B200OV_DT_I8,
                 "input_to_nhwc: unknown element type %d", dtype);
  if (n == 0) return B200OV_OK;
  switch (dtype) {
    case B200OV_DT_F16:
      return input_to_nhwc_typed(static_cast<const __half*>(x), y, n, c, hw, y_ld, has_scale, scale_vec, scale_s, has_shift,
                                 shift_vec, shift_s, as_stream(stream));
    case B200OV_DT_U8:
      return input_to_nhwc_typed(static_cast<const uint8_t*>(x), y, n, c, hw, y_ld, has_scale, scale_vec, scale_s, has_shift,
                                 shift_vec, shift_s, as_stream(stream));
    case B200OV_DT_I8:
      return input_to_nhwc_typed(static_cast<const int8_t*>(x), y, n, c, hw, y_ld, has_scale, scale_vec, scale_s, has_shift,
                                 shift_vec, shift_s, as_stream(stream));
    default:
      return input_to_nhwc_typed(static_cast<const float*>(x), y, n, c, hw, y_ld, has_scale, scale_vec, scale_s, has_shift,
                                 shift_vec, shift_s, as_stream(stream));
  }
}

int b200ov_transpose_st(const void* x, int x_dtype, void* y, int y_dtype, int batch, int rows, int cols, int x_ld, int y_ld,
                        void* stream) {
  B200OV_REQUIRE(x && y && batch >= 0 && rows > 0 && cols > 0 && x_ld >= cols && y_ld >= rows, "transpose: bad argument");
  cudaStream_t s = as_stream(stream);
  const bool xh = x_dtype == B200OV_DT_F16, yh = y_dtype == B200OV_DT_F16;
  B200OV_REQUIRE((xh || x_dtype == B200OV_DT_F32) && (yh || y_dtype == B200OV_DT_F32), "transpose: bad storage type");
  if (xh && yh) return launch_transpose_st(static_cast<const __half*>(x), static_cast<__half*>(y), batch, rows, cols, x_ld, y_ld, s);
  if (xh) return launch_transpose_st(static_cast<const __half*>(x), static_cast<float*>(y), batch, rows, cols, x_ld, y_ld, s);
  if (yh) return launch_transpose_st(static_cast<const float*>(x), static_cast<__half*>(y), batch, rows, cols, x_ld, y_ld, s);
  return launch_transpose_st(static_cast<const float*>(x), static_cast<float*>(y), batch, rows, cols, x_ld, y_ld, s);
}

int b200ov_copy2d_st(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t rows, int cols, int src_ld, int dst_ld,
                     void* stream) {
  B200OV_REQUIRE(src && dst && rows >= 0 && cols > 0 && src_ld >= cols && dst_ld >= cols, "copy2d: bad argument");
  const bool sh = src_dtype == B200OV_DT_F16, dh = dst_dtype == B200OV_DT_F16;
  B200OV_REQUIRE((sh || src_dtype == B200OV_DT_F32) && (dh || dst_dtype == B200OV_DT_F32), "copy2d: bad storage type");
  if (rows == 0) return B200OV_OK;
  if (!sh && !dh) return b200ov_copy2d(static_cast<const float*>(src), static_cast<float*>(dst), rows, cols, src_ld, dst_ld, stream);
  cudaStream_t s = as_stream(stream);
  if (sh && dh && cols % 2 == 0 && src_ld % 2 == 0 && dst_ld % 2 == 0 && (reinterpret_cast<uintptr_t>(src) & 3u) == 0 &&
      (reinterpret_cast<uintptr_t>(dst) & 3u) == 0)       // half -> half: a float copy of half the width
    return b200ov_copy2d(static_cast<const float*>(src), static_cast<float*>(dst), rows, cols / 2, src_ld / 2, dst_ld / 2, stream);
  const int g = bw_grid(rows * cols, 256);
  if (sh && dh) launch_k(copy2d_cast_kernel<__half, __half>, g, 256, 0, s, static_cast<const __half*>(src), static_cast<__half*>(dst), rows, cols, src_ld, dst_ld);
  else if (sh) launch_k(copy2d_cast_kernel<__half, float>, g, 256, 0, s, static_cast<const __half*>(src), static_cast<float*>(dst), rows, cols, src_ld, dst_ld);
  else launch_k(copy2d_cast_kernel<float, __half>, g, 256, 0, s, static_cast<const float*>(src), static_cast<__half*>(dst), rows, cols, src_ld, dst_ld);
  B200OV_LAUNCH_CHECK("copy2d_cast_kernel");
  return B200OV_OK;
}

int b200ov_widen(const void* x, int dtype, float* y, int64_t count, void* stream) {
  B200OV_REQUIRE(x && y && count >= 0, "widen: bad argument");
  if (count == 0) return B200OV_OK;
  cudaStream_t s = as_stream(stream);
  const int g = bw_grid(count, 256);
  switch (dtype) {
    case B200OV_DT_F32: B200OV_CUDA(cudaMemcpyAsync(y, x, (size_t)count * 4, cudaMemcpyDeviceToDevice, s)); return B200OV_OK;
    case B200OV_DT_F16: launch_k(widen_kernel<__half>, g, 256, 0, s, static_cast<const __half*>(x), y, count); break;
    case B200OV_DT_U8: launch_k(widen_kernel<uint8_t>, g, 256, 0, s, static_cast<const uint8_t*>(x), y, count); break;
    case B200OV_DT_I8: launch_k(widen_kernel<int8_t>, g, 256, 0, s, static_cast<const int8_t*>(x), y, count); break;
    default: return set_error(B200OV_ERR_INVALID, "widen: unknown element type %d", dtype);
  }
  B200OV_LAUNCH_CHECK("widen_kernel");
  return B200OV_OK;
}

int b200ov_concat_rows(int nparts, const float* const* srcs, const int* cols, float* dst, int64_t rows, void* stream) {
  B200OV_REQUIRE(srcs && cols && dst && rows >= 0 && nparts >= 1 && nparts <= B200OV_CONCAT_MAX_PARTS, "concat_rows: bad argument");
  ConcatP p;
  memset(&p, 0, sizeof(p));
  p.nparts = nparts;
  bool vec = aligned16(dst);
  long long off = 0;
  for (int i = 0; i < nparts; ++i) {
    B200OV_REQUIRE(srcs[i] && cols[i] > 0, "concat_rows: bad part %d", i);
    p.src[i] = srcs[i]; p.cols[i] = cols[i]; p.off[i] = (int)off;
    off += cols[i];
    vec = vec && cols[i] % 4 == 0 && aligned16(srcs[i]);
  }
  B200OV_REQUIRE(off <= 0x7fffffffLL, "concat_rows: row too long");
  p.off[nparts] = p.total = (int)off;
  if (rows == 0) return B200OV_OK;
  cudaStream_t s = as_stream(stream);
  const int v = vec ? 4 : 1;
  const int gy = (int)(rows < 4096 ? rows : 4096);
  int gx = ceil_div(p.total / v, 256 * 4);                    // ~4 elements per thread and row
  const int cap = ceil_div(bw_grid(rows * (p.total / v), 256), gy);
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  if (vec) launch_k(concat_rows_kernel<4>, dim3(gx, gy), 256, 0, s, p, dst, (long long)rows);
  else launch_k(concat_rows_kernel<1>, dim3(gx, gy), 256, 0, s, p, dst, (long long)rows);
  B200OV_LAUNCH_CHECK("concat_rows_kernel");
  return B200OV_OK;
}

int b200ov_copy2d(const float* src, float* dst, int64_t rows, int cols, int src_ld, int dst_ld, void* stream) {
  B200OV_REQUIRE(src && dst && rows >= 0 && cols > 0 && src_ld >= cols && dst_ld >= cols, "copy2d: bad argument");
  if (rows == 0) return B200OV_OK;
  const bool vec = (cols % 4 == 0) && (src_ld % 4 == 0) && (dst_ld % 4 == 0) && aligned16(src) && aligned16(dst);
  cudaStream_t s = as_stream(stream);
  if (vec) launch_k(copy2d_kernel<4>, bw_grid(rows * (cols / 4), 256), 256, 0, s, src, dst, rows, cols, src_ld, dst_ld);
  else launch_k(copy2d_kernel<1>, bw_grid(rows * cols, 256), 256, 0, s, src, dst, rows, cols, src_ld, dst_ld);
  B200OV_LAUNCH_CHECK("copy2d_kernel");
  return B200OV_OK;
}

}  // extern "C"
