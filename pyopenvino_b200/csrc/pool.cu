// MaxPool / AvgPool (NHWC) with the reference's exact window rules and an optional per-channel
// scale + shift epilogue (the folded BatchNorm that follows the pools of mnist_bn).
//
// MaxPool  (MaxPool.py:41-72): the input is zero-padded (np.pad 'constant', :53) and the padding
//          takes part in the max; in ceil mode windows that overhang the padded tensor are clipped
//          (min(h, ...), :69).
// AvgPool  (AvgPool.py:41-59): no padding is applied and the window is clipped at h-1 / w-1
//          (:56), so the 7x7 GoogLeNet pool averages rows/cols 0..5 only.
// Bandwidth-bound: one thread owns V consecutive channels of one output pixel (128-bit accesses).
#include "common.cuh"

namespace b200ov {

struct PoolP {
  int n, h, w, c, kh, kw, sh, sw, pt, pl, pb, pr, oh, ow, x_ld, y_ld, mode;
};

template <int V>
__device__ __forceinline__ void loadv(const float* p, float (&v)[V]) {
  if constexpr (V == 4) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
    v[0] = __ldg(p);
  }
}
template <int V>
__device__ __forceinline__ void storev(float* p, const float (&v)[V]) {
  if constexpr (V == 4) *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  else *p = v[0];
}

template <int V>
__global__ void __launch_bounds__(256) pool_kernel(PoolP p, const float* __restrict__ x, const float* __restrict__ scale,
                                                   const float* __restrict__ shift, float* __restrict__ y) {
  const int cg = p.c / V;
  const long long total = (long long)p.n * p.oh * p.ow * cg;
  const int hp = p.h + p.pt + p.pb, wpad = p.w + p.pl + p.pr;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(idx % cg);
    long long pix = idx / cg;
    const int ox = (int)(pix % p.ow);
    long long t = pix / p.ow;
    const int oy = (int)(t % p.oh);
    const int img = (int)(t / p.oh);
    const int c0 = g * V;
    const float* ximg = x + (long long)img * p.h * p.w * p.x_ld + c0;
    float res[V];
    if (p.mode == B200OV_POOL_MAX) {
#pragma unroll
      for (int j = 0; j < V; ++j) res[j] = -INFINITY;
      const int py0 = oy * p.sh, px0 = ox * p.sw;
      const int py1 = min(hp, py0 + p.kh), px1 = min(wpad, px0 + p.kw);
      for (int py = py0; py < py1; ++py) {
        const int iy = py - p.pt;
        const bool row_in = iy >= 0 && iy < p.h;
        for (int px = px0; px < px1; ++px) {
          const int ix = px - p.pl;
          float v[V];
          if (row_in && ix >= 0 && ix < p.w) {
            loadv<V>(ximg + ((long long)iy * p.w + ix) * p.x_ld, v);
          } else {
#pragma unroll
            for (int j = 0; j < V; ++j) v[j] = 0.f;   // the zero padding participates
          }
#pragma unroll
          for (int j = 0; j < V; ++j) res[j] = fmaxf(res[j], v[j]);
        }
      }
    } else {
      const int y0 = oy * p.sh, x0 = ox * p.sw;
      const int y1 = min(p.h - 1, y0 + p.kh), x1 = min(p.w - 1, x0 + p.kw);
#pragma unroll
      for (int j = 0; j < V; ++j) res[j] = 0.f;
      int cnt = 0;
      for (int iy = y0; iy < y1; ++iy)
        for (int ix = x0; ix < x1; ++ix) {
          float v[V];
          loadv<V>(ximg + ((long long)iy * p.w + ix) * p.x_ld, v);
#pragma unroll
          for (int j = 0; j < V; ++j) res[j] = __fadd_rn(res[j], v[j]);
          ++cnt;
        }
#pragma unroll
      for (int j = 0; j < V; ++j) res[j] = __fdiv_rn(res[j], (float)cnt);   // cnt == 0 -> NaN like np.average([])
    }
    if (scale != nullptr) {
      float s[V];
      loadv<V>(scale + c0, s);
#pragma unroll
      for (int j = 0; j < V; ++j) res[j] = __fmul_rn(res[j], s[j]);
    }
    if (shift != nullptr) {
      float s[V];
      loadv<V>(shift + c0, s);
#pragma unroll
      for (int j = 0; j < V; ++j) res[j] = __fadd_rn(res[j], s[j]);
    }
    storev<V>(y + pix * p.y_ld + c0, res);
  }
}

}  // namespace b200ov

using namespace b200ov;

extern "C" int b200ov_pool2d(const b200ov_pool_desc* d, const float* x, const float* scale, const float* shift,
                             float* y, void* stream) {
  B200OV_REQUIRE(d && x && y, "pool2d: null argument");
  B200OV_REQUIRE(d->n >= 0 && d->h > 0 && d->w > 0 && d->c > 0 && d->kh > 0 && d->kw > 0 && d->sh > 0 && d->sw > 0 &&
                     d->oh > 0 && d->ow > 0 && d->pt >= 0 && d->pl >= 0 && d->pb >= 0 && d->pr >= 0,
                 "pool2d: bad geometry");
  B200OV_REQUIRE(d->x_ld >= d->c && d->y_ld >= d->c, "pool2d: channel pitch smaller than channel count");
  B200OV_REQUIRE(d->mode == B200OV_POOL_MAX || d->mode == B200OV_POOL_AVG_REF, "pool2d: bad mode");
  if (d->mode == B200OV_POOL_MAX)
    B200OV_REQUIRE((d->oh - 1) * d->sh < d->h + d->pt + d->pb && (d->ow - 1) * d->sw < d->w + d->pl + d->pr,
                   "pool2d: a window starts outside the padded input");
  PoolP p{d->n, d->h, d->w, d->c, d->kh, d->kw, d->sh, d->sw, d->pt, d->pl, d->pb, d->pr, d->oh, d->ow, d->x_ld, d->y_ld,
          d->mode};
  if (d->n == 0) return B200OV_OK;
  const bool vec = (d->c % 4 == 0) && (d->x_ld % 4 == 0) && (d->y_ld % 4 == 0) && aligned16(x) && aligned16(y) &&
                   (scale == nullptr || aligned16(scale)) && (shift == nullptr || aligned16(shift));
  long long total = (long long)d->n * d->oh * d->ow * (vec ? d->c / 4 : d->c);
  int grid = bw_grid(total, 256);
  if (vec) pool_kernel<4><<<grid, 256, 0, as_stream(stream)>>>(p, x, scale, shift, y);
  else pool_kernel<1><<<grid, 256, 0, as_stream(stream)>>>(p, x, scale, shift, y);
  B200OV_LAUNCH_CHECK("pool_kernel");
  return B200OV_OK;
}
