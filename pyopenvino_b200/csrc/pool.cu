// MaxPool / AvgPool (NHWC) with the reference's exact window rules and an optional per-channel
// scale + shift epilogue (the folded BatchNorm that follows the pools of mnist_bn).
//
// MaxPool  (MaxPool.py:41-72): the input is zero-padded (np.pad 'constant', :53) and the padding
//          takes part in the max; in ceil mode windows that overhang the padded tensor are clipped
//          (min(h, ...), :69).
// AvgPool  (AvgPool.py:41-59): no padding is applied and the window is clipped at h-1 / w-1
//          (:56), so the 7x7 GoogLeNet pool averages rows/cols 0..5 only.
// Bandwidth-bound.  Two kernels:
//   pool_max_strip_kernel : the MaxPool hot case (compile-time window / stride).  A thread owns 4 consecutive
//       channels (128-bit accesses) of a vertical strip of TH output pixels and keeps the per-row maxima of
//       the rows two neighbouring windows share in registers, so a 3x3 stride-1 pool issues 3 loads per
//       output instead of 9 (the LSU / L1 wavefront rate, not HBM, is what limits the naive form).
//       32-bit index arithmetic with multiply-high division (fastdiv.cuh).
//   pool_kernel           : every other window (and AvgPool), one thread per output, runtime loops.
#include "common.cuh"
#include "fastdiv.cuh"

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

struct PoolStripP {
  int h, w, pt, pl, hp, wpad, oh, ow, x_ld, y_ld, th;
  uint32_t total;                // work items = n * strips * ceil(ow / TW) * cg
  FastDiv d_cg, d_owt, d_strips;
};

__device__ __forceinline__ float4 max4(float4 a, float4 b) {
  return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
}

// MaxPool, window KH x KW, stride S in both directions, 4 channels x TW adjacent output columns per thread.
// Per input row a thread loads the (TW-1)*S + KW columns its TW windows cover once (1.5 loads per output for a
// 3x3 stride-1 pool with TW = 4 instead of 3) and keeps the per-row maxima two vertically adjacent windows share.
template <int KH, int KW, int S, int TW>
__global__ void __launch_bounds__(256) pool_max_strip_kernel(PoolStripP p, const float* __restrict__ x,
                                                             const float* __restrict__ scale,
                                                             const float* __restrict__ shift, float* __restrict__ y) {
  constexpr int NCOL = (TW - 1) * S + KW;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < p.total; idx += stride) {
    uint32_t q, g, q2, oxt, img, strip;
    p.d_cg.divmod(idx, q, g);
    p.d_owt.divmod(q, q2, oxt);
    p.d_strips.divmod(q2, img, strip);
    const int c0 = (int)g * 4;
    const int ox0 = (int)oxt * TW;
    const int oy0 = (int)strip * p.th;
    const int oy1 = min(p.oh, oy0 + p.th);
    const float* ximg = x + (size_t)img * p.h * p.w * p.x_ld + c0;
    float* yp = y + ((size_t)(img * p.oh + oy0) * p.ow + ox0) * p.y_ld + c0;
    const int px0 = ox0 * S;
    const int ix0 = px0 - p.pl;
    // column state (fixed for the strip): inside the padded tensor / inside the real tensor / needed at all
    bool col_in_pad[NCOL], col_in_x[NCOL];
#pragma unroll
    for (int cidx = 0; cidx < NCOL; ++cidx) {
      col_in_pad[cidx] = px0 + cidx < p.wpad;
      const bool needed = ox0 + (cidx >= KW ? (cidx - KW) / S + 1 : 0) < p.ow;   // first output column that uses it exists
      col_in_x[cidx] = needed && ix0 + cidx >= 0 && ix0 + cidx < p.w;
    }
    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sf = make_float4(0.f, 0.f, 0.f, 0.f);
    if (scale != nullptr) sc = __ldg(reinterpret_cast<const float4*>(scale + c0));
    if (shift != nullptr) sf = __ldg(reinterpret_cast<const float4*>(shift + c0));
    // per output column: maximum over its window columns of padded row py; -inf when the row lies outside the padded tensor
    auto row_max = [&](int py, float4 (&m)[TW]) {
#pragma unroll
      for (int t = 0; t < TW; ++t) m[t] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      if (py >= p.hp) return;
      const int iy = py - p.pt;
      const bool row_in = iy >= 0 && iy < p.h;
      const float* xr = ximg + ((long long)(row_in ? iy : 0) * p.w + ix0) * p.x_ld;
      float4 v[NCOL];
#pragma unroll
      for (int cidx = 0; cidx < NCOL; ++cidx) {
        v[cidx] = make_float4(0.f, 0.f, 0.f, 0.f);                     // the zero padding participates
        if (row_in && col_in_x[cidx]) v[cidx] = __ldg(reinterpret_cast<const float4*>(xr + cidx * p.x_ld));
      }
#pragma unroll
      for (int t = 0; t < TW; ++t)
#pragma unroll
        for (int kx = 0; kx < KW; ++kx)
          if (col_in_pad[t * S + kx]) m[t] = max4(m[t], v[t * S + kx]);
    };
    // Two output rows per iteration: their 2*S new input rows are all requested before the first max is taken.
    constexpr int KEEP = KH > S ? KH - S : 0;        // rows shared by vertically adjacent windows
    constexpr int NR = KH + S;                       // padded rows under two vertically adjacent windows
    float4 rm[NR][TW];
#pragma unroll
    for (int r = 0; r < KEEP; ++r) row_max(oy0 * S + r, rm[r]);
    for (int oy = oy0; oy < oy1; oy += 2) {
      const bool two = oy + 1 < oy1;
#pragma unroll
      for (int r = KEEP; r < NR; ++r) row_max((two || r < KH) ? oy * S + r : p.hp, rm[r]);
#pragma unroll
      for (int v2 = 0; v2 < 2; ++v2) {
        if (v2 == 1 && !two) break;
#pragma unroll
        for (int t = 0; t < TW; ++t) {
          float4 o = rm[v2 * S][t];
#pragma unroll
          for (int r = 1; r < KH; ++r) o = max4(o, rm[v2 * S + r][t]);
          if (scale != nullptr) { o.x = __fmul_rn(o.x, sc.x); o.y = __fmul_rn(o.y, sc.y); o.z = __fmul_rn(o.z, sc.z); o.w = __fmul_rn(o.w, sc.w); }
          if (shift != nullptr) { o.x = __fadd_rn(o.x, sf.x); o.y = __fadd_rn(o.y, sf.y); o.z = __fadd_rn(o.z, sf.z); o.w = __fadd_rn(o.w, sf.w); }
          if (ox0 + t < p.ow) *reinterpret_cast<float4*>(yp + t * p.y_ld) = o;
        }
        yp += (size_t)p.ow * p.y_ld;
      }
#pragma unroll
      for (int r = 0; r < KEEP; ++r)
#pragma unroll
        for (int t = 0; t < TW; ++t) rm[r][t] = rm[r + 2 * S][t];
    }
  }
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
  if (vec && d->mode == B200OV_POOL_MAX && d->kh == d->kw && (d->kh == 2 || d->kh == 3) && d->sh == d->sw && (d->sh == 1 || d->sh == 2)) {
    PoolStripP q;
    q.h = d->h; q.w = d->w; q.pt = d->pt; q.pl = d->pl; q.hp = d->h + d->pt + d->pb;
    q.wpad = d->w + d->pl + d->pr; q.oh = d->oh; q.ow = d->ow; q.x_ld = d->x_ld; q.y_ld = d->y_ld;
    const int th = d->sh == 1 ? 4 : 8;          // strip height: measured sweep (4 / 8 / 14 / 28) per stride
    q.th = d->oh < th ? d->oh : th;
    const int tw = d->sh == 1 ? 4 : 2;
    const int strips = ceil_div(d->oh, q.th), cg = d->c / 4, owt = ceil_div(d->ow, tw);
    const long long items = (long long)d->n * strips * owt * cg;
    if (items < 0x7fffffffLL) {
      q.total = (uint32_t)items;
      q.d_cg = FastDiv(cg); q.d_owt = FastDiv(owt); q.d_strips = FastDiv(strips);
      const int g = bw_grid(items, 256);
      cudaStream_t s = as_stream(stream);
      if (d->kh == 3 && d->sh == 1) pool_max_strip_kernel<3, 3, 1, 4><<<g, 256, 0, s>>>(q, x, scale, shift, y);
      else if (d->kh == 3) pool_max_strip_kernel<3, 3, 2, 2><<<g, 256, 0, s>>>(q, x, scale, shift, y);
      else if (d->sh == 1) pool_max_strip_kernel<2, 2, 1, 4><<<g, 256, 0, s>>>(q, x, scale, shift, y);
      else pool_max_strip_kernel<2, 2, 2, 2><<<g, 256, 0, s>>>(q, x, scale, shift, y);
      B200OV_LAUNCH_CHECK("pool_max_strip_kernel");
      return B200OV_OK;
    }
  }
  long long total = (long long)d->n * d->oh * d->ow * (vec ? d->c / 4 : d->c);
  int grid = bw_grid(total, 256);
  if (vec) pool_kernel<4><<<grid, 256, 0, as_stream(stream)>>>(p, x, scale, shift, y);
  else pool_kernel<1><<<grid, 256, 0, as_stream(stream)>>>(p, x, scale, shift, y);
  B200OV_LAUNCH_CHECK("pool_kernel");
  return B200OV_OK;
}
