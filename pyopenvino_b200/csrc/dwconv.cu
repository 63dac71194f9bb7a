// Depthwise GroupConvolution (NHWC) with fused bias + activation.
//
// Reference: GroupConvolution.py:53-79, `out[g,oy,ox] = np.sum(in_pad[g, window] * w[g,0,0])`.
// Bandwidth-bound (0.9-2.2 FLOP/B): one thread owns V consecutive channels of one output pixel,
// so a warp reads/writes 32*V consecutive floats of the NHWC row (128-bit accesses for V = 4).
// The kh*kw products are rounded individually (no FMA contraction) and summed in the order numpy's
// pairwise float32 reduction uses, so the pre-bias value is bit-identical to the reference.
#include "common.cuh"

namespace b200ov {

struct DwP {
  int n, h, w, c, kh, kw, sh, sw, pt, pl, oh, ow, x_ld, y_ld, act;
  float lo, hi;
};

template <int V> struct Vec;
template <> struct Vec<4> {
  float v[4];
  __device__ static Vec load(const float* p) { float4 t = __ldg(reinterpret_cast<const float4*>(p)); return {{t.x, t.y, t.z, t.w}}; }
  __device__ void store(float* p) const { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};
template <> struct Vec<1> {
  float v[1];
  __device__ static Vec load(const float* p) { return {{__ldg(p)}}; }
  __device__ void store(float* p) const { *p = v[0]; }
};

// numpy `pairwise_sum` order for one run of KK float32 terms (KK <= 128):
//   KK < 8  : sequential
//   KK >= 8 : eight interleaved partial sums over the leading 8*floor(KK/8) terms, combined as
//             ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the remaining terms added one by one.
template <int V>
struct PairwiseAcc {
  float r[8][V];
  float tail[V];
  int count = 0;
  int kk;
  __device__ explicit PairwiseAcc(int kk_) : kk(kk_) {}
  __device__ void push(const float (&t)[V]) {
    const int full = (kk >= 8) ? (kk / 8) * 8 : 0;
    if (kk < 8) {
#pragma unroll
      for (int j = 0; j < V; ++j) tail[j] = (count == 0) ? t[j] : __fadd_rn(tail[j], t[j]);
    } else if (count < 8) {
#pragma unroll
      for (int j = 0; j < V; ++j) r[count][j] = t[j];
    } else if (count < full) {
#pragma unroll
      for (int j = 0; j < V; ++j) r[count & 7][j] = __fadd_rn(r[count & 7][j], t[j]);
    } else {
      if (count == full) {
#pragma unroll
        for (int j = 0; j < V; ++j)
          tail[j] = __fadd_rn(__fadd_rn(__fadd_rn(r[0][j], r[1][j]), __fadd_rn(r[2][j], r[3][j])),
                              __fadd_rn(__fadd_rn(r[4][j], r[5][j]), __fadd_rn(r[6][j], r[7][j])));
      }
#pragma unroll
      for (int j = 0; j < V; ++j) tail[j] = __fadd_rn(tail[j], t[j]);
    }
    ++count;
  }
  __device__ void finish(float (&out)[V]) {
    const int full = (kk >= 8) ? (kk / 8) * 8 : 0;
    if (kk >= 8 && count == full) {
#pragma unroll
      for (int j = 0; j < V; ++j)
        tail[j] = __fadd_rn(__fadd_rn(__fadd_rn(r[0][j], r[1][j]), __fadd_rn(r[2][j], r[3][j])),
                            __fadd_rn(__fadd_rn(r[4][j], r[5][j]), __fadd_rn(r[6][j], r[7][j])));
    }
#pragma unroll
    for (int j = 0; j < V; ++j) out[j] = tail[j];
  }
};

// KH/KW > 0: compile-time window (fully unrolled, the 3x3 hot case); 0: runtime window.
template <int V, int KH, int KW>
__global__ void __launch_bounds__(256) dwconv_kernel(DwP p, const float* __restrict__ x, const float* __restrict__ wp,
                                                     const float* __restrict__ bias, float* __restrict__ y) {
  const int kh = KH > 0 ? KH : p.kh;
  const int kw = KW > 0 ? KW : p.kw;
  const int cg = p.c / V;
  const long long total = (long long)p.n * p.oh * p.ow * cg;
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
    const int iy0 = oy * p.sh - p.pt, ix0 = ox * p.sw - p.pl;
    float res[V];
    if constexpr (KH == 3 && KW == 3) {
      float pr[9][V];
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int iy = iy0 + ky, ix = ix0 + kx;
          Vec<V> wv = Vec<V>::load(wp + (ky * 3 + kx) * p.c + c0);
          if (iy >= 0 && iy < p.h && ix >= 0 && ix < p.w) {
            Vec<V> xv = Vec<V>::load(ximg + ((long long)iy * p.w + ix) * p.x_ld);
#pragma unroll
            for (int j = 0; j < V; ++j) pr[ky * 3 + kx][j] = __fmul_rn(xv.v[j], wv.v[j]);
          } else {
#pragma unroll
            for (int j = 0; j < V; ++j) pr[ky * 3 + kx][j] = __fmul_rn(0.f, wv.v[j]);
          }
        }
#pragma unroll
      for (int j = 0; j < V; ++j) {
        float s01 = __fadd_rn(pr[0][j], pr[1][j]), s23 = __fadd_rn(pr[2][j], pr[3][j]);
        float s45 = __fadd_rn(pr[4][j], pr[5][j]), s67 = __fadd_rn(pr[6][j], pr[7][j]);
        res[j] = __fadd_rn(__fadd_rn(__fadd_rn(s01, s23), __fadd_rn(s45, s67)), pr[8][j]);
      }
    } else {
      PairwiseAcc<V> acc(kh * kw);
      for (int ky = 0; ky < kh; ++ky)
        for (int kx = 0; kx < kw; ++kx) {
          const int iy = iy0 + ky, ix = ix0 + kx;
          Vec<V> wv = Vec<V>::load(wp + (ky * kw + kx) * p.c + c0);
          float term[V];
          if (iy >= 0 && iy < p.h && ix >= 0 && ix < p.w) {
            Vec<V> xv = Vec<V>::load(ximg + ((long long)iy * p.w + ix) * p.x_ld);
#pragma unroll
            for (int j = 0; j < V; ++j) term[j] = __fmul_rn(xv.v[j], wv.v[j]);
          } else {
#pragma unroll
            for (int j = 0; j < V; ++j) term[j] = __fmul_rn(0.f, wv.v[j]);
          }
          acc.push(term);
        }
      acc.finish(res);
    }
    Vec<V> out;
    if (bias != nullptr) {
      Vec<V> bv = Vec<V>::load(bias + c0);
#pragma unroll
      for (int j = 0; j < V; ++j) out.v[j] = apply_act(__fadd_rn(res[j], bv.v[j]), p.act, p.lo, p.hi);
    } else {
#pragma unroll
      for (int j = 0; j < V; ++j) out.v[j] = apply_act(res[j], p.act, p.lo, p.hi);
    }
    out.store(y + pix * p.y_ld + c0);
  }
}

__global__ void pack_dw_weights_kernel(const float* __restrict__ w, float* __restrict__ out, int c, int kk) {
  int total = c * kk;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    int tap = idx / c, ch = idx % c;
    out[idx] = w[ch * kk + tap];
  }
}

}  // namespace b200ov

using namespace b200ov;

extern "C" {

int b200ov_pack_dw_weights(const float* w_g11hw, float* w_packed, int c, int kh, int kw, void* stream) {
  B200OV_REQUIRE(w_g11hw && w_packed && c > 0 && kh > 0 && kw > 0, "pack_dw_weights: bad argument");
  pack_dw_weights_kernel<<<bw_grid((long long)c * kh * kw, 256), 256, 0, as_stream(stream)>>>(w_g11hw, w_packed, c, kh * kw);
  B200OV_LAUNCH_CHECK("pack_dw_weights_kernel");
  return B200OV_OK;
}

int b200ov_dwconv2d(const b200ov_dwconv_desc* d, const float* x, const float* w_packed, const float* bias, float* y,
                    void* stream) {
  B200OV_REQUIRE(d && x && w_packed && y, "dwconv2d: null argument");
  B200OV_REQUIRE(d->n >= 0 && d->h > 0 && d->w > 0 && d->c > 0 && d->kh > 0 && d->kw > 0 && d->sh > 0 && d->sw > 0 &&
                     d->oh > 0 && d->ow > 0 && d->pt >= 0 && d->pl >= 0,
                 "dwconv2d: bad geometry");
  B200OV_REQUIRE(d->x_ld >= d->c && d->y_ld >= d->c, "dwconv2d: channel pitch smaller than channel count");
  B200OV_REQUIRE(d->kh * d->kw <= 128, "dwconv2d: window larger than 128 taps is not supported");
  B200OV_REQUIRE(d->act >= B200OV_ACT_NONE && d->act <= B200OV_ACT_SIGMOID, "dwconv2d: bad activation");
  DwP p{d->n, d->h, d->w, d->c, d->kh, d->kw, d->sh, d->sw, d->pt, d->pl, d->oh, d->ow, d->x_ld, d->y_ld, d->act,
        d->act_lo, d->act_hi};
  if (d->n == 0) return B200OV_OK;
  cudaStream_t s = as_stream(stream);
  const bool vec = (d->c % 4 == 0) && (d->x_ld % 4 == 0) && (d->y_ld % 4 == 0) && aligned16(x) && aligned16(y) &&
                   aligned16(w_packed) && (bias == nullptr || aligned16(bias));
  const bool k3 = d->kh == 3 && d->kw == 3;
  long long total = (long long)d->n * d->oh * d->ow * (vec ? d->c / 4 : d->c);
  int grid = bw_grid(total, 256);
  if (vec && k3) dwconv_kernel<4, 3, 3><<<grid, 256, 0, s>>>(p, x, w_packed, bias, y);
  else if (vec) dwconv_kernel<4, 0, 0><<<grid, 256, 0, s>>>(p, x, w_packed, bias, y);
  else if (k3) dwconv_kernel<1, 3, 3><<<grid, 256, 0, s>>>(p, x, w_packed, bias, y);
  else dwconv_kernel<1, 0, 0><<<grid, 256, 0, s>>>(p, x, w_packed, bias, y);
  B200OV_LAUNCH_CHECK("dwconv_kernel");
  return B200OV_OK;
}

}  // extern "C"
