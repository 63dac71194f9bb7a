// FP32 CUDA-core implicit-GEMM convolution (NHWC) with fused bias + activation epilogue.
//
// GEMM view (reference: Convolution.py:57-87, im2col + np.dot):
//   D[M = n*oh*ow pixels][N = cout] = A[M][K = kh*kw*cin] * W[K][N]
// A is never materialised: each CTA gathers its BM x 16 slice of the im2col matrix straight from
// the NHWC activation (zero fill == the reference's np.pad, Convolution.py:63) into shared memory,
// register-prefetching the next slice while the current one is multiplied (2-stage pipeline).
// This is the exact-FP32 path: it serves the tiny-K stem layers (C_in = 1 / 3), odd shapes, and
// is the in-library cross-check for the tcgen05 3xTF32 path (gemm_tcgen05.cu).
#include <stdlib.h>

#include "common.cuh"
#include "f16split.cuh"
#include "fastdiv.cuh"

namespace b200ov {

struct ConvP {
  int n, h, w, cin, cout, kh, kw, sh, sw, pt, pl, oh, ow, x_ld, y_ld, ldw;
  int M, K, ohow, nb_n;
  int act;
  float lo, hi;
};

constexpr int BK = 16;
constexpr int NT = 256;

template <int BM, int BN, int TM, int TN, bool VEC>
__global__ void __launch_bounds__(NT) conv_ffma_kernel(ConvP p, const float* __restrict__ x,
                                                       const float* __restrict__ wp,
                                                       const float* __restrict__ bias,
                                                       float* __restrict__ y) {
  B200OV_PDL_SYNC();
  static_assert((BM / TM) * (BN / TN) == NT, "thread tiling must cover the CTA tile");
  constexpr int APAD = 4;
  __shared__ __align__(16) float As[2][BK][BM + APAD];
  __shared__ __align__(16) float Bs[2][BK][BN];

  const int tid = threadIdx.x;
  const int m_blk = blockIdx.x / p.nb_n;
  const int n_blk = blockIdx.x % p.nb_n;
  const int m0 = m_blk * BM;
  const int n0 = n_blk * BN;

  // ---- A gather bookkeeping ------------------------------------------------------------------
  // VEC : thread owns (row = idx/4, kvec = idx%4) for idx = tid + i*NT  -> one float4 along ci
  // !VEC: thread owns row = tid % BM and k columns (tid / BM) + i * (NT / BM)
  constexpr int A_ITERS = VEC ? (BM * BK / 4) / NT : (BM * BK) / NT;
  static_assert(A_ITERS >= 1, "tile too small");
  constexpr int A_ROWS = VEC ? A_ITERS : 1;
  int a_iy0[A_ROWS], a_ix0[A_ROWS];
  long long a_img[A_ROWS];
  bool a_ok[A_ROWS];
#pragma unroll
  for (int i = 0; i < A_ROWS; ++i) {
    int row = VEC ? (tid + i * NT) / 4 : tid % BM;
    int m = m0 + row;
    a_ok[i] = m < p.M;
    int mm = a_ok[i] ? m : 0;
    int img = mm / p.ohow;
    int r = mm - img * p.ohow;
    int oy = r / p.ow;
    int ox = r - oy * p.ow;
    a_iy0[i] = oy * p.sh - p.pt;
    a_ix0[i] = ox * p.sw - p.pl;
    a_img[i] = (long long)img * p.h * p.w * p.x_ld;
  }

  float4 a_reg4[VEC ? A_ITERS : 1];
  float a_reg[VEC ? 1 : A_ITERS];
  constexpr int B_ITERS = (BK * BN / 4 + NT - 1) / NT;
  float4 b_reg[B_ITERS];

  auto load_tile = [&](int k0) {
    if (VEC) {
      const int kv = tid & 3;
      const int k = k0 + kv * 4;
      int ky = 0, kx = 0, ci = k;
      if (p.kh * p.kw != 1) {
        int tap = k / p.cin;
        ci = k - tap * p.cin;
        ky = tap / p.kw;
        kx = tap - ky * p.kw;
      }
#pragma unroll
      for (int i = 0; i < A_ITERS; ++i) {
        int iy = a_iy0[i] + ky, ix = a_ix0[i] + kx;
        bool ok = a_ok[i] && k < p.K && iy >= 0 && iy < p.h && ix >= 0 && ix < p.w;
        a_reg4[i] = ok ? __ldg(reinterpret_cast<const float4*>(x + a_img[i] + ((long long)iy * p.w + ix) * p.x_ld + ci))
                       : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {
#pragma unroll
      for (int i = 0; i < A_ITERS; ++i) {
        const int k = k0 + tid / BM + i * (NT / BM);
        int tap = k / p.cin;
        int ci = k - tap * p.cin;
        int ky = tap / p.kw;
        int kx = tap - ky * p.kw;
        int iy = a_iy0[0] + ky, ix = a_ix0[0] + kx;
        bool ok = a_ok[0] && k < p.K && iy >= 0 && iy < p.h && ix >= 0 && ix < p.w;
        a_reg[i] = ok ? __ldg(x + a_img[0] + ((long long)iy * p.w + ix) * p.x_ld + ci) : 0.f;
      }
    }
#pragma unroll
    for (int i = 0; i < B_ITERS; ++i) {
      int idx = tid + i * NT;
      if (idx < BK * BN / 4) {
        int kk = idx / (BN / 4), nv = idx % (BN / 4);
        b_reg[i] = __ldg(reinterpret_cast<const float4*>(wp + (long long)(k0 + kk) * p.ldw + n0 + nv * 4));
      }
    }
  };

  auto store_tile = [&](int buf) {
    if (VEC) {
      const int kv = tid & 3;
#pragma unroll
      for (int i = 0; i < A_ITERS; ++i) {
        int row = (tid + i * NT) / 4;
        As[buf][kv * 4 + 0][row] = a_reg4[i].x;
        As[buf][kv * 4 + 1][row] = a_reg4[i].y;
        As[buf][kv * 4 + 2][row] = a_reg4[i].z;
        As[buf][kv * 4 + 3][row] = a_reg4[i].w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < A_ITERS; ++i) As[buf][tid / BM + i * (NT / BM)][tid % BM] = a_reg[i];
    }
#pragma unroll
    for (int i = 0; i < B_ITERS; ++i) {
      int idx = tid + i * NT;
      if (idx < BK * BN / 4) {
        int kk = idx / (BN / 4), nv = idx % (BN / 4);
        *reinterpret_cast<float4*>(&Bs[buf][kk][nv * 4]) = b_reg[i];
      }
    }
  };

  const int tx = tid % (BN / TN);
  const int ty = tid / (BN / TN);
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int ktiles = (p.K + BK - 1) / BK;
  load_tile(0);
  store_tile(0);
  __syncthreads();
  int cur = 0;
  for (int t = 0; t < ktiles; ++t) {
    if (t + 1 < ktiles) load_tile((t + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        float4 v = *reinterpret_cast<const float4*>(&As[cur][kk][ty * TM + i]);
        a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
      }
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        float4 v = *reinterpret_cast<const float4*>(&Bs[cur][kk][tx * TN + j]);
        b[j] = v.x; b[j + 1] = v.y; b[j + 2] = v.z; b[j + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (t + 1 < ktiles) {
      store_tile(cur ^ 1);
      __syncthreads();
      cur ^= 1;
    }
  }

  // ---- epilogue: + bias, activation, NHWC store ------------------------------------------------
  const int nbase = n0 + tx * TN;
  float bv[TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) bv[j] = (bias != nullptr && nbase + j < p.cout) ? __ldg(bias + nbase + j) : 0.f;
  const bool vec_store = ((p.y_ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(y) & 15u) == 0);
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int m = m0 + ty * TM + i;
    if (m >= p.M) continue;
    float* yrow = y + (long long)m * p.y_ld;
#pragma unroll
    for (int j = 0; j < TN; j += 4) {
      int nn = nbase + j;
      float v0 = apply_act(acc[i][j] + bv[j], p.act, p.lo, p.hi);
      float v1 = apply_act(acc[i][j + 1] + bv[j + 1], p.act, p.lo, p.hi);
      float v2 = apply_act(acc[i][j + 2] + bv[j + 2], p.act, p.lo, p.hi);
      float v3 = apply_act(acc[i][j + 3] + bv[j + 3], p.act, p.lo, p.hi);
      if (vec_store && nn + 3 < p.cout) {
        *reinterpret_cast<float4*>(yrow + nn) = make_float4(v0, v1, v2, v3);
      } else {
        if (nn < p.cout) yrow[nn] = v0;
        if (nn + 1 < p.cout) yrow[nn + 1] = v1;
        if (nn + 2 < p.cout) yrow[nn + 2] = v2;
        if (nn + 3 < p.cout) yrow[nn + 3] = v3;
      }
    }
  }
}

__global__ void pack_conv_weights_kernel(const float* __restrict__ w, float* __restrict__ out, int cout, int cin,
                                         int kh, int kw, int K, int rows, int ldw) {
  long long total = (long long)rows * ldw;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int k = (int)(idx / ldw), n = (int)(idx % ldw);
    float v = 0.f;
    if (k < K && n < cout) {
      int tap = k / cin, ci = k % cin;
      int ky = tap / kw, kx = tap % kw;
      v = w[(((long long)n * cin + ci) * kh + ky) * kw + kx];
    }
    out[idx] = v;
  }
}

template <int BM, int BN, int TM, int TN>
static int launch_cfg(const ConvP& p0, bool vec, const float* x, const float* wp, const float* bias, float* y,
                      cudaStream_t s) {
  ConvP p = p0;
  p.nb_n = ceil_div(p.cout, BN);
  long long blocks = (long long)ceil_div(p.M, BM) * p.nb_n;
  if (blocks > 0x7fffffffLL) return set_error(B200OV_ERR_INVALID, "conv2d grid too large");
  if (vec)
    launch_k(conv_ffma_kernel<BM, BN, TM, TN, true>, (unsigned)blocks, NT, 0, s, p, x, wp, bias, y);
  else
    launch_k(conv_ffma_kernel<BM, BN, TM, TN, false>, (unsigned)blocks, NT, 0, s, p, x, wp, bias, y);
  B200OV_LAUNCH_CHECK("conv_ffma_kernel");
  return B200OV_OK;
}

// C_in = 1 stems (MNIST: 28 x 28 x 1 -> 32 / 64 channels, K = 9 / 25): the layer is an output-write stream (K FMAs per 4 bytes
// written), which the tiled implicit GEMM above turns into 16-wide K slices of mostly padding (mnist_bn conv2d at batch 1024:
// 0.101 ms for 103 MB of output).  Direct form: C_out / 4 consecutive threads own one output pixel (a float4 of channels each,
// so a warp writes 512 contiguous bytes), the taps come from L1 and the filter from shared memory.  Same FP32 FMA arithmetic
// class as conv_ffma_kernel (taps accumulated in (ky, kx) order, bias added last).
constexpr int C1_MAX_COUT = 64;
// HL: the result is written as the (hi, lo) FP16 pairs the f16x2 contraction that reads it would otherwise compute per tap
// (B200OV_DT_HL; same bits downstream, see conv_f16x2.cu).
template <int KH, int KW, bool HL>
__global__ void __launch_bounds__(256) conv_c1_direct_kernel(ConvP p, FastDiv d_ohow, FastDiv d_ow, const float* __restrict__ x,
                                                             const float* __restrict__ wp, const float* __restrict__ bias,
                                                             float* __restrict__ y) {
  B200OV_PDL_SYNC();
  // thread -> (pixel lane, channel group): the group is fixed per thread (256 % (C_out / 4) == 0), so its KH x KW x 4 filter
  // values and bias stay in registers; every index is 32-bit (the host checks the sizes)
  const int cg = p.cout >> 2, c0 = ((int)threadIdx.x % cg) * 4, ppb = 256 / cg;
  float4 w4[KH * KW];
#pragma unroll
  for (int t = 0; t < KH * KW; ++t) w4[t] = __ldg(reinterpret_cast<const float4*>(wp + t * p.ldw + c0));
  const float4 b4 = bias != nullptr ? make_float4(__ldg(bias + c0), __ldg(bias + c0 + 1), __ldg(bias + c0 + 2), __ldg(bias + c0 + 3))
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
  for (uint32_t m = blockIdx.x * ppb + threadIdx.x / cg; m < (uint32_t)p.M; m += gridDim.x * ppb) {
    uint32_t img, r, oy, ox;
    d_ohow.divmod(m, img, r);
    d_ow.divmod(r, oy, ox);
    const int iy0 = (int)oy * p.sh - p.pt, ix0 = (int)ox * p.sw - p.pl;
    const float* xi = x + img * (uint32_t)(p.h * p.w) + iy0 * p.w + ix0;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int ky = 0; ky < KH; ++ky) {
      const bool row_ok = (unsigned)(iy0 + ky) < (unsigned)p.h;
#pragma unroll
      for (int kx = 0; kx < KW; ++kx) {
        const bool ok = row_ok && (unsigned)(ix0 + kx) < (unsigned)p.w;       // zero padding (Convolution.py:63)
        const float v = ok ? __ldg(xi + ky * p.w + kx) : 0.f;
        const float4 w = w4[ky * KW + kx];
        acc.x = fmaf(v, w.x, acc.x); acc.y = fmaf(v, w.y, acc.y); acc.z = fmaf(v, w.z, acc.z); acc.w = fmaf(v, w.w, acc.w);
      }
    }
    acc.x = apply_act(acc.x + b4.x, p.act, p.lo, p.hi); acc.y = apply_act(acc.y + b4.y, p.act, p.lo, p.hi);
    acc.z = apply_act(acc.z + b4.z, p.act, p.lo, p.hi); acc.w = apply_act(acc.w + b4.w, p.act, p.lo, p.hi);
    if constexpr (HL) acc = encode_hl4(acc.x, acc.y, acc.z, acc.w);
    *reinterpret_cast<float4*>(y + (size_t)m * p.y_ld + c0) = acc;
  }
}

// shapes the direct C_in = 1 kernel takes (3x3 / 5x5, C_out a multiple of 4 with a power-of-two number of channel groups)
bool conv2d_c1_direct_ok(const b200ov_conv_desc* d, const void* x, const void* wp, const void* y) {
  const int cg = d->cout / 4;
  return d->cin == 1 && d->x_ld == 1 && d->x_dtype == B200OV_DT_F32 && d->cout % 4 == 0 && d->cout <= C1_MAX_COUT && cg > 0 &&
         (cg & (cg - 1)) == 0 && d->y_ld % 4 == 0 && aligned16(y) && aligned16(wp) && x != nullptr &&
         (long long)d->n * d->h * d->w < 0x7fffffffLL && ((d->kh == 3 && d->kw == 3) || (d->kh == 5 && d->kw == 5)) &&
         (d->y_dtype == B200OV_DT_F32 || (d->y_dtype == B200OV_DT_HL && d->act != B200OV_ACT_SIGMOID)) &&
         getenv("B200OV_NO_C1_DIRECT") == nullptr;
}

int conv2d_ffma(const b200ov_conv_desc* d, const float* x, const float* wp, const float* bias, float* y,
                cudaStream_t s) {
  ConvP p;
  p.n = d->n; p.h = d->h; p.w = d->w; p.cin = d->cin; p.cout = d->cout; p.kh = d->kh; p.kw = d->kw;
  p.sh = d->sh; p.sw = d->sw; p.pt = d->pt; p.pl = d->pl; p.oh = d->oh; p.ow = d->ow;
  p.x_ld = d->x_ld; p.y_ld = d->y_ld; p.ldw = d->ldw;
  p.ohow = d->oh * d->ow;
  long long M = (long long)d->n * p.ohow;
  if (M > 0x7fffffffLL) return set_error(B200OV_ERR_INVALID, "conv2d: too many output pixels");
  p.M = (int)M;
  p.K = d->kh * d->kw * d->cin;
  p.act = d->act; p.lo = d->act_lo; p.hi = d->act_hi;
  p.nb_n = 1;
  if (p.M == 0) return B200OV_OK;
  if (conv2d_c1_direct_ok(d, x, wp, y)) {
    const int cg1 = d->cout / 4, ppb = 256 / cg1;
    const int grid = bw_grid((long long)ceil_div(p.M, ppb) * 256, 256);
    const bool hl = d->y_dtype == B200OV_DT_HL;
    const FastDiv d_ohow(p.ohow), d_ow(d->ow);
    if (d->kh == 3 && hl) launch_k(conv_c1_direct_kernel<3, 3, true>, grid, 256, 0, s, p, d_ohow, d_ow, x, wp, bias, y);
    else if (d->kh == 3) launch_k(conv_c1_direct_kernel<3, 3, false>, grid, 256, 0, s, p, d_ohow, d_ow, x, wp, bias, y);
    else if (hl) launch_k(conv_c1_direct_kernel<5, 5, true>, grid, 256, 0, s, p, d_ohow, d_ow, x, wp, bias, y);
    else launch_k(conv_c1_direct_kernel<5, 5, false>, grid, 256, 0, s, p, d_ohow, d_ow, x, wp, bias, y);
    B200OV_LAUNCH_CHECK("conv_c1_direct_kernel");
    return B200OV_OK;
  }
  if (d->y_dtype != B200OV_DT_F32) return set_error(B200OV_ERR_UNSUPPORTED, "conv2d: the FP32 FMA kernels write float32 feature maps only");
  const bool vec = (d->cin % 4 == 0) && (d->x_ld % 4 == 0) && aligned16(x);
  const int sms = props().sm_count;
  if (d->cout <= 32) return launch_cfg<128, 32, 4, 4>(p, vec, x, wp, bias, y, s);
  long long blocks128 = (long long)ceil_div(p.M, 128) * ceil_div(d->cout, 64);
  if (blocks128 < 2LL * sms) return launch_cfg<64, 64, 4, 4>(p, vec, x, wp, bias, y, s);
  return launch_cfg<128, 64, 8, 4>(p, vec, x, wp, bias, y, s);
}

}  // namespace b200ov

namespace b200ov {
void tf32_weight_dims(int cout, int cin, int kh, int kw, int* coutp, long long* kpad);
int pack_tf32_weights(const float* w_oihw, float* out, int cout, int cin, int kh, int kw, cudaStream_t s);
bool has_tf32_section(int cin) { return cin % 4 == 0 && cin >= 8; }
long long f16_section_floats(int cout, int cin, int kh, int kw);
int pack_f16_weights(const float* w_oihw, float* out, int cout, int cin, int kh, int kw, cudaStream_t s);
}  // namespace b200ov

using namespace b200ov;

extern "C" {

int b200ov_conv_weight_dims(int cout, int cin, int kh, int kw, int* rows, int* ldw, int64_t* total_floats) {
  B200OV_REQUIRE(cout > 0 && cin > 0 && kh > 0 && kw > 0, "bad filter dims");
  const int r = round_up(kh * kw * cin, 16), l = round_up(cout, 64);
  if (rows) *rows = r;
  if (ldw) *ldw = l;
  if (total_floats) {
    long long total = (long long)r * l;
    if (has_tf32_section(cin)) {
      int coutp;
      long long kpad;
      tf32_weight_dims(cout, cin, kh, kw, &coutp, &kpad);
      total += 2LL * coutp * kpad;
    }
    total += f16_section_floats(cout, cin, kh, kw);
    *total_floats = total;
  }
  return B200OV_OK;
}

int b200ov_pack_conv_weights(const float* w_oihw, float* w_packed, int cout, int cin, int kh, int kw, void* stream) {
  B200OV_REQUIRE(w_oihw && w_packed, "null weight pointer");
  int rows, ldw;
  int rc = b200ov_conv_weight_dims(cout, cin, kh, kw, &rows, &ldw, nullptr);
  if (rc) return rc;
  long long total = (long long)rows * ldw;
  pack_conv_weights_kernel<<<bw_grid(total, 256), 256, 0, as_stream(stream)>>>(w_oihw, w_packed, cout, cin, kh, kw,
                                                                              kh * kw * cin, rows, ldw);
  B200OV_LAUNCH_CHECK("pack_conv_weights_kernel");
  if (has_tf32_section(cin)) {
    rc = pack_tf32_weights(w_oihw, w_packed + total, cout, cin, kh, kw, as_stream(stream));
    if (rc) return rc;
    int coutp;
    long long kpad;
    tf32_weight_dims(cout, cin, kh, kw, &coutp, &kpad);
    total += 2LL * coutp * kpad;
  }
  return pack_f16_weights(w_oihw, w_packed + total, cout, cin, kh, kw, as_stream(stream));
}

}  // extern "C"
