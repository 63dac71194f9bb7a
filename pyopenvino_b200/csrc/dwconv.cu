// Depthwise GroupConvolution (NHWC) with fused bias + activation.
//
// Reference: GroupConvolution.py:53-79, `out[g,oy,ox] = np.sum(in_pad[g, window] * w[g,0,0])`.
// Bandwidth-bound (0.9-2.2 FLOP/B).  The 3x3 hot case uses a register-blocked strip kernel (below): the naive
// one-thread-per-output form issues 9 loads per output and is limited by the LSU / L1 wavefront rate, not HBM.
// Index arithmetic is 32-bit with multiply-high division (fastdiv.cuh).
// The generic kernel rounds the kh*kw products individually (no FMA contraction) and sums them in the order
// numpy's pairwise float32 reduction uses, so its pre-bias value is bit-identical to the reference.
#include <stdlib.h>

#include "common.cuh"
#include "f16split.cuh"
#include "fastdiv.cuh"
#include "tc_ptx.cuh"
#include "tma_util.cuh"
#include "vec4io.cuh"

namespace b200ov {

struct DwP {
  int n, h, w, c, kh, kw, sh, sw, pt, pl, oh, ow, x_ld, y_ld, act;
  float lo, hi;
  int owt;                 // output column tiles per row = ceil(ow / TW)
  uint32_t total;          // work items
  FastDiv d_cg, d_owt, d_oh;
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

// ---- 3x3 hot path ------------------------------------------------------------------------------------
struct DwStripP {
  int h, w, c, pt, pl, oh, ow, x_ld, y_ld, act, th;
  float lo, hi;
  uint32_t total;                // work items = n * strips * ceil(ow / 2) * (c / 4)
  FastDiv d_cg, d_owt, d_strips;
};

// Packed FP32 pairs (sm_100 FFMA2): two fused multiply-adds per instruction.  This fast path accumulates
// the nine taps as an FMA chain starting from the bias, so it agrees with the reference's
// np.sum(patch * flt) + bias to a few ulp (tolerance class 1e-5 + 1e-4|ref|), not bit for bit; ptxas
// contracts mul.rn.f32x2 + add.rn.f32x2 pairs anyway, so an "unfused" packed form cannot be expressed.
// `math == B200OV_DW_EXACT` selects the pairwise kernel below, which is bit-identical to numpy.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ float2 unpack2(f32x2 v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}

// S = stride (both directions).  A thread owns 4 channels x 2 adjacent output columns x a vertical strip of
// `th` output rows and keeps the 3 x (S + 3) input window in registers: moving down one output row loads
// S new input rows, i.e. (S + 3) * S / 2 128-bit loads per output instead of 9.  The row buffers rotate
// by name (the loop is unrolled over one rotation period), and for S = 1 a fourth buffer receives the next
// input row before the current output row is computed, so a load is always in flight behind the arithmetic.
__device__ __align__(16) float g_dw_zeros[4];   // source of the np.pad zeros: padded taps load from here (no branches)

template <int ACT>
__device__ __forceinline__ float act_t(float v, float lo, float hi) {
  if constexpr (ACT == B200OV_ACT_RELU) return v < 0.f ? 0.f : v;
  else if constexpr (ACT == B200OV_ACT_CLAMP) return fminf(fmaxf(v, lo), hi);
  else return v;
}

template <int S, int ACT>
__global__ void __launch_bounds__(128, 4) dwconv3x3_strip_kernel(DwStripP p, const float* __restrict__ x,
                                                              const float* __restrict__ wp,
                                                              const float* __restrict__ bias, float* __restrict__ y) {
  constexpr int NC = S + 3;
  constexpr int NB = S == 1 ? 4 : 3;     // row buffers
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < p.total; idx += stride) {
    uint32_t q, g, q2, oxt, img, strip;
    p.d_cg.divmod(idx, q, g);
    p.d_owt.divmod(q, q2, oxt);
    p.d_strips.divmod(q2, img, strip);
    const int c0 = (int)g * 4;
    const int ox0 = (int)oxt * 2;
    const int oy0 = (int)strip * p.th, oy1 = min(p.oh, oy0 + p.th);
    const int ix0 = ox0 * S - p.pl;
    const float* ximg = x + (size_t)img * p.h * p.w * p.x_ld + c0;
    float* yp = y + ((size_t)(img * p.oh + oy0) * p.ow + ox0) * p.y_ld + c0;
    ulonglong2 wt[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) wt[t] = __ldg(reinterpret_cast<const ulonglong2*>(wp + t * p.c + c0));
    ulonglong2 bv = make_ulonglong2(0ull, 0ull);
    if (bias != nullptr) bv = __ldg(reinterpret_cast<const ulonglong2*>(bias + c0));
    const bool second = ox0 + 1 < p.ow;
    int col_off[NC];                     // element offset of each window column inside an input row, -1 = padding
#pragma unroll
    for (int cidx = 0; cidx < NC; ++cidx) {
      const bool ok = ix0 + cidx >= 0 && ix0 + cidx < p.w && (cidx < 3 || second);   // columns >= 3 only feed the 2nd output
      col_off[cidx] = ok ? (ix0 + cidx) * p.x_ld : -1;
    }
    const ulonglong2* zeros = reinterpret_cast<const ulonglong2*>(g_dw_zeros);
    ulonglong2 R[NB][NC];
    // interior strips (no padded tap, both output columns exist) skip every bounds test and pointer select
    const int iy_first = oy0 * S - p.pt, iy_last = (oy1 - 1) * S - p.pt + 2;
    const bool interior = second && iy_first >= 0 && iy_last < p.h && ix0 >= 0 && ix0 + NC <= p.w;
    const float* xcol0 = ximg + (long long)ix0 * p.x_ld;          // only dereferenced for valid taps
    const int row_pitch = p.w * p.x_ld;
    auto load_row = [&](int iy, ulonglong2 (&dst)[NC]) {
      if (interior) {
        const float* xr = xcol0 + (long long)iy * row_pitch;
#pragma unroll
        for (int cidx = 0; cidx < NC; ++cidx) dst[cidx] = __ldg(reinterpret_cast<const ulonglong2*>(xr + cidx * p.x_ld));
      } else {
        const bool row_ok = iy >= 0 && iy < p.h;
        const float* xr = ximg + (long long)(row_ok ? iy : 0) * row_pitch;
#pragma unroll
        for (int cidx = 0; cidx < NC; ++cidx) {
          const ulonglong2* src = (row_ok && col_off[cidx] >= 0) ? reinterpret_cast<const ulonglong2*>(xr + col_off[cidx]) : zeros;
          dst[cidx] = __ldg(src);
        }
      }
    };
    // one output row (2 pixels x 4 channels) from the three buffered input rows
    auto compute = [&](const ulonglong2 (&r0)[NC], const ulonglong2 (&r1)[NC], const ulonglong2 (&r2)[NC], float* yrow) {
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        f32x2 lo = bv.x, hi = bv.y;        // channels (0,1) and (2,3)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) { lo = fma2(r0[t * S + kx].x, wt[kx].x, lo); hi = fma2(r0[t * S + kx].y, wt[kx].y, hi); }
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) { lo = fma2(r1[t * S + kx].x, wt[3 + kx].x, lo); hi = fma2(r1[t * S + kx].y, wt[3 + kx].y, hi); }
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) { lo = fma2(r2[t * S + kx].x, wt[6 + kx].x, lo); hi = fma2(r2[t * S + kx].y, wt[6 + kx].y, hi); }
        const float2 a = unpack2(lo), b = unpack2(hi);
        const float4 o = make_float4(act_t<ACT>(a.x, p.lo, p.hi), act_t<ACT>(a.y, p.lo, p.hi), act_t<ACT>(b.x, p.lo, p.hi),
                                     act_t<ACT>(b.y, p.lo, p.hi));
        if (t == 0 || second) *reinterpret_cast<float4*>(yrow + t * p.y_ld) = o;
      }
    };
    const size_t yrow_stride = (size_t)p.ow * p.y_ld;
    const int iy_base = oy0 * S - p.pt;
    if constexpr (S == 1) {
      load_row(iy_base, R[0]);
      load_row(iy_base + 1, R[1]);
      load_row(iy_base + 2, R[2]);
      for (int oy = oy0; oy < oy1; oy += 4) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (oy + k < oy1) {
            if (oy + k + 1 < oy1) load_row(iy_base + (oy + k + 1 - oy0) + 2, R[(k + 3) & 3]);   // next bottom row
            compute(R[k & 3], R[(k + 1) & 3], R[(k + 2) & 3], yp);
            yp += yrow_stride;
          }
        }
      }
    } else {
      load_row(iy_base, R[0]);
      for (int oy = oy0; oy < oy1; oy += 3) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          if (oy + k < oy1) {
            const int iy = iy_base + (oy + k - oy0) * 2;
            load_row(iy + 1, R[(2 * k + 1) % 3]);
            load_row(iy + 2, R[(2 * k + 2) % 3]);
            compute(R[(2 * k) % 3], R[(2 * k + 1) % 3], R[(2 * k + 2) % 3], yp);
            yp += yrow_stride;
          }
        }
      }
    }
  }
}

// ---- 3x3, TMA-staged tiles (the hot path) -------------------------------------------------------------------------------
// Same tiling as the TMA MaxPool kernel (pool.cu): a work item is NIMG images x TR output rows x TW output columns
// (NIMG * TW <= 32) x 32 channels; its input window, halo included, arrives by ONE 4-D bulk tensor copy into a 3-stage
// shared-memory ring, out-of-bounds elements zero-filled by TMA = the reference's zero padding (GroupConvolution.py:62-63).
// A consumer thread owns one output column x 4 channels and walks down the rows with the 3x3 input window in registers
// (S new input rows per output row: 3*S 128-bit LDS instead of 9), nine packed FFMA2 pairs from the bias like the strip
// kernel (identical arithmetic and summation order, so both give the same bits).
struct DwTmaP {
  int n, c, oh, ow, y_ld;
  int pt, pl;
  int tw, tr, nimg, bw, bh;
  int stage_bytes, box_bytes;
  float lo, hi;
  int y_hl;                            // 1: y is written as FP16 (hi, lo) pairs, 16 bytes per 4 channels (B200OV_DT_HL; T = float only)
  uint32_t items;
  FastDiv d_cchunks, d_coltiles, d_rowtiles, d_tw;
};
constexpr int DW_TMA_CONSUMERS = 512;                       // 16 warps, all consumers; thread 0 also feeds the ring (a 17th warp
constexpr int DW_TMA_THREADS = DW_TMA_CONSUMERS;            // would cost every thread registers: 5 warps per SM sub-partition)
constexpr int DW_TMA_STAGES = 4;
constexpr int DW_TMA_STAGE_BYTES = 56 * 1024;       // per stage; 4 stages + barriers < 227 KB

__device__ __forceinline__ f32x2 pack2(float a, float b) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}

template <int S, int ACT, typename T>
__global__ void __launch_bounds__(DW_TMA_THREADS, 1) dwconv3x3_tma_kernel(const DwTmaP p, const __grid_constant__ CUtensorMap map_x,
                                                                         const float* __restrict__ wp, const float* __restrict__ bias,
                                                                         T* __restrict__ y) {
  using IO = Vec4IO<T>;
  using namespace ptx;
  extern __shared__ uint8_t dw_smem_raw[];
  const uint32_t base = (smem_u32(dw_smem_raw) + 127u) & ~127u;
  const uint8_t* base_ptr = dw_smem_raw + (base - smem_u32(dw_smem_raw));
  const uint32_t bars = base + DW_TMA_STAGES * p.stage_bytes;            // full[STAGES], empty[STAGES]
  auto bar_full = [&](uint32_t s) { return bars + 8 * s; };
  auto bar_empty = [&](uint32_t s) { return bars + 8 * (DW_TMA_STAGES + s); };
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < DW_TMA_STAGES; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), DW_TMA_CONSUMERS / 32);
    }
    fence_mbar_init();
    prefetch_tensormap(&map_x);
  }
  __syncthreads();
  B200OV_PDL_SYNC();                     // the prologue above may overlap the previous kernel's tail
  const uint32_t my_items = p.items > blockIdx.x ? (p.items - 1 - blockIdx.x) / gridDim.x + 1 : 0;
  auto decode = [&](uint32_t k, int& cc, int& ct, int& rt, int& ig) {
    const uint32_t item = blockIdx.x + k * gridDim.x;
    uint32_t q, q2, a, b, c3;
    p.d_cchunks.divmod(item, q, a);
    p.d_coltiles.divmod(q, q2, b);
    p.d_rowtiles.divmod(q2, c3, q);
    cc = (int)a; ct = (int)b; rt = (int)q; ig = (int)c3;
  };
  // Thread 0 feeds the ring: item j may be issued once all 16 warps have drained the stage's previous item (empty
  // barrier).  Before it waits for item k itself it makes sure k has been issued (blocking on the laggards if it must),
  // then issues ahead opportunistically (non-blocking probes) so that up to STAGES - 1 tiles are in flight.
  uint32_t issued = 0;
  auto issue = [&](uint32_t j) {
    int cc, ct, rt, ig;
    decode(j, cc, ct, rt, ig);
    const uint32_t s = j % DW_TMA_STAGES;
    mbar_arrive_expect_tx(bar_full(s), (uint32_t)p.box_bytes);
    tma::load_4d(base + s * p.stage_bytes, &map_x, cc * 32, ct * p.tw * S - p.pl, rt * p.tr * S - p.pt, ig * p.nimg, bar_full(s));
  };
  auto feed = [&](uint32_t k) {
    while (issued < my_items) {
      const uint32_t s = issued % DW_TMA_STAGES;
      if (issued >= DW_TMA_STAGES) {
        const uint32_t par = ((issued / DW_TMA_STAGES) - 1) & 1;
        if (issued <= k) mbar_wait(bar_empty(s), par);
        else if (!mbar_test(bar_empty(s), par)) break;
      }
      issue(issued++);
    }
  };

  const int half = tid >> 8, t = tid & 255;
  const int cg = t & 7, lane_col = t >> 3;
  uint32_t img_l, ox_l;
  p.d_tw.divmod((uint32_t)lane_col, img_l, ox_l);
  const bool lane_ok = (int)img_l < p.nimg;
  const int rows_half = (p.tr + 1) >> 1;
  const int r_begin = half * rows_half, r_end = min(p.tr, r_begin + rows_half);
  int cur_c0 = -1;
  ulonglong2 wt[9], bv = make_ulonglong2(0ull, 0ull);
  constexpr int KEEP = 3 - S;                               // input rows two vertically adjacent windows share

  for (uint32_t k = 0; k < my_items; ++k) {
    int cc, ct, rt, ig;
    decode(k, cc, ct, rt, ig);
    const uint32_t s = k % DW_TMA_STAGES;
    const int c0 = cc * 32 + cg * 4;
    const int img = ig * p.nimg + (int)img_l;
    const int ox = ct * p.tw + (int)ox_l;
    const bool active = lane_ok && img < p.n && ox < p.ow && c0 < p.c;
    if (active && c0 != cur_c0) {                          // weights of this channel quad (while the tile is still in flight)
      cur_c0 = c0;
#pragma unroll
      for (int tp = 0; tp < 9; ++tp) wt[tp] = __ldg(reinterpret_cast<const ulonglong2*>(wp + tp * p.c + c0));
      bv = bias != nullptr ? __ldg(reinterpret_cast<const ulonglong2*>(bias + c0)) : make_ulonglong2(0ull, 0ull);
    }
    if (tid == 0) feed(k);
    mbar_wait(bar_full(s), (k / DW_TMA_STAGES) & 1);
    if (active) {
      const T* tile = reinterpret_cast<const T*>(base_ptr + s * p.stage_bytes) +
                      ((size_t)img_l * p.bh * p.bw + (size_t)ox_l * S) * 32 + cg * 4;
      auto load_row = [&](int lr, ulonglong2 (&dst)[3]) {
        const T* rp = tile + (size_t)lr * p.bw * 32;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          if constexpr (sizeof(T) == 4) {
            dst[kx] = *reinterpret_cast<const ulonglong2*>(rp + kx * 32);
          } else {
            const float4 v = IO::ld(rp + kx * 32);
            dst[kx] = make_ulonglong2(pack2(v.x, v.y), pack2(v.z, v.w));
          }
        }
      };
      ulonglong2 R[3][3];                                   // three input rows x three columns, rotating BY NAME:
#pragma unroll                                              // the loop below is unrolled over one rotation period
      for (int i = 0; i < KEEP; ++i) load_row(r_begin * S + i, R[i]);
      const int oy0 = rt * p.tr;
      T* yp = y + (((size_t)img * p.oh + oy0 + r_begin) * p.ow + ox) * p.y_ld + c0;
      const size_t yrow = (size_t)p.ow * p.y_ld;
      const int r_stop = min(r_end, p.oh - oy0);
      for (int r = r_begin; r < r_stop; r += 3) {
#pragma unroll
        for (int ph = 0; ph < 3; ++ph) {
          if (r + ph < r_stop) {
            const int b0 = (ph * S) % 3;                    // compile-time after unrolling: window row ky lives in R[(b0 + ky) % 3]
#pragma unroll
            for (int i = KEEP; i < 3; ++i) load_row((r + ph) * S + i, R[(b0 + i) % 3]);
            f32x2 lo = bv.x, hi = bv.y;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) {
                lo = fma2(R[(b0 + ky) % 3][kx].x, wt[ky * 3 + kx].x, lo);
                hi = fma2(R[(b0 + ky) % 3][kx].y, wt[ky * 3 + kx].y, hi);
              }
            const float2 a = unpack2(lo), b = unpack2(hi);
            float4 o = make_float4(act_t<ACT>(a.x, p.lo, p.hi), act_t<ACT>(a.y, p.lo, p.hi), act_t<ACT>(b.x, p.lo, p.hi),
                                   act_t<ACT>(b.y, p.lo, p.hi));
            if constexpr (sizeof(T) == 4) {
              if (p.y_hl) o = encode_hl4(o.x, o.y, o.z, o.w);      // the pointwise contraction's operand form (same 16 bytes)
            }
            IO::st(yp, o);
            yp += yrow;
          }
        }
      }
    }
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(bar_empty(s));         // this warp is done with stage s
  }
}

static bool dw_tma_plan(const b200ov_dwconv_desc* d, DwTmaP& q, int esize) {
  tma::TilePlan t;
  if (!tma::plan_tiles(d->n, d->oh, d->ow, 3, d->sh, 32 * esize, DW_TMA_STAGE_BYTES, t)) return false;
  q.tw = t.tw; q.tr = t.tr; q.nimg = t.nimg; q.bw = t.bw; q.bh = t.bh;
  q.box_bytes = t.box_bytes;
  q.stage_bytes = round_up(t.box_bytes, 128);
  const int cchunks = ceil_div(d->c, 32), img_groups = ceil_div(d->n, q.nimg);
  const long long items = (long long)img_groups * t.row_tiles * t.col_tiles * cchunks;
  if (items <= 0 || items > 0x7fffffffLL) return false;
  q.items = (uint32_t)items;
  q.n = d->n; q.c = d->c; q.oh = d->oh; q.ow = d->ow; q.y_ld = d->y_ld; q.pt = d->pt; q.pl = d->pl;
  q.lo = d->act_lo; q.hi = d->act_hi;
  q.y_hl = d->y_dtype == B200OV_DT_HL ? 1 : 0;
  q.d_cchunks = FastDiv(cchunks); q.d_coltiles = FastDiv(t.col_tiles); q.d_rowtiles = FastDiv(t.row_tiles); q.d_tw = FastDiv(q.tw);
  return true;
}

template <int S, int ACT, typename T>
static int launch_dw_tma(const DwTmaP& q, const CUtensorMap& map, const float* wp, const float* bias, T* y, cudaStream_t s) {
  auto kern = dwconv3x3_tma_kernel<S, ACT, T>;
  static bool configured = false;
  if (!configured) {
    B200OV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, DW_TMA_STAGES * DW_TMA_STAGE_BYTES + 16 * DW_TMA_STAGES + 256));
    configured = true;
  }
  const int need = DW_TMA_STAGES * q.stage_bytes + 16 * DW_TMA_STAGES + 256;
  const int grid = (int)(q.items < (uint32_t)props().sm_count ? q.items : (uint32_t)props().sm_count);
  launch_k(kern, grid, DW_TMA_THREADS, need, s, q, map, wp, bias, y);
  B200OV_LAUNCH_CHECK("dwconv3x3_tma_kernel");
  return B200OV_OK;
}

// ---- generic window (runtime kh x kw), one output per thread --------------------------------------------
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
  __device__ void combine() {
#pragma unroll
    for (int j = 0; j < V; ++j)
      tail[j] = __fadd_rn(__fadd_rn(__fadd_rn(r[0][j], r[1][j]), __fadd_rn(r[2][j], r[3][j])),
                          __fadd_rn(__fadd_rn(r[4][j], r[5][j]), __fadd_rn(r[6][j], r[7][j])));
  }
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
      if (count == full) combine();
#pragma unroll
      for (int j = 0; j < V; ++j) tail[j] = __fadd_rn(tail[j], t[j]);
    }
    ++count;
  }
  __device__ void finish(float (&out)[V]) {
    const int full = (kk >= 8) ? (kk / 8) * 8 : 0;
    if (kk >= 8 && count == full) combine();
#pragma unroll
    for (int j = 0; j < V; ++j) out[j] = tail[j];
  }
};

template <int V>
__global__ void __launch_bounds__(256) dwconv_generic_kernel(DwP p, const float* __restrict__ x, const float* __restrict__ wp,
                                                             const float* __restrict__ bias, float* __restrict__ y) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < p.total; idx += stride) {
    uint32_t q, g;
    p.d_cg.divmod(idx, q, g);
    uint32_t q2, ox;
    p.d_owt.divmod(q, q2, ox);                             // TW = 1 here: owt == ow
    uint32_t img, oy;
    p.d_oh.divmod(q2, img, oy);
    const int c0 = (int)g * V;
    const float* ximg = x + (size_t)img * p.h * p.w * p.x_ld + c0;
    const int iy0 = (int)oy * p.sh - p.pt, ix0 = (int)ox * p.sw - p.pl;
    PairwiseAcc<V> acc(p.kh * p.kw);
    for (int ky = 0; ky < p.kh; ++ky)
      for (int kx = 0; kx < p.kw; ++kx) {
        const int iy = iy0 + ky, ix = ix0 + kx;
        Vec<V> wv = Vec<V>::load(wp + (ky * p.kw + kx) * p.c + c0);
        float term[V];
        if (iy >= 0 && iy < p.h && ix >= 0 && ix < p.w) {
          Vec<V> xv = Vec<V>::load(ximg + ((size_t)iy * p.w + ix) * p.x_ld);
#pragma unroll
          for (int j = 0; j < V; ++j) term[j] = __fmul_rn(xv.v[j], wv.v[j]);
        } else {
#pragma unroll
          for (int j = 0; j < V; ++j) term[j] = __fmul_rn(0.f, wv.v[j]);
        }
        acc.push(term);
      }
    float res[V];
    acc.finish(res);
    Vec<V> out;
    if (bias != nullptr) {
      Vec<V> bv = Vec<V>::load(bias + c0);
#pragma unroll
      for (int j = 0; j < V; ++j) out.v[j] = apply_act(__fadd_rn(res[j], bv.v[j]), p.act, p.lo, p.hi);
    } else {
#pragma unroll
      for (int j = 0; j < V; ++j) out.v[j] = apply_act(res[j], p.act, p.lo, p.hi);
    }
    out.store(y + ((size_t)(img * p.oh + oy) * p.ow + ox) * p.y_ld + c0);
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

int b200ov_dwconv2d(const b200ov_dwconv_desc* d, const void* x_raw, const float* w_packed, const float* bias, void* y_raw,
                    void* stream) {
  const float* x = static_cast<const float*>(x_raw);
  float* y = static_cast<float*>(y_raw);
  B200OV_REQUIRE(d && x && w_packed && y, "dwconv2d: null argument");
  B200OV_REQUIRE(d->n >= 0 && d->h > 0 && d->w > 0 && d->c > 0 && d->kh > 0 && d->kw > 0 && d->sh > 0 && d->sw > 0 &&
                     d->oh > 0 && d->ow > 0 && d->pt >= 0 && d->pl >= 0,
                 "dwconv2d: bad geometry");
  B200OV_REQUIRE(d->x_ld >= d->c && d->y_ld >= d->c, "dwconv2d: channel pitch smaller than channel count");
  B200OV_REQUIRE(d->kh * d->kw <= 128, "dwconv2d: window larger than 128 taps is not supported");
  B200OV_REQUIRE(d->act >= B200OV_ACT_NONE && d->act <= B200OV_ACT_SIGMOID, "dwconv2d: bad activation");
  B200OV_REQUIRE(d->math == B200OV_DW_AUTO || d->math == B200OV_DW_EXACT, "dwconv2d: bad math mode");
  if (d->n == 0) return B200OV_OK;
  cudaStream_t s = as_stream(stream);
  const bool vec = (d->c % 4 == 0) && (d->x_ld % 4 == 0) && (d->y_ld % 4 == 0) && aligned16(x) && aligned16(y) &&
                   aligned16(w_packed) && (bias == nullptr || aligned16(bias));
  const int V = vec ? 4 : 1;
  const int cg = d->c / V;
  B200OV_REQUIRE(d->dtype == B200OV_DT_F32 || d->dtype == B200OV_DT_F16, "dwconv2d: bad storage type");
  B200OV_REQUIRE(d->y_dtype == 0 || d->y_dtype == d->dtype || (d->y_dtype == B200OV_DT_HL && d->dtype == B200OV_DT_F32),
                 "dwconv2d: bad output storage type");
  const bool want_hl = d->y_dtype == B200OV_DT_HL;
  if (d->dtype == B200OV_DT_F16) {
    // FP16 feature maps: the TMA tile kernel only (3x3, stride 1 / 2, packed-FMA arithmetic)
    const bool ok = d->c % 4 == 0 && d->x_ld % 8 == 0 && d->y_ld % 4 == 0 && aligned16(x_raw) && aligned_vec4<__half>(y_raw) &&
                    aligned16(w_packed) && (bias == nullptr || aligned16(bias)) && d->math != B200OV_DW_EXACT && d->kh == 3 &&
                    d->kw == 3 && d->sh == d->sw && (d->sh == 1 || d->sh == 2) && d->act <= B200OV_ACT_CLAMP;
    DwTmaP tq;
    CUtensorMap map;
    if (!ok || !dw_tma_plan(d, tq, 2) ||
        tma::make_map_nhwc(&map, x_raw, 2, d->n, d->h, d->w, d->c, d->x_ld, 32, tq.bw, tq.bh, tq.nimg) != B200OV_OK)
      return set_error(B200OV_ERR_UNSUPPORTED, "dwconv2d: this shape has no FP16-storage kernel");
    __half* yh = static_cast<__half*>(y_raw);
#define B200OV_DWH(S_, A_) return launch_dw_tma<S_, A_, __half>(tq, map, w_packed, bias, yh, s)
    if (d->sh == 1) {
      if (d->act == B200OV_ACT_NONE) B200OV_DWH(1, B200OV_ACT_NONE);
      if (d->act == B200OV_ACT_RELU) B200OV_DWH(1, B200OV_ACT_RELU);
      B200OV_DWH(1, B200OV_ACT_CLAMP);
    } else {
      if (d->act == B200OV_ACT_NONE) B200OV_DWH(2, B200OV_ACT_NONE);
      if (d->act == B200OV_ACT_RELU) B200OV_DWH(2, B200OV_ACT_RELU);
      B200OV_DWH(2, B200OV_ACT_CLAMP);
    }
#undef B200OV_DWH
  }
  const bool hot = vec && d->math != B200OV_DW_EXACT && d->kh == 3 && d->kw == 3 && d->sh == d->sw && (d->sh == 1 || d->sh == 2);
  static const bool no_tma = getenv("B200OV_DW_NO_TMA") != nullptr && atoi(getenv("B200OV_DW_NO_TMA")) != 0;     // developer knob (A/B)
  if (hot && !no_tma && d->act <= B200OV_ACT_CLAMP && (long long)d->n * d->oh * d->ow * d->c >= (1 << 16)) {
    DwTmaP tq;
    CUtensorMap map;
    if (dw_tma_plan(d, tq, 4) && tma::make_map_nhwc(&map, x, 4, d->n, d->h, d->w, d->c, d->x_ld, 32, tq.bw, tq.bh, tq.nimg) == B200OV_OK) {
#define B200OV_DWT(S_, A_) return launch_dw_tma<S_, A_, float>(tq, map, w_packed, bias, y, s)
      if (d->sh == 1) {
        if (d->act == B200OV_ACT_NONE) B200OV_DWT(1, B200OV_ACT_NONE);
        if (d->act == B200OV_ACT_RELU) B200OV_DWT(1, B200OV_ACT_RELU);
        B200OV_DWT(1, B200OV_ACT_CLAMP);
      } else {
        if (d->act == B200OV_ACT_NONE) B200OV_DWT(2, B200OV_ACT_NONE);
        if (d->act == B200OV_ACT_RELU) B200OV_DWT(2, B200OV_ACT_RELU);
        B200OV_DWT(2, B200OV_ACT_CLAMP);
      }
#undef B200OV_DWT
    }
  }
  // only the tile kernel above writes the (hi, lo) pair form: the caller falls back to an FP32 output
  if (want_hl) return set_error(B200OV_ERR_UNSUPPORTED, "dwconv2d: this shape has no kernel that writes (hi, lo) pairs");
  if (hot) {
    DwStripP q;
    q.h = d->h; q.w = d->w; q.c = d->c; q.pt = d->pt; q.pl = d->pl; q.oh = d->oh; q.ow = d->ow; q.x_ld = d->x_ld;
    q.y_ld = d->y_ld; q.act = d->act; q.lo = d->act_lo; q.hi = d->act_hi;
    q.th = d->oh < 8 ? d->oh : 8;
    const int strips = ceil_div(d->oh, q.th), owt = ceil_div(d->ow, 2);
    const long long items = (long long)d->n * strips * owt * cg;
    if (items < 0x7fffffffLL) {
      q.total = (uint32_t)items;
      q.d_cg = FastDiv(cg); q.d_owt = FastDiv(owt); q.d_strips = FastDiv(strips);
      const int g = bw_grid(items, 128, 16);
#define B200OV_DW_LAUNCH(S_, A_) dwconv3x3_strip_kernel<S_, A_><<<g, 128, 0, s>>>(q, x, w_packed, bias, y)
      if (d->act <= B200OV_ACT_CLAMP) {
        if (d->sh == 1) {
          if (d->act == B200OV_ACT_NONE) B200OV_DW_LAUNCH(1, B200OV_ACT_NONE);
          else if (d->act == B200OV_ACT_RELU) B200OV_DW_LAUNCH(1, B200OV_ACT_RELU);
          else B200OV_DW_LAUNCH(1, B200OV_ACT_CLAMP);
        } else {
          if (d->act == B200OV_ACT_NONE) B200OV_DW_LAUNCH(2, B200OV_ACT_NONE);
          else if (d->act == B200OV_ACT_RELU) B200OV_DW_LAUNCH(2, B200OV_ACT_RELU);
          else B200OV_DW_LAUNCH(2, B200OV_ACT_CLAMP);
        }
#undef B200OV_DW_LAUNCH
        B200OV_LAUNCH_CHECK("dwconv3x3_strip_kernel");
        return B200OV_OK;
      }
    }
  }
  DwP p;
  p.n = d->n; p.h = d->h; p.w = d->w; p.c = d->c; p.kh = d->kh; p.kw = d->kw; p.sh = d->sh; p.sw = d->sw; p.pt = d->pt;
  p.pl = d->pl; p.oh = d->oh; p.ow = d->ow; p.x_ld = d->x_ld; p.y_ld = d->y_ld; p.act = d->act; p.lo = d->act_lo; p.hi = d->act_hi;
  p.owt = d->ow;
  long long total = (long long)d->n * d->oh * p.owt * cg;
  B200OV_REQUIRE(total < 0x7fffffffLL, "dwconv2d: problem too large for 32-bit indexing");
  p.total = (uint32_t)total;
  p.d_cg = FastDiv(cg); p.d_owt = FastDiv(p.owt); p.d_oh = FastDiv(d->oh);
  const int grid = bw_grid(total, 256);
  if (vec) dwconv_generic_kernel<4><<<grid, 256, 0, s>>>(p, x, w_packed, bias, y);
  else dwconv_generic_kernel<1><<<grid, 256, 0, s>>>(p, x, w_packed, bias, y);
  B200OV_LAUNCH_CHECK("dwconv_kernel");
  return B200OV_OK;
}

}  // extern "C"
