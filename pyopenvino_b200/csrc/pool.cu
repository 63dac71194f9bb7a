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
#include <stdlib.h>

#include "common.cuh"
#include "fastdiv.cuh"
#include "tc_ptx.cuh"
#include "tma_util.cuh"
#include "vec4io.cuh"

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

// ---- MaxPool, TMA-staged tiles (the hot path) ----------------------------------------------------------------------------
// A work item is a tile of the OUTPUT: NIMG images x TR rows x TW columns (NIMG * TW <= 32) x 32 channels.  One elected
// thread fetches the input window of the tile -- ((TR-1)*S + K) x ((TW-1)*S + K) pixels x 32 channels per image, halo
// included -- with a single 4-D bulk tensor copy (cp.async.bulk.tensor) into a 3-stage shared-memory ring; the box
// origin may be negative / overhang the tensor and TMA fills those elements with zeros, which is precisely the
// reference's np.pad(..., 'constant') whose zeros take part in the max (MaxPool.py:53).  Positions beyond the PADDED
// tensor (ceil-mode overhang, clipped by min(h, ...) in MaxPool.py:69) are masked out by the consumers.
// 512 consumer threads = 2 row halves x 32 column lanes x 8 channel quads: a thread walks down one output column with
// rolling horizontal maxima (K - S input rows are shared by vertically adjacent windows), 128-bit LDS (a warp reads 4
// whole 128-byte pixel chunks: conflict-free) and one 128-bit global store per output.  Loads for the next two tiles
// are always in flight (~100-190 KB per SM), so HBM latency is covered without occupancy or registers.
struct PoolTmaP {
  int n, c, oh, ow, y_ld;
  int pt, pl, hp, wpad;
  int tw, tr, nimg, bw, bh;
  int stage_bytes, box_bytes;
  uint32_t items;
  FastDiv d_cchunks, d_coltiles, d_rowtiles, d_tw;
};

constexpr int POOL_TMA_CONSUMERS = 512;                     // 16 consumer warps
constexpr int POOL_TMA_THREADS = POOL_TMA_CONSUMERS + 32;   // + the TMA producer warp
constexpr int POOL_TMA_STAGES = 4;
constexpr int POOL_TMA_STAGE_BYTES = 56 * 1024;       // per stage; 4 stages + barriers < 227 KB

template <int K, int S, typename T>
__global__ void __launch_bounds__(POOL_TMA_THREADS, 1) pool_max_tma_kernel(const PoolTmaP p, const __grid_constant__ CUtensorMap map_x,
                                                                          const float* __restrict__ scale,
                                                                          const float* __restrict__ shift, T* __restrict__ y) {
  using IO = Vec4IO<T>;
  using namespace ptx;
  extern __shared__ uint8_t pool_smem_raw[];
  const uint32_t base = (smem_u32(pool_smem_raw) + 127u) & ~127u;
  const uint8_t* base_ptr = pool_smem_raw + (base - smem_u32(pool_smem_raw));
  const uint32_t bars = base + POOL_TMA_STAGES * p.stage_bytes;          // full[STAGES], empty[STAGES]
  auto bar_full = [&](uint32_t s) { return bars + 8 * s; };
  auto bar_empty = [&](uint32_t s) { return bars + 8 * (POOL_TMA_STAGES + s); };
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < POOL_TMA_STAGES; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), POOL_TMA_CONSUMERS / 32);
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
  if (tid >= POOL_TMA_CONSUMERS) {
    // ---- producer warp: one lane keeps the ring full; it only ever waits for a stage to be drained by all 16 consumer warps
    if (tid == POOL_TMA_CONSUMERS) {
      for (uint32_t k = 0; k < my_items; ++k) {
        const uint32_t s = k % POOL_TMA_STAGES;
        if (k >= POOL_TMA_STAGES) mbar_wait(bar_empty(s), ((k / POOL_TMA_STAGES) - 1) & 1);
        int cc, ct, rt, ig;
        decode(k, cc, ct, rt, ig);
        mbar_arrive_expect_tx(bar_full(s), (uint32_t)p.box_bytes);
        tma::load_4d(base + s * p.stage_bytes, &map_x, cc * 32, ct * p.tw * S - p.pl, rt * p.tr * S - p.pt, ig * p.nimg, bar_full(s));
      }
    }
    return;
  }

  // fixed per-thread geometry
  const int half = tid >> 8, t = tid & 255;
  const int cg = t & 7, lane_col = t >> 3;
  uint32_t img_l, ox_l;
  p.d_tw.divmod((uint32_t)lane_col, img_l, ox_l);
  const bool lane_ok = (int)img_l < p.nimg;
  const int rows_half = (p.tr + 1) >> 1;
  const int r_begin = half * rows_half, r_end = min(p.tr, r_begin + rows_half);
  constexpr int KEEP = K > S ? K - S : 0;
  const float4 ninf = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);

  for (uint32_t k = 0; k < my_items; ++k) {
    int cc, ct, rt, ig;
    decode(k, cc, ct, rt, ig);
    const uint32_t s = k % POOL_TMA_STAGES;
    mbar_wait(bar_full(s), (k / POOL_TMA_STAGES) & 1);
    const int c0 = cc * 32 + cg * 4;
    const int img = ig * p.nimg + (int)img_l;
    const int ox = ct * p.tw + (int)ox_l;
    if (lane_ok && img < p.n && ox < p.ow && c0 < p.c && (int)ox_l < p.tw) {
      const T* tile = reinterpret_cast<const T*>(base_ptr + s * p.stage_bytes) +
                      ((size_t)img_l * p.bh * p.bw + (size_t)ox_l * S) * 32 + cg * 4;
      bool col_ok[K];
#pragma unroll
      for (int kx = 0; kx < K; ++kx) col_ok[kx] = ox * S + kx < p.wpad;
      float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sf = make_float4(0.f, 0.f, 0.f, 0.f);
      if (scale != nullptr) sc = __ldg(reinterpret_cast<const float4*>(scale + c0));
      if (shift != nullptr) sf = __ldg(reinterpret_cast<const float4*>(shift + c0));
      const int oy0 = rt * p.tr;
      // horizontal maximum of tile row `lr` (padded row oy0*S + lr); -inf when the row lies beyond the padded tensor
      auto hmax = [&](int lr) -> float4 {
        if (oy0 * S + lr >= p.hp) return ninf;
        const T* rp = tile + (size_t)lr * p.bw * 32;
        float4 m = ninf;
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
          const float4 v = IO::ld(rp + kx * 32);
          if (col_ok[kx]) m = max4(m, v);
        }
        return m;
      };
      float4 keep[KEEP > 0 ? KEEP : 1];
#pragma unroll
      for (int i = 0; i < KEEP; ++i) keep[i] = hmax(r_begin * S + i);
      T* yp = y + (((size_t)img * p.oh + oy0 + r_begin) * p.ow + ox) * p.y_ld + c0;
      const size_t yrow = (size_t)p.ow * p.y_ld;
      for (int r = r_begin; r < r_end && oy0 + r < p.oh; ++r) {
        float4 o = ninf;
#pragma unroll
        for (int i = 0; i < KEEP; ++i) o = max4(o, keep[i]);
        float4 fresh[K - KEEP];
#pragma unroll
        for (int i = 0; i < K - KEEP; ++i) {
          fresh[i] = hmax(r * S + KEEP + i);
          o = max4(o, fresh[i]);
        }
        // rows shared with the next window: the last KEEP of this window's K rows
#pragma unroll
        for (int i = 0; i < KEEP; ++i) {
          const int src = S + i;                        // index into this window's rows 0..K-1
          keep[i] = src < KEEP ? keep[src] : fresh[src - KEEP];
        }
        if (scale != nullptr) { o.x = __fmul_rn(o.x, sc.x); o.y = __fmul_rn(o.y, sc.y); o.z = __fmul_rn(o.z, sc.z); o.w = __fmul_rn(o.w, sc.w); }
        if (shift != nullptr) { o.x = __fadd_rn(o.x, sf.x); o.y = __fadd_rn(o.y, sf.y); o.z = __fadd_rn(o.z, sf.z); o.w = __fadd_rn(o.w, sf.w); }
        IO::st(yp, o);
        yp += yrow;
      }
    }
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(bar_empty(s));       // this warp is done with stage s
  }
}

// Tile geometry for the TMA kernel; false when the shape does not fit (the strip / generic kernels take over).
static bool pool_tma_plan(const b200ov_pool_desc* d, PoolTmaP& q, int esize) {
  tma::TilePlan t;
  if (!tma::plan_tiles(d->n, d->oh, d->ow, d->kh, d->sh, 32 * esize, POOL_TMA_STAGE_BYTES, t)) return false;
  q.tw = t.tw; q.tr = t.tr; q.nimg = t.nimg; q.bw = t.bw; q.bh = t.bh;
  q.box_bytes = t.box_bytes;
  q.stage_bytes = round_up(t.box_bytes, 128);
  const int cchunks = ceil_div(d->c, 32), img_groups = ceil_div(d->n, q.nimg);
  const long long items = (long long)img_groups * t.row_tiles * t.col_tiles * cchunks;
  if (items <= 0 || items > 0x7fffffffLL) return false;
  q.items = (uint32_t)items;
  q.n = d->n; q.c = d->c; q.oh = d->oh; q.ow = d->ow; q.y_ld = d->y_ld; q.pt = d->pt; q.pl = d->pl;
  q.hp = d->h + d->pt + d->pb; q.wpad = d->w + d->pl + d->pr;
  q.d_cchunks = FastDiv(cchunks); q.d_coltiles = FastDiv(t.col_tiles); q.d_rowtiles = FastDiv(t.row_tiles); q.d_tw = FastDiv(q.tw);
  return true;
}

template <int K, int S, typename T>
static int launch_pool_tma(const PoolTmaP& q, const CUtensorMap& map, const float* scale, const float* shift, T* y, cudaStream_t s) {
  auto kern = pool_max_tma_kernel<K, S, T>;
  static bool configured = false;
  const int smem = POOL_TMA_STAGES * POOL_TMA_STAGE_BYTES + 16 * POOL_TMA_STAGES + 256;
  if (!configured) {
    B200OV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  const int need = POOL_TMA_STAGES * q.stage_bytes + 16 * POOL_TMA_STAGES + 256;
  const int grid = (int)(q.items < (uint32_t)props().sm_count ? q.items : (uint32_t)props().sm_count);
  launch_k(kern, grid, POOL_TMA_THREADS, need, s, q, map, scale, shift, y);
  B200OV_LAUNCH_CHECK("pool_max_tma_kernel");
  return B200OV_OK;
}

template <int V, typename T>
__device__ __forceinline__ void loadv_t(const T* p, float (&v)[V]) {
  if constexpr (V == 4) {
    const float4 t = Vec4IO<T>::ldg(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
    v[0] = __ldg(reinterpret_cast<const float*>(p));
  }
}
template <int V, typename T>
__device__ __forceinline__ void storev_t(T* p, const float (&v)[V]) {
  if constexpr (V == 4) Vec4IO<T>::st(p, make_float4(v[0], v[1], v[2], v[3]));
  else *reinterpret_cast<float*>(p) = v[0];
}

template <int V, typename T = float>
__global__ void __launch_bounds__(256) pool_kernel(PoolP p, const T* __restrict__ x, const float* __restrict__ scale,
                                                   const float* __restrict__ shift, T* __restrict__ y) {
  B200OV_PDL_SYNC();
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
    const T* ximg = x + (long long)img * p.h * p.w * p.x_ld + c0;
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
            loadv_t<V, T>(ximg + ((long long)iy * p.w + ix) * p.x_ld, v);
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
          loadv_t<V, T>(ximg + ((long long)iy * p.w + ix) * p.x_ld, v);
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
    storev_t<V, T>(y + pix * p.y_ld + c0, res);
  }
}

}  // namespace b200ov

using namespace b200ov;

// FP16 feature maps (d->dtype == B200OV_DT_F16): the TMA tile kernel for the MaxPool hot cases, the generic kernel otherwise
static int pool2d_f16(const b200ov_pool_desc* d, const __half* x, const float* scale, const float* shift, __half* y, cudaStream_t s) {
  const bool vec = (d->c % 4 == 0) && (d->x_ld % 4 == 0) && (d->y_ld % 4 == 0) && aligned_vec4<__half>(x) && aligned_vec4<__half>(y) &&
                   (scale == nullptr || aligned16(scale)) && (shift == nullptr || aligned16(shift));
  if (!vec) return set_error(B200OV_ERR_UNSUPPORTED, "pool2d: FP16 feature maps need C %% 4 == 0 and 8-byte aligned pixels");
  const bool hot = d->mode == B200OV_POOL_MAX && d->kh == d->kw && (d->kh == 2 || d->kh == 3) && d->sh == d->sw && (d->sh == 1 || d->sh == 2);
  if (hot && d->x_ld % 8 == 0 && aligned16(x)) {
    PoolTmaP tq;
    CUtensorMap map;
    if (pool_tma_plan(d, tq, 2) && tma::make_map_nhwc(&map, x, 2, d->n, d->h, d->w, d->c, d->x_ld, 32, tq.bw, tq.bh, tq.nimg) == B200OV_OK) {
      if (d->kh == 3 && d->sh == 1) return launch_pool_tma<3, 1, __half>(tq, map, scale, shift, y, s);
      if (d->kh == 3) return launch_pool_tma<3, 2, __half>(tq, map, scale, shift, y, s);
      if (d->sh == 1) return launch_pool_tma<2, 1, __half>(tq, map, scale, shift, y, s);
      return launch_pool_tma<2, 2, __half>(tq, map, scale, shift, y, s);
    }
  }
  PoolP p{d->n, d->h, d->w, d->c, d->kh, d->kw, d->sh, d->sw, d->pt, d->pl, d->pb, d->pr, d->oh, d->ow, d->x_ld, d->y_ld, d->mode};
  const long long total = (long long)d->n * d->oh * d->ow * (d->c / 4);
  launch_k(pool_kernel<4, __half>, bw_grid(total, 256), 256, 0, s, p, x, scale, shift, y);
  B200OV_LAUNCH_CHECK("pool_kernel");
  return B200OV_OK;
}

extern "C" int b200ov_pool2d(const b200ov_pool_desc* d, const void* x_raw, const float* scale, const float* shift,
                             void* y_raw, void* stream) {
  const float* x = static_cast<const float*>(x_raw);
  float* y = static_cast<float*>(y_raw);
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
  B200OV_REQUIRE(d->dtype == B200OV_DT_F32 || d->dtype == B200OV_DT_F16, "pool2d: bad storage type");
  if (d->dtype == B200OV_DT_F16)
    return pool2d_f16(d, static_cast<const __half*>(x_raw), scale, shift, static_cast<__half*>(y_raw), as_stream(stream));
  const bool vec = (d->c % 4 == 0) && (d->x_ld % 4 == 0) && (d->y_ld % 4 == 0) && aligned16(x) && aligned16(y) &&
                   (scale == nullptr || aligned16(scale)) && (shift == nullptr || aligned16(shift));
  const bool hot = vec && d->mode == B200OV_POOL_MAX && d->kh == d->kw && (d->kh == 2 || d->kh == 3) && d->sh == d->sw && (d->sh == 1 || d->sh == 2);
  static const bool no_tma = getenv("B200OV_POOL_NO_TMA") != nullptr && atoi(getenv("B200OV_POOL_NO_TMA")) != 0;     // developer knob (A/B)
  if (hot && !no_tma && (long long)d->n * d->oh * d->ow * d->c >= (1 << 16)) {
    PoolTmaP tq;
    CUtensorMap map;
    if (pool_tma_plan(d, tq, 4) &&
        tma::make_map_nhwc(&map, x, 4, d->n, d->h, d->w, d->c, d->x_ld, 32, tq.bw, tq.bh, tq.nimg) == B200OV_OK) {
      cudaStream_t s = as_stream(stream);
      if (d->kh == 3 && d->sh == 1) return launch_pool_tma<3, 1, float>(tq, map, scale, shift, y, s);
      if (d->kh == 3) return launch_pool_tma<3, 2, float>(tq, map, scale, shift, y, s);
      if (d->sh == 1) return launch_pool_tma<2, 1, float>(tq, map, scale, shift, y, s);
      return launch_pool_tma<2, 2, float>(tq, map, scale, shift, y, s);
    }
  }
  if (hot) {
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
  if (vec) launch_k(pool_kernel<4, float>, grid, 256, 0, as_stream(stream), p, x, scale, shift, y);
  else launch_k(pool_kernel<1, float>, grid, 256, 0, as_stream(stream), p, x, scale, shift, y);
  B200OV_LAUNCH_CHECK("pool_kernel");
  return B200OV_OK;
}
