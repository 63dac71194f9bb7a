// Host-side tensor-map (TMA descriptor) construction and the N-D bulk tensor load, shared by the bandwidth kernels
// that stage NHWC tiles (with their halo) in shared memory.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "common.cuh"

namespace b200ov {
namespace tma {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// NHWC feature map [n][h][w][ld] (channels [0, c) of each pixel) as a 4-D tensor (C, W, H, N); box = (bc, bw, bh, bn).
// Out-of-bounds elements of a box are filled with zeros: exactly the np.pad(..., 'constant') the reference applies.
// `esize` = bytes per element (4: float32, 2: float16).
inline int make_map_nhwc(CUtensorMap* map, const void* base, int esize, int n, int h, int w, int c, int ld, int bc, int bw, int bh,
                         int bn) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) return set_error(B200OV_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)ld * esize, (cuuint64_t)w * ld * esize, (cuuint64_t)h * w * ld * esize};
  cuuint32_t box[4] = {(cuuint32_t)bc, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(map, esize == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims,
                  strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(B200OV_ERR_CUDA, "cuTensorMapEncodeTiled (NHWC tile) failed (%d)", (int)r);
  return B200OV_OK;
}

// Tile geometry shared by the TMA-staged MaxPool / depthwise kernels.  A tile = nimg images x tr output rows x tw output
// columns (nimg * tw <= 32 column lanes) x 32 channels; its input box is nimg x bh x bw pixels.  Columns are cut so
// that the 32 column lanes are well used without too much halo (score = lanes used x useful fraction of the box width);
// rows so that a stage stays within `budget` bytes.
struct TilePlan {
  int tw, tr, nimg, bw, bh, col_tiles, row_tiles, box_bytes;
};
inline bool plan_tiles(int n, int oh, int ow, int K, int S, int chunk_bytes, int budget, TilePlan& t) {
  if (oh <= 0 || ow <= 0 || n <= 0) return false;
  int best_ct = 0;
  double best = -1.0;
  const int ct0 = ceil_div(ow, 32);
  for (int ct = ct0; ct <= ct0 + 3 && ct <= ow; ++ct) {
    const int tw = ceil_div(ow, ct);
    int nimg = 32 / tw;
    if (nimg > n) nimg = n;
    const double score = (double)(nimg * tw) / 32.0 * (double)(tw * S) / (double)((tw - 1) * S + K);
    if (score > best + 1e-9) { best = score; best_ct = ct; }
  }
  t.col_tiles = best_ct;
  t.tw = ceil_div(ow, best_ct);
  t.nimg = 32 / t.tw;
  if (t.nimg > n) t.nimg = n;
  if (t.nimg < 1) t.nimg = 1;
  t.bw = (t.tw - 1) * S + K;
  int tr = oh;
  while (tr > 1 && t.nimg * ((tr - 1) * S + K) * t.bw * chunk_bytes > budget) --tr;
  if (t.nimg > 1 && t.nimg * ((tr - 1) * S + K) * t.bw * chunk_bytes > budget) {      // several images do not fit even one row
    t.nimg = 1;
    tr = oh;
    while (tr > 1 && ((tr - 1) * S + K) * t.bw * chunk_bytes > budget) --tr;
  }
  t.row_tiles = ceil_div(oh, tr);
  t.tr = ceil_div(oh, t.row_tiles);
  t.bh = (t.tr - 1) * S + K;
  t.box_bytes = t.nimg * t.bh * t.bw * chunk_bytes;
  return t.box_bytes <= budget && t.bw <= 256 && t.bh <= 256;
}

__device__ __forceinline__ void load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

}  // namespace tma
}  // namespace b200ov
