// SSD DetectionOutput on the device, bit-compatible with the reference's python loops
// (DetectionOutput.py:162-260) for the configuration it supports: share_location, normalized priors,
// two-row proposals (boxes + variances), one image per record block.
//
// With a workspace (b200ov_detection_output_ws) step 1 runs first as a grid-wide kernel (a warp per prior of EVERY image:
// 64 x 1917 x 91 scores are 45 MB, HBM work for all 148 SMs instead of 60 serial rounds of one CTA per image); the CTA of an
// image then starts from the (score, class) pairs.
//
// One CTA per image:
//   1. warp per prior: top-1 class over num_classes scores (ties -> higher class index, what
//      np.argsort(...)[::-1][0] gives with numpy 2.x); keep if score > confidence_threshold and class != 0
//      (DetectionOutput.py:69-94, 196-201)
//   2. decode kept priors (CENTER_SIZE / CORNER, DetectionOutput.py:97-151): float32 arithmetic in the
//      reference's operation order with NO fma contraction; exp() evaluated in double then demoted to
//      float32 before the multiply (numpy 2 scalar promotion of `math.exp(...) * prior_width`)
//   3. class-agnostic all-pairs NMS in closed form (DetectionOutput.py:38-63):
//        drop k  <=>  exists j != k: IoU(k, j) > thr and (score[j] > score[k] or (score[j] == score[k] and j < k))
//      (suppressed boxes keep suppressing, exactly like the reference's double loop)
//   4. clip to [0,1] (clip_after_nms), rank survivors by score (descending) and emit
//      [rank, class, score, xmin, ymin, xmax, ymax] records, terminator [-1,0,...] when fewer than keep_top_k.
#include "common.cuh"

namespace b200ov {

struct DetP {
  int num_priors, num_classes, keep_top_k;
  int code_center_size, variance_in_target, clip_before, clip_after;
  float conf_thr, nms_thr;
};

__device__ __forceinline__ float clip01(float v) { return fmaxf(0.f, fminf(1.f, v)); }

__device__ __forceinline__ float iou_ref(const float4 a, const float4 b) {
  const float area_a = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
  const float area_b = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
  const float iw = __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x));
  const float ih = __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y));
  if (iw < 0.f || ih < 0.f) return 0.f;
  const float inter = __fmul_rn(iw, ih);
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));   // 0/0 -> NaN -> never > thr
}

// top-1 class of one prior, computed by a whole warp (ties -> higher class index)
__device__ __forceinline__ void top1_of_prior(const float* __restrict__ c_pr, int num_classes, int lane, float& best, int& best_c) {
  best = -INFINITY;
  best_c = -1;
  for (int c = lane; c < num_classes; c += 32) {
    float v = __ldg(c_pr + c);
    if (v >= best) { best = v; best_c = c; }             // within a lane classes ascend: >= keeps the higher index
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, best, o);
    int oc = __shfl_xor_sync(0xffffffffu, best_c, o);
    if (ov > best || (ov == best && oc > best_c)) { best = ov; best_c = oc; }
  }
}

// decoded (and optionally clipped) box of prior `pr` (DetectionOutput.py:97-151, see the file header)
__device__ __forceinline__ float4 decode_prior(const DetP& p, const float* __restrict__ prior, const float* __restrict__ var,
                                               const float* __restrict__ loc_i, int pr) {
  const float pxmin = prior[pr * 4 + 0], pymin = prior[pr * 4 + 1], pxmax = prior[pr * 4 + 2], pymax = prior[pr * 4 + 3];
  const float l0 = loc_i[pr * 4 + 0], l1 = loc_i[pr * 4 + 1], l2 = loc_i[pr * 4 + 2], l3 = loc_i[pr * 4 + 3];
  const float v0 = var[pr * 4 + 0], v1 = var[pr * 4 + 1], v2 = var[pr * 4 + 2], v3 = var[pr * 4 + 3];
  float4 b;
  if (p.code_center_size) {
    const float pw = __fsub_rn(pxmax, pxmin), ph = __fsub_rn(pymax, pymin);
    const float pcx = __fdiv_rn(__fadd_rn(pxmin, pxmax), 2.f), pcy = __fdiv_rn(__fadd_rn(pymin, pymax), 2.f);
    float cx, cy, bw, bh;
    if (p.variance_in_target) {
      cx = __fadd_rn(__fmul_rn(l0, pw), pcx);
      cy = __fadd_rn(__fmul_rn(l1, ph), pcy);
      bw = __fmul_rn((float)exp((double)l2), pw);
      bh = __fmul_rn((float)exp((double)l3), ph);
    } else {
      cx = __fadd_rn(__fmul_rn(__fmul_rn(v0, l0), pw), pcx);
      cy = __fadd_rn(__fmul_rn(__fmul_rn(v1, l1), ph), pcy);
      bw = __fmul_rn((float)exp((double)__fmul_rn(v2, l2)), pw);
      bh = __fmul_rn((float)exp((double)__fmul_rn(v3, l3)), ph);
    }
    const float hw = __fdiv_rn(bw, 2.f), hh = __fdiv_rn(bh, 2.f);
    b = make_float4(__fsub_rn(cx, hw), __fsub_rn(cy, hh), __fadd_rn(cx, hw), __fadd_rn(cy, hh));
  } else {
    if (p.variance_in_target)
      b = make_float4(__fadd_rn(pxmin, l0), __fadd_rn(pymin, l1), __fadd_rn(pxmax, l2), __fadd_rn(pymax, l3));
    else
      b = make_float4(__fadd_rn(pxmin, __fmul_rn(v0, l0)), __fadd_rn(pymin, __fmul_rn(v1, l1)),
                      __fadd_rn(pxmax, __fmul_rn(v2, l2)), __fadd_rn(pymax, __fmul_rn(v3, l3)));
  }
  if (p.clip_before) b = make_float4(clip01(b.x), clip01(b.y), clip01(b.z), clip01(b.w));
  return b;
}

// step 1 for the whole batch: (best score, best class as int bits) per prior
__global__ void __launch_bounds__(256) detection_top1_kernel(long long total_priors, int num_classes, const float* __restrict__ conf,
                                                             float2* __restrict__ top1) {
  B200OV_PDL_SYNC();
  const int lane = threadIdx.x & 31;
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long pr = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); pr < total_priors; pr += nwarps) {
    float best;
    int best_c;
    top1_of_prior(conf + pr * num_classes, num_classes, lane, best, best_c);
    if (lane == 0) top1[pr] = make_float2(best, __int_as_float(best_c));
  }
}

// dynamic smem: box[num_priors] (float4) | score[num_priors] | label[num_priors] | pidx[num_priors] | keep[num_priors]
__global__ void __launch_bounds__(1024) detection_output_kernel(DetP p, const float* __restrict__ loc,
                                                                const float* __restrict__ conf,
                                                                const float* __restrict__ proposals,
                                                                float* __restrict__ out, const float2* __restrict__ top1) {
  B200OV_PDL_SYNC();
  extern __shared__ __align__(16) uint8_t det_smem[];
  float4* box = reinterpret_cast<float4*>(det_smem);
  float* score = reinterpret_cast<float*>(box + p.num_priors);
  int* label = reinterpret_cast<int*>(score + p.num_priors);
  int* pidx = label + p.num_priors;
  int* keep = pidx + p.num_priors;
  __shared__ int n_cand, n_keep;

  const int img = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const float* loc_i = loc + (long long)img * p.num_priors * 4;
  const float* conf_i = conf + (long long)img * p.num_priors * p.num_classes;
  const float* prior = proposals;                          // [num_priors][4]
  const float* var = proposals + (long long)p.num_priors * 4;
  float* out_i = out + (long long)img * p.keep_top_k * 7;

  if (tid == 0) { n_cand = 0; n_keep = 0; }
  for (int i = tid; i < p.keep_top_k * 7; i += blockDim.x) out_i[i] = 0.f;
  __syncthreads();

  // ---- 1 + 2: top-1 class per prior, threshold, decode ------------------------------------------------
  if (top1 != nullptr) {
    // step 1 already done for the whole batch (detection_top1_kernel): a thread per prior
    const float2* t_i = top1 + (long long)img * p.num_priors;
    for (int pr = tid; pr < p.num_priors; pr += blockDim.x) {
      const float2 t = __ldg(t_i + pr);
      const int best_c = __float_as_int(t.y);
      if (t.x > p.conf_thr && best_c != 0) {
        const int slot = atomicAdd(&n_cand, 1);
        box[slot] = decode_prior(p, prior, var, loc_i, pr);
        score[slot] = t.x;
        label[slot] = best_c;
        pidx[slot] = pr;
      }
    }
  } else {
    for (int pr = warp; pr < p.num_priors; pr += nwarps) {
      float best;
      int best_c;
      top1_of_prior(conf_i + (long long)pr * p.num_classes, p.num_classes, lane, best, best_c);
      if (lane == 0 && best > p.conf_thr && best_c != 0) {
        const int slot = atomicAdd(&n_cand, 1);
        box[slot] = decode_prior(p, prior, var, loc_i, pr);
        score[slot] = best;
        label[slot] = best_c;
        pidx[slot] = pr;
      }
    }
  }
  __syncthreads();
  const int n = n_cand;

  // ---- 3: all-pairs NMS, closed form ------------------------------------------------------------------
  for (int k = tid; k < n; k += blockDim.x) {
    const float4 bk = box[k];
    const float sk = score[k];
    const int ik = pidx[k];
    int alive = 1;
    for (int j = 0; j < n; ++j) {
      if (j == k) continue;
      const float sj = score[j];
      if (!(sj > sk || (sj == sk && pidx[j] < ik))) continue;
      // reference evaluates iou(lower prior index, higher prior index); the formula is symmetric except for
      // the operand order of the area sum, so keep that order
      const float v = (pidx[j] < ik) ? iou_ref(box[j], bk) : iou_ref(bk, box[j]);
      if (v > p.nms_thr) { alive = 0; break; }
    }
    keep[k] = alive;
  }
  __syncthreads();

  // ---- 4: clip, rank by score (descending), emit records ------------------------------------------------
  for (int k = tid; k < n; k += blockDim.x) {
    if (!keep[k]) continue;
    atomicAdd(&n_keep, 1);
    const float sk = score[k];
    const int ik = pidx[k];
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      if (!keep[j] || j == k) continue;
      const float sj = score[j];
      // np.argsort(score)[::-1]: higher score first; equal scores: later prior first (reversed stable order)
      if (sj > sk || (sj == sk && pidx[j] > ik)) ++rank;
    }
    if (rank < p.keep_top_k) {
      float4 b = box[k];
      if (p.clip_after) b = make_float4(clip01(b.x), clip01(b.y), clip01(b.z), clip01(b.w));
      float* rec = out_i + rank * 7;
      rec[0] = (float)rank;
      rec[1] = (float)label[k];
      rec[2] = sk;
      rec[3] = b.x; rec[4] = b.y; rec[5] = b.z; rec[6] = b.w;
    }
  }
  __syncthreads();
  if (tid == 0 && n_keep < p.keep_top_k) out_i[n_keep * 7] = -1.f;     // record terminator
}

}  // namespace b200ov

using namespace b200ov;

static size_t det_ws_bytes(const b200ov_detection_desc* d) { return (size_t)d->n * d->num_priors * sizeof(float2); }

extern "C" int b200ov_detection_output_workspace(const b200ov_detection_desc* d, size_t* bytes) {
  B200OV_REQUIRE(d && bytes && d->n >= 0 && d->num_priors > 0, "detection_output_workspace: bad argument");
  *bytes = det_ws_bytes(d);
  return B200OV_OK;
}

extern "C" int b200ov_detection_output_ws(const b200ov_detection_desc* d, const float* loc, const float* conf,
                                          const float* proposals, float* out, void* workspace, size_t workspace_bytes,
                                          void* stream) {
  B200OV_REQUIRE(d && loc && conf && proposals && out, "detection_output: null argument");
  B200OV_REQUIRE(d->n >= 0 && d->num_priors > 0 && d->num_classes > 0 && d->keep_top_k > 0, "detection_output: bad sizes");
  if (d->n == 0) return B200OV_OK;
  DetP p{d->num_priors, d->num_classes, d->keep_top_k, d->code_center_size, d->variance_encoded_in_target,
         d->clip_before_nms, d->clip_after_nms, d->confidence_threshold, d->nms_threshold};
  size_t smem = (size_t)d->num_priors * (16 + 4 + 4 + 4 + 4);
  B200OV_REQUIRE(smem <= 200 * 1024, "detection_output: %d priors do not fit in shared memory", d->num_priors);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    B200OV_CUDA(cudaFuncSetAttribute(detection_output_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  float2* top1 = nullptr;
  if (workspace != nullptr) {
    B200OV_REQUIRE(workspace_bytes >= det_ws_bytes(d) && (reinterpret_cast<uintptr_t>(workspace) & 7u) == 0,
                   "detection_output: workspace too small or misaligned (%zu bytes)", workspace_bytes);
    top1 = static_cast<float2*>(workspace);
    const long long total = (long long)d->n * d->num_priors;
    launch_k(detection_top1_kernel, bw_grid(total * 32, 256), 256, 0, as_stream(stream), total, d->num_classes, conf, top1);
    B200OV_LAUNCH_CHECK("detection_top1_kernel");
  }
  launch_k(detection_output_kernel, d->n, 1024, smem, as_stream(stream), p, loc, conf, proposals, out, static_cast<const float2*>(top1));
  B200OV_LAUNCH_CHECK("detection_output_kernel");
  return B200OV_OK;
}

extern "C" int b200ov_detection_output(const b200ov_detection_desc* d, const float* loc, const float* conf,
                                       const float* proposals, float* out, void* stream) {
  return b200ov_detection_output_ws(d, loc, conf, proposals, out, nullptr, 0, stream);
}
