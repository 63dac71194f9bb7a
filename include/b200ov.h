/*
 * b200ov.h -- C ABI of libb200ov.so: the B200 (sm_100a) kernels behind pyOpenVINO's op-plugin
 * compute() contract.
 *
 * Every entry point is `extern "C"`, takes plain pointers / sizes / POD descriptors and a
 * cudaStream_t passed as `void*`; no C++ or torch types cross this boundary.  All functions
 * return B200OV_OK (0) or an error code; b200ov_last_error() returns the message of the last
 * failure on the calling thread.  Nothing here falls back to the CPU.
 *
 * Device tensor conventions
 *   - activations: float32, physical layout NHWC with a channel pitch `ld` (floats between
 *     consecutive pixels, ld >= C).  ld > C lets a producer write straight into a channel slice
 *     of a Concat buffer.  The logical (IR) shape stays NCHW; the Python host tracks the tag.
 *   - 2-D tensors (MatMul operands, SoftMax rows): row-major.
 *   - conv weights: packed once per network by b200ov_pack_conv_weights() into one buffer with two
 *     sections: (1) FP32 section [K = kh*kw*cin rows (padded to 16)][ldw = cout padded to 64] for the
 *     CUDA-core kernel; (2) when cin % 4 == 0 and cin >= 8, the tcgen05 section: tf32 hi and lo planes,
 *     each [cout padded to 32][kh*kw*(cin padded to 32)] K-major, read by TMA.
 *
 * Each function cites the reference interface it replaces (paths relative to the
 * yas-sim/pyopenvino root).
 */
#ifndef B200OV_H
#define B200OV_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200OV_VERSION 100

enum {
  B200OV_OK = 0,
  B200OV_ERR_INVALID = 1,     /* bad descriptor / argument */
  B200OV_ERR_CUDA = 2,        /* a CUDA runtime / driver call failed */
  B200OV_ERR_UNSUPPORTED = 3  /* valid request this build has no kernel for */
};

/* fused activation applied last in every epilogue */
enum {
  B200OV_ACT_NONE = 0,
  B200OV_ACT_RELU = 1,    /* ReLU.py:9-12   : x < 0 ? 0 : x            */
  B200OV_ACT_CLAMP = 2,   /* Clamp.py:9-12  : min(max(x, lo), hi)      */
  B200OV_ACT_SIGMOID = 3  /* Sigmoid.py:10-13: 1 / (1 + exp(-x))       */
};

/* arithmetic used by the dense contractions */
enum {
  B200OV_MATH_AUTO = 0,     /* pick per shape: F16X2 when eligible, else 3xTF32, else FP32 FFMA */
  B200OV_MATH_FP32 = 1,     /* CUDA-core FP32 FFMA implicit GEMM                                */
  B200OV_MATH_TF32X3 = 2,   /* tcgen05 kind::tf32, hi/lo split, 3 MMAs, FP32 accumulate in TMEM */
  B200OV_MATH_TF32 = 3,     /* tcgen05 kind::tf32 single pass (1e-3 tolerance class)            */
  B200OV_MATH_F16X2 = 4,    /* tcgen05 kind::f16, FP16 hi/lo split, 3 MMAs, A operand in TMEM;  *
                             * FP32-accurate for |values| < 65504, overflow raises the status word */
  B200OV_MATH_SAFE = 5      /* AUTO without F16X2 (full FP32 exponent range): 3xTF32, else FFMA  */
};

/* Element types.  Host inputs may arrive as any of them (Parameter.py:13 casts whatever array-like it is given with
 * `np.array(param).reshape(shape).astype(precision)`; draw-and-infer.py:56-60 feeds uint8): the bytes cross PCIe in their
 * native width and are widened on the device (exact, so bit-identical to the host cast).  F32 / F16 are also the two
 * STORAGE types of NHWC feature maps in HBM: F32 is the reference's precision (models/*.xml ports are FP32) and the
 * default; F16 is the opt-in storage mode (`load_network(..., storage='f16')`): every kernel still computes in FP32 and
 * rounds to nearest even on store, halving the bytes of the bandwidth-bound layers.  The reference's plugins are
 * dtype-generic (common_def.py:18-19) and its original IRs were FP16 (GroupConvolution.py:136-143). */
enum { B200OV_DT_F32 = 0, B200OV_DT_F16 = 1, B200OV_DT_U8 = 2, B200OV_DT_I8 = 3,
       B200OV_DT_HL = 4   /* the contraction's own operand form, same 4 bytes per value as FP32: per 4 channels (16 bytes)
                             [hi(c0,c1) hi(c2,c3) lo(c0,c1) lo(c2,c3)] FP16 pairs with c = hi + 2^-11 lo -- exactly the 22
                             significant bits the contraction's split keeps, so a contraction reading it produces the same
                             bits as on the FP32 tensor.  Written by b200ov_input_to_nhwc_split (network input of a <= 4
                             channel stem) and by b200ov_conv2d / _multi / b200ov_dwconv2d for outputs whose only readers
                             are contractions (y_dtype); only b200ov_conv2d / _multi read it */
};

/* ---- library / device -------------------------------------------------------------------- */
int b200ov_version(void);
const char* b200ov_last_error(void);
/* Select the device for the calling thread and cache its properties. */
int b200ov_init(int device);
int b200ov_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem_bytes);

/* ---- memory, copies, streams (the Python host usually brings torch-owned pointers instead) -- */
int b200ov_malloc(void** dptr, size_t bytes);
int b200ov_free(void* dptr);
int b200ov_host_alloc(void** hptr, size_t bytes);   /* pinned */
int b200ov_host_free(void* hptr);
int b200ov_memcpy_h2d(void* dst, const void* src, size_t bytes, void* stream);
int b200ov_memcpy_d2h(void* dst, const void* src, size_t bytes, void* stream);
int b200ov_memcpy_d2d(void* dst, const void* src, size_t bytes, void* stream);
int b200ov_memset(void* dst, int value, size_t bytes, void* stream);
int b200ov_stream_create(void** stream);
int b200ov_stream_destroy(void* stream);
int b200ov_stream_sync(void* stream);

/* CUDA-graph capture of a launch sequence (replaces the per-node python dispatch loop of
 * Executable_Network.run_tasks, inference_engine.py:259-292, on the replay path). */
int b200ov_graph_begin(void* stream);
int b200ov_graph_end(void* stream, void** graph_exec);
int b200ov_graph_launch(void* graph_exec, void* stream);
int b200ov_graph_destroy(void* graph_exec);

/* ---- Convolution / MatMul ------------------------------------------------------------------ */
typedef struct {
  int32_t n, h, w, cin;        /* input  [n][h][w][cin], pixel pitch x_ld                        */
  int32_t cout, kh, kw;        /* filter                                                         */
  int32_t sh, sw;              /* strides                                                        */
  int32_t pt, pl;              /* pads_begin (top, left); zero padding, Convolution.py:63,104    */
  int32_t oh, ow;              /* output size from calc_output_shape, Convolution.py:21-49       */
  int32_t x_ld, y_ld;          /* channel pitch of input / output pixels (floats)                */
  int32_t ldw;                 /* pitch of the packed weight rows (floats)                       */
  int32_t act;                 /* B200OV_ACT_*                                                   */
  float act_lo, act_hi;        /* Clamp bounds                                                   */
  int32_t math;                /* B200OV_MATH_*                                                  */
  int32_t x_dtype, y_dtype;    /* storage type of x / y: B200OV_DT_F32 (0, default) or B200OV_DT_F16; F16 needs the
                                  B200OV_MATH_F16X2 path (an FP16 input IS its own hi part: 2 MMAs per product)  */
  int32_t pre_pool;            /* B200OV_PREPOOL_*: an operator applied to x on the way into the contraction, so that its
                                  result never exists in memory (the 3x3 stride-1 MaxPool in front of an inception
                                  module's pool_proj 1x1 convolution)                                             */
} b200ov_conv_desc;

/* pre_pool values.  MAX3X3S1: x is first passed through MaxPool kernel 3x3, strides 1, pads_begin = pads_end = 1 with the
 * reference's window rule (zero padding takes part in the max, MaxPool.py:41-72); needs a 1x1 / stride-1 / unpadded
 * convolution of FP32 feature maps on the B200OV_MATH_F16X2 path (B200OV_ERR_UNSUPPORTED otherwise: the caller then
 * issues b200ov_pool2d + b200ov_conv2d).  The result is bit-identical to those two calls. */
#define B200OV_PREPOOL_NONE 0
#define B200OV_PREPOOL_MAX3X3S1 1

/* OIHW -> packed [kh*kw*cin (pad 16)][ldw] (row index ordered ky, kx, ci).  `ldw` and the row
 * count come from b200ov_conv_weight_dims(); `w_packed` must hold `total_floats`.  Replaces `kernel.reshape(kn,-1).T`, Convolution.py:83. */
int b200ov_conv_weight_dims(int cout, int cin, int kh, int kw, int* rows, int* ldw, int64_t* total_floats);
int b200ov_pack_conv_weights(const float* w_oihw, float* w_packed, int cout, int cin, int kh, int kw, void* stream);

/* y = act(conv(x, w) + bias).  Replaces Convolution.compute (Convolution.py:149-176) and, through
 * the epilogue, the Add (Add.py:9-14) and ReLU / Clamp nodes that follow it.  `bias` may be NULL. */
int b200ov_conv2d(const b200ov_conv_desc* d, const void* x, const float* w_packed, const float* bias,
                  void* y, void* stream);

/* Sibling convolutions that read the same feature map with the same geometry (the 1x1 / 3x3_reduce / 5x5_reduce
 * branches of an inception module) as ONE contraction: `w_packed` / `bias` are the packed form of the weights
 * concatenated along C_out, each member starting at a column that is a multiple of 32 (zero rows / zeros in the gaps),
 * d->cout = width of that fused matrix (d->y_ld is ignored).  Output columns [col0, col0 + cout) go to tensor `y`
 * with channel pitch `y_ld`.  1..3 segments; needs the B200OV_MATH_F16X2 path (B200OV_ERR_UNSUPPORTED otherwise,
 * the caller then issues the members one by one).  Replaces the corresponding Convolution.compute calls
 * (Convolution.py:149-176) + Add + ReLU. */
typedef struct {
  void* y;
  int32_t col0, cout, y_ld;
  int32_t y_dtype;             /* B200OV_DT_F32 (0: the descriptor's y_dtype) or B200OV_DT_HL: this output is read by
                                  contractions only (cout % 4 == 0, 16-byte aligned pixels)                       */
} b200ov_conv_seg;
int b200ov_conv2d_multi(const b200ov_conv_desc* d, const void* x, const float* w_packed, const float* bias,
                        int nseg, const b200ov_conv_seg* segs, void* stream);

/* Y[m][n] = act(sum_k A[m][k] * Bkn[k][n] + bias[n]); Bkn is the packed form of a 1x1 conv weight
 * (b200ov_pack_conv_weights with kh = kw = 1 on B[n][k]).  Replaces MatMul.compute for
 * transpose_a=false / transpose_b=true (MatMul.py:9-17); other flag combinations are brought to
 * this form by the host with b200ov_transpose. */
int b200ov_matmul(int m, int n, int k, const float* a, int lda, const float* b_packed, int ldw,
                  const float* bias, int act, float act_lo, float act_hi, int math,
                  float* y, int ldy, void* stream);

/* Split-K form for products with few output tiles (6272 -> 512 at batch 1024 is 32 tiles on 148 SMs): several CTAs share
 * one output tile, each over a slice of K, writing raw partial sums to `workspace`; a second kernel adds them in split
 * order (deterministic), then bias and activation.  b200ov_matmul_workspace() returns the bytes b200ov_matmul_ws() wants
 * for this shape (0: no split, plain b200ov_matmul behaviour; a NULL workspace is always accepted).  MatMul.py:9-17. */
int b200ov_matmul_workspace(int m, int n, int k, size_t* bytes);
int b200ov_matmul_ws(int m, int n, int k, const float* a, int lda, const float* b_packed, int ldw,
                     const float* bias, int act, float act_lo, float act_hi, int math,
                     float* y, int ldy, void* workspace, size_t workspace_bytes, void* stream);

/* Device address of the library's sticky status word (uint32).  Bit 0 is set by a B200OV_MATH_F16X2
 * contraction whose output held a non-finite value (an operand beyond the FP16 range, or inf/NaN data).
 * The executor clears it before an inference (cudaMemsetAsync), reads it back with the results and, if set,
 * repeats the inference with B200OV_MATH_TF32X3.  The reference has no such condition: its numpy kernels
 * compute in FP32 throughout (Convolution.py:83-84). */
int b200ov_status_word(void** device_ptr);
/* Clear the status word / copy it to (pinned) host memory, both asynchronously on `stream`. */
int b200ov_status_reset(void* stream);
int b200ov_status_fetch(uint32_t* host_out, void* stream);

/* ---- depthwise GroupConvolution ------------------------------------------------------------ */
typedef struct {
  int32_t n, h, w, c;
  int32_t kh, kw, sh, sw, pt, pl, oh, ow;
  int32_t x_ld, y_ld;
  int32_t act;
  float act_lo, act_hi;
  int32_t math;                /* B200OV_DW_*                                                    */
  int32_t dtype;               /* storage type of x and y: B200OV_DT_F32 (0, default) or B200OV_DT_F16 (3x3 only) */
  int32_t y_dtype;             /* 0 / dtype: y stored like x.  B200OV_DT_HL (x FP32): y is read by contractions only (the
                                  pointwise convolution after a depthwise one) and is written in their (hi, lo) operand
                                  form; B200OV_ERR_UNSUPPORTED when the shape has no kernel for it (retry with 0)   */
} b200ov_dwconv_desc;

enum {
  B200OV_DW_AUTO = 0,   /* 3x3: packed-FP32 FMA chain from the bias (a few ulp from the reference)          */
  B200OV_DW_EXACT = 1   /* products rounded individually, numpy pairwise order: bit-identical to np.sum() */
};

/* [C][1][1][kh][kw] -> [kh*kw][C] */
int b200ov_pack_dw_weights(const float* w_g11hw, float* w_packed, int c, int kh, int kw, void* stream);
/* Replaces GroupConvolution.compute (GroupConvolution.py:114-137, depthwise case) + Add + Clamp.
 * math = B200OV_DW_EXACT: products are summed in numpy's pairwise order without FMA contraction, so the
 * pre-bias value is bit-identical to `np.sum(patch*flt)` (GroupConvolution.py:78); B200OV_DW_AUTO may use
 * an FMA chain (3x3 windows), which meets the FP32 tolerance class but is not bit-identical. */
int b200ov_dwconv2d(const b200ov_dwconv_desc* d, const void* x, const float* w_packed, const float* bias,
                    void* y, void* stream);

/* ---- pooling --------------------------------------------------------------------------------- */
enum { B200OV_POOL_MAX = 0, B200OV_POOL_AVG_REF = 1 };
typedef struct {
  int32_t n, h, w, c;
  int32_t kh, kw, sh, sw;
  int32_t pt, pl, pb, pr;      /* pads_begin / pads_end                                          */
  int32_t oh, ow;
  int32_t x_ld, y_ld;
  int32_t mode;                /* MAX: zero padding takes part, overhang clipped (MaxPool.py:53,69)
                                  AVG_REF: no padding, window clipped at h-1 / w-1 (AvgPool.py:56) */
  int32_t dtype;               /* storage type of x and y: B200OV_DT_F32 (0, default) or B200OV_DT_F16        */
} b200ov_pool_desc;
/* Replaces MaxPool.compute (MaxPool.py:111-135) / AvgPool.compute (AvgPool.py:94-118).  Optional
 * per-channel epilogue y = y*scale[c] + shift[c] (the folded BatchNorm Multiply+Add that follows
 * the pools of mnist_bn); either pointer may be NULL. */
int b200ov_pool2d(const b200ov_pool_desc* d, const void* x, const float* scale, const float* shift,
                  void* y, void* stream);

/* ---- elementwise tail -------------------------------------------------------------------------- */
/* y[i] = act((x[i] * s) + b) over rows x C elements with channel = i % C (NHWC pixels or 2-D rows).
 *   scale_vec / shift_vec : per-channel vectors or NULL; when NULL the scalar scale_s / shift_s is
 *   used, and has_scale / has_shift == 0 skips the step entirely (keeps results bit-identical to
 *   the separate numpy ops).  Replaces Add.py:9-14, Multiply.py:9-17 (broadcast operand cases),
 *   ReLU.py:9-12, Clamp.py:9-12, Sigmoid.py:10-13 as standalone nodes. */
int b200ov_affine_act(const float* x, float* y, int64_t rows, int c, int x_ld, int y_ld,
                      int has_scale, const float* scale_vec, float scale_s,
                      int has_shift, const float* shift_vec, float shift_s,
                      int act, float act_lo, float act_hi, void* stream);
/* y = a (+|*) b for two tensors of identical shape; op 0 = add, 1 = multiply. */
int b200ov_binary(int op, const float* a, const float* b, float* y, int64_t count, void* stream);
/* Row softmax, y[r][:] = exp(x[r][:]) / sum(exp(x[r][:])) (max-shifted form).  Replaces
 * SoftMax.compute (SoftMax.py:10-14); batched meaning = one row per image. */
int b200ov_softmax(const float* x, float* y, int rows, int cols, void* stream);
/* Across-channel LRN on NHWC pixels, LRN.py:10-22 (alpha NOT divided by size). */
int b200ov_lrn(const float* x, float* y, int64_t pixels, int c, int x_ld, int y_ld, int size,
               float alpha, float beta, float bias, void* stream);

/* b200ov_lrn on a feature map stored as `dtype` (B200OV_DT_F32 / B200OV_DT_F16), input and output alike. */
int b200ov_lrn_st(const void* x, void* y, int dtype, int64_t pixels, int c, int x_ld, int y_ld, int size,
                  float alpha, float beta, float bias, void* stream);

/* ---- layout glue ----------------------------------------------------------------------------- */
/* [batch][rows][cols] -> [batch][cols][rows] with an output pitch (NCHW<->NHWC, Transpose.py:9-13,
 * MatMul transpose flags).  Optional y = x*scale[c]+shift[c] on the fly when the source is NCHW
 * (c = row index): the Parameter -> mean/scale pre-processing of GoogLeNet / SSD. */
int b200ov_transpose(const float* x, float* y, int batch, int rows, int cols, int x_ld, int y_ld,
                     void* stream);
int b200ov_nchw_to_nhwc_affine(const float* x, float* y, int n, int c, int hw, int y_ld,
                               int has_scale, const float* scale_vec, float scale_s,
                               int has_shift, const float* shift_vec, float shift_s, void* stream);
/* b200ov_nchw_to_nhwc_affine for an input of element type `dtype` (NCHW, dense). */
int b200ov_input_to_nhwc(const void* x, int dtype, float* y, int n, int c, int hw, int y_ld,
                         int has_scale, const float* scale_vec, float scale_s,
                         int has_shift, const float* shift_vec, float shift_s, void* stream);
/* The same with the FP16 (hi, lo) split of the f16x2 contraction done here, once per input pixel, instead of once per
 * filter tap in the stem convolution's producer warps (y: B200OV_DT_HL, pixel pitch 16 bytes; needs C <= 4). */
int b200ov_input_to_nhwc_split(const void* x, int dtype, void* y, int n, int c, int hw,
                               int has_scale, const float* scale_vec, float scale_s,
                               int has_shift, const float* shift_vec, float shift_s, void* stream);
/* y[i] = (float)x[i]: a non-image (not 4-D) input of element type `dtype` (Parameter.py:13). */
int b200ov_widen(const void* x, int dtype, float* y, int64_t count, void* stream);
/* b200ov_transpose / b200ov_copy2d between storage types (F32 / F16): the FP16 storage mode's NHWC feature maps leave
 * the device-resident path through these (NHWC f16 -> plain NCHW f32 in front of Reshape / MatMul / Result). */
int b200ov_transpose_st(const void* x, int x_dtype, void* y, int y_dtype, int batch, int rows, int cols,
                        int x_ld, int y_ld, void* stream);
int b200ov_copy2d_st(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t rows, int cols,
                     int src_ld, int dst_ld, void* stream);
/* rows x cols strided copy (Concat.py:9-13 when producers could not write in place). */
int b200ov_copy2d(const float* src, float* dst, int64_t rows, int cols, int src_ld, int dst_ld, void* stream);
/* Concat of `nparts` (<= B200OV_CONCAT_MAX_PARTS) dense row blocks along the columns in one launch: dst[r][off_p + c] = srcs[p][r][c],
 * dst dense with sum(cols) columns (Concat.py:9-13, `np.concatenate(..., axis)` with everything before `axis` folded into rows). */
#define B200OV_CONCAT_MAX_PARTS 8
int b200ov_concat_rows(int nparts, const float* const* srcs, const int* cols, float* dst, int64_t rows, void* stream);

/* ---- SSD DetectionOutput ------------------------------------------------------------------------ */
typedef struct {
  int32_t n;                          /* images                                                       */
  int32_t num_priors, num_classes;    /* 1917, 91 for ssd_mobilenet_v1_coco                            */
  int32_t keep_top_k;                 /* records per image                                            */
  int32_t code_center_size;           /* 1 = caffe.PriorBoxParameter.CENTER_SIZE, 0 = CORNER          */
  int32_t variance_encoded_in_target;
  int32_t clip_before_nms, clip_after_nms;
  float confidence_threshold, nms_threshold;
} b200ov_detection_desc;
/* Replaces DetectionOutput.compute (DetectionOutput.py:162-300) for share_location / normalized priors:
 * top-1 class per prior, threshold, box decode, class-agnostic all-pairs NMS, score-ordered records
 * [rank, class, score, xmin, ymin, xmax, ymax] with a [-1,0,..] terminator.  loc [n][priors*4],
 * conf [n][priors*classes], proposals [2][priors*4] (boxes, variances), out [n*keep_top_k][7].
 * Decisions and record order are bit-compatible with the reference's float32 python arithmetic. */
int b200ov_detection_output(const b200ov_detection_desc* d, const float* loc, const float* conf,
                            const float* proposals, float* out, void* stream);
/* The same with a caller-provided scratch buffer of b200ov_detection_output_workspace() bytes (8-byte aligned): the top-1 class
 * per prior (DetectionOutput.py:69-94) then runs as a grid-wide kernel over the whole batch before the per-image CTAs.
 * Same records bit for bit. */
int b200ov_detection_output_workspace(const b200ov_detection_desc* d, size_t* bytes);
int b200ov_detection_output_ws(const b200ov_detection_desc* d, const float* loc, const float* conf,
                               const float* proposals, float* out, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200OV_H */
