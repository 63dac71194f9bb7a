// b200ov_conv2d / b200ov_matmul: descriptor validation and kernel-family dispatch.
#include "common.cuh"

namespace b200ov {
int conv2d_ffma(const b200ov_conv_desc* d, const float* x, const float* wp, const float* bias, float* y, cudaStream_t s);
bool conv2d_c1_direct_ok(const b200ov_conv_desc* d, const void* x, const void* wp, const void* y);
// tcgen05 path (gemm_tcgen05.cu): returns B200OV_ERR_UNSUPPORTED when the shape is not eligible
int conv2d_tcgen05(const b200ov_conv_desc* d, const float* x, const float* wp, const float* bias, float* y,
                   cudaStream_t s, bool probe_only);
// persistent tcgen05 FP16-split path (conv_f16x2.cu)
bool f16x2_eligible(const b200ov_conv_desc* d, const void* x);
int conv2d_f16x2(const b200ov_conv_desc* d, const void* x, const float* wt, const float* bias, void* y, cudaStream_t s);
int conv2d_f16x2_multi(const b200ov_conv_desc* d, const void* x, const float* wt, const float* bias, int nseg,
                       const b200ov_conv_seg* segs, cudaStream_t s, int ksplit, int ws_rows);
int f16x2_splitk_plan(int m, int cout, int cin, int kh, int kw, int* ws_rows, int* ws_ld);
int conv2d_f16x2_splitk(const b200ov_conv_desc* d, const void* x, const float* wt, const float* bias, float* y, float* ws,
                        size_t ws_bytes, cudaStream_t s);
unsigned int* f16x2_status_word();
bool has_tf32_section(int cin);
void tf32_weight_dims(int cout, int cin, int kh, int kw, int* coutp, long long* kpad);
}  // namespace b200ov

using namespace b200ov;

static int validate(const b200ov_conv_desc* d, const void* x, const float* wp, void* y) {
  B200OV_REQUIRE(d && x && wp && y, "conv2d: null argument");
  B200OV_REQUIRE(d->n >= 0 && d->h > 0 && d->w > 0 && d->cin > 0 && d->cout > 0, "conv2d: bad tensor dims");
  B200OV_REQUIRE(d->kh > 0 && d->kw > 0 && d->sh > 0 && d->sw > 0 && d->pt >= 0 && d->pl >= 0, "conv2d: bad filter geometry");
  B200OV_REQUIRE(d->oh > 0 && d->ow > 0, "conv2d: bad output size");
  B200OV_REQUIRE(d->x_ld >= d->cin && d->y_ld >= d->cout, "conv2d: channel pitch smaller than channel count");
  B200OV_REQUIRE(d->ldw >= d->cout && d->ldw % 64 == 0, "conv2d: weights are not in packed form (ldw %d)", d->ldw);
  B200OV_REQUIRE(d->act >= B200OV_ACT_NONE && d->act <= B200OV_ACT_SIGMOID, "conv2d: bad activation");
  // the last window must start inside the zero-padded image
  B200OV_REQUIRE((d->oh - 1) * d->sh - d->pt < d->h && (d->ow - 1) * d->sw - d->pl < d->w,
                 "conv2d: output size inconsistent with input / stride / pads");
  return B200OV_OK;
}

// packed weight buffer: [FP32 section | tf32 hi/lo planes (when eligible) | f16 hi/lo planes]
static const float* tf32_section(const b200ov_conv_desc* d, const float* w_packed) {
  return w_packed + (long long)round_up(d->kh * d->kw * d->cin, 16) * d->ldw;
}
static const float* f16_section(const b200ov_conv_desc* d, const float* w_packed) {
  const float* p = tf32_section(d, w_packed);
  if (has_tf32_section(d->cin)) {
    int coutp;
    long long kpad;
    tf32_weight_dims(d->cout, d->cin, d->kh, d->kw, &coutp, &kpad);
    p += 2LL * coutp * kpad;
  }
  return p;
}

extern "C" {

int b200ov_conv2d(const b200ov_conv_desc* d, const void* x_raw, const float* w_packed, const float* bias, void* y_raw,
                  void* stream) {
  int rc = validate(d, x_raw, w_packed, y_raw);
  if (rc) return rc;
  cudaStream_t s = as_stream(stream);
  if (d->pre_pool != B200OV_PREPOOL_NONE) {
    // the fused MaxPool lives in the f16x2 contraction's A producers only; no other kernel may silently drop it
    B200OV_REQUIRE(d->pre_pool == B200OV_PREPOOL_MAX3X3S1, "conv2d: unknown pre_pool %d", d->pre_pool);
    if ((d->math != B200OV_MATH_AUTO && d->math != B200OV_MATH_F16X2) || d->x_dtype != B200OV_DT_F32 || d->y_dtype != B200OV_DT_F32 ||
        !f16x2_eligible(d, x_raw))
      return set_error(B200OV_ERR_UNSUPPORTED, "conv2d: pre_pool needs the f16x2 path on FP32 feature maps");
    return conv2d_f16x2(d, x_raw, f16_section(d, w_packed), bias, y_raw, s);
  }
  if (d->x_dtype == B200OV_DT_F32 && d->y_dtype == B200OV_DT_HL && d->cin == 1 && (d->math == B200OV_MATH_AUTO || d->math == B200OV_MATH_FP32) &&
      conv2d_c1_direct_ok(d, x_raw, w_packed, y_raw))
    // C_in = 1 stem whose only readers are contractions: the direct FP32 kernel writes their (hi, lo) operand form
    return conv2d_ffma(d, static_cast<const float*>(x_raw), w_packed, bias, static_cast<float*>(y_raw), s);
  if (d->x_dtype != B200OV_DT_F32 || d->y_dtype != B200OV_DT_F32) {
    // FP16 feature maps: only the f16x2 contraction reads / writes them
    B200OV_REQUIRE((d->x_dtype == B200OV_DT_F32 || d->x_dtype == B200OV_DT_F16 || d->x_dtype == B200OV_DT_HL) &&
                       (d->y_dtype == B200OV_DT_F32 || d->y_dtype == B200OV_DT_F16 || d->y_dtype == B200OV_DT_HL),
                   "conv2d: bad storage type");
    if ((d->math != B200OV_MATH_AUTO && d->math != B200OV_MATH_F16X2) || !f16x2_eligible(d, x_raw))
      return set_error(B200OV_ERR_UNSUPPORTED, "conv2d: FP16 feature maps need the f16x2 path (cin %% 8 == 0, 16-byte aligned pixels)");
    return conv2d_f16x2(d, x_raw, f16_section(d, w_packed), bias, y_raw, s);
  }
  const float* x = static_cast<const float*>(x_raw);
  float* y = static_cast<float*>(y_raw);
  switch (d->math) {
    case B200OV_MATH_FP32:
      return conv2d_ffma(d, x, w_packed, bias, y, s);
    case B200OV_MATH_TF32X3:
    case B200OV_MATH_TF32:
      return conv2d_tcgen05(d, x, tf32_section(d, w_packed), bias, y, s, false);
    case B200OV_MATH_F16X2:
      return conv2d_f16x2(d, x, f16_section(d, w_packed), bias, y, s);
    case B200OV_MATH_AUTO:
      if (f16x2_eligible(d, x)) return conv2d_f16x2(d, x, f16_section(d, w_packed), bias, y, s);
      /* fall through */
    case B200OV_MATH_SAFE: {
      if (conv2d_tcgen05(d, x, nullptr, bias, y, s, true) == B200OV_OK) {
        b200ov_conv_desc dd = *d;
        dd.math = B200OV_MATH_TF32X3;
        return conv2d_tcgen05(&dd, x, tf32_section(d, w_packed), bias, y, s, false);
      }
      return conv2d_ffma(d, x, w_packed, bias, y, s);
    }
    default:
      return set_error(B200OV_ERR_INVALID, "conv2d: unknown math mode %d", d->math);
  }
}

int b200ov_conv2d_multi(const b200ov_conv_desc* d, const void* x, const float* w_packed, const float* bias, int nseg,
                        const b200ov_conv_seg* segs, void* stream) {
  B200OV_REQUIRE(d && x && w_packed && segs, "conv2d_multi: null argument");
  B200OV_REQUIRE(d->n >= 0 && d->h > 0 && d->w > 0 && d->cin > 0 && d->cout > 0 && d->kh > 0 && d->kw > 0 && d->sh > 0 && d->sw > 0 &&
                     d->pt >= 0 && d->pl >= 0 && d->oh > 0 && d->ow > 0 && d->x_ld >= d->cin,
                 "conv2d_multi: bad geometry");
  B200OV_REQUIRE(d->ldw >= d->cout && d->ldw % 64 == 0, "conv2d_multi: weights are not in packed form (ldw %d)", d->ldw);
  B200OV_REQUIRE(d->act >= B200OV_ACT_NONE && d->act <= B200OV_ACT_CLAMP, "conv2d_multi: bad activation");
  if (d->math != B200OV_MATH_AUTO && d->math != B200OV_MATH_F16X2)
    return set_error(B200OV_ERR_UNSUPPORTED, "conv2d_multi: only the f16x2 path supports several output tensors");
  return conv2d_f16x2_multi(d, x, f16_section(d, w_packed), bias, nseg, segs, as_stream(stream), 1, 0);
}

int b200ov_matmul(int m, int n, int k, const float* a, int lda, const float* b_packed, int ldw, const float* bias,
                  int act, float act_lo, float act_hi, int math, float* y, int ldy, void* stream) {
  return b200ov_matmul_ws(m, n, k, a, lda, b_packed, ldw, bias, act, act_lo, act_hi, math, y, ldy, nullptr, 0, stream);
}

int b200ov_matmul_workspace(int m, int n, int k, size_t* bytes) {
  B200OV_REQUIRE(bytes && m >= 0 && n > 0 && k > 0, "matmul_workspace: bad argument");
  *bytes = 0;
  if (m == 0 || k % 8 != 0) return B200OV_OK;
  int ws_rows, ws_ld;
  const int ksplit = f16x2_splitk_plan(m, n, k, 1, 1, &ws_rows, &ws_ld);
  if (ksplit > 1) *bytes = (size_t)ksplit * ws_rows * ws_ld * sizeof(float);
  return B200OV_OK;
}

int b200ov_matmul_ws(int m, int n, int k, const float* a, int lda, const float* b_packed, int ldw, const float* bias,
                     int act, float act_lo, float act_hi, int math, float* y, int ldy, void* workspace, size_t workspace_bytes,
                     void* stream) {
  B200OV_REQUIRE(m >= 0 && n > 0 && k > 0, "matmul: bad dims");
  b200ov_conv_desc d;
  memset(&d, 0, sizeof(d));
  // rows of A are the "pixels" of a 1x1 convolution over a 1 x m image with k channels
  d.n = 1; d.h = 1; d.w = m > 0 ? m : 1; d.cin = k; d.cout = n; d.kh = 1; d.kw = 1; d.sh = 1; d.sw = 1;
  d.pt = 0; d.pl = 0; d.oh = 1; d.ow = m > 0 ? m : 1; d.x_ld = lda; d.y_ld = ldy; d.ldw = ldw;
  d.act = act; d.act_lo = act_lo; d.act_hi = act_hi; d.math = math;
  if (m == 0) return B200OV_OK;
  if (workspace != nullptr && workspace_bytes > 0 && (math == B200OV_MATH_AUTO || math == B200OV_MATH_F16X2) && f16x2_eligible(&d, a) &&
      act != B200OV_ACT_SIGMOID) {
    int rc = validate(&d, a, b_packed, y);
    if (rc) return rc;
    return conv2d_f16x2_splitk(&d, a, f16_section(&d, b_packed), bias, y, static_cast<float*>(workspace), workspace_bytes,
                               as_stream(stream));
  }
  return b200ov_conv2d(&d, a, b_packed, bias, y, stream);
}

/* Sticky device status word (bit 0: a non-finite value left an f16x2 contraction, i.e. an operand exceeded the
 * FP16 range or the data held inf/NaN).  The caller resets it (cudaMemsetAsync) before an inference and reads
 * it back with the results; if set, re-run with B200OV_MATH_TF32X3. */
int b200ov_status_word(void** device_ptr) {
  B200OV_REQUIRE(device_ptr, "status_word: null argument");
  unsigned int* p = f16x2_status_word();
  if (p == nullptr) return set_error(B200OV_ERR_CUDA, "status_word: cudaGetSymbolAddress failed");
  *device_ptr = p;
  return B200OV_OK;
}

int b200ov_status_reset(void* stream) {
  unsigned int* p = f16x2_status_word();
  if (p == nullptr) return set_error(B200OV_ERR_CUDA, "status_reset: cudaGetSymbolAddress failed");
  B200OV_CUDA(cudaMemsetAsync(p, 0, sizeof(unsigned int), as_stream(stream)));
  return B200OV_OK;
}

int b200ov_status_fetch(uint32_t* host_out, void* stream) {
  B200OV_REQUIRE(host_out, "status_fetch: null argument");
  unsigned int* p = f16x2_status_word();
  if (p == nullptr) return set_error(B200OV_ERR_CUDA, "status_fetch: cudaGetSymbolAddress failed");
  B200OV_CUDA(cudaMemcpyAsync(host_out, p, sizeof(unsigned int), cudaMemcpyDeviceToHost, as_stream(stream)));
  return B200OV_OK;
}

}  // extern "C"
