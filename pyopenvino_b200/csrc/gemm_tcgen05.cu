// tcgen05 / TMEM / TMA implicit-GEMM path (placeholder until the kernel lands in this file).
#include "common.cuh"

namespace b200ov {

int conv2d_tcgen05(const b200ov_conv_desc* d, const float* x, const float* wp, const float* bias, float* y,
                   cudaStream_t s, bool probe_only) {
  (void)d; (void)x; (void)wp; (void)bias; (void)y; (void)s; (void)probe_only;
  return set_error(B200OV_ERR_UNSUPPORTED, "tcgen05 path not built for this shape");
}

}  // namespace b200ov
