// Shared helpers for libb200ov (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/b200ov.h"

namespace b200ov {

// thread-local message behind b200ov_last_error()
char* err_buf();
int set_error(int code, const char* fmt, ...);

struct DeviceProps {
  int device = -1;
  int sm_count = 0;
  int cc_major = 0, cc_minor = 0;
  size_t total_mem = 0;
};
const DeviceProps& props();

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

#define B200OV_CUDA(call)                                                                        \
  do {                                                                                           \
    cudaError_t e__ = (call);                                                                    \
    if (e__ != cudaSuccess)                                                                      \
      return b200ov::set_error(B200OV_ERR_CUDA, "%s failed: %s (%s:%d)", #call,                  \
                               cudaGetErrorString(e__), __FILE__, __LINE__);                     \
  } while (0)

#define B200OV_REQUIRE(cond, ...)                                                                \
  do {                                                                                           \
    if (!(cond)) return b200ov::set_error(B200OV_ERR_INVALID, __VA_ARGS__);                      \
  } while (0)

// launch check: catches bad configurations at the call site without synchronising
#define B200OV_LAUNCH_CHECK(name)                                                                \
  do {                                                                                           \
    cudaError_t e__ = cudaGetLastError();                                                        \
    if (e__ != cudaSuccess)                                                                      \
      return b200ov::set_error(B200OV_ERR_CUDA, "launch of %s failed: %s", name,                 \
                               cudaGetErrorString(e__));                                         \
  } while (0)

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

// Activation applied last in every epilogue.  ReLU keeps numpy's `where(x<0,0,x)` behaviour for
// -0.0 and NaN (ReLU.py:11); Clamp follows np.clip for finite data.
__device__ __forceinline__ float apply_act(float v, int act, float lo, float hi) {
  switch (act) {
    case B200OV_ACT_RELU: return v < 0.f ? 0.f : v;
    case B200OV_ACT_CLAMP: return fminf(fmaxf(v, lo), hi);
    case B200OV_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    default: return v;
  }
}

// ---- programmatic dependent launch (PDL) -------------------------------------------------------------------------------
// Every kernel of the per-inference launch sequence is launched with programmatic stream serialization: its CTAs may be
// scheduled (and run their prologue: barrier init, TMEM allocation, tensor-map prefetch, weight loads) while the previous
// kernel of the stream is still draining, instead of after launch latency + a full drain.  The contract each kernel keeps:
//   * B200OV_PDL_SYNC() before its first read of a tensor another kernel produced AND before its first global write
//     (per-inference buffers are recycled, so the previous kernel may still be reading what this one overwrites);
//     constants (weights, biases, descriptors) may be read before it;
//   * the trigger comes after the wait, so at most one dependent grid is ever launched ahead (no chains of waiting grids).
// `griddepcontrol.wait` returns at once when the kernel was launched without the attribute.  B200OV_PDL=0 disables it.
#define B200OV_PDL_SYNC()                                            \
  do {                                                               \
    asm volatile("griddepcontrol.wait;" ::: "memory");               \
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); \
  } while (0)

bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Grid size for grid-stride bandwidth kernels: enough CTAs to fill every SM a few times over,
// a multiple of the SM count (148 on B200), never more than the work needs.
inline int bw_grid(int64_t work_items, int threads, int ctas_per_sm = 8) {
  int64_t need = ceil_div64(work_items, threads);
  int64_t cap = (int64_t)props().sm_count * ctas_per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

}  // namespace b200ov
