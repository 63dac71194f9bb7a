// Library / device / memory / stream / CUDA-graph entry points of libb200ov.
#include <stdarg.h>

#include "common.cuh"

namespace b200ov {

static thread_local char g_err[512] = "";
static DeviceProps g_props;

char* err_buf() { return g_err; }

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

const DeviceProps& props() {
  if (g_props.device < 0) {
    // lazily bind to the current device so kernels work even if b200ov_init was skipped
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) {
      cudaDeviceProp p;
      if (cudaGetDeviceProperties(&p, dev) == cudaSuccess) {
        g_props.device = dev;
        g_props.sm_count = p.multiProcessorCount;
        g_props.cc_major = p.major;
        g_props.cc_minor = p.minor;
        g_props.total_mem = p.totalGlobalMem;
      }
    }
    if (g_props.sm_count <= 0) g_props.sm_count = 148;
  }
  return g_props;
}

// Programmatic dependent launch is opt-in (B200OV_PDL=1): see common.cuh.
bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("B200OV_PDL");
    on = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return on == 1;
}

}  // namespace b200ov

using namespace b200ov;

extern "C" {

int b200ov_version(void) { return B200OV_VERSION; }

const char* b200ov_last_error(void) { return err_buf(); }

int b200ov_init(int device) {
  int count = 0;
  B200OV_CUDA(cudaGetDeviceCount(&count));
  B200OV_REQUIRE(device >= 0 && device < count, "device %d out of range (have %d)", device, count);
  B200OV_CUDA(cudaSetDevice(device));
  cudaDeviceProp p;
  B200OV_CUDA(cudaGetDeviceProperties(&p, device));
  if (p.major != 10)
    return set_error(B200OV_ERR_UNSUPPORTED, "libb200ov is built for sm_100a only; device %d is sm_%d%d",
                     device, p.major, p.minor);
  g_props.device = device;
  g_props.sm_count = p.multiProcessorCount;
  g_props.cc_major = p.major;
  g_props.cc_minor = p.minor;
  g_props.total_mem = p.totalGlobalMem;
  return B200OV_OK;
}

int b200ov_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem_bytes) {
  const DeviceProps& p = props();
  B200OV_REQUIRE(p.device >= 0, "no CUDA device bound");
  if (sm_count) *sm_count = p.sm_count;
  if (cc_major) *cc_major = p.cc_major;
  if (cc_minor) *cc_minor = p.cc_minor;
  if (total_mem_bytes) *total_mem_bytes = p.total_mem;
  return B200OV_OK;
}

int b200ov_malloc(void** dptr, size_t bytes) {
  B200OV_REQUIRE(dptr != nullptr, "null out pointer");
  B200OV_CUDA(cudaMalloc(dptr, bytes ? bytes : 1));
  return B200OV_OK;
}
int b200ov_free(void* dptr) {
  B200OV_CUDA(cudaFree(dptr));
  return B200OV_OK;
}
int b200ov_host_alloc(void** hptr, size_t bytes) {
  B200OV_REQUIRE(hptr != nullptr, "null out pointer");
  B200OV_CUDA(cudaMallocHost(hptr, bytes ? bytes : 1));
  return B200OV_OK;
}
int b200ov_host_free(void* hptr) {
  B200OV_CUDA(cudaFreeHost(hptr));
  return B200OV_OK;
}
int b200ov_memcpy_h2d(void* dst, const void* src, size_t bytes, void* stream) {
  B200OV_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, as_stream(stream)));
  return B200OV_OK;
}
int b200ov_memcpy_d2h(void* dst, const void* src, size_t bytes, void* stream) {
  B200OV_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, as_stream(stream)));
  return B200OV_OK;
}
int b200ov_memcpy_d2d(void* dst, const void* src, size_t bytes, void* stream) {
  B200OV_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, as_stream(stream)));
  return B200OV_OK;
}
int b200ov_memset(void* dst, int value, size_t bytes, void* stream) {
  B200OV_CUDA(cudaMemsetAsync(dst, value, bytes, as_stream(stream)));
  return B200OV_OK;
}
int b200ov_stream_create(void** stream) {
  B200OV_REQUIRE(stream != nullptr, "null out pointer");
  cudaStream_t s;
  B200OV_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  *stream = s;
  return B200OV_OK;
}
int b200ov_stream_destroy(void* stream) {
  B200OV_CUDA(cudaStreamDestroy(as_stream(stream)));
  return B200OV_OK;
}
int b200ov_stream_sync(void* stream) {
  B200OV_CUDA(cudaStreamSynchronize(as_stream(stream)));
  return B200OV_OK;
}

int b200ov_graph_begin(void* stream) {
  B200OV_CUDA(cudaStreamBeginCapture(as_stream(stream), cudaStreamCaptureModeThreadLocal));
  return B200OV_OK;
}
int b200ov_graph_end(void* stream, void** graph_exec) {
  B200OV_REQUIRE(graph_exec != nullptr, "null out pointer");
  cudaGraph_t g = nullptr;
  B200OV_CUDA(cudaStreamEndCapture(as_stream(stream), &g));
  cudaGraphExec_t ge = nullptr;
  cudaError_t e = cudaGraphInstantiate(&ge, g, 0);
  cudaGraphDestroy(g);
  if (e != cudaSuccess) return set_error(B200OV_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
  *graph_exec = ge;
  return B200OV_OK;
}
int b200ov_graph_launch(void* graph_exec, void* stream) {
  B200OV_CUDA(cudaGraphLaunch(reinterpret_cast<cudaGraphExec_t>(graph_exec), as_stream(stream)));
  return B200OV_OK;
}
int b200ov_graph_destroy(void* graph_exec) {
  B200OV_CUDA(cudaGraphExecDestroy(reinterpret_cast<cudaGraphExec_t>(graph_exec)));
  return B200OV_OK;
}

}  // extern "C"
