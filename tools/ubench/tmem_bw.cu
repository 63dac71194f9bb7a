// Micro-benchmark: TMEM load / store throughput per SM (tcgen05.ld / tcgen05.st), 4 warps, one CTA per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu ; run on a B200.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_16x256b_x8(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_16x256b_x4(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}

// mode 0: 32x32b.x32 loads (4 per iteration = 128 columns), mode 1: 16x256b.x8 loads (2 lane groups x 2 = 64 columns ... ),
// mode 2: 16x256b.x4 stores (32 columns x 16 lanes per instruction)
__global__ void __launch_bounds__(128, 1) tmem_bw_kernel(int mode, int iters, unsigned long long* out, uint32_t* sink) {
  __shared__ uint32_t tmem_base_smem;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base_smem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tmem_base_smem + ((uint32_t)(warp * 32) << 16);
  uint32_t acc = 0;
  // initialise TMEM so loads return defined data
  {
    uint32_t z[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) z[i] = threadIdx.x + i;
    for (int c = 0; c < 512; c += 32) {
      tmem_st_16x256b_x4(base + c, z);
      tmem_st_16x256b_x4(base + (16u << 16) + c, z);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  long long t0 = clock64();
  if (mode == 0) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(base + q * 32, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) acc += v[j];
      }
    }
  } else if (mode == 1) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t v[32];
        tmem_ld_16x256b_x8(base + ((uint32_t)((q & 1) * 16) << 16) + (q >> 1) * 64, v);   // 16 lanes x 64 columns
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) acc += v[j];
      }
    }
  } else if (mode == 3) {   // loads without waiting between them (4 in flight)
    for (int it = 0; it < iters; ++it) {
      uint32_t v[4][32];
#pragma unroll
      for (int q = 0; q < 4; ++q) tmem_ld_32x32b_x32(base + q * 32, v[q]);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int j = 0; j < 32; ++j) acc += v[q][j];
    }
  } else {
    uint32_t z[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) z[i] = threadIdx.x * 3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        tmem_st_16x256b_x4(base + ((uint32_t)((q & 1) * 16) << 16) + (q >> 1) * 32, z);   // 16 lanes x 32 columns
        z[q] += it;
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
  sink[blockIdx.x * 128 + threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_smem), "r"(512u) : "memory");
}

int main() {
  unsigned long long* out;
  uint32_t* sink;
  cudaMalloc(&out, 148 * 8);
  cudaMalloc(&sink, 148 * 128 * 4);
  const int iters = 2000;
  const char* names[4] = {"ld 32x32b.x32 (wait each)", "ld 16x256b.x8 (wait each)", "st 16x256b.x4", "ld 32x32b.x32 (4 in flight)"};
  const double bytes_per_iter[4] = {128.0 * 128 * 4, 4.0 * 16 * 64 * 4 * 4, 4.0 * 16 * 32 * 4 * 4, 128.0 * 128 * 4};
  for (int mode = 0; mode < 4; ++mode) {
    tmem_bw_kernel<<<148, 128>>>(mode, iters, out, sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
    unsigned long long h[148];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-30s: %8.1f clk/iter, %7.1f B/clk/SM\n", names[mode], (double)h[0] / iters, bytes_per_iter[mode] * iters / (double)h[0]);
  }
  return 0;
}
