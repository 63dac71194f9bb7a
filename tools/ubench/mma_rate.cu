// Micro-benchmark: tcgen05.mma kind::f16 M=128 N=128 K=16 issue/execute rate, one CTA per SM, one issuing thread.
// mode 0: A in shared memory (SS), one accumulator   mode 1: A in TMEM (TS), one accumulator
// mode 2: TS, three accumulators in the order main, cross, cross (what conv_f16x2_kernel issues)
// mode 3: TS, N = 256 per MMA, one accumulator
// mode 4: as mode 2 plus one tcgen05.commit (to an mbarrier nobody waits on) after every 6 MMAs
// mode 5: as mode 2 plus three commits after every 6 MMAs
// mode 6: TS, N = 64, one accumulator                  mode 7: TS, N = 64, main/cross/cross + 1 commit/6
// mode 8 / 9: N = 128 / 64, bursts of 6 MMAs + commit into an IDLE pipe (wait for the commit after every burst):
//             column 1 = clk to issue one burst, column 2 = clk per burst including its completion
// mode 10: as mode 4, plus a volatile shared-memory load after the commits of every slot (the ready-flag poll)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../pyopenvino_b200/csrc/tc_ptx.cuh"
using namespace b200ov::ptx;

__device__ __forceinline__ uint32_t ld_volatile_shared(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.volatile.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}

__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// noise (other warps of the SM working while the MMA warp issues; the MMA stream is mode 4's):
//   1: tcgen05.st 16x256b.x4 into TMEM columns 448..479   2: tcgen05.ld 32x32b.x32 of accumulator columns
//   3: 128-bit shared-memory stores                      4: 256-bit global loads (L1-resident 64 KB window per warp)
__global__ void __launch_bounds__(256, 1) mma_rate_kernel(int mode, int iters, unsigned long long* out, int fill, int noise, const float* gbuf) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint32_t tmem_ptr;
  __shared__ uint64_t bar;
  __shared__ uint64_t dummy[4];
  __shared__ volatile uint32_t stop_flag;
  if (threadIdx.x == 0) stop_flag = 0;
  const int warp = threadIdx.x >> 5;
  // zero the operand tiles (A: 128 rows x 128 B, B: 256 rows x 128 B)
  // operand data: zeros, or (fill != 0) pseudo-random finite halfs in [-2, 2) -- the tensor pipe is power-managed, and
  // its sustained rate depends on how many bits toggle
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 256) {
    uint32_t v = 0;
    if (fill) {
      uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
      h ^= h >> 13; h *= 0x5bd1e995u; h ^= h >> 15;
      v = (h & 0xbfffbfffu) & ~0x40004000u;          // clear the top exponent bit of both halfs: |x| < 2
    }
    reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = v;
  }
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&dummy[i]), 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_ptr), 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_ptr;
  if (fill) {                                      // A operand ring in TMEM (columns 384..447): random halfs as well
    uint32_t v[16];
    for (int j = 0; j < 16; ++j) {
      uint32_t h = (threadIdx.x * 16 + j) * 2246822519u + blockIdx.x * 977u;
      h ^= h >> 13; h *= 0x5bd1e995u; h ^= h >> 15;
      v[j] = h & 0xbfffbfffu & ~0x40004000u;
    }
    for (int c = 0; c < 2; ++c)
      for (int g = 0; g < 2; ++g) tmem_st_16x256b_x4(tmem + ((uint32_t)(32 * warp + 16 * g) << 16) + 384 + 32 * c, v);
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (warp >= 4 && noise) {
    const int q = warp & 3;
    uint32_t v[32];
    for (int j = 0; j < 32; ++j) v[j] = threadIdx.x * 33 + j;
    float acc = 0.f;
    uint32_t n = 0;
    while (stop_flag == 0) {
      if (noise == 1) {
        uint32_t w[16];
        for (int j = 0; j < 16; ++j) w[j] = v[j] + n;
        tmem_st_16x256b_x4(tmem + ((uint32_t)(32 * q) << 16) + 448, w);
        tmem_st_16x256b_x4(tmem + ((uint32_t)(32 * q + 16) << 16) + 448, w);
        tmem_st_wait();
      } else if (noise == 2) {
        tmem_ld_32x32b_x32(tmem + ((uint32_t)(32 * q) << 16) + (n & 3) * 32, v);
      } else if (noise == 3) {
        uint8_t* sp = smem_raw + (base - smem_u32(smem_raw)) + 49152 + (threadIdx.x - 128) * 16;
        for (int j = 0; j < 8; ++j) asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(sp + j * 2048)), "r"(v[0] + n), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
      } else {
        const float* gp = gbuf + (size_t)blockIdx.x * 65536 + (size_t)(warp - 4) * 16384 + (threadIdx.x & 31) * 8;
        for (int j = 0; j < 8; ++j) {
          float a0, a1, a2, a3, a4, a5, a6, a7;
          asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];" : "=f"(a0), "=f"(a1), "=f"(a2), "=f"(a3), "=f"(a4), "=f"(a5), "=f"(a6), "=f"(a7) : "l"(gp + ((n * 8 + j) & 63) * 256));
          acc += a0 + a7;
        }
      }
      ++n;
    }
    if (acc == 123.f || v[5] == 0x7fffffffu) out[998] = n;
  }
  if (warp == 1 && elect_one_sync()) {
    const uint64_t a_desc = make_smem_desc_sw128(base), b_desc = make_smem_desc_sw128(base + 16384);
    const int n = mode == 3 ? 256 : (mode == 6 || mode == 7 || mode == 9) ? 64 : mode == 14 ? 96 : mode == 15 ? 192 : mode == 16 ? 32 : 128;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
    const long long t0 = clock64();
    long long issue_clk = 0;
    uint32_t sink = 0;
    for (int it = 0; it < iters; ++it) {
      if (mode == 8 || mode == 9) {
        const long long a0 = clock64();
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          umma_f16_ts(tmem, tmem + 384 + 8 * k, b_desc, idesc, 1u);
          umma_f16_ts(tmem + 256, tmem + 400 + 8 * k, b_desc, idesc, 1u);
          umma_f16_ts(tmem + 256, tmem + 384 + 8 * k, b_desc + 2, idesc, 1u);
        }
        umma_commit(smem_u32(&dummy[0]));
        issue_clk += clock64() - a0;
        mbar_wait(smem_u32(&dummy[0]), it & 1);
        continue;
      }
      if (mode == 11 || mode == 12 || mode == 13) {
        // 11: 6 MMAs, commit, then ~200 cycles of dependent ALU work (the next slot's bookkeeping)
        // 12: 6 MMAs, the ALU work, then the commit          13: ALU work only between bursts, no commit
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          umma_f16_ts(tmem, tmem + 384 + 8 * k, b_desc, idesc, 1u);
          umma_f16_ts(tmem + 256, tmem + 400 + 8 * k, b_desc, idesc, 1u);
          umma_f16_ts(tmem + 256, tmem + 384 + 8 * k, b_desc + 2, idesc, 1u);
        }
        if (mode == 11) umma_commit(smem_u32(&dummy[0]));
#pragma unroll 1
        for (int j = 0; j < 7; ++j) sink = sink * 1664525u + 1013904223u + (sink >> 7);      // dependent chain, ~5 clk per step
        if (mode == 12) umma_commit(smem_u32(&dummy[0]));
        continue;
      }
      if (mode == 0) {
#pragma unroll
        for (int k = 0; k < 6; ++k) umma_f16_ss(tmem, a_desc, b_desc, idesc, 1u);
      } else if (mode == 1) {
#pragma unroll
        for (int k = 0; k < 6; ++k) umma_f16_ts(tmem, tmem + 384, b_desc, idesc, 1u);
      } else if (mode == 2 || mode == 4 || mode == 5 || mode == 7 || mode == 10) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          umma_f16_ts(tmem, tmem + 384 + 8 * k, b_desc, idesc, 1u);
          umma_f16_ts(tmem + 256, tmem + 400 + 8 * k, b_desc, idesc, 1u);
          umma_f16_ts(tmem + 256, tmem + 384 + 8 * k, b_desc + 2, idesc, 1u);
        }
        if (mode >= 4) umma_commit(smem_u32(&dummy[0]));
        if (mode == 10) sink += ld_volatile_shared(smem_u32(&tmem_ptr));
        if (mode == 5) { umma_commit(smem_u32(&dummy[1])); umma_commit(smem_u32(&dummy[2])); }
      } else {
#pragma unroll
        for (int k = 0; k < 6; ++k) umma_f16_ts(tmem, tmem + 384, b_desc, idesc, 1u);      // modes 3, 6, 14, 15, 16: one accumulator, N as above
      }
    }
    long long t1 = clock64();               // all issued (back-pressured by the queue)
    if (mode == 8 || mode == 9) t1 = t0 + issue_clk * 6;
    if (sink == 0x12345678u) out[999] = sink;
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    const long long t2 = clock64();               // all executed
    stop_flag = 1;
    out[2 * blockIdx.x] = (unsigned long long)(t1 - t0);
    out[2 * blockIdx.x + 1] = (unsigned long long)(t2 - t0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  unsigned long long* out;
  cudaMalloc(&out, 1000 * 16);
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 32768 + 2048);
  const int iters = 2000;
  const char* names[17] = {"SS  N=128, one accumulator", "TS  N=128, one accumulator", "TS  N=128, main/cross/cross", "TS  N=256, one accumulator", "TS  main/cross/cross + 1 commit/6", "TS  main/cross/cross + 3 commits/6",
                           "TS  N=64, one accumulator", "TS  N=64 main/cross/cross + 1 commit/6", "N=128 burst of 6 into idle pipe (x6)", "N=64 burst of 6 into idle pipe (x6)",
                           "TS  m/c/c + commit + shared load", "m/c/c, commit, ~200 clk ALU chain", "m/c/c, ~200 clk ALU chain, commit", "m/c/c, ~200 clk ALU chain, no commit", "TS  N=96, one accumulator", "TS  N=192, one accumulator", "TS  N=32, one accumulator"};
  float* gbuf;
  cudaMalloc(&gbuf, 148 * 65536 * sizeof(float));
  cudaMemset(gbuf, 0, 148 * 65536 * sizeof(float));
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 32768 + 16384 + 2048);
  const char* noise_names[5] = {"", "tcgen05.st", "tcgen05.ld", "STS.128", "LDG.256"};
  for (int noise = 1; noise < 5; ++noise) {
    mma_rate_kernel<<<148, 256, 16384 + 32768 + 16384 + 2048>>>(4, iters * 5, out, 1, noise, gbuf);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("noise %d: %s\n", noise, cudaGetErrorString(e)); return 1; }
    unsigned long long h[2];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("TS m/c/c + commit, 4 warps of %-10s noise: %7.1f clk per MMA executed\n", noise_names[noise], (double)h[1] / (6.0 * iters * 5));
  }
  for (int fill = 0; fill < 2; ++fill)
  for (int mode = 0; mode < 17; ++mode) {
    if (mode == 0 && fill) printf("---- random operand data ----\n");
    mma_rate_kernel<<<148, 256, 16384 + 32768 + 16384 + 2048>>>(mode, iters * (fill ? 10 : 1), out, fill, 0, gbuf);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
    unsigned long long h[2];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-40s: %7.1f clk per MMA issued, %7.1f clk per MMA executed\n", names[mode], (double)h[0] / (6.0 * iters * (fill ? 10 : 1)), (double)h[1] / (6.0 * iters * (fill ? 10 : 1)));
  }
  return 0;
}
