// tcgen05 / TMEM / TMA implicit-GEMM convolution + MatMul for sm_100a, FP32-accurate via 3xTF32.
//
// GEMM view (reference: Convolution.py:57-87):  D[M pixels][N cout] = A[M][K] * W[N][K]^T,
// K ordered (ky, kx, ci) with ci padded to 32 per tap.  One CTA computes a 128 x BLOCK_N tile:
//
//   warps 0-3  producers : each thread owns one output pixel (one A row).  Per K block it copies the
//                          pixel's 32-channel run of the current tap straight from the NHWC feature map
//                          into shared memory with 16-byte cp.async (zero-fill = the reference's np.pad),
//                          writing the 128B-swizzled K-major layout UMMA expects.  Thread 0 also issues
//                          the TMA loads of the pre-split weight tiles (hi and lo planes).
//   warps 4-7  converters: split the raw FP32 A tile into tf32 hi / lo parts (a = hi + lo + O(2^-22 a)),
//                          in place + a second buffer, then fence to the async proxy.  After the K loop
//                          the same warps run the epilogue: tcgen05.ld the FP32 accumulator from TMEM,
//                          add bias, apply ReLU / Clamp, store NHWC (optionally into a Concat slice).
//   warp 8     MMA issuer: one elected thread issues tcgen05.mma kind::tf32 (M=128, N=BLOCK_N, K=8):
//                          lo*hi + hi*lo + hi*hi per K step, accumulating in TMEM; tcgen05.commit frees
//                          the stage / signals the epilogue.
//
// Accuracy: the dropped lo*lo term and the second rounding are ~2^-21 relative per product.  The tensor
// core truncates (does not round) when it adds into the FP32 accumulator, which biases long K sums by
// ~0.5 ulp(acc) per MMA; two measures keep that below the FP32 tolerance (measured: 1.4e-5 -> ~2e-6):
//   * the two small cross terms (lo*hi, hi*lo) accumulate into their own TMEM accumulator, so only the
//     hi*hi MMAs touch the large one;
//   * the large accumulator is drained every CHUNK K blocks (ping-pong between two TMEM buffers) into
//     per-thread FP32 registers with round-to-nearest adds ("promotion"), so truncation never
//     compounds over more than 8 MMAs.
// Tests hold the result to |d| <= 1e-5 + 1e-4|ref| against the oracle.
#include <cuda.h>

#include "common.cuh"

namespace b200ov {

namespace tc {

#ifdef B200OV_TC_TRACE
// developer-only pipeline trace (build with B200OV_EXTRA_NVCC_FLAGS=-DB200OV_TC_TRACE): clock64 per role / K block
__device__ unsigned long long g_trace[8][128];
#define TC_TRACE(ev, idx) do { if (blockIdx.x == gridDim.x / 2 && (idx) < 128) g_trace[ev][idx] = clock64(); } while (0)
#else
#define TC_TRACE(ev, idx) do { } while (0)
#endif

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 32;                         // 32 floats = one 128-byte swizzle row
constexpr int A_TILE_BYTES = BLOCK_M * BLOCK_K * 4;  // 16 KB
constexpr int NUM_PRODUCERS = 128;
constexpr int NUM_CONVERTERS = 128;
constexpr int NUM_THREADS = NUM_PRODUCERS + NUM_CONVERTERS + 32;
constexpr int CHUNK = 2;                             // K blocks per promotion chunk (64 K elements, 8 hi*hi MMAs)

struct Params {
  int n, h, w, cin, cout, kh, kw, sh, sw, pt, pl, oh, ow, x_ld, y_ld;
  int M, ohow, cin_blocks, num_k_blocks, tiles_n;
  int act;
  float lo, hi;
};

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a protocol bug must abort the kernel, not hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  long long t0 = clock64();
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ uint32_t to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
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
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128B-swizzled operand tile (rows of 128 bytes, 8-row groups of 1024 bytes).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);        // start address
  d |= (uint64_t)1 << 16;                            // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset: next 8-row group
  d |= (uint64_t)1 << 46;                            // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
  return d;
}

template <int BLOCK_N>
__device__ __forceinline__ constexpr uint32_t instr_desc() {
  return (1u << 4)                       // D format  : F32
         | (2u << 7)                     // A format  : TF32
         | (2u << 10)                    // B format  : TF32
         | ((uint32_t)(BLOCK_N >> 3) << 17)
         | ((uint32_t)(BLOCK_M >> 4) << 24);   // A, B K-major; no negate
}

template <int BLOCK_N, int STAGES>
struct Smem {
  static constexpr int B_TILE_BYTES = BLOCK_N * BLOCK_K * 4;
  static constexpr int A_HI = 0;
  static constexpr int A_LO = A_HI + STAGES * A_TILE_BYTES;
  static constexpr int B_HI = A_LO + STAGES * A_TILE_BYTES;
  static constexpr int B_LO = B_HI + STAGES * B_TILE_BYTES;
  // barriers: a_full, b_full, conv_done, empty [STAGES]; chunk_done[2]; acc_free[2]
  static constexpr int BARS = B_LO + STAGES * B_TILE_BYTES;
  static constexpr int TMEM_PTR = BARS + (4 * STAGES + 4) * 8;
  static constexpr int TOTAL = TMEM_PTR + 16 + 1024;            // + slack for the 1024-byte alignment of the base
};

// X3 = true: 3xTF32 (FP32-accurate).  X3 = false: single-pass TF32 (operands rounded to tf32).
template <int BLOCK_N, int STAGES, bool X3>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tcgen05_kernel(const Params p, const float* __restrict__ x, const float* __restrict__ bias, float* __restrict__ y,
                    const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo) {
  using L = Smem<BLOCK_N, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  auto bar_a_full = [&](int s) { return base + L::BARS + 8 * s; };
  auto bar_b_full = [&](int s) { return base + L::BARS + 8 * (STAGES + s); };
  auto bar_conv = [&](int s) { return base + L::BARS + 8 * (2 * STAGES + s); };
  auto bar_empty = [&](int s) { return base + L::BARS + 8 * (3 * STAGES + s); };
  auto bar_chunk_done = [&](int i) { return base + L::BARS + 8 * (4 * STAGES + i); };
  auto bar_acc_free = [&](int i) { return base + L::BARS + 8 * (4 * STAGES + 2 + i); };
  // TMEM columns: [0, N) and [N, 2N) ping-pong hi*hi accumulators, [2N, 3N) cross-term accumulator
  constexpr uint32_t TMEM_COLS = (3 * BLOCK_N <= 128) ? 128 : (3 * BLOCK_N <= 256 ? 256 : 512);
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(base_ptr + L::TMEM_PTR);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int m_blk = blockIdx.x / p.tiles_n, n_blk = blockIdx.x % p.tiles_n;
  const int m0 = m_blk * BLOCK_M, n0 = n_blk * BLOCK_N;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_a_full(s), NUM_PRODUCERS);
      mbar_init(bar_b_full(s), 1);
      mbar_init(bar_conv(s), NUM_CONVERTERS);
      mbar_init(bar_empty(s), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_chunk_done(i), 1);
      mbar_init(bar_acc_free(i), NUM_CONVERTERS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_lo) : "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(base + L::TMEM_PTR),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) TC_TRACE(7, 0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_ptr_smem;
  if (tid == 0) TC_TRACE(7, 1);

  if (warp < 4) {
    // ================= producers: im2col gather of A (cp.async) + TMA of the weight tiles =============
    // Thread t copies 16-byte chunk (t & 7) of rows (t >> 3) + 16*i, i = 0..7: eight neighbouring threads
    // fetch one pixel's whole 128-byte channel run, so every LDGSTS warp instruction moves 4 full lines
    // (one shared-memory wavefront per line instead of one per thread).
    const int chunk = tid & 7;
    int iy0[8], ix0[8];
    long long img_off[8];
    uint32_t dst_off[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = (tid >> 3) + 16 * i;
      const int m = m0 + row;
      const bool row_ok = m < p.M;
      const int mm = row_ok ? m : 0;
      const int img = mm / p.ohow;
      const int r = mm - img * p.ohow;
      const int oy = r / p.ow, ox = r - oy * p.ow;
      iy0[i] = row_ok ? oy * p.sh - p.pt : -(1 << 28);       // an invalid row never passes the bounds test
      ix0[i] = ox * p.sw - p.pl;
      img_off[i] = (long long)img * p.h * p.w * p.x_ld;
      dst_off[i] = row * 128 + ((chunk ^ (row & 7)) << 4);
    }
    int kb = 0;
    const int taps = p.kh * p.kw;
    for (int tap = 0; tap < taps; ++tap) {
      // per tap: source pointer and validity of each of this thread's 8 rows (hoisted out of the channel loop)
      const int ky = tap / p.kw, kx = tap - ky * p.kw;
      const float* src[8];
      uint32_t valid = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int iy = iy0[i] + ky, ix = ix0[i] + kx;
        const bool v = iy >= 0 && iy < p.h && ix >= 0 && ix < p.w;
        valid |= (v ? 1u : 0u) << i;
        src[i] = x + (v ? img_off[i] + ((long long)iy * p.w + ix) * p.x_ld + chunk * 4 : 0);
      }
      for (int cb = 0; cb < p.cin_blocks; ++cb, ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(bar_empty(s), ph ^ 1);
        if (tid == 0) TC_TRACE(0, kb);
        const uint32_t ok_mask = (cb * BLOCK_K + chunk * 4 < p.cin) ? valid : 0u;
        const uint32_t dst = base + L::A_HI + s * A_TILE_BYTES;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const bool v = (ok_mask >> i) & 1u;
          cp_async_16(dst + dst_off[i], v ? src[i] + cb * BLOCK_K : x, v ? 16u : 0u);
        }
        cp_async_mbar_arrive_noinc(bar_a_full(s));
        if (tid == 0) TC_TRACE(1, kb);
        if (tid == 0) {
          mbar_arrive_expect_tx(bar_b_full(s), X3 ? 2 * L::B_TILE_BYTES : L::B_TILE_BYTES);
          tma_load_2d(base + L::B_HI + s * L::B_TILE_BYTES, &map_hi, kb * BLOCK_K, n0, bar_b_full(s));
          if (X3) tma_load_2d(base + L::B_LO + s * L::B_TILE_BYTES, &map_lo, kb * BLOCK_K, n0, bar_b_full(s));
        }
      }
    }
  } else if (warp < 8) {
    // ================= converters: FP32 -> tf32 hi / lo split of the A tile =============================
    const int c = tid - NUM_PRODUCERS;
    const int ew = warp - 4;                         // TMEM lanes [32*ew, 32*ew + 32)
    const uint32_t tmem_lane = tmem_acc + ((uint32_t)(ew * 32) << 16);
    const int num_chunks = (p.num_k_blocks + CHUNK - 1) / CHUNK;
    float acc[BLOCK_N];
#pragma unroll
    for (int j = 0; j < BLOCK_N; ++j) acc[j] = 0.f;
    // drain one ping-pong accumulator into the FP32 registers (round-to-nearest adds)
    auto promote = [&](int chunk) {
      const int buf = chunk & 1;
      mbar_wait(bar_chunk_done(buf), (chunk >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int q = 0; q < BLOCK_N / 32; ++q) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_lane + buf * BLOCK_N + q * 32, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[q * 32 + j] += __uint_as_float(v[j]);
      }
      tc_fence_before();
      mbar_arrive(bar_acc_free(buf));
    };
    for (int kb = 0; kb < p.num_k_blocks; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      mbar_wait(bar_a_full(s), ph);
      if (c == 0) TC_TRACE(2, kb);
      float4* a_hi = reinterpret_cast<float4*>(base_ptr + L::A_HI + s * A_TILE_BYTES);
      float4* a_lo = reinterpret_cast<float4*>(base_ptr + L::A_LO + s * A_TILE_BYTES);
#pragma unroll
      for (int i = 0; i < A_TILE_BYTES / 16 / NUM_CONVERTERS; ++i) {
        const int idx = i * NUM_CONVERTERS + c;
        float4 v = a_hi[idx];
        if (X3) {
          // hi = the top 19 bits (exact tf32), lo = v - hi (exact in FP32; the tensor core keeps its top 19 bits):
          // v = hi + lo + O(2^-20 |v|), two ALU ops per element
          const float h0 = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u), h1 = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
          const float h2 = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u), h3 = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
          a_hi[idx] = make_float4(h0, h1, h2, h3);
          a_lo[idx] = make_float4(v.x - h0, v.y - h1, v.z - h2, v.w - h3);
        } else {
          uint32_t h0 = to_tf32(v.x), h1 = to_tf32(v.y), h2 = to_tf32(v.z), h3 = to_tf32(v.w);
          a_hi[idx] = make_float4(__uint_as_float(h0), __uint_as_float(h1), __uint_as_float(h2), __uint_as_float(h3));
        }
      }
      fence_proxy_async();
      mbar_arrive(bar_conv(s));
      if (c == 0) TC_TRACE(3, kb);
      // the MMA warp runs at most STAGES blocks behind us: chunk (current - 2) retired long ago
      if (kb % CHUNK == 0 && kb / CHUNK >= 2) promote(kb / CHUNK - 2);
      if (c == 0) TC_TRACE(4, kb);
    }
    // ================= epilogue: remaining chunks + cross terms -> (+bias, act) -> NHWC global ==========
    if (num_chunks >= 2) promote(num_chunks - 2);
    promote(num_chunks - 1);                         // also implies every MMA of this tile has retired
    if (X3) {
#pragma unroll
      for (int q = 0; q < BLOCK_N / 32; ++q) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_lane + 2 * BLOCK_N + q * 32, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[q * 32 + j] += __uint_as_float(v[j]);
      }
    }
    if (c == 0) TC_TRACE(7, 2);
    const int m = m0 + ew * 32 + lane;
    const bool vec_ok = ((p.y_ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(y) & 15u) == 0);
#pragma unroll
    for (int q = 0; q < BLOCK_N / 32; ++q) {
      const int nb = n0 + q * 32;
      if (m < p.M && nb < p.cout) {
        float* yrow = y + (long long)m * p.y_ld + nb;
        if (vec_ok && nb + 32 <= p.cout) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 b4 = bias ? __ldg(reinterpret_cast<const float4*>(bias + nb + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 o;
            o.x = apply_act(acc[q * 32 + j] + b4.x, p.act, p.lo, p.hi);
            o.y = apply_act(acc[q * 32 + j + 1] + b4.y, p.act, p.lo, p.hi);
            o.z = apply_act(acc[q * 32 + j + 2] + b4.z, p.act, p.lo, p.hi);
            o.w = apply_act(acc[q * 32 + j + 3] + b4.w, p.act, p.lo, p.hi);
            *reinterpret_cast<float4*>(yrow + j) = o;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (nb + j < p.cout) {
              float bj = bias ? __ldg(bias + nb + j) : 0.f;
              yrow[j] = apply_act(acc[q * 32 + j] + bj, p.act, p.lo, p.hi);
            }
          }
        }
      }
    }
  } else {
    // ================= MMA issuer =====================================================================
    constexpr uint32_t idesc = instr_desc<BLOCK_N>();
    const uint32_t tmem_cross = tmem_acc + 2 * BLOCK_N;
    for (int kb = 0; kb < p.num_k_blocks; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      const int chunk = kb / CHUNK, buf = chunk & 1;
      if (lane == 0) TC_TRACE(0, 64 + kb);
      if (kb % CHUNK == 0) mbar_wait(bar_acc_free(buf), ((chunk >> 1) & 1) ^ 1);   // promotion of chunk-2 done
      if (lane == 0) TC_TRACE(1, 64 + kb);
      mbar_wait(bar_conv(s), ph);
      if (lane == 0) TC_TRACE(2, 64 + kb);
      mbar_wait(bar_b_full(s), ph);
      tc_fence_after();
      if (lane == 0) TC_TRACE(5, kb);
      if (lane == 0) {
        const uint64_t a_hi = make_smem_desc(base + L::A_HI + s * A_TILE_BYTES);
        const uint64_t a_lo = make_smem_desc(base + L::A_LO + s * A_TILE_BYTES);
        const uint64_t b_hi = make_smem_desc(base + L::B_HI + s * L::B_TILE_BYTES);
        const uint64_t b_lo = make_smem_desc(base + L::B_LO + s * L::B_TILE_BYTES);
        const uint32_t tmem_main = tmem_acc + buf * BLOCK_N;
#pragma unroll
        for (int k = 0; k < BLOCK_K / 8; ++k) {
          const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);     // 8 tf32 = 32 bytes along K inside the swizzle row
          if (X3) {
            umma_tf32(tmem_cross, a_lo + adv, b_hi + adv, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            umma_tf32(tmem_cross, a_hi + adv, b_lo + adv, idesc, 1u);
          }
          umma_tf32(tmem_main, a_hi + adv, b_hi + adv, idesc, (kb % CHUNK > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(bar_empty(s));                               // frees the stage when these MMAs retire
        if (kb % CHUNK == CHUNK - 1 || kb == p.num_k_blocks - 1) umma_commit(bar_chunk_done(buf));
        TC_TRACE(6, kb);
      }
      __syncwarp();
    }
  }

  if (tid == NUM_PRODUCERS) TC_TRACE(7, 3);
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(TMEM_COLS) : "memory");
  }
}

// OIHW -> [plane hi | plane lo], each [cout padded to 32][kh*kw*cin32] (K-major rows), tf32-rounded.
__global__ void pack_tf32_weights_kernel(const float* __restrict__ w, float* __restrict__ out, int cout, int cin, int kh,
                                         int kw, int coutp, int cin32) {
  const long long kpad = (long long)kh * kw * cin32;
  const long long plane = (long long)coutp * kpad;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < plane;
       idx += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(idx / kpad);
    const long long kp = idx - (long long)n * kpad;
    const int tap = (int)(kp / cin32), c = (int)(kp - (long long)tap * cin32);
    float v = 0.f;
    if (n < cout && c < cin) {
      const int ky = tap / kw, kx = tap - ky * kw;
      v = w[(((long long)n * cin + c) * kh + ky) * kw + kx];
    }
    uint32_t hi, lo;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(v));
    lo = __float_as_uint(v - __uint_as_float(hi));     // exact residual; the tensor core keeps its top 19 bits
    out[idx] = __uint_as_float(hi);
    out[plane + idx] = __uint_as_float(lo);
  }
}

// ---- host side -------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

static int make_weight_map(CUtensorMap* map, const float* plane, long long kpad, int coutp, int block_n) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) return set_error(B200OV_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)kpad, (cuuint64_t)coutp};
  cuuint64_t strides[1] = {(cuuint64_t)kpad * 4};
  cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)block_n};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(plane), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(B200OV_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return B200OV_OK;
}

template <int BLOCK_N, int STAGES, bool X3>
static int launch(const Params& p, const float* x, const float* bias, float* y, const CUtensorMap& mh, const CUtensorMap& ml,
                  cudaStream_t s) {
  using L = Smem<BLOCK_N, STAGES>;
  auto kern = conv_tcgen05_kernel<BLOCK_N, STAGES, X3>;
  static bool configured = false;
  if (!configured) {
    B200OV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    configured = true;
  }
  long long blocks = (long long)ceil_div(p.M, BLOCK_M) * p.tiles_n;
  if (blocks > 0x7fffffffLL) return set_error(B200OV_ERR_INVALID, "conv2d grid too large");
  kern<<<(unsigned)blocks, NUM_THREADS, L::TOTAL, s>>>(p, x, bias, y, mh, ml);
  B200OV_LAUNCH_CHECK("conv_tcgen05_kernel");
  return B200OV_OK;
}

}  // namespace tc

#ifdef B200OV_TC_TRACE
extern "C" int b200ov_debug_trace(unsigned long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, tc::g_trace, sizeof(tc::g_trace)) == cudaSuccess ? 0 : 2;
}
#endif

bool tcgen05_eligible(const b200ov_conv_desc* d, const float* x) {
  return (d->cin % 4 == 0) && (d->x_ld % 4 == 0) && aligned16(x) && d->cin >= 8;
}

void tf32_weight_dims(int cout, int cin, int kh, int kw, int* coutp, long long* kpad) {
  *coutp = round_up(cout, 32);
  *kpad = (long long)kh * kw * round_up(cin, tc::BLOCK_K);
}

int pack_tf32_weights(const float* w_oihw, float* out, int cout, int cin, int kh, int kw, cudaStream_t s) {
  int coutp;
  long long kpad;
  tf32_weight_dims(cout, cin, kh, kw, &coutp, &kpad);
  tc::pack_tf32_weights_kernel<<<bw_grid((long long)coutp * kpad, 256), 256, 0, s>>>(w_oihw, out, cout, cin, kh, kw, coutp,
                                                                                      round_up(cin, tc::BLOCK_K));
  B200OV_LAUNCH_CHECK("pack_tf32_weights_kernel");
  return B200OV_OK;
}

// `wt` points at the tf32 section of the packed weights: [hi plane | lo plane].
int conv2d_tcgen05(const b200ov_conv_desc* d, const float* x, const float* wt, const float* bias, float* y,
                   cudaStream_t s, bool probe_only) {
  if (!tcgen05_eligible(d, x))
    return set_error(B200OV_ERR_UNSUPPORTED, "tcgen05 path needs cin %% 4 == 0, cin >= 8 and a 16-byte aligned NHWC input");
  if (probe_only) return B200OV_OK;
  tc::Params p;
  p.n = d->n; p.h = d->h; p.w = d->w; p.cin = d->cin; p.cout = d->cout; p.kh = d->kh; p.kw = d->kw; p.sh = d->sh;
  p.sw = d->sw; p.pt = d->pt; p.pl = d->pl; p.oh = d->oh; p.ow = d->ow; p.x_ld = d->x_ld; p.y_ld = d->y_ld;
  p.ohow = d->oh * d->ow;
  long long M = (long long)d->n * p.ohow;
  if (M > 0x7fffffffLL) return set_error(B200OV_ERR_INVALID, "conv2d: too many output pixels");
  p.M = (int)M;
  if (p.M == 0) return B200OV_OK;
  p.cin_blocks = ceil_div(d->cin, tc::BLOCK_K);
  p.num_k_blocks = d->kh * d->kw * p.cin_blocks;
  p.act = d->act; p.lo = d->act_lo; p.hi = d->act_hi;
  int coutp;
  long long kpad;
  tf32_weight_dims(d->cout, d->cin, d->kh, d->kw, &coutp, &kpad);
  const float* hi_plane = wt;
  const float* lo_plane = wt + (long long)coutp * kpad;
  const bool x3 = d->math != B200OV_MATH_TF32;
  const int block_n = d->cout > 64 ? 128 : (d->cout > 32 ? 64 : 32);
  p.tiles_n = ceil_div(d->cout, block_n);
  CUtensorMap mh, ml;
  int rc = tc::make_weight_map(&mh, hi_plane, kpad, coutp, block_n);
  if (rc) return rc;
  rc = tc::make_weight_map(&ml, lo_plane, kpad, coutp, block_n);
  if (rc) return rc;
  if (block_n == 128) return x3 ? tc::launch<128, 3, true>(p, x, bias, y, mh, ml, s) : tc::launch<128, 3, false>(p, x, bias, y, mh, ml, s);
  if (block_n == 64) return x3 ? tc::launch<64, 4, true>(p, x, bias, y, mh, ml, s) : tc::launch<64, 4, false>(p, x, bias, y, mh, ml, s);
  return x3 ? tc::launch<32, 4, true>(p, x, bias, y, mh, ml, s) : tc::launch<32, 4, false>(p, x, bias, y, mh, ml, s);
}

}  // namespace b200ov
