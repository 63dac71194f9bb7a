// Persistent tcgen05 / TMEM / TMA implicit-GEMM convolution + MatMul for sm_100a, FP32-accurate through a
// two-term FP16 split ("f16x2"):  a = a_hi + 2^-11 a_lo',  w = w_hi + 2^-11 w_lo'  (each part an FP16 number,
// 11 + 11 significant bits), and
//
//     sum a*w  =  sum a_hi*w_hi  +  2^-11 * sum (a_lo'*w_hi + a_hi*w_lo')   + O(2^-22 |a||w|)
//
// i.e. three kind::f16 MMAs per product: the same accuracy class as 3xTF32 at twice the tensor-pipe rate and
// half the operand bytes.  FP16 has a 5-bit exponent, so inputs beyond +-65504 overflow: the epilogue raises a
// sticky status flag when it sees a non-finite output and the engine re-runs the inference on the 3xTF32 path
// (gemm_tcgen05.cu).  Values below the FP16 normal range are still covered: the scaled residual keeps the
// absolute error below 2^-36.
//
// GEMM view (reference: Convolution.py:57-87):  D[M pixels][N cout] = A[M][K] * W[N][K]^T.  K is cut into
// units of 8 channels of one filter tap (cin padded to 8), 4 units = one 32-wide "slot".
//
// Warp roles (512 threads = 4 warpgroups with setmaxnreg budgets, one persistent CTA per SM, static tile schedule,
// 128 x BLOCK_N output tiles, BLOCK_N = 32 / 64 / 96 / 128):
//   warp 12     B loader : TMA loads of the pre-split FP16 weight tiles (hi / lo planes, 64 K-elements per
//                          stage, 128B swizzle) into a shared-memory ring.
//   warp 13     MMA      : issues tcgen05.mma.kind::f16 with the A operand in TENSOR MEMORY and B in shared memory;
//                          tcgen05.commit releases A slots / B stages / accumulators.  The whole warp runs the loop
//                          converged (election inside the asm blocks, uniform datapath) and waits on NAMED BARRIERS
//                          only: a shared-memory operation of this warp queues behind the producers' gathers.
//   warp 14     gate     : polls the mbarriers the MMA warp depends on (weight stage landed; accumulators drained,
//                          when the epilogue does not signal the MMA warp directly) and forwards them as arrivals
//                          on the stage's named barrier.
//   warp 15     pixel-tile loader (POOL / staged variants only, idle otherwise): one TMA box of the input map per
//                          slot into a shared-memory ring the A producers read instead of gathering from global
//                          memory (MaxPool 3x3/s1 folded into a 1x1 convolution; streamed 1x1 inputs).
//   warps 0-7   A        : im2col gather straight from the NHWC feature map into registers (256-bit loads,
//                          4 threads cover one pixel's 32-channel run), FP32 -> FP16 hi/lo split (F2FP, FMUL2,
//                          mixed-precision FHFMA), and tcgen05.st into a TMEM ring of 4 or 8 slots.  Two sets of
//                          four warps work on alternating pairs of slots, running ahead across tile boundaries.
//                          The A operand never touches shared memory: the MMA reads of B alone already use ~60%
//                          of the 128 B/clk shared-memory port.
//   warps 8-11  epilogue : drain the hi*hi accumulator every 256 K-elements into FP32 registers ("promotion",
//                          see below), add the cross terms, bias, activation, stage the tile in shared memory
//                          and write it with TMA stores (coalesced, asynchronous), overlapping the next tile.
// Host-side re-descriptions on top of the same kernel: C_in <= 4 stems as 8-channel convolutions over super-pixels, with
// `group` horizontally adjacent output pixels per GEMM row (conv2d_f16x2_multi); sibling 1x1 convolutions as one GEMM
// with up to three output tensors; split-K for MatMuls with few tiles.
// DESIGN.md sections 5.1 - 5.7 have the measurements behind each of these choices.
//
// Accuracy engineering (as in gemm_tcgen05.cu): the tensor core truncates when it adds into the FP32
// accumulator.  The hi*hi accumulator therefore ping-pongs between two TMEM buffers and is added into
// registers with round-to-nearest every CHUNK slots; the cross terms (2^-11 smaller) keep their own
// accumulator for the whole tile.  Tests hold the result to |d| <= 1e-5 + 1e-4|ref| against the oracle.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "f16split.cuh"
#include "fastdiv.cuh"
#include "tc_ptx.cuh"

namespace b200ov {

namespace f16 {

using namespace ptx;

constexpr int BLOCK_M = 128;
constexpr int SLOT_K = 32;             // K elements per A slot (4 units of 8 channels)
// TMEM columns: [0, NBUF*N) hi*hi accumulator(s), then NCROSS cross-term accumulators of N columns, then the A ring,
// 32 columns per slot (16 hi + 16 lo), at most 8 slots.
#ifndef B200OV_F16_NBUF128
#define B200OV_F16_NBUF128 2
#endif
constexpr int nbuf(int block_n) { return block_n > 64 ? B200OV_F16_NBUF128 : 2; }
// cross-term accumulators: two (alternating per tile) where TMEM has room, so that the MMA warp starts the next tile
// while the epilogue still reads the previous one; with 128 columns per buffer there is room for one only
constexpr int ncross(int block_n) { return block_n <= 96 ? 2 : 1; }
constexpr int a_col0(int block_n) { return (nbuf(block_n) + ncross(block_n)) * block_n; }
constexpr int a_slots(int block_n) { return (512 - a_col0(block_n)) / 32 < 8 ? (512 - a_col0(block_n)) / 32 : 8; }
constexpr int STAGE_K = 64;            // K elements per B stage (2 slots): one 128-byte swizzle row of halfs
#ifndef B200OV_F16_CHUNK
#define B200OV_F16_CHUNK 8
#endif
constexpr int CHUNK2 = B200OV_F16_CHUNK;              // slots per promotion chunk (256 K elements, 16 hi*hi MMAs; tools/acc_probe.py: max err/tol 0.19 up to K = 6272)
constexpr int chunk_slots(int block_n) { return nbuf(block_n) == 1 ? 8 : CHUNK2; }
constexpr int NUM_SETS = 2;             // producer warp sets
constexpr int SET_THREADS = 128;       // threads that build one A slot
constexpr int NUM_EPILOGUE = 128;
constexpr int NUM_THREADS = 128 + NUM_SETS * SET_THREADS + NUM_EPILOGUE;      // 512: four warpgroups
// Warp numbering: the SM sub-partition arbiter prefers the HIGHEST warp id among eligible warps, so the three control
// warps (they issue little but everything waits on them) sit at the top and the arithmetic-heavy producers at the
// bottom.  With the control warps in warps 0-2 the gate needed 500-900 cycles to forward one barrier and the MMA warp
// ~300 cycles between two slots (timeline trace, tools/f16_trace.py).
constexpr int W_PRODUCER0 = 0;                               // warps 0-7  : producer sets (warpgroups 0, 1)
constexpr int W_EPILOGUE0 = 4 * NUM_SETS;                    // warps 8-11 : epilogue (warpgroup 2)
constexpr int W_TMA = W_EPILOGUE0 + 4, W_MMA = W_TMA + 1, W_GATE = W_TMA + 2;    // warpgroup 3: 12, 13, 14 (15 idle)
// register budget per warpgroup (setmaxnreg): 40 + 2 * 136 + 200 = 512 = 4 * 128 (the launch allocation)
constexpr int REGS_CONTROL = 40, REGS_PRODUCER = 136, REGS_EPILOGUE = 200;
// mbarrier arrivals of a whole warp set (stage / accumulator released): every thread (1, the default), or one elected lane per
// warp behind a __syncwarp (32: -DB200OV_F16_ELECTED_ARRIVE).  32 lanes arriving on one mbarrier are 32 serialised shared-memory
// atomics, but the elected form measured no faster (GoogLeNet 70.4-70.7 k images/s either way, and ncu's shared-memory
// bank-conflict count of the staged kernels is the same with both: tools/gpu_r2s.sh, gpu_r2v.sh / gpu_r2bb.sh), so the
// protocol with the longer clean record stays the default.
#ifdef B200OV_F16_ELECTED_ARRIVE
constexpr int ARRIVE_DIV = 32;
#else
constexpr int ARRIVE_DIV = 1;
#endif
constexpr int EPI_BAR_ID = 1;
constexpr int ID_A0 = 2;                 // named barriers 2 .. 2 + A_SLOTS - 1: A slot written (producer set + MMA warp)
constexpr float LO_SCALE = 2048.f;     // 2^11
constexpr float LO_UNSCALE = 1.f / 2048.f;

struct Params {
  int h, w, cin, cout, sh, sw, pt, pl, x_ld, y_ld;
  int M, num_slots, units, tiles_n, num_tiles;
  int act;
  float lo, hi;
  int tma_store;                       // 1: output written by TMA (every y 16-byte aligned, y_ld % 4 == 0)
  // output segments: fused columns [seg_col0, seg_col0 + seg_cout) go to tensor seg_y (pitch seg_yld); seg_col0 is a
  // multiple of 32.  One segment for an ordinary convolution, up to three for sibling 1x1 convolutions run as one GEMM.
  int nseg;
  int seg_hl;                          // bit i: segment i is written as FP16 (hi, scaled lo) pairs (B200OV_DT_HL) for a contraction to read
  int seg_col0[3], seg_cout[3], seg_yld[3];
  float* seg_y[3];
  int prefetch;                        // 1: L2 prefetch of the set's next slot pair (long channel runs, HBM-bound layers)
  int wide_loads;                      // 1: 256-bit gathers (x 32-byte aligned, x_ld % 8 == 0)
  int pair4;                           // 1: cin <= 4 with a pixel pitch of 4 floats: a unit is two horizontally
                                       //    adjacent filter taps x 4 channels (d_upt then divides by units per filter ROW)
  // split-K (MatMul with few output tiles): tile index = (m_blk * tiles_n + n_blk) * ksplit + ks; split ks covers slots
  // [ks * num_slots, (ks + 1) * num_slots) of the K range (num_slots = slots per split, even) and writes its raw partial
  // sums to rows [ks * ws_rows + m, ...) of the workspace, which a second kernel reduces in a fixed order.
  int ksplit, ws_rows;
  // POOL variant (1x1 convolution of MaxPool3x3/s1/p1(x)): per slot, the 128 tile pixels plus a halo of w + 1 pixels on either
  // side of their linear NHW range (pool_rows = 128 + 2w + 2 rows of 32 channels = 128 bytes) are staged in shared memory
  // by TMA from the [pixels][channels] view of x, in a ring of pool_stages stages of pool_stage_bytes.
  int pool_rows, pool_stage_bytes, pool_stages;
  FastDiv d_pool_stages;
  int bias_n;                          // bias[(column) % bias_n]: the real C_out (pixel-group stems repeat the bias per group member)
  int a_hl;                            // 1: x already holds the FP16 (hi, scaled lo) pairs (network input written by the layout
                                       //    kernel, b200ov_input_to_nhwc_split): the producers only route words, no split
  FastDiv d_ohow, d_ow, d_upt, d_kw, d_tiles_n, d_slots, d_ksplit;
};

// tile index -> (row block, column block, K split)
__device__ __forceinline__ void tile_coord(const Params& p, uint32_t tile, uint32_t& m_blk, uint32_t& n_blk, uint32_t& ks) {
  uint32_t q;
  p.d_ksplit.divmod(tile, q, ks);
  p.d_tiles_n.divmod(q, m_blk, n_blk);
}

__device__ __align__(32) float g_zero_run[8];      // source of the np.pad zeros

#ifdef B200OV_F16_TRACE
// developer-only (build with B200OV_EXTRA_NVCC_FLAGS=-DB200OV_F16_TRACE): cycles CTA 0 spends in each wait, per role
__device__ long long g_f16_trace[4][8];
// per-item timeline of CTA 0 (first 128 items): [event][item], events: 0 loads issued, 1 slot free (a_empty passed),
// 2 published (a_full arrive), 3 gate published ready, 4 MMA issue start, 5 MMA issue end, 6 epilogue got chunk, 7 epilogue released chunk
__device__ long long g_f16_timeline[8][128];
#define F16_STAMP(ev, idx) do { if (blockIdx.x == 0 && (idx) < 128) g_f16_timeline[ev][idx] = clock64(); } while (0)
#define F16_TRACE_DECL long long tr_[8] = {0, 0, 0, 0, 0, 0, 0, 0}; const long long tr_start_ = clock64();
#define F16_WAIT(k, bar, par) do { const long long t0_ = clock64(); mbar_wait(bar, par); tr_[k] += clock64() - t0_; } while (0)
#define F16_TIMED(k, stmt) do { const long long t0_ = clock64(); stmt; tr_[k] += clock64() - t0_; } while (0)
#define F16_TRACE_STORE(role, cond) do { if (blockIdx.x == 0 && (cond)) { tr_[7] = clock64() - tr_start_; for (int i_ = 0; i_ < 8; ++i_) g_f16_trace[role][i_] = tr_[i_]; } } while (0)
#else
#define F16_TRACE_DECL
#define F16_WAIT(k, bar, par) mbar_wait(bar, par)
#define F16_TIMED(k, stmt) stmt
#define F16_TRACE_STORE(role, cond)
#define F16_STAMP(ev, idx)
#endif

template <int BLOCK_N, int SB>
struct Smem {
  static constexpr int B_PLANE_BYTES = BLOCK_N * 128;                  // one stage of one plane: BLOCK_N rows x 64 halfs
  static constexpr int B_HI = 0;
  static constexpr int B_LO = B_HI + SB * B_PLANE_BYTES;
  static constexpr int STG_BLOCKS = BLOCK_N == 96 ? 3 : (BLOCK_N >= 64 ? 2 : 1);   // 32-column blocks staged per round
  static constexpr int STAGING = B_LO + SB * B_PLANE_BYTES;            // 4 warps x STG_BLOCKS blocks x 4 KB
  static constexpr int STAGING_BYTES = 4 * STG_BLOCKS * 4096;
  static constexpr int BIAS = STAGING + STAGING_BYTES;                 // BLOCK_N floats
  static constexpr int BARS = BIAS + BLOCK_N * 4;
  // b_full[SB], b_empty[SB], (unused)[A_SLOTS], a_empty[A_SLOTS], main_full[2], main_empty[2], cross_full[2], cross_empty[2]
  static constexpr int A_SLOTS = a_slots(BLOCK_N);
  static constexpr int NUM_BARS = 2 * SB + 2 * A_SLOTS + 8;
  static constexpr int TMEM_PTR = BARS + NUM_BARS * 8;
  static constexpr int TOTAL = TMEM_PTR + 16 + 1024;                   // + slack for the 1024-byte alignment of the base
  // POOL variant only: full[8] / empty[8] barriers of the pixel-tile ring, then the ring itself (1024-byte aligned stages)
  static constexpr int POOL_BARS = TMEM_PTR + 16;
  static constexpr int POOL_RING = (POOL_BARS + 16 * 8 + 1023) / 1024 * 1024;
  static constexpr int POOL_MAX_STAGES = 8;
};

__device__ __forceinline__ constexpr uint32_t instr_desc(int block_n) {
  return (1u << 4)                                   // D format: F32;  A, B format F16 (0), both K-major, no negate
         | ((uint32_t)(block_n >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
}

typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack_f32x2(float a, float b) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ float2 unpack_f32x2(f32x2 v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 pack_u32x2(uint32_t a, uint32_t b) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// (c0, c1) -> packed FP16 hi pair and scaled-residual lo pair:  c = hi + 2^-11 lo  up to 2^-22 |c|.
// (c - hi) is exact in FP32 and so is its 2^11 scaling, hence fma(hi, -2^11, 2^11 c) is exact too.  The FMA is the
// mixed-precision one (FHFMA: f16 x f16 + f32 -> f32), which reads hi straight from the packed pair: five instructions
// per pair (F2FP, FMUL2, 2 FHFMA, F2FP) -- the producers' instruction count is what bounds this kernel.
__device__ __forceinline__ void split_pair(float c0, float c1, uint32_t& hi, uint32_t& lo) {
#ifdef B200OV_F16_EXP_NOCONVERT
  hi = __float_as_uint(c0); lo = __float_as_uint(c1);
  return;
#endif
  split_pair_f16x2(c0, c1, hi, lo);                  // f16split.cuh (shared with the kernels that WRITE the pair form)
}

struct Run8 {
  float v[8];
};
__device__ __forceinline__ Run8 load_run8(const float* p, bool wide) {
  Run8 r;
#ifdef B200OV_F16_EXP_NOLOAD
  for (int i = 0; i < 8; ++i) r.v[i] = __int_as_float(((int)(size_t)p & 0xffff) + 0x3f800000 + i);
  return r;
#endif
  if (wide) {
    asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
                 : "l"(p));
  } else {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  }
  return r;
}

// A16: the input feature map is stored as FP16 (a == a_hi: no split, 4 MMAs per slot); O16: the output is stored as FP16.
// POOL: 1x1 convolution of MaxPool3x3/s1/p1(x) (b200ov_conv_desc.pre_pool): the producers take the 9-tap max of every
// 8-channel run before the split, so the pooled tensor never exists (MaxPool.py:41-72: the zero padding takes part).
// POOL == 2 ("staged"): a plain 1x1 / stride-1 convolution whose pixel tiles take the same road -- TMA into the shared-memory ring
// (no halo), 128-bit shared loads in the producers -- so that the latency of a feature map streamed from HBM / L2 is hidden by
// the ring (5..7 tiles of 16 KB in flight per SM) instead of by what two producer sets keep in flight in registers.
// OHL: some output segment is written as (hi, lo) pairs (Params::seg_hl) -- a variant of its own so that the plain epilogue
// keeps its register allocation (with 128-column tiles it is at the limit: the pair encoder cost 80 more spill bytes and
// 10 % on K-short layers when it shared the code).
template <int BLOCK_N, int SB, bool WIDE, bool PAIR, bool A16, bool O16, int POOL = 0, bool OHL = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_f16x2_kernel(const Params p, const void* __restrict__ x_raw, const float* __restrict__ bias,
                  unsigned int* __restrict__ status, const __grid_constant__ CUtensorMap map_hi,
                  const __grid_constant__ CUtensorMap map_lo, const __grid_constant__ CUtensorMap map_y0,
                  const __grid_constant__ CUtensorMap map_y1, const __grid_constant__ CUtensorMap map_y2,
                  const __grid_constant__ CUtensorMap map_x) {
  using L = Smem<BLOCK_N, SB>;
  constexpr int A_SLOTS = a_slots(BLOCK_N);
  constexpr int A_COL0 = a_col0(BLOCK_N);
  constexpr int NBUF = nbuf(BLOCK_N);
  constexpr int CHUNK = chunk_slots(BLOCK_N);
  constexpr int NCROSS = ncross(BLOCK_N);
  constexpr int ID_B0 = ID_A0 + A_SLOTS;   // named barriers of the B stages (gate warp + MMA warp)
  // With one cross buffer the epilogue's "accumulator drained" signals are on the critical path of every tile
  // boundary: they go straight to the MMA warp on named barriers (epilogue warps + MMA warp) instead of through
  // an mbarrier and the gate warp.
  constexpr bool DIRECT_EMPTY = NCROSS == 1;
  constexpr int ID_X = ID_B0 + SB, ID_M0 = ID_X + 1;
  static_assert(ID_B0 + SB + (DIRECT_EMPTY ? 1 + NBUF : 0) <= 16, "out of named barriers");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  auto bar_b_full = [&](int s) { return base + L::BARS + 8 * s; };
  auto bar_b_empty = [&](int s) { return base + L::BARS + 8 * (SB + s); };
  auto bar_a_empty = [&](int s) { return base + L::BARS + 8 * (2 * SB + A_SLOTS + s); };
  auto bar_main_full = [&](int i) { return base + L::BARS + 8 * (2 * SB + 2 * A_SLOTS + i); };
  auto bar_main_empty = [&](int i) { return base + L::BARS + 8 * (2 * SB + 2 * A_SLOTS + 2 + i); };
  auto bar_cross_full = [&](int i) { return base + L::BARS + 8 * (2 * SB + 2 * A_SLOTS + 4 + i); };
  auto bar_cross_empty = [&](int i) { return base + L::BARS + 8 * (2 * SB + 2 * A_SLOTS + 6 + i); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(base_ptr + L::TMEM_PTR);
  auto bar_pool_full = [&](int s) { return base + L::POOL_BARS + 8 * s; };
  auto bar_pool_empty = [&](int s) { return base + L::POOL_BARS + 8 * (L::POOL_MAX_STAGES + s); };

  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;      // broadcast: tells ptxas the role branches are warp-uniform

  if (tid == 0) {
    for (int s = 0; s < SB; ++s) {
      mbar_init(bar_b_full(s), 1);
      mbar_init(bar_b_empty(s), 1);
    }
    for (int s = 0; s < A_SLOTS; ++s) {
      mbar_init(bar_a_empty(s), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_main_full(i), 1);
      mbar_init(bar_main_empty(i), NUM_EPILOGUE / ARRIVE_DIV);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_cross_full(i), 1);
      mbar_init(bar_cross_empty(i), NUM_EPILOGUE / ARRIVE_DIV);
    }
    if constexpr (POOL) {
      for (int i = 0; i < p.pool_stages; ++i) {
        mbar_init(bar_pool_full(i), 1);
        mbar_init(bar_pool_empty(i), SET_THREADS / ARRIVE_DIV);
      }
      prefetch_tensormap(&map_x);
    }
    fence_mbar_init();
    prefetch_tensormap(&map_hi);
    prefetch_tensormap(&map_lo);
    prefetch_tensormap(&map_y0);
  }
  if (warp == W_MMA) tmem_alloc(base + L::TMEM_PTR, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch) may run while the
  // previous kernel of the stream drains; nothing before this line reads or writes a tensor.
  B200OV_PDL_SYNC();

  const int my_tiles = (p.num_tiles > (int)blockIdx.x) ? (p.num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int num_stages = (p.num_slots + 1) >> 1;      // B stages per tile

  if (warp >= W_TMA) {
    // control warpgroup: B loader, MMA issuer, gate, one idle warp.  Hand the registers to the others.
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_CONTROL));
    if (warp == W_TMA) {
      // ================= B loader ====================================================================
      {
        F16_TRACE_DECL
        uint32_t bcount = 0;
        for (int tl = 0; tl < my_tiles; ++tl) {
          const uint32_t tile = blockIdx.x + (uint32_t)tl * gridDim.x;
          uint32_t m_blk_, n_blk_, ksp_;
          tile_coord(p, tile, m_blk_, n_blk_, ksp_);
          const int n0 = (int)n_blk_ * BLOCK_N;
          const int stage0 = (int)ksp_ * num_stages;
          for (int ks = stage0; ks < stage0 + num_stages; ++ks, ++bcount) {
            const int s = bcount % SB;
            F16_WAIT(0, bar_b_empty(s), ((bcount / SB) & 1) ^ 1);
            if (elect_one_sync()) {
              mbar_arrive_expect_tx(bar_b_full(s), 2 * L::B_PLANE_BYTES);
              tma_load_2d(base + L::B_HI + s * L::B_PLANE_BYTES, &map_hi, ks * STAGE_K, n0, bar_b_full(s));
              tma_load_2d(base + L::B_LO + s * L::B_PLANE_BYTES, &map_lo, ks * STAGE_K, n0, bar_b_full(s));
            }
            __syncwarp();
          }
        }
        F16_TRACE_STORE(0, lane == 0);
      }
    } else if (warp == W_MMA) {
      // ================= MMA issuer ==================================================================
      constexpr uint32_t idesc = instr_desc(BLOCK_N);
      F16_TRACE_DECL
      // The whole warp runs the loop with warp-uniform bookkeeping (ring positions advance by increments, no division),
      // and one elected lane issues.  What this thread executes per slot is the critical path of the kernel: six MMAs
      // are 384 tensor-pipe cycles, and an earlier version of this loop (per-slot descriptor structs built in a
      // divergent single-lane branch, ~50 dependent ALU ops + R2UR moves + a spilled local) needed ~700 cycles per
      // slot, so the pipe idled half of the time (tools/f16_trace.py).
      const uint64_t desc_hi0 = make_smem_desc_sw128(base + L::B_HI), desc_lo0 = make_smem_desc_sw128(base + L::B_LO);
      constexpr uint32_t STAGE_DESC = (uint32_t)(L::B_PLANE_BYTES >> 4);
      static_assert((A_SLOTS & (A_SLOTS - 1)) == 0, "A ring size must be a power of two");
      // One iteration = one B stage = two slots = 12 MMAs (768 tensor-pipe cycles).  The pipe runs only ~2 MMAs ahead of
      // the issuing thread (tools/ubench/mma_rate.cu: 190 cycles of other work between bursts of 6 cost 78), so what the
      // thread does between two bursts -- barrier waits, descriptor arithmetic, R2UR moves, commits -- is mostly exposed;
      // per pair it is paid once instead of twice.
      // The CTA allocates all 512 TMEM columns, so the allocation starts at lane 0, column 0: with the base a
      // compile-time constant every TMEM address of this loop is warp-uniform arithmetic (no vector -> uniform moves).
      if (tmem_base != 0) __trap();
      uint32_t as = 0, bs = 0, buf = 0, ready = 0, chunks = 0;
#ifdef B200OV_F16_TRACE
      long long t_prev_end_ = 0;
#endif
      for (int tl = 0; tl < my_tiles; ++tl) {
        uint32_t in_chunk = 0;
        const int xb = NCROSS == 2 ? (tl & 1) : 0;
        const uint32_t tmem_cross = (NBUF + xb) * BLOCK_N;
        for (int slot = 0; slot < p.num_slots; slot += 2) {
          const bool two = slot + 1 < p.num_slots;
          const bool last = slot + 2 >= p.num_slots;
          const bool end_chunk = in_chunk == CHUNK - 2 || last;
          const uint32_t as1 = (as + 1) & (A_SLOTS - 1);
          // Waits of this warp are NAMED BARRIERS, not mbarriers or shared-memory flags: every shared-memory operation
          // of this warp queues behind the producers' gathers in the SM's memory pipe and took 200-300 cycles
          // (tools/f16_trace.py), more than the six MMAs of a slot take to issue.  The gate warp turns "B stage landed,
          // accumulators drained" into an arrival on the stage's barrier; the producers arrive on the A slot's barrier.
          // Reuse of a barrier id is safe: nobody can arrive for the next round of a slot / stage before this warp's
          // commit for the current round, which follows its bar.sync.
          F16_TIMED(2, named_bar_sync(ID_B0 + bs, 64));
          if constexpr (DIRECT_EMPTY) {
            if (in_chunk == 0 && chunks >= (uint32_t)NBUF) F16_TIMED(0, named_bar_sync(ID_M0 + buf, NUM_EPILOGUE + 32));   // promotion of chunk - NBUF done
            if (slot == 0 && tl > 0) F16_TIMED(0, named_bar_sync(ID_X, NUM_EPILOGUE + 32));                              // previous tile's cross terms read
          }
          F16_TIMED(3, named_bar_sync(ID_A0 + as, SET_THREADS + 32));
          if (two) F16_TIMED(3, named_bar_sync(ID_A0 + as1, SET_THREADS + 32));
#ifdef B200OV_F16_TRACE
          const long long t_f0_ = clock64();
          if (t_prev_end_ != 0) tr_[6] += t_f0_ - t_prev_end_;       // whole gap since the previous burst
#endif
          tc_fence_after();
#ifdef B200OV_F16_TRACE
          const long long t_issue0_ = clock64();
          tr_[4] += t_issue0_ - t_f0_;
          if (lane == 0) { F16_STAMP(4, ready); F16_STAMP(4, ready + 1); }
#endif
          const uint32_t a0 = A_COL0 + as * 32, a1 = A_COL0 + as1 * 32;       // TMEM base is 0 (checked above)
          const uint32_t d_main = buf * BLOCK_N;
          const uint64_t b_hi = desc_hi0 + (uint64_t)(bs * STAGE_DESC);
          const uint64_t b_lo = desc_lo0 + (uint64_t)(bs * STAGE_DESC);
          if constexpr (A16) {
            umma_f16a_slot(d_main, tmem_cross, a0, b_hi, b_lo, idesc, in_chunk > 0 ? 1u : 0u, slot > 0 ? 1u : 0u, bar_a_empty(as));
            if (two) umma_f16a_slot(d_main, tmem_cross, a1, b_hi + 4, b_lo + 4, idesc, 1u, 1u, bar_a_empty(as1));
          } else {
            umma_f16x2_slot(d_main, tmem_cross, a0, b_hi, b_lo, idesc, in_chunk > 0 ? 1u : 0u, slot > 0 ? 1u : 0u, bar_a_empty(as));
            if (two) umma_f16x2_slot(d_main, tmem_cross, a1, b_hi + 4, b_lo + 4, idesc, 1u, 1u, bar_a_empty(as1));   // +64 B along K
          }
          umma_commit_elect(bar_b_empty(bs), 1u);
          umma_commit_elect(bar_main_full(buf), end_chunk ? 1u : 0u);
          umma_commit_elect(bar_cross_full(xb), last ? 1u : 0u);
          __syncwarp();
#ifdef B200OV_F16_TRACE
          t_prev_end_ = clock64();
          tr_[5] += t_prev_end_ - t_issue0_;
          if (lane == 0) { F16_STAMP(5, ready); F16_STAMP(5, ready + 1); }
#endif
          ready += two ? 2 : 1;
          as = two ? (as + 2) & (A_SLOTS - 1) : as1;
          bs = bs + 1 == SB ? 0 : bs + 1;
          if (end_chunk) { buf = buf + 1 == NBUF ? 0 : buf + 1; in_chunk = 0; ++chunks; } else { in_chunk += 2; }
        }
      }
      __syncwarp();
      F16_TRACE_STORE(1, lane == 0);
    } else if (warp == W_GATE) {
      // ================= gate ========================================================================
      // Polls the mbarriers the MMA warp depends on, one B stage ahead of it, and forwards each stage as an arrival
      // on the stage's named barrier (an mbarrier try_wait in the issuing warp itself cost 200-300 cycles).
      static_assert(CHUNK % 2 == 0, "a B stage (two slots) must not straddle promotion chunks");
      uint32_t bcount = 0, chunkcount = 0;
      for (int tl = 0; tl < my_tiles; ++tl) {
        for (int slot = 0; slot < p.num_slots; slot += 2) {          // one B stage = two slots
          const bool last = slot + 2 >= p.num_slots;
          const uint32_t bs = bcount % SB;
          mbar_wait(bar_b_full(bs), (bcount / SB) & 1);
          if constexpr (!DIRECT_EMPTY) {
            if (slot % CHUNK == 0) mbar_wait(bar_main_empty(chunkcount % NBUF), ((chunkcount / NBUF) & 1) ^ 1);   // promotion of chunk-NBUF done
            if (slot == 0) mbar_wait(bar_cross_empty(tl & 1), ((tl >> 1) & 1) ^ 1);                      // cross terms of tile - 2 read
          }
          __syncwarp();
          named_bar_arrive(ID_B0 + bs, 64);
          if (lane == 0) F16_STAMP(3, (uint32_t)(tl * p.num_slots + slot));
          ++bcount;
          if ((slot + 2) % CHUNK == 0 || last) ++chunkcount;
        }
      }
    } else if constexpr (POOL) {
      // ================= pixel-tile loader (POOL variant; this warp idles otherwise) ===================
      // One TMA box per slot: 32 channels x pool_rows consecutive pixels starting w + 1 pixels before the tile's first
      // pixel (rows outside the tensor arrive as zeros; the producers never use them).
      int st = 0;
      uint32_t phase = 0;
      for (int tl = 0; tl < my_tiles; ++tl) {
        const uint32_t tile = blockIdx.x + (uint32_t)tl * gridDim.x;
        uint32_t m_blk_, n_blk_, ksp_;
        tile_coord(p, tile, m_blk_, n_blk_, ksp_);
        const int r0 = (int)m_blk_ * BLOCK_M - (POOL == 1 ? p.w + 1 : 0);
        for (int slot = 0; slot < p.num_slots; ++slot) {
          mbar_wait(bar_pool_empty(st), phase ^ 1);
          if (elect_one_sync()) {
#ifdef B200OV_POOL_EXP_NOTMA
            mbar_arrive(bar_pool_full(st));
#else
            mbar_arrive_expect_tx(bar_pool_full(st), (uint32_t)p.pool_rows * 128u);
            tma_load_2d(base + L::POOL_RING + st * p.pool_stage_bytes, &map_x, slot * SLOT_K, r0, bar_pool_full(st));
#endif
          }
          __syncwarp();
          if (++st == p.pool_stages) { st = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp < W_EPILOGUE0) {
    // ================= A producers: gather -> split -> TMEM ==========================================
    // Two sets of four warps; a set owns every other PAIR of consecutive slots ("items").  All loads of a
    // warp share one scoreboard slot, so a register ring inside a warp cannot overlap load latency with
    // the split; the overlap comes from the other set (and the other warps of the SM sub-partition)
    // working on the neighbouring pair in the meantime.
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_PRODUCER));
    const int set = (warp - W_PRODUCER0) >> 2;
    const int q = warp & 3;                          // TMEM lane quarter this warp may access
    const int u4 = lane & 3, rsub = lane >> 2;
    const uint32_t total_items = (uint32_t)my_tiles * (uint32_t)p.num_slots;
    // row state of the tile the gather is currently in: rows 32q + 16g + rsub + 8h  (r = 2g + h)
    using TA = typename std::conditional<A16, __half, float>::type;
    const TA* x = reinterpret_cast<const TA*>(x_raw);
    const TA* rbase[4];
    int riy[4], rix[4];
    int prow[4], pflag[4];             // POOL: box row of the window centre; bits: 1 up, 2 down, 4 left, 8 right neighbour inside the image, 16 row < M
    uint32_t cur_tl = 0xffffffffu, cur_slot0 = 0;
    uint32_t pool_release[2] = {0, 0};   // POOL: "stage consumed" barriers of the items between issue_loads and convert_store
    F16_TRACE_DECL
    auto issue_loads = [&](uint32_t item, Run8 (&dst)[4], const int which = 0) {
      uint32_t tl, slot;
      p.d_slots.divmod(item, tl, slot);
      if (tl != cur_tl) {
        cur_tl = tl;
        const uint32_t tile = blockIdx.x + tl * gridDim.x;
        uint32_t m_blk_, n_blk_, ksp_;
        tile_coord(p, tile, m_blk_, n_blk_, ksp_);
        cur_slot0 = ksp_ * (uint32_t)p.num_slots;
        const int m0 = (int)m_blk_ * BLOCK_M;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          // POOL: the thread's four rows are four CONSECUTIVE pixels (tile row 16g + 8h + rsub <-> pixel 4 rsub + 2g + h of the
          // warp's 32), so their 3x3 windows share columns; the epilogue undoes the permutation when it stages the tile
          const int m = POOL == 1 ? m0 + 32 * q + 4 * rsub + r : m0 + 32 * q + 16 * (r >> 1) + rsub + 8 * (r & 1);
          uint32_t img, rem, oy, ox;
          p.d_ohow.divmod((uint32_t)(m < p.M ? m : 0), img, rem);
          p.d_ow.divmod(rem, oy, ox);
          riy[r] = m < p.M ? (int)oy * p.sh - p.pt : -(1 << 28);      // a row past M never passes the bounds test
          rix[r] = (int)ox * p.sw - p.pl;
          rbase[r] = x + ((long long)((int)img * p.h + riy[r]) * p.w + rix[r]) * p.x_ld;
          if constexpr (POOL) {
            prow[r] = (m - m0) + (POOL == 1 ? p.w + 1 : 0);
            pflag[r] = m < p.M ? (16 | (oy > 0 ? 1 : 0) | ((int)oy + 1 < p.h ? 2 : 0) | (ox > 0 ? 4 : 0) | ((int)ox + 1 < p.w ? 8 : 0)) : 0;
          }
        }
      }
      const uint32_t unit = (cur_slot0 + slot) * 4 + u4;
      const bool uvalid = unit < (uint32_t)p.units;
      if constexpr (POOL) {
        // 1x1 convolution (one tap, a slot = 32 channels) of the 3x3 / stride 1 / pad 1 max-pooled map, taken from the slot's
        // pixel tile in shared memory (TMA, 128B swizzle: 16-byte chunk c of box row r sits at chunk c ^ (r & 7)).  A tap
        // outside the image is aliased to the window centre (max is idempotent) and the reference's zero padding enters as
        // one max with 0 for border pixels (MaxPool.py:41-72); rows past M produce zeros.
        uint32_t round, st;
        p.d_pool_stages.divmod(item, round, st);
        mbar_wait(bar_pool_full(st), round & 1);
        const uint32_t sbase = base + L::POOL_RING + st * p.pool_stage_bytes;
        if constexpr (POOL == 2) {
          // staged 1x1: the thread's four rows (the register-gather kernel's row mapping) x 8 channels, two 128-bit loads each.
          // The eight lanes of a shared-memory wavefront are two neighbouring rows x four units: their swizzle terms differ in
          // bit 0, so they cover all eight 16-byte columns (no bank conflict).  Rows past M and channels past C_in are TMA's
          // zero fill.
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int row = prow[r];
            const uint32_t a = sbase + (uint32_t)row * 128u + ((uint32_t)((2 * u4) ^ (row & 7)) << 4);
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(dst[r].v[0]), "=f"(dst[r].v[1]), "=f"(dst[r].v[2]), "=f"(dst[r].v[3]) : "r"(a));
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(dst[r].v[4]), "=f"(dst[r].v[5]), "=f"(dst[r].v[6]), "=f"(dst[r].v[7]) : "r"(a ^ 16u));
          }
          pool_release[which] = bar_pool_empty(st);
          return;
        }
        // The four windows cover 6 pixel columns (linear neighbours p - 1 .. p + 4 of the thread's first pixel p) x 3 image rows:
        // 18 pixel positions instead of 36.  Column maxima first (vertical), then each output takes its three columns.  A
        // vertical tap outside the image is aliased to the centre row; a horizontal neighbour that belongs to another image row
        // (x = 0 / x = w - 1; the linear neighbour is then the previous / next row's pixel) is clamped to <= 0 with a min, which
        // is exact because such a pixel takes the zero padding into its max anyway (MaxPool.py:41-72).  Rows past M sit beyond
        // the tensor map's pixel dimension and read TMA's zero fill.
        float4 cm[6][2];
        const int hswap = rsub & 1;
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          const int f = pflag[c == 0 ? 0 : (c == 5 ? 3 : c - 1)];
          const int rowc = prow[0] + c - 1;
          float4 t[3][2];
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            const int row = rowc + (dy == 0 ? ((f & 1) ? -p.w : 0) : (dy == 1 ? 0 : ((f & 2) ? p.w : 0)));
            // Bank layout: the eight lanes of a 128-bit shared-memory wavefront are two pixel groups (rsub, rsub + 1: box rows 4
            // apart, so their swizzle terms differ by 4) x four units.  With every lane reading its lower 16-byte half first
            // the two groups hit the same four chunk columns (a two-way conflict on each of the 36 loads: ncu, r2n); odd
            // groups therefore read their upper half first and the halves are swapped back when the outputs are formed.
            const uint32_t a = sbase + (uint32_t)row * 128u + ((uint32_t)((2 * u4 + hswap) ^ (row & 7)) << 4);
#ifdef B200OV_POOL_EXP_1TAP
            if (dy != 1) { t[dy][0] = t[dy][1] = make_float4(0.f, 0.f, 0.f, 0.f); continue; }
#endif
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t[dy][0].x), "=f"(t[dy][0].y), "=f"(t[dy][0].z), "=f"(t[dy][0].w) : "r"(a));
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t[dy][1].x), "=f"(t[dy][1].y), "=f"(t[dy][1].z), "=f"(t[dy][1].w) : "r"(a ^ 16u));
          }
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            cm[c][hf].x = fmaxf(fmaxf(t[0][hf].x, t[1][hf].x), t[2][hf].x); cm[c][hf].y = fmaxf(fmaxf(t[0][hf].y, t[1][hf].y), t[2][hf].y);
            cm[c][hf].z = fmaxf(fmaxf(t[0][hf].z, t[1][hf].z), t[2][hf].z); cm[c][hf].w = fmaxf(fmaxf(t[0][hf].w, t[1][hf].w), t[2][hf].w);
          }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int f = pflag[r];
          const float lim_l = (f & 4) ? INFINITY : 0.f, lim_r = (f & 8) ? INFINITY : 0.f, edge = (f & 15) == 15 ? -INFINITY : 0.f;
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
#define B200OV_OUT(F_) fmaxf(fmaxf(fmaxf(fminf(cm[r][hf].F_, lim_l), cm[r + 1][hf].F_), fminf(cm[r + 2][hf].F_, lim_r)), edge)
            dst[r].v[4 * hf + 0] = B200OV_OUT(x); dst[r].v[4 * hf + 1] = B200OV_OUT(y);
            dst[r].v[4 * hf + 2] = B200OV_OUT(z); dst[r].v[4 * hf + 3] = B200OV_OUT(w);
#undef B200OV_OUT
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {            // odd pixel groups loaded channels 4..7 first
            const float lo4 = dst[r].v[i], hi4 = dst[r].v[4 + i];
            dst[r].v[i] = hswap ? hi4 : lo4;
            dst[r].v[4 + i] = hswap ? lo4 : hi4;
          }
        }
        // The stage is released in convert_store, after the tcgen05.st that consume every value loaded here: an arrive placed
        // here is not ordered behind the LDS still queued in the (saturated) shared-memory pipe -- it overtook them and the
        // loader's refill of the stage then raced the last loads of the window (seen as wrong fourth pixels of a group on the
        // first item of a tile, when the loader is already waiting for the stage).
        pool_release[which] = bar_pool_empty(st);
      } else if constexpr (!PAIR) {
        uint32_t tap, cu, ky, kx;
        p.d_upt.divmod(unit, tap, cu);
        p.d_kw.divmod(tap, ky, kx);
        const int off = ((int)ky * p.w + (int)kx) * p.x_ld + (int)cu * 8;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const bool ok = uvalid && (unsigned)(riy[r] + (int)ky) < (unsigned)p.h && (unsigned)(rix[r] + (int)kx) < (unsigned)p.w;
          if constexpr (A16) {
            // 8 channels = 16 bytes of halfs; the four words are the packed (hi) operand pairs as they are
            const uint4 t = __ldg(reinterpret_cast<const uint4*>(ok ? static_cast<const void*>(rbase[r] + off) : static_cast<const void*>(g_zero_run)));
            dst[r].v[0] = __uint_as_float(t.x); dst[r].v[1] = __uint_as_float(t.y); dst[r].v[2] = __uint_as_float(t.z); dst[r].v[3] = __uint_as_float(t.w);
          } else {
            dst[r] = load_run8(ok ? reinterpret_cast<const float*>(rbase[r] + off) : g_zero_run, WIDE);
            // the same rows, 16 units (4 slots) further along the channels of this tap: what this set gathers next
#ifdef B200OV_F16_PREFETCH_BUILD     // developer build only: the predicated-off prefetches still cost 13 issue slots per item
            if (p.prefetch && u4 == 0 && ok && cu + 16 < p.d_upt.d)
              asm volatile("prefetch.global.L2 [%0];" ::"l"(rbase[r] + off + 128));
#endif
          }
        }
      } else {
        // two adjacent taps (ky, 2*kxp) and (ky, 2*kxp + 1): with a pixel pitch of 4 floats they are 8 contiguous
        // floats, but each half has its own bounds test (the image edge can fall between them)
        uint32_t ky, kxp;
        p.d_upt.divmod(unit, ky, kxp);
        const int kx = 2 * (int)kxp;
        const int off = ((int)ky * p.w + kx) * 4;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const bool row_ok = uvalid && (unsigned)(riy[r] + (int)ky) < (unsigned)p.h;
          const bool ok0 = row_ok && (unsigned)(rix[r] + kx) < (unsigned)p.w;
          const bool ok1 = row_ok && (unsigned)(rix[r] + kx + 1) < (unsigned)p.w;
          const float4 a = __ldg(reinterpret_cast<const float4*>(ok0 ? reinterpret_cast<const float*>(rbase[r] + off) : g_zero_run));
          const float4 b = __ldg(reinterpret_cast<const float4*>(ok1 ? reinterpret_cast<const float*>(rbase[r] + off + 4) : g_zero_run));
          dst[r].v[0] = a.x; dst[r].v[1] = a.y; dst[r].v[2] = a.z; dst[r].v[3] = a.w;
          dst[r].v[4] = b.x; dst[r].v[5] = b.y; dst[r].v[6] = b.z; dst[r].v[7] = b.w;
        }
      }
    };
    auto convert_store = [&](uint32_t item, const Run8 (&src)[4], const int which = 0) {
      const int as = item % A_SLOTS;
      F16_WAIT(0, bar_a_empty(as), ((item / A_SLOTS) & 1) ^ 1);
      if (q == 0 && lane == 0) F16_STAMP(1, item);
      tc_fence_after();
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const Run8& r0 = src[2 * g];               // lane rsub
        const Run8& r1 = src[2 * g + 1];           // lane rsub + 8
        if constexpr (A16) {
          uint32_t h[8];
          h[0] = __float_as_uint(r0.v[0]); h[1] = __float_as_uint(r0.v[1]); h[2] = __float_as_uint(r1.v[0]); h[3] = __float_as_uint(r1.v[1]);
          h[4] = __float_as_uint(r0.v[2]); h[5] = __float_as_uint(r0.v[3]); h[6] = __float_as_uint(r1.v[2]); h[7] = __float_as_uint(r1.v[3]);
          tmem_st_16x256b_x2(tmem_base + ((uint32_t)(32 * q + 16 * g) << 16) + A_COL0 + as * 32, h);
          continue;
        }
        uint32_t v[16];
        if (p.a_hl) {
          // per pixel (4 channels, 16 bytes): [hi(c0,c1) hi(c2,c3) lo(c0,c1) lo(c2,c3)]; a run = two pixels
          v[0] = __float_as_uint(r0.v[0]); v[1] = __float_as_uint(r0.v[1]); v[8] = __float_as_uint(r0.v[2]); v[9] = __float_as_uint(r0.v[3]);
          v[4] = __float_as_uint(r0.v[4]); v[5] = __float_as_uint(r0.v[5]); v[12] = __float_as_uint(r0.v[6]); v[13] = __float_as_uint(r0.v[7]);
          v[2] = __float_as_uint(r1.v[0]); v[3] = __float_as_uint(r1.v[1]); v[10] = __float_as_uint(r1.v[2]); v[11] = __float_as_uint(r1.v[3]);
          v[6] = __float_as_uint(r1.v[4]); v[7] = __float_as_uint(r1.v[5]); v[14] = __float_as_uint(r1.v[6]); v[15] = __float_as_uint(r1.v[7]);
          tmem_st_16x256b_x4(tmem_base + ((uint32_t)(32 * q + 16 * g) << 16) + A_COL0 + as * 32, v);
          continue;
        }
        split_pair(r0.v[0], r0.v[1], v[0], v[8]);
        split_pair(r0.v[2], r0.v[3], v[1], v[9]);
        split_pair(r1.v[0], r1.v[1], v[2], v[10]);
        split_pair(r1.v[2], r1.v[3], v[3], v[11]);
        split_pair(r0.v[4], r0.v[5], v[4], v[12]);
        split_pair(r0.v[6], r0.v[7], v[5], v[13]);
        split_pair(r1.v[4], r1.v[5], v[6], v[14]);
        split_pair(r1.v[6], r1.v[7], v[7], v[15]);
        tmem_st_16x256b_x4(tmem_base + ((uint32_t)(32 * q + 16 * g) << 16) + A_COL0 + as * 32, v);
      }
      if constexpr (POOL) {
        // (after the stores: see issue_loads; ARRIVE_DIV above)
        if constexpr (ARRIVE_DIV == 32) __syncwarp();
        if (ARRIVE_DIV == 1 || lane == 0) mbar_arrive(pool_release[which]);
      }
      F16_TIMED(2, tmem_st_wait());
      tc_fence_before();
      __syncwarp();
      named_bar_arrive(ID_A0 + as, SET_THREADS + 32);
      if (q == 0 && lane == 0) F16_STAMP(2, item);
    };
    Run8 d0[4], d1[4];
    for (uint32_t item = 2 * set; item < total_items; item += 2 * NUM_SETS) {
      const bool two = item + 1 < total_items;
      if constexpr (POOL == 1) {
        // shared-memory gather: nothing to overlap with a second item's loads, and one buffer leaves the registers to the window
        issue_loads(item, d0);
        convert_store(item, d0);
        if (two) { issue_loads(item + 1, d0); convert_store(item + 1, d0); }
        continue;
      }
      F16_TIMED(1, issue_loads(item, d0, 0); if (two) issue_loads(item + 1, d1, 1));
      if (q == 0 && lane == 0) { F16_STAMP(0, item); F16_STAMP(0, item + 1); }
      F16_TIMED(3, convert_store(item, d0, 0));
      if (two) F16_TIMED(4, convert_store(item + 1, d1, 1));
    }
    F16_TRACE_STORE(2, warp == W_PRODUCER0 && lane == 0);
  } else {
    // ================= epilogue ======================================================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_EPILOGUE));
    const int q = warp & 3;
    const int e = tid - 32 * W_EPILOGUE0;                      // 0..127
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(32 * q) << 16);
    float* sbias = reinterpret_cast<float*>(base_ptr + L::BIAS);
    const uint32_t stage_u32 = base + L::STAGING + q * L::STG_BLOCKS * 4096;
    uint8_t* stage_ptr = base_ptr + L::STAGING + q * L::STG_BLOCKS * 4096;
    uint32_t chunkcount = 0;
    // staging row of this thread's accumulator lane: the pixel order (POOL: the producers' permutation undone, see there)
    const int srow = POOL == 1 ? 4 * (lane & 7) + (lane >> 3) : lane;
    f32x2 chk = pack_f32x2(0.f, 0.f);
    const float act_lo = p.act == B200OV_ACT_NONE ? -INFINITY : (p.act == B200OV_ACT_RELU ? 0.f : p.lo);
    const float act_hi = p.act == B200OV_ACT_CLAMP ? p.hi : INFINITY;
    F16_TRACE_DECL
    for (int tl = 0; tl < my_tiles; ++tl) {
      const uint32_t tile = blockIdx.x + (uint32_t)tl * gridDim.x;
      uint32_t m_blk, n_blk, ksp;
      tile_coord(p, tile, m_blk, n_blk, ksp);
      const int m0 = (int)m_blk * BLOCK_M, n0 = (int)n_blk * BLOCK_N;
      const int row_off = (int)ksp * p.ws_rows;                 // split-K: this split's block of workspace rows
      named_bar_sync(EPI_BAR_ID, NUM_EPILOGUE);                 // everyone is done with the previous tile's bias
      if (e < BLOCK_N) sbias[e] = (bias != nullptr && n0 + e < p.cout) ? __ldg(bias + (n0 + e) % p.bias_n) : 0.f;
      named_bar_sync(EPI_BAR_ID, NUM_EPILOGUE);
      // FP32 accumulators as packed pairs (FADD2 / FFMA2 halve the issue slots), initialised with the bias
      f32x2 acc[BLOCK_N / 2];
#pragma unroll
      for (int j = 0; j < BLOCK_N / 4; ++j) {
        const ulonglong2 b4 = *reinterpret_cast<const ulonglong2*>(sbias + 4 * j);
        acc[2 * j] = b4.x;
        acc[2 * j + 1] = b4.y;
      }
      const int num_chunks = (p.num_slots + CHUNK - 1) / CHUNK;
      auto promote = [&]() {
        const int buf = chunkcount % NBUF;
        F16_WAIT(0, bar_main_full(buf), (chunkcount / NBUF) & 1);
        if (warp == W_EPILOGUE0 && lane == 0) F16_STAMP(6, chunkcount);
        tc_fence_after();
#pragma unroll
        for (int qb = 0; qb < BLOCK_N / 32; ++qb) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_lane + buf * BLOCK_N + qb * 32, v);
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[qb * 16 + j] = add2(acc[qb * 16 + j], pack_u32x2(v[2 * j], v[2 * j + 1]));
        }
        tc_fence_before();
        if constexpr (DIRECT_EMPTY) { __syncwarp(); named_bar_arrive(ID_M0 + buf, NUM_EPILOGUE + 32); }
        else { if constexpr (ARRIVE_DIV == 32) __syncwarp(); if (ARRIVE_DIV == 1 || lane == 0) mbar_arrive(bar_main_empty(buf)); }
        if (warp == W_EPILOGUE0 && lane == 0) F16_STAMP(7, chunkcount);
        ++chunkcount;
      };
      for (int c = 0; c < num_chunks - 1; ++c) promote();
      // The cross terms are complete together with the last chunk: read them first so the MMA warp can start
      // the next tile's cross accumulation while the last chunk is still being promoted.
      const int xb = NCROSS == 2 ? (tl & 1) : 0;
      F16_WAIT(1, bar_cross_full(xb), NCROSS == 2 ? (tl >> 1) & 1 : tl & 1);
      tc_fence_after();
      const f32x2 unscale = pack_f32x2(LO_UNSCALE, LO_UNSCALE);
#pragma unroll
      for (int qb = 0; qb < BLOCK_N / 32; ++qb) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_lane + (NBUF + xb) * BLOCK_N + qb * 32, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[qb * 16 + j] = fma2(pack_u32x2(v[2 * j], v[2 * j + 1]), unscale, acc[qb * 16 + j]);
      }
      tc_fence_before();
      if constexpr (DIRECT_EMPTY) { __syncwarp(); named_bar_arrive(ID_X, NUM_EPILOGUE + 32); }
      else { if constexpr (ARRIVE_DIV == 32) __syncwarp(); if (ARRIVE_DIV == 1 || lane == 0) mbar_arrive(bar_cross_empty(xb)); }
      promote();
      // activation (None / ReLU / Clamp as one clamp with infinite bounds), staged in shared memory in the
      // 128B-swizzled box layout TMA expects, STG_BLOCKS 32-column blocks per round.  chk turns NaN as soon
      // as one output is inf / NaN.
      const f32x2 zero2 = pack_f32x2(0.f, 0.f);
      // which output tensor a 32-column block belongs to (segments start at multiples of 32 columns)
      auto segment_of = [&](int nb, int& local) -> int {
        int sg = -1;
#pragma unroll
        for (int i = 0; i < 3; ++i)
          if (i < p.nseg && nb >= p.seg_col0[i] && nb < p.seg_col0[i] + p.seg_cout[i]) { sg = i; local = nb - p.seg_col0[i]; }
        return sg;
      };
      if constexpr (O16) {
        // FP16 feature map out: a thread owns one output row (pixel) and writes its 32-column blocks as 64 contiguous
        // bytes, straight from registers.  A value that does not fit FP16 raises the status word like a non-finite one
        // (x * 2^112 * 1.001 overflows to inf from |x| >= 65471).
        const int m = m0 + 32 * q + lane;
        const f32x2 big = pack_f32x2(5.1974e33f, 5.1974e33f);
#pragma unroll
        for (int qb = 0; qb < BLOCK_N / 32; ++qb) {
          int local = 0;
          const int sg = segment_of(n0 + qb * 32, local);
          uint32_t hw[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            chk = fma2(acc[qb * 16 + j], zero2, chk);              // inf / NaN before the activation hides it (fmaxf drops NaN)
            const float2 u = unpack_f32x2(acc[qb * 16 + j]);
            const float o0 = fminf(fmaxf(u.x, act_lo), act_hi), o1 = fminf(fmaxf(u.y, act_lo), act_hi);
            chk = fma2(mul2(pack_f32x2(o0, o1), big), zero2, chk);
            const __half2 h = __floats2half2_rn(o0, o1);
            hw[j] = *reinterpret_cast<const uint32_t*>(&h);
          }
          if (sg < 0 || m >= p.M) continue;
          __half* dst = reinterpret_cast<__half*>(p.seg_y[sg]) + (long long)(row_off + m) * p.seg_yld[sg] + local;
          const int ncols = min(32, p.seg_cout[sg] - local);
          if (ncols == 32 && p.tma_store) {                 // (tma_store: every y 16-byte aligned with an aligned pitch)
#pragma unroll
            for (int v4 = 0; v4 < 4; ++v4)
              *reinterpret_cast<uint4*>(dst + v4 * 8) = make_uint4(hw[v4 * 4], hw[v4 * 4 + 1], hw[v4 * 4 + 2], hw[v4 * 4 + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const __half2 h = *reinterpret_cast<const __half2*>(&hw[j]);
              if (2 * j < ncols) dst[2 * j] = __low2half(h);
              if (2 * j + 1 < ncols) dst[2 * j + 1] = __high2half(h);
            }
          }
        }
        continue;                                            // next tile
      }
#pragma unroll
      for (int round = 0; round < (BLOCK_N / 32) / L::STG_BLOCKS; ++round) {
        F16_TIMED(2, tma_store_wait_read());                     // this thread's earlier stores have read the staging buffer
        __syncwarp();
#pragma unroll
        // A segment another contraction reads is written in that kernel's operand form: per 4 channels (16 bytes)
        // [hi(c0,c1) hi(c2,c3) lo(c0,c1) lo(c2,c3)] -- exactly what its producers would compute from the FP32 values, once
        // per element here instead of once per filter tap and column tile there (OHL variant only).
        auto stage_blocks = [&](auto with_hl) {
#pragma unroll
          for (int sb = 0; sb < L::STG_BLOCKS; ++sb) {
            const int qb = round * L::STG_BLOCKS + sb;
            bool as_hl = false;
            if constexpr (decltype(with_hl)::value) {
              int local_hl = 0;
              const int sg_hl = segment_of(n0 + qb * 32, local_hl);
              as_hl = sg_hl >= 0 && ((p.seg_hl >> sg_hl) & 1);
            }
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) {
              const f32x2 a0 = acc[qb * 16 + c4 * 2], a1 = acc[qb * 16 + c4 * 2 + 1];
              chk = fma2(a0, zero2, chk);
              chk = fma2(a1, zero2, chk);
              const float2 u0 = unpack_f32x2(a0), u1 = unpack_f32x2(a1);
              float4 o;
              o.x = fminf(fmaxf(u0.x, act_lo), act_hi); o.y = fminf(fmaxf(u0.y, act_lo), act_hi);
              o.z = fminf(fmaxf(u1.x, act_lo), act_hi); o.w = fminf(fmaxf(u1.y, act_lo), act_hi);
              if constexpr (decltype(with_hl)::value) {
                if (as_hl) o = encode_hl4(o.x, o.y, o.z, o.w);
              }
              *reinterpret_cast<float4*>(stage_ptr + sb * 4096 + srow * 128 + ((c4 ^ (srow & 7)) << 4)) = o;
            }
          }
        };
        stage_blocks(std::integral_constant<bool, OHL>{});
        if (p.tma_store) {
          fence_proxy_async();
          __syncwarp();
          if (elect_one_sync()) {
#pragma unroll
            for (int sb = 0; sb < L::STG_BLOCKS; ++sb) {
              int local = 0;
              const int sg = segment_of(n0 + (round * L::STG_BLOCKS + sb) * 32, local);
              if (sg >= 0 && m0 + 32 * q < p.M) {
                const CUtensorMap* mp = sg == 0 ? &map_y0 : (sg == 1 ? &map_y1 : &map_y2);
                tma_store_2d(mp, stage_u32 + sb * 4096, local, row_off + m0 + 32 * q);
              }
            }
            tma_store_commit();
          }
        } else {
          // pitch or alignment TMA cannot express: coalesced copy, lane = channel
          __syncwarp();
#pragma unroll
          for (int sb = 0; sb < L::STG_BLOCKS; ++sb) {
            int local = 0;
            const int sg = segment_of(n0 + (round * L::STG_BLOCKS + sb) * 32, local);
            if (sg < 0) continue;
            const int n = local + lane;
            float* ys = p.seg_y[sg];
            const int yld = p.seg_yld[sg], ncout = p.seg_cout[sg];
            for (int r = 0; r < 32; ++r) {
              const int m = m0 + 32 * q + r;
              const float v = *reinterpret_cast<const float*>(stage_ptr + sb * 4096 + r * 128 + (((lane >> 2) ^ (r & 7)) << 4) + (lane & 3) * 4);
              if (m < p.M && n < ncout) ys[(long long)(row_off + m) * yld + n] = v;
            }
          }
          __syncwarp();
        }
      }
    }
    tma_store_wait_all();
    F16_TRACE_STORE(3, warp == W_EPILOGUE0 && lane == 0);
    const float2 ck = unpack_f32x2(chk);
    if ((ck.x != ck.x || ck.y != ck.y) && status != nullptr) atomicOr(status, 1u);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc(tmem_base, 512);
}

// OIHW -> [plane hi | plane lo], each [coutp][kpad] halfs, K ordered (slot, kappa) with the in-slot permutation
// the A producers use: kappa = 16*b + 4*u + j  <->  unit 4*slot + u, channel 4*b + j of that unit.
// Pixel-group form (group > 1, pair layout only): row n' = j * cout + n of the packed matrix is filter n seen from the j-th of
// `group` horizontally adjacent output pixels -- the same taps moved `shift` tap pairs to the right inside a window of
// upt = upt0 + (group - 1) * shift pairs, zeros elsewhere (see conv2d_f16x2_multi).
__global__ void pack_f16_weights_kernel(const float* __restrict__ w, __half* __restrict__ out, int cout, int cin, int kh,
                                        int kw, int coutp, int kpad, int upt, int units, int pair4, int kx0, int group = 1,
                                        int shift = 0) {
  const long long plane = (long long)coutp * kpad;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < plane;
       idx += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(idx / kpad);
    const int kp = (int)(idx - (long long)n * kpad);
    const int slot = kp >> 5, kappa = kp & 31;
    const int unit = slot * 4 + ((kappa & 15) >> 2);
    const int cj = ((kappa >> 4) << 2) + (kappa & 3);
    float v = 0.f;
    if (n < cout * group && unit < units) {
      if (pair4) {                       // unit = (filter row ky, tap pair kxp): taps kx0 + 2*kxp, kx0 + 2*kxp + 1, 4 channels each
        const int j = n / cout, nf = n - j * cout;
        const int ky = unit / upt, kx = kx0 + 2 * (unit - ky * upt - j * shift) + (cj >> 2), c = cj & 3;
        if (kx >= 0 && kx < kw && c < cin && unit - ky * upt >= j * shift) v = w[(((long long)nf * cin + c) * kh + ky) * kw + kx];
      } else {
        const int tap = unit / upt, c = (unit - tap * upt) * 8 + cj;
        if (c < cin) {
          const int ky = tap / kw, kx = tap - ky * kw;
          v = w[(((long long)n * cin + c) * kh + ky) * kw + kx];
        }
      }
    }
    const __half h = __float2half_rn(v);
    const __half l = __float2half_rn((v - __half2float(h)) * LO_SCALE);
    out[idx] = h;
    out[plane + idx] = l;
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

static int make_map_2d(CUtensorMap* map, CUtensorMapDataType dt, int esize, const void* ptr, long long dim0, long long dim1,
                       long long stride1_bytes, int box0, int box1) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) return set_error(B200OV_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)dim0, (cuuint64_t)dim1};
  cuuint64_t strides[1] = {(cuuint64_t)stride1_bytes};
  cuuint32_t box[2] = {(cuuint32_t)box0, (cuuint32_t)box1};
  cuuint32_t estr[2] = {1, 1};
  (void)esize;
  CUresult r = fn(map, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(B200OV_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return B200OV_OK;
}

constexpr int MAX_SMEM = 227 * 1024;

template <int BLOCK_N, int SB, bool WIDE, bool PAIR, bool A16, bool O16, int POOL = 0, bool OHL = false>
static int launch(const Params& p0, const void* x, const float* bias, unsigned int* status, const CUtensorMap& mh,
                  const CUtensorMap& ml, const CUtensorMap* my, cudaStream_t s, const CUtensorMap* mx = nullptr) {
  using L = Smem<BLOCK_N, SB>;
  auto kern = conv_f16x2_kernel<BLOCK_N, SB, WIDE, PAIR, A16, O16, POOL, OHL>;
  static bool configured = false;
  if (!configured) {
    B200OV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, POOL ? MAX_SMEM : L::TOTAL));
    configured = true;
  }
  Params p = p0;
  int smem = L::TOTAL;
  if constexpr (POOL) {
    // the pixel-tile ring takes what the weight ring and the output staging leave of the SM's shared memory
    int stages = (MAX_SMEM - 1024 - L::POOL_RING) / p.pool_stage_bytes;
    if (stages > L::POOL_MAX_STAGES) stages = L::POOL_MAX_STAGES;
    // (staged 1x1: each producer set holds the stages of two items at a time)
    if (stages < (POOL == 2 ? 4 : 2)) return set_error(B200OV_ERR_UNSUPPORTED, "conv2d: pixel-tile ring does not fit in shared memory");
    p.pool_stages = stages;
    p.d_pool_stages = FastDiv(stages);
    smem = L::POOL_RING + stages * p.pool_stage_bytes + 1024;
  }
  const int grid = p.num_tiles < props().sm_count ? p.num_tiles : props().sm_count;
  launch_k(kern, grid, NUM_THREADS, smem, s, p, x, bias, status, mh, ml, my[0], my[1], my[2], mx != nullptr ? *mx : mh);
  B200OV_LAUNCH_CHECK("conv_f16x2_kernel");
  return B200OV_OK;
}

__device__ unsigned int g_status_word;     // sticky: bit 0 = a non-finite value left an f16x2 contraction

}  // namespace f16

// cin <= 4: "pair" layout (a unit = two adjacent taps of one filter row x 4 channels; upt = units per filter row),
// stored twice: pairs starting at tap 0 and pairs starting at tap -1 (a zero tap), see conv2d_f16x2_multi;
// otherwise a unit = 8 channels of one tap (upt = units per tap).
void f16_weight_dims(int cout, int cin, int kh, int kw, int* coutp, int* kpad, int* upt, int* units) {
  const bool pair4 = cin <= 4;
  const int u = pair4 ? ceil_div(kw + 1, 2) : ceil_div(cin, 8);
  const int n_units = pair4 ? kh * u : kh * kw * u;
  if (coutp) *coutp = round_up(cout, 8);
  if (kpad) *kpad = round_up(n_units, 8) * 8;         // whole B stages of 64 K-elements
  if (upt) *upt = u;
  if (units) *units = n_units;
}

// Pixel-group forms of a C_in <= 4 stem with horizontal stride 2 (conv2d_f16x2_multi): `group` horizontally adjacent output
// pixels are one GEMM row with group * cout columns over a window of upt + group - 1 tap pairs.  Groups of 2 .. max are
// packed (the call picks the largest one that divides the output width); f16_group_max returns 1 when there is none.
int f16_group_max(int cout, int cin) {
  if (cin > 4 || cout % 8 != 0 || cout > 64) return 1;
  int group = 128 / round_up(cout, 32);
  if (group > 4) group = 4;
  return group;
}
void f16_group_dims(int group, int cout, int kh, int kw, int* coutp, int* kpad, int* upt, int* units) {
  const int u = ceil_div(kw + 1, 2) + group - 1;
  if (coutp) *coutp = group * cout;
  if (kpad) *kpad = round_up(kh * u, 8) * 8;
  if (upt) *upt = u;
  if (units) *units = kh * u;
}
// halfs from the start of the f16 section to group form `group`, tap alignment `al`
static long long f16_group_offset(int group, int al, int cout, int cin, int kh, int kw) {
  int coutp, kpad;
  f16_weight_dims(cout, cin, kh, kw, &coutp, &kpad, nullptr, nullptr);
  long long off = 4LL * coutp * kpad;
  for (int g = 2; g <= group; ++g) {
    int gc, gk;
    f16_group_dims(g, cout, kh, kw, &gc, &gk, nullptr, nullptr);
    if (g < group) off += 4LL * gc * gk;
    else off += 2LL * al * gc * gk;
  }
  return off;
}

long long f16_section_floats(int cout, int cin, int kh, int kw) {
  int coutp, kpad;
  f16_weight_dims(cout, cin, kh, kw, &coutp, &kpad, nullptr, nullptr);
  long long n = (long long)coutp * kpad * (cin <= 4 ? 2 : 1);   // two planes of halfs (x two tap alignments for the pair layout)
  const int gmax = f16_group_max(cout, cin);
  if (gmax > 1) n = f16_group_offset(gmax + 1, 0, cout, cin, kh, kw) / 2;      // + every pixel-group form, both alignments
  return n;
}

int pack_f16_weights(const float* w_oihw, float* out, int cout, int cin, int kh, int kw, cudaStream_t s) {
  int coutp, kpad, upt, units;
  f16_weight_dims(cout, cin, kh, kw, &coutp, &kpad, &upt, &units);
  f16::pack_f16_weights_kernel<<<bw_grid((long long)coutp * kpad, 256), 256, 0, s>>>(w_oihw, reinterpret_cast<__half*>(out), cout,
                                                                                    cin, kh, kw, coutp, kpad, upt, units,
                                                                                    cin <= 4 ? 1 : 0, 0);
  B200OV_LAUNCH_CHECK("pack_f16_weights_kernel");
  if (cin <= 4) {
    f16::pack_f16_weights_kernel<<<bw_grid((long long)coutp * kpad, 256), 256, 0, s>>>(
        w_oihw, reinterpret_cast<__half*>(out) + 2LL * coutp * kpad, cout, cin, kh, kw, coutp, kpad, upt, units, 1, -1);
    B200OV_LAUNCH_CHECK("pack_f16_weights_kernel");
    for (int group = 2; group <= f16_group_max(cout, cin); ++group) {
      int gcoutp, gkpad, gupt, gunits;
      f16_group_dims(group, cout, kh, kw, &gcoutp, &gkpad, &gupt, &gunits);
      for (int al = 0; al < 2; ++al) {
        f16::pack_f16_weights_kernel<<<bw_grid((long long)gcoutp * gkpad, 256), 256, 0, s>>>(
            w_oihw, reinterpret_cast<__half*>(out) + f16_group_offset(group, al, cout, cin, kh, kw), cout, cin, kh, kw, gcoutp, gkpad,
            gupt, gunits, 1, -al, group, 1);
        B200OV_LAUNCH_CHECK("pack_f16_weights_kernel");
      }
    }
  }
  return B200OV_OK;
}

// The gather reads 8-float runs: either cin is a multiple of 8 (a run = 8 channels of one tap), or cin <= 4 with a
// pixel pitch of exactly 4 floats (a run = two adjacent taps; the producer of x zero-fills the pad lanes, e.g. the
// network-input layout kernel writes a 3-channel image with pitch 4).
bool f16x2_eligible(const b200ov_conv_desc* d, const void* x) {
  if (d->x_dtype == B200OV_DT_HL) {
    if (d->cin <= 4)                   // pre-split network input: the stem's 8-channel super-pixel view only
      return d->x_ld == 4 && d->sw % 2 == 0 && d->w % 2 == 0 && (reinterpret_cast<uintptr_t>(x) & 31u) == 0 && d->act != B200OV_ACT_SIGMOID;
    // a feature map another contraction's epilogue wrote as (hi, lo) pairs, 16 bytes per 4 channels
    return (d->x_ld % 4 == 0) && aligned16(x) && d->cin % 8 == 0 && d->act != B200OV_ACT_SIGMOID;
  }
  if (d->x_dtype == B200OV_DT_F16)     // FP16 feature map in: 8-channel runs of 16 bytes
    return (d->x_ld % 8 == 0) && aligned16(x) && d->cin % 8 == 0 && d->act != B200OV_ACT_SIGMOID;
  return (d->x_ld % 4 == 0) && aligned16(x) && (d->cin % 8 == 0 || (d->cin <= 4 && d->x_ld == 4)) &&
         d->act != B200OV_ACT_SIGMOID;
}

#ifdef B200OV_F16_TRACE
extern "C" int b200ov_debug_f16_trace(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, f16::g_f16_trace, sizeof(f16::g_f16_trace)) == cudaSuccess ? 0 : 2;
}
extern "C" int b200ov_debug_f16_timeline(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, f16::g_f16_timeline, sizeof(f16::g_f16_timeline)) == cudaSuccess ? 0 : 2;
}
#endif

unsigned int* f16x2_status_word() {
  unsigned int* p = nullptr;
  if (cudaGetSymbolAddress(reinterpret_cast<void**>(&p), f16::g_status_word) != cudaSuccess) return nullptr;
  return p;
}

// `wt` points at the f16 section of the packed weights: [hi plane | lo plane] of halfs.  d->cout is the width of the
// (fused) weight matrix; the output columns are routed to `nseg` tensors.
// Split-K plan for a contraction with too few output tiles to fill the GPU (MatMul: 6272 -> 512 at batch 1024 is 32
// tiles on 148 SMs, and every tile walks all 196 K slots): ksplit CTAs share one output tile, each over 1/ksplit of K.
// Returns ksplit (1 = do not split) and the workspace geometry [ksplit * ws_rows][ws_ld] floats.
int f16x2_splitk_plan(int m, int cout, int cin, int kh, int kw, int* ws_rows, int* ws_ld) {
  int coutp, kpad, upt, units;
  f16_weight_dims(cout, cin, kh, kw, &coutp, &kpad, &upt, &units);
  const int num_slots = ceil_div(units, 4);
  int block_n = cout > 64 ? 128 : (cout > 32 ? 64 : 32);
  if (cout > 64 && ceil_div(cout, 96) == ceil_div(cout, 128)) block_n = 96;
  const long long tiles = (long long)ceil_div(m, f16::BLOCK_M) * ceil_div(cout, block_n);
  const int sms = props().sm_count > 0 ? props().sm_count : 148;
  int ksplit = 1;
  if (tiles * 2 <= sms && num_slots >= 16) {
    ksplit = (int)(sms / tiles);
    if (ksplit > num_slots / 8) ksplit = num_slots / 8;          // at least 8 slots (256 K elements) per split
    if (ksplit > 16) ksplit = 16;
    if (ksplit < 1) ksplit = 1;
  }
  if (const char* e = getenv("B200OV_F16_KSPLIT")) { const int v = atoi(e); if (v >= 1 && v <= 32) ksplit = v; }   // developer knob
  if (ws_rows) *ws_rows = round_up(m, f16::BLOCK_M);
  if (ws_ld) *ws_ld = round_up(cout, 4);
  return ksplit;
}

int conv2d_f16x2_multi(const b200ov_conv_desc* d, const void* x, const float* wt, const float* bias, int nseg,
                       const b200ov_conv_seg* segs, cudaStream_t s, int ksplit, int ws_rows) {
  const bool a16 = d->x_dtype == B200OV_DT_F16, o16 = d->y_dtype == B200OV_DT_F16;
  if (!f16x2_eligible(d, x))
    return set_error(B200OV_ERR_UNSUPPORTED, "f16x2 path needs 16-byte aligned NHWC input with cin %% 8 == 0 (or cin <= 4 at a pixel pitch of 4) and no fused Sigmoid");
  if (nseg < 1 || nseg > 3) return set_error(B200OV_ERR_INVALID, "conv2d: 1..3 output segments");
  f16::Params p;
  memset(&p, 0, sizeof(p));
  p.h = d->h; p.w = d->w; p.cin = d->cin; p.cout = d->cout; p.sh = d->sh; p.sw = d->sw; p.pt = d->pt; p.pl = d->pl;
  p.x_ld = d->x_ld; p.y_ld = d->y_ld;
  const long long M = (long long)d->n * d->oh * d->ow;
  if (M > 0x7fffffffLL) return set_error(B200OV_ERR_INVALID, "conv2d: too many output pixels");
  p.M = (int)M;
  if (p.M == 0) return B200OV_OK;
  p.nseg = nseg;
  p.tma_store = 1;
  for (int i = 0; i < nseg; ++i) {
    if (segs[i].y == nullptr || segs[i].cout <= 0 || segs[i].col0 < 0 || segs[i].col0 % 32 != 0 || segs[i].col0 + segs[i].cout > d->cout ||
        segs[i].y_ld < segs[i].cout)
      return set_error(B200OV_ERR_INVALID, "conv2d: bad output segment %d", i);
    p.seg_col0[i] = segs[i].col0; p.seg_cout[i] = segs[i].cout; p.seg_yld[i] = segs[i].y_ld; p.seg_y[i] = static_cast<float*>(segs[i].y);
    if (segs[i].y_dtype == B200OV_DT_HL) {
      if (o16 || segs[i].cout % 4 != 0 || segs[i].y_ld % 4 != 0 || !aligned16(segs[i].y))
        return set_error(B200OV_ERR_INVALID, "conv2d: an (hi, lo)-pair output needs whole 16-byte channel groups (segment %d)", i);
      p.seg_hl |= 1 << i;
    } else if (segs[i].y_dtype != B200OV_DT_F32 && segs[i].y_dtype != d->y_dtype) {
      return set_error(B200OV_ERR_INVALID, "conv2d: bad storage type of output segment %d", i);
    }
    if (segs[i].y_ld % (o16 ? 8 : 4) != 0 || !aligned16(segs[i].y)) p.tma_store = 0;     // (FP16 out: "vector stores allowed")
  }
  int coutp, kpad, upt, units;
  f16_weight_dims(d->cout, d->cin, d->kh, d->kw, &coutp, &kpad, &upt, &units);
  p.units = units;
  p.num_slots = ceil_div(units, 4);
  if (ksplit < 1) ksplit = 1;
  if (ksplit > 1) {
    if (nseg != 1 || bias != nullptr || d->act != B200OV_ACT_NONE || ws_rows < (int)M || o16 || p.seg_hl != 0)
      return set_error(B200OV_ERR_INVALID, "conv2d: split-K writes raw partial sums of one tensor (no bias / activation)");
    p.num_slots = round_up(ceil_div(p.num_slots, ksplit), 2);     // whole weight stages; the last split's tail reads zeros
  }
  p.ksplit = ksplit; p.ws_rows = ksplit > 1 ? ws_rows : 0;
  p.act = d->act; p.lo = d->act_lo; p.hi = d->act_hi;
  p.bias_n = d->cout;
  p.pair4 = d->cin <= 4 && !a16;
  const __half* hi_plane = reinterpret_cast<const __half*>(wt);
  int kw_eff = d->kw, ow_eff = d->ow;
  if (p.pair4 && d->sw % 2 == 0 && d->w % 2 == 0 && (reinterpret_cast<uintptr_t>(x) & 31u) == 0) {
    // Stem with an even horizontal stride on an even-width image (pixel pitch 4 floats): two adjacent pixels are one
    // 32-byte aligned "super-pixel" of 8 channels, and with the tap pairs aligned to even pixel columns (pairs start at
    // tap -(pl & 1); the packed weights hold both alignments) the stem is an ordinary 8-channel convolution over the
    // [h][w/2][8] view: kernel kh x upt, horizontal stride sw/2, left padding ceil(pl/2).  The producers then run the
    // regular path (one 256-bit load per row and unit, one bounds test) instead of the two-half pair gather.
    p.pair4 = 0;
    p.w = d->w / 2; p.x_ld = 8; p.sw = d->sw / 2; p.pl = (d->pl + 1) / 2;
    // Pixel groups: with C_out <= 64 a 128-row tile would issue N <= 64 MMAs, which cost as much pipe time as N ~ 100
    // (52 cycles against 64 for N = 128), and every output pixel would gather its own kh x upt window.  `group` horizontally
    // adjacent output pixels (stride 2 = one super-pixel apart) share all but group - 1 of their tap-pair columns, so they
    // run as ONE GEMM row: group * C_out columns over a window of upt + group - 1 pairs, member j's filter shifted j pairs
    // (zeros elsewhere; packed at load).  The [M / group][group * C_out] result is the same memory as [M][C_out].  7x7 / s2
    // stem, C_out 64: 9 slots per two pixels instead of 14, N = 128; SSD 3x3 / s2 stem, C_out 32: 4 slots per four pixels
    // instead of 8.
    int group = f16_group_max(d->cout, d->cin);
    if (const char* e = getenv("B200OV_F16_STEM_GROUP")) { const int v = atoi(e); if (v >= 1 && v < group) group = v; }   // developer knob
    while (group > 1 && d->ow % group != 0) --group;
    if (group > 1 && d->sw == 2 && nseg == 1 && segs[0].cout == d->cout && segs[0].y_ld == d->cout && p.seg_hl == 0 && !o16 &&
        ksplit == 1) {
      int gcoutp, gkpad, gupt, gunits;
      f16_group_dims(group, d->cout, d->kh, d->kw, &gcoutp, &gkpad, &gupt, &gunits);
      hi_plane += f16_group_offset(group, d->pl & 1, d->cout, d->cin, d->kh, d->kw);
      coutp = gcoutp; kpad = gkpad; units = gunits; upt = gupt;
      p.units = units;
      p.num_slots = ceil_div(units, 4);
      p.cout = group * d->cout;
      p.M = p.M / group;
      p.sw = group;                                    // super-pixels per GEMM row
      p.seg_cout[0] = p.cout; p.seg_yld[0] = p.cout;
      ow_eff = d->ow / group;
    } else if (d->pl & 1) {
      hi_plane += 2LL * coutp * kpad;
    }
    kw_eff = upt;
    upt = 1;
  }
  // Tile width: 128 columns, or 96 where that needs no more tiles (C_out in (64, 96], (128, 192], (256, 288]): an
  // N = 96 MMA takes 56 cycles against 64 for N = 128 (tools/ubench/mma_rate.cu), a 96-wide tile leaves TMEM room
  // for a second cross-term accumulator, and e.g. C_out = 192 is two full tiles instead of one and a half.
  int block_n = p.cout > 64 ? 128 : (p.cout > 32 ? 64 : 32);
  if (p.cout > 64 && ceil_div(p.cout, 96) == ceil_div(p.cout, 128)) block_n = 96;
  if (const char* e = getenv("B200OV_F16_FORCE_N")) {                  // developer knob (tile-width experiments)
    const int v = atoi(e);
    if (v == 32 || v == 64 || v == 96 || v == 128) block_n = v;
  }
  p.tiles_n = ceil_div(p.cout, block_n);
  const long long tiles = (long long)ceil_div(p.M, f16::BLOCK_M) * p.tiles_n * ksplit;
  if (tiles * p.num_slots > 0x7fffffffLL) return set_error(B200OV_ERR_INVALID, "conv2d: problem too large");
  p.num_tiles = (int)tiles;
  // Optional L2 prefetch of what a producer set gathers next (same rows, 16 units further along the channel run), only
  // where that stays inside one filter tap and the four lanes of a pixel cover one 128-byte line.  +4..10 % on layers
  // with long channel runs when the input comes from HBM (micro-benchmarks with a flushed L2), nothing inside the models,
  // where the producing layer left the input in L2 -- so it is opt-in: B200OV_F16_PREFETCH=1 in a -DB200OV_F16_PREFETCH_BUILD build.
  { const char* e = getenv("B200OV_F16_PREFETCH"); p.prefetch = (e && atoi(e) != 0 && !p.pair4 && upt % 4 == 0 && upt > 16) ? 1 : 0; }
  const bool pool = d->pre_pool == B200OV_PREPOOL_MAX3X3S1;
  if (d->pre_pool != B200OV_PREPOOL_NONE &&
      !(pool && d->kh == 1 && d->kw == 1 && d->sh == 1 && d->sw == 1 && d->pt == 0 && d->pl == 0 && d->oh == d->h && d->ow == d->w &&
        d->x_dtype == B200OV_DT_F32 && !o16 && d->cin % 8 == 0 && ksplit == 1))
    return set_error(B200OV_ERR_UNSUPPORTED, "conv2d: pre_pool needs a 1x1 / stride-1 / unpadded convolution of FP32 feature maps");
  p.a_hl = d->x_dtype == B200OV_DT_HL ? 1 : 0;
  if (p.a_hl && p.pair4) return set_error(B200OV_ERR_UNSUPPORTED, "conv2d: a pre-split input needs the super-pixel stem path");
  p.wide_loads = !a16 && !p.pair4 && (p.x_ld % 8 == 0) && ((reinterpret_cast<uintptr_t>(x) & 31u) == 0);
  if (a16) p.prefetch = 0;
  p.d_ohow = FastDiv(d->oh * ow_eff); p.d_ow = FastDiv(ow_eff); p.d_upt = FastDiv(upt); p.d_kw = FastDiv(kw_eff);
  p.d_tiles_n = FastDiv(p.tiles_n); p.d_slots = FastDiv(p.num_slots); p.d_ksplit = FastDiv(ksplit);
  const __half* lo_plane = hi_plane + (long long)coutp * kpad;
  CUtensorMap mh, ml, my[3];
  int rc = f16::make_map_2d(&mh, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, hi_plane, kpad, coutp, (long long)kpad * 2, f16::STAGE_K, block_n);
  if (rc) return rc;
  rc = f16::make_map_2d(&ml, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, lo_plane, kpad, coutp, (long long)kpad * 2, f16::STAGE_K, block_n);
  if (rc) return rc;
  for (int i = 0; i < 3; ++i) {
    if (p.tma_store && i < nseg && !o16) {
      rc = f16::make_map_2d(&my[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, segs[i].y, p.seg_cout[i],
                            ksplit > 1 ? (long long)ksplit * ws_rows : (long long)p.M, (long long)p.seg_yld[i] * 4, 32, 32);
      if (rc) return rc;
    } else {
      my[i] = mh;    // never dereferenced
    }
  }
  unsigned int* status = f16x2_status_word();
  if (pool && p.seg_hl != 0) return set_error(B200OV_ERR_UNSUPPORTED, "conv2d: pre_pool with an (hi, lo) output is not built");
  // Staged 1x1 (POOL == 2): pixel tiles of a plain 1x1 / stride-1 convolution (or MatMul) over an FP32 map go through the same
  // TMA ring as the pooled variant, without the halo.  B200OV_F16_STAGE=0 switches it off (A/B measurements).
  bool hl_stage = false;                 // (hi, lo)-pair inputs too (the producers then only route words): opt-in, B200OV_F16_STAGE=2
  if (const char* e = getenv("B200OV_F16_STAGE")) hl_stage = atoi(e) == 2;
  bool stage = !pool && d->pre_pool == B200OV_PREPOOL_NONE && d->kh == 1 && d->kw == 1 && d->sh == 1 && d->sw == 1 && d->pt == 0 &&
               d->pl == 0 && d->oh == d->h && d->ow == d->w && (d->x_dtype == B200OV_DT_F32 || (d->x_dtype == B200OV_DT_HL && hl_stage)) && !o16 && !a16 && d->cin % 8 == 0 &&
               ksplit == 1 && !p.pair4 && (p.x_ld % 4 == 0) && d->cin >= f16::SLOT_K;
  if (const char* e = getenv("B200OV_F16_STAGE")) { if (atoi(e) == 0) stage = false; }
  if (stage) {
    p.pool_rows = f16::BLOCK_M;
    p.pool_stage_bytes = f16::BLOCK_M * 128;
    CUtensorMap mx;
    rc = f16::make_map_2d(&mx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, x, d->cin, (long long)p.M, (long long)d->x_ld * 4, f16::SLOT_K, p.pool_rows);
    if (rc) return rc;
#define B200OV_F16_LAUNCH_ST(N_, SB_) \
    (p.seg_hl != 0 ? f16::launch<N_, SB_, false, false, false, false, 2, true>(p, x, bias, status, mh, ml, my, s, &mx) \
                   : f16::launch<N_, SB_, false, false, false, false, 2, false>(p, x, bias, status, mh, ml, my, s, &mx))
    if (block_n == 128) return B200OV_F16_LAUNCH_ST(128, 2);
    if (block_n == 96) return B200OV_F16_LAUNCH_ST(96, 2);
    if (block_n == 64) return B200OV_F16_LAUNCH_ST(64, 4);
    return B200OV_F16_LAUNCH_ST(32, 4);
#undef B200OV_F16_LAUNCH_ST
  }
  if (pool) {
    // [pixels][channels] view of x; box = 32 channels x (128 + 2w + 2) pixels, 128B swizzle, zeros outside the tensor
    p.pool_rows = f16::BLOCK_M + 2 * d->w + 2;
    if (p.pool_rows > 256) return set_error(B200OV_ERR_UNSUPPORTED, "conv2d: pre_pool needs an image width of at most 63 pixels");
    p.pool_stage_bytes = round_up(p.pool_rows * 128, 1024);
    CUtensorMap mx;
    rc = f16::make_map_2d(&mx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, x, d->cin, (long long)p.M, (long long)d->x_ld * 4, f16::SLOT_K, p.pool_rows);
    if (rc) return rc;
    // a two-stage weight ring leaves room for the pixel tiles (the layer is bandwidth-bound: the MMA warp never waits on B)
    if (block_n == 128) return f16::launch<128, 2, false, false, false, false, 1>(p, x, bias, status, mh, ml, my, s, &mx);
    if (block_n == 96) return f16::launch<96, 2, false, false, false, false, 1>(p, x, bias, status, mh, ml, my, s, &mx);
    if (block_n == 64) return f16::launch<64, 4, false, false, false, false, 1>(p, x, bias, status, mh, ml, my, s, &mx);
    return f16::launch<32, 4, false, false, false, false, 1>(p, x, bias, status, mh, ml, my, s, &mx);
  }
  if (p.seg_hl != 0) {
    // (hi, lo) output: FP32 feature maps in (or the pair form itself), regular gather
    if (a16 || o16 || p.pair4) return set_error(B200OV_ERR_UNSUPPORTED, "conv2d: no (hi, lo)-output kernel for this input");
#define B200OV_F16_LAUNCH_HL(N_, SB_) \
    (p.wide_loads ? f16::launch<N_, SB_, true, false, false, false, 0, true>(p, x, bias, status, mh, ml, my, s) \
                  : f16::launch<N_, SB_, false, false, false, false, 0, true>(p, x, bias, status, mh, ml, my, s))
    if (block_n == 128) return B200OV_F16_LAUNCH_HL(128, 4);
    if (block_n == 96) return B200OV_F16_LAUNCH_HL(96, 4);
    if (block_n == 64) return B200OV_F16_LAUNCH_HL(64, 6);
    return B200OV_F16_LAUNCH_HL(32, 6);
#undef B200OV_F16_LAUNCH_HL
  }
#define B200OV_F16_LAUNCH2(N_, SB_, W_, P_, A_) \
  (o16 ? f16::launch<N_, SB_, W_, P_, A_, true>(p, x, bias, status, mh, ml, my, s) \
       : f16::launch<N_, SB_, W_, P_, A_, false>(p, x, bias, status, mh, ml, my, s))
#define B200OV_F16_LAUNCH(N_, SB_) \
  (a16 ? B200OV_F16_LAUNCH2(N_, SB_, false, false, true) \
       : p.pair4 ? B200OV_F16_LAUNCH2(N_, SB_, false, true, false) \
                 : p.wide_loads ? B200OV_F16_LAUNCH2(N_, SB_, true, false, false) : B200OV_F16_LAUNCH2(N_, SB_, false, false, false))
  if (block_n == 128) return B200OV_F16_LAUNCH(128, 4);
  if (block_n == 96) return B200OV_F16_LAUNCH(96, 4);
  if (block_n == 64) return B200OV_F16_LAUNCH(64, 6);
  return B200OV_F16_LAUNCH(32, 6);
#undef B200OV_F16_LAUNCH
#undef B200OV_F16_LAUNCH2
}

int conv2d_f16x2(const b200ov_conv_desc* d, const void* x, const float* wt, const float* bias, void* y, cudaStream_t s) {
  b200ov_conv_seg seg;
  seg.y = y; seg.col0 = 0; seg.cout = d->cout; seg.y_ld = d->y_ld; seg.y_dtype = d->y_dtype == B200OV_DT_HL ? B200OV_DT_HL : B200OV_DT_F32;
  return conv2d_f16x2_multi(d, x, wt, bias, 1, &seg, s, 1, 0);
}

// ---- split-K: partial sums -> y = act(bias + sum over splits, in split order) -------------------------------------------
template <int V>
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ ws, const float* __restrict__ bias,
                                                            float* __restrict__ y, int m, int n, int ws_rows, int ws_ld, int ldy,
                                                            int ksplit, int act, float lo, float hi) {
  B200OV_PDL_SYNC();
  const int ng = n / V;
  const long long total = (long long)m * ng;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(idx / ng), c0 = (int)(idx - (long long)row * ng) * V;
    float acc[V];
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = bias != nullptr ? __ldg(bias + c0 + j) : 0.f;
    for (int k = 0; k < ksplit; ++k) {
      const float* src = ws + ((long long)k * ws_rows + row) * ws_ld + c0;
      if constexpr (V == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(src));
        acc[0] = __fadd_rn(acc[0], t.x); acc[1] = __fadd_rn(acc[1], t.y); acc[2] = __fadd_rn(acc[2], t.z); acc[3] = __fadd_rn(acc[3], t.w);
      } else {
        acc[0] = __fadd_rn(acc[0], __ldg(src));
      }
    }
#pragma unroll
    for (int j = 0; j < V; ++j) y[(long long)row * ldy + c0 + j] = apply_act(acc[j], act, lo, hi);
  }
}

int conv2d_f16x2_splitk(const b200ov_conv_desc* d, const void* x, const float* wt, const float* bias, float* y, float* ws,
                        size_t ws_bytes, cudaStream_t s) {
  int ws_rows, ws_ld;
  const long long M = (long long)d->n * d->oh * d->ow;
  const int ksplit = f16x2_splitk_plan((int)M, d->cout, d->cin, d->kh, d->kw, &ws_rows, &ws_ld);
  if (ksplit <= 1 || ws == nullptr) return conv2d_f16x2(d, x, wt, bias, y, s);
  if (ws_bytes < (size_t)ksplit * ws_rows * ws_ld * sizeof(float) || !aligned16(ws))
    return set_error(B200OV_ERR_INVALID, "split-K workspace too small or misaligned (%zu bytes)", ws_bytes);
  b200ov_conv_desc dd = *d;
  dd.act = B200OV_ACT_NONE;
  b200ov_conv_seg seg;
  seg.y = ws; seg.col0 = 0; seg.cout = d->cout; seg.y_ld = ws_ld; seg.y_dtype = B200OV_DT_F32;
  int rc = conv2d_f16x2_multi(&dd, x, wt, nullptr, 1, &seg, s, ksplit, ws_rows);
  if (rc) return rc;
  const bool vec = d->cout % 4 == 0 && d->y_ld % 4 == 0 && aligned16(y) && (bias == nullptr || aligned16(bias));
  if (vec)
    launch_k(splitk_reduce_kernel<4>, bw_grid(M * (d->cout / 4), 256), 256, 0, s, ws, bias, y, (int)M, d->cout, ws_rows, ws_ld, d->y_ld, ksplit,
                                                                         d->act, d->act_lo, d->act_hi);
  else
    launch_k(splitk_reduce_kernel<1>, bw_grid(M * d->cout, 256), 256, 0, s, ws, bias, y, (int)M, d->cout, ws_rows, ws_ld, d->y_ld, ksplit, d->act,
                                                                   d->act_lo, d->act_hi);
  B200OV_LAUNCH_CHECK("splitk_reduce_kernel");
  return B200OV_OK;
}

}  // namespace b200ov
