// Inline-PTX wrappers shared by the tcgen05 kernels (sm_100a): mbarrier, TMA, tcgen05 fences / MMA / TMEM access.
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace b200ov {
namespace ptx {

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
  long long t0 = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if (t0 == 0) t0 = clock64();
    else if (clock64() - t0 > 4000000000LL) __trap();
  }
}
// one non-blocking probe: true if the phase with this parity has completed
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// true in exactly one lane of a converged warp; ptxas then emits the uniform-datapath instructions
// (UTCHMMA, UTCBAR, UTMALDG) of the guarded block once, without a per-lane election loop around each
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(uint32_t id, uint32_t threads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T, kind::f16 (A row-major M x 16 halfs packed two per TMEM column)
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// The six MMAs of one A slot (two K steps of hi*hi, lo*hi, hi*lo) and the commit that releases the slot, issued by one
// elected lane from CONVERGED code: with the election inside the asm block the surrounding loop stays warp-uniform, and
// ptxas keeps its state and all operand arithmetic on the uniform datapath.  Wrapped in `if (elect_one_sync())` the
// same code is a divergent single-lane region: every operand is computed in vector registers and moved with R2UR, and
// an R2UR into a uniform register that a queued UTCHMMA still has to read stalls on the scoreboard -- 27% of the MMA
// warp's time in the ncu source view of that version.
__device__ __forceinline__ void umma_f16x2_slot(uint32_t d_main, uint32_t d_cross, uint32_t a0, uint64_t b_hi, uint64_t b_lo,
                                                uint32_t idesc, uint32_t acc_main, uint32_t acc_cross, uint32_t bar_a_empty) {
  asm volatile(
      "{\n\t.reg .pred pm, pc, le;\n\t.reg .b32 a8, a16, a24;\n\t.reg .b64 h2, l2;\n\t"
      "setp.ne.b32 pm, %6, 0;\n\t"
      "setp.ne.b32 pc, %7, 0;\n\t"
      "elect.sync _|le, 0xffffffff;\n\t"
      "add.u32 a8, %2, 8;\n\t"
      "add.u32 a16, %2, 16;\n\t"
      "add.u32 a24, %2, 24;\n\t"
      "add.u64 h2, %3, 2;\n\t"
      "add.u64 l2, %4, 2;\n\t"
      "@le tcgen05.mma.cta_group::1.kind::f16 [%0], [%2], %3, %5, pm;\n\t"
      "@le tcgen05.mma.cta_group::1.kind::f16 [%1], [a16], %3, %5, pc;\n\t"
      "@le tcgen05.mma.cta_group::1.kind::f16 [%1], [%2], %4, %5, 1;\n\t"
      "@le tcgen05.mma.cta_group::1.kind::f16 [%0], [a8], h2, %5, 1;\n\t"
      "@le tcgen05.mma.cta_group::1.kind::f16 [%1], [a24], h2, %5, 1;\n\t"
      "@le tcgen05.mma.cta_group::1.kind::f16 [%1], [a8], l2, %5, 1;\n\t"
      "@le tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%8];\n\t"
      "}"
      ::"r"(d_main), "r"(d_cross), "r"(a0), "l"(b_hi), "l"(b_lo), "r"(idesc), "r"(acc_main), "r"(acc_cross), "r"(bar_a_empty)
      : "memory");
}
// The same for an A operand that already IS FP16 (feature maps stored as halfs): a = a_hi exactly, so the a_lo * w_hi
// products vanish and a slot costs four MMAs (two K steps of a*w_hi -> main, a*w_lo -> cross) instead of six.
__device__ __forceinline__ void umma_f16a_slot(uint32_t d_main, uint32_t d_cross, uint32_t a0, uint64_t b_hi, uint64_t b_lo,
                                               uint32_t idesc, uint32_t acc_main, uint32_t acc_cross, uint32_t bar_a_empty) {
  asm volatile(
      "{\n\t.reg .pred pm, pc, le;\n\t.reg .b32 a8;\n\t.reg .b64 h2, l2;\n\t"
      "setp.ne.b32 pm, %6, 0;\n\t"
      "setp.ne.b32 pc, %7, 0;\n\t"
      "elect.sync _|le, 0xffffffff;\n\t"
      "add.u32 a8, %2, 8;\n\t"
      "add.u64 h2, %3, 2;\n\t"
      "add.u64 l2, %4, 2;\n\t"
      "@le tcgen05.mma.cta_group::1.kind::f16 [%0], [%2], %3, %5, pm;\n\t"
      "@le tcgen05.mma.cta_group::1.kind::f16 [%1], [%2], %4, %5, pc;\n\t"
      "@le tcgen05.mma.cta_group::1.kind::f16 [%0], [a8], h2, %5, 1;\n\t"
      "@le tcgen05.mma.cta_group::1.kind::f16 [%1], [a8], l2, %5, 1;\n\t"
      "@le tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%8];\n\t"
      "}"
      ::"r"(d_main), "r"(d_cross), "r"(a0), "l"(b_hi), "l"(b_lo), "r"(idesc), "r"(acc_main), "r"(acc_cross), "r"(bar_a_empty)
      : "memory");
}
// tcgen05.commit by one elected lane of a converged warp, if `cond` (warp-uniform) is non-zero
__device__ __forceinline__ void umma_commit_elect(uint32_t bar, uint32_t cond) {
  asm volatile(
      "{\n\t.reg .pred le, pc;\n\t"
      "elect.sync _|le, 0xffffffff;\n\t"
      "setp.ne.and.b32 pc, %1, 0, le;\n\t"
      "@pc tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
      "}"
      ::"r"(bar), "r"(cond)
      : "memory");
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
// 16 lanes x 32 columns: thread t holds, for its 4 column blocks j of 8 columns, lanes t/4 and t/4 + 8,
// columns 8j + 2(t%4) + {0,1}: v[4j+0..1] = lane t/4, v[4j+2..3] = lane t/4 + 8.
__device__ __forceinline__ void tmem_st_16x256b_x4(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
// 16 lanes x 16 columns (two column blocks of 8): v[4j+0..1] = lane t/4, v[4j+2..3] = lane t/4 + 8 of block j
__device__ __forceinline__ void tmem_st_16x256b_x2(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// K-major, 128B-swizzled shared-memory operand tile (rows of 128 bytes, 8-row groups of 1024 bytes).
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);        // start address
  d |= (uint64_t)1 << 16;                            // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset: next 8-row group
  d |= (uint64_t)1 << 46;                            // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
  return d;
}

}  // namespace ptx
}  // namespace b200ov
