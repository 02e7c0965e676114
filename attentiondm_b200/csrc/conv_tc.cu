// int8 implicit-GEMM convolution on the 5th-gen tensor cores (sm_100a):
//   TMA (cp.async.bulk.tensor, SWIZZLE_128B) -> shared -> tcgen05.mma kind::i8
//   -> s32 accumulators in TMEM -> tcgen05.ld -> fused epilogue -> NHWC fp32.
// Reference op: QConv2d.forward -> F.conv2d, utils/quant_util.py:383-385.
//
// GEMM view.  Codes live in HBM as rows of Cp bytes.  For a 3x3/pad-1 conv the
// rows are the halo layout [B][H+2][W+2] (ring = code of 0.0), so tap (kh, kw) is
// the SAME matrix shifted by kh*(W+2)+kw rows: the A tile of every k-block is one
// plain 2-D TMA box {128 channels, 128 rows} at row m0 + shift, no im2col, no
// border logic; rows past the end / channels past Cp are zero-filled by TMA.
// Weights are [O][taps*Cp] (K-major), tile {128, BN}.  One CTA computes a
// 128-row x BN-output tile: M = 128 (TMEM lanes), N = BN <= 256 (TMEM columns).
//
// Two kernels: qconv_i8_halo_kernel (the default: one halo tile per output tile feeds all nine taps, weights
// resident in shared memory when they fit) and qconv_i8_tc_persistent_kernel (ring-fed, one TMA box per
// (tap, channel block); the fallback for feature maps whose halo tile exceeds 512 rows).  Both are persistent
// (one CTA per SM, static round-robin over tiles) with the accumulator double-buffered in TMEM.
#include <cuda.h>

#include <stdlib.h>

#include <mutex>

#include "common.cuh"
#include "conv_common.cuh"

namespace attndm {

constexpr int TC_BM = 128;        // rows per tile  (UMMA M)
constexpr int TC_BK = 128;        // bytes of K per stage (one SWIZZLE_128B row)
constexpr int TC_UMMA_K = 32;     // K per tcgen05.mma for 8-bit operands
constexpr int TC_EPI_WARPS_P = 8;                       // persistent kernel: epilogue warps
constexpr int TC_THREADS_P = 64 + 32 * TC_EPI_WARPS_P;
constexpr int TC_MAX_STAGES = 8;

// ---- PTX wrappers -----------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded wait: a lost arrival traps (fails the launch) instead of hanging the GPU.  mbarrier.try_wait suspends the
// thread in hardware until the phase completes or a system time limit passes, so the loop around it issues only a
// few instructions per poll; the bound is a poll count (reading the clock in the loop costs more issue slots than
// the poll itself).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  uint32_t polls = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && ++polls > (1u << 24)) __trap();
  } while (!done);
}
// Waits of the warps that are NOT on the critical path (epilogue, geometry, producers).  They used to back off with
// nanosleep between polls; the sleeps (up to 2x the requested time each) chained through the producer -> MMA ->
// epilogue handshakes and set the tile rate of the whole kernel (30 us with neither MMAs nor epilogue work), so
// they now rely on try_wait's own hardware suspend like the critical waits.
#ifndef ATTNDM_TC_RELAX_NS
#define ATTNDM_TC_RELAX_NS 0
#endif
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
  if (ATTNDM_TC_RELAX_NS == 0) { mbar_wait(bar, parity); return; }
  uint32_t done;
  uint32_t polls = 0;
  for (;;) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(ATTNDM_TC_RELAX_NS);
    if (++polls > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// The issuing warps run their loops converged (all 32 lanes) and only the instruction itself is
// predicated on an elected lane.  Issuing from inside an `if (lane == 0)` branch makes the compiler
// wrap every UTCIMMA / UTMALDG (which take uniform registers) in a per-thread "waterfall" loop with
// R2UR moves: measured ~200 cycles per MMA issue instead of the ~64 the tensor pipe needs.
__device__ __forceinline__ void tma_load_2d_elect(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n\t}"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// L2 prefetch of a TMA box (no shared memory, no barrier): the halo tiles of the tiles a CTA will process two and
// three iterations from now are pulled from HBM into L2 early, so that the real load (issued only when one of
// the two shared-memory halo buffers frees up) pays L2 latency instead of DRAM latency.  Measured on the
// 128->128 3x3 @32x32 layer: the MMA-only time per tile drops from 2.9 us (= DRAM latency of the halo load)
// towards the 1.2-1.7 us the tensor pipe needs.
__device__ __forceinline__ void tma_prefetch_2d_elect(const CUtensorMap* map, int c0, int c1) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];\n\t}"
      ::"l"(map), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_elect(uint32_t bar, uint32_t bytes) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}"
      ::"r"(bar), "r"(bytes)
      : "memory");
}
__device__ __forceinline__ void tcgen05_commit_elect(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
      ::"r"(bar)
      : "memory");
}
__device__ __forceinline__ void umma_i8_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// leader-predicated variants: the elected lane is chosen ONCE per warp (elect.sync costs ~20 cycles each)
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t is_leader;
  asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(is_leader));
  return is_leader;
}
__device__ __forceinline__ void umma_i8_if(uint32_t leader, uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(leader)
      : "memory");
}
// Up to four K steps (32 bytes each) of one 128-byte k-block in ONE asm block: the descriptors are the constant
// high word 0x40004040 (SBO = 1024 B, version 1, SWIZZLE_128B) over a low word ((addr >> 4) | LBO) that simply
// advances by 2 per step, so the issuing warp executes ~6 instructions per MMA instead of a C-level loop that
// the compiler lowers through vector registers and R2UR moves (measured: 115 cycles per MMA issued, against
// the 64 the tensor pipe needs).
__device__ __forceinline__ void umma_i8_x4_if(uint32_t leader, uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo,
                                              uint32_t idesc, uint32_t accumulate, int nk) {
  asm volatile(
      "{\n\t"
      ".reg .pred q, p, tr, k2, k3, k4;\n\t"
      ".reg .b64 ad, bd;\n\t"
      ".reg .b32 al, bl;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.eq.u32 tr, 0, 0;\n\t"
      "setp.gt.and.s32 k2, %6, 1, q;\n\t"
      "setp.gt.and.s32 k3, %6, 2, q;\n\t"
      "setp.gt.and.s32 k4, %6, 3, q;\n\t"
      "mov.b64 ad, {%1, %7};\n\t"
      "mov.b64 bd, {%2, %7};\n\t"
      "@q tcgen05.mma.cta_group::1.kind::i8 [%0], ad, bd, %3, p;\n\t"
      "add.u32 al, %1, 2;\n\t"
      "add.u32 bl, %2, 2;\n\t"
      "mov.b64 ad, {al, %7};\n\t"
      "mov.b64 bd, {bl, %7};\n\t"
      "@k2 tcgen05.mma.cta_group::1.kind::i8 [%0], ad, bd, %3, tr;\n\t"
      "add.u32 al, %1, 4;\n\t"
      "add.u32 bl, %2, 4;\n\t"
      "mov.b64 ad, {al, %7};\n\t"
      "mov.b64 bd, {bl, %7};\n\t"
      "@k3 tcgen05.mma.cta_group::1.kind::i8 [%0], ad, bd, %3, tr;\n\t"
      "add.u32 al, %1, 6;\n\t"
      "add.u32 bl, %2, 6;\n\t"
      "mov.b64 ad, {al, %7};\n\t"
      "mov.b64 bd, {bl, %7};\n\t"
      "@k4 tcgen05.mma.cta_group::1.kind::i8 [%0], ad, bd, %3, tr;\n\t"
      "}"
      ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(leader), "r"(nk), "r"(0x40004040u)
      : "memory");
}
// ---- CTA pair (cta_group::2): two CTAs of a cluster, i.e. the two SMs of a TPC, execute ONE M = 256 MMA.  Each CTA holds
// its own 128 rows of A and HALF of B (N/2 weight rows); the leader CTA (cluster rank 0) issues the MMAs for both and
// the accumulator rows of a CTA land in its own TMEM.  Per SM the operand fetch drops from 8 KB to 6 KB per MMA (the
// single-CTA N = 128 MMA saturates the 128 B/clk shared-memory port: measured 75-80 cycles per MMA instead of 64 next to
// the epilogue's traffic) and the resident weights from 144 KB to 72 KB, which buys more halo buffers.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;          // clears the CTA-rank bit of a shared::cluster address: rank 0's copy
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load issued by either CTA of the pair into ITS OWN shared memory; the transaction bytes are counted on the LEADER's barrier
__device__ __forceinline__ void tma_load_2d_pair_elect(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n\t}"
      ::"r"(dst), "l"(map), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
// The same two loads through a 5-D view of the weight matrix (make_map_b_perm): coordinates {k, 0, 0, 0, unit of 16 rows}
__device__ __forceinline__ void tma_load_5d_elect(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c4) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %5, %5, %5, %4}], [%2];\n\t}"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c4), "r"(0)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d_pair_elect(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c4) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %5, %5, %5, %4}], [%2];\n\t}"
      ::"r"(dst), "l"(map), "r"(bar & kPeerBitMask), "r"(c0), "r"(c4), "r"(0)
      : "memory");
}
// arrive on the barrier at the same offset in the leader CTA (from either CTA)
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, 0;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(bar)
      : "memory");
}
// commit of the pair's MMAs: one arrival on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void tcgen05_commit_pair_if(uint32_t leader, uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t.reg .b16 m;\n\t"
      "setp.ne.b32 q, %1, 0;\n\t"
      "mov.b16 m, 3;\n\t"
      "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}"
      ::"r"(bar), "r"(leader)
      : "memory");
}
__device__ __forceinline__ void umma_i8_lo_pair(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 ad, bd;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 ad, {%1, %5};\n\t"
      "mov.b64 bd, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], ad, bd, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(0x40004040u)
      : "memory");
}
// One MMA from the two descriptor low words (high word 0x40004040: SBO = 1024 B, version 1, SWIZZLE_128B), for code
// that runs inside `if (elect_one())`: under a branch on elect.sync ptxas keeps every operand in uniform registers
// and emits a bare UTCIMMA with ~3 uniform-datapath instructions around it.  The predicated forms above cost ~20
// instructions per MMA (VOTEU.ANY / R2UR / UMOV under predicates): the issuing warp, not the tensor pipe, then sets
// the MMA rate (measured 75-87 cycles per MMA instead of 64).
__device__ __forceinline__ void umma_i8_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 ad, bd;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 ad, {%1, %5};\n\t"
      "mov.b64 bd, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], ad, bd, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(0x40004040u)
      : "memory");
}
// All MMAs of one output tile with resident weights (called by the elected lane only).  a_lo0 / b_lo0: descriptor low
// words of the halo buffer and of the weight block; tap (kh, kw) slides A by (kh*Wp + kw) rows of 128 B = 8 units.
template <bool PAIR>
__device__ __forceinline__ void umma_tile_resident(uint32_t tmem_d, uint32_t a_lo0, uint32_t b_lo0, uint32_t idesc, int kdim,
                                                   int wp8, int ncb, uint32_t a_step, uint32_t b_step, int ksteps_last) {
  uint32_t b_lo = b_lo0, accumulate = 0;
#pragma unroll 1
  for (int kh = 0; kh < kdim; ++kh) {
    const uint32_t a_row = a_lo0 + (uint32_t)(kh * wp8);
#pragma unroll 1
    for (int kw = 0; kw < kdim; ++kw) {
      uint32_t a_lo = a_row + (uint32_t)(kw * 8);
#pragma unroll 1
      for (int cb = 0; cb < ncb; ++cb) {
        const int nk = (cb == ncb - 1) ? ksteps_last : 4;      // warp-uniform: cheap uniform predicates
        if (PAIR) {
          umma_i8_lo_pair(tmem_d, a_lo, b_lo, idesc, accumulate);
          if (nk > 1) umma_i8_lo_pair(tmem_d, a_lo + 2, b_lo + 2, idesc, 1u);
          if (nk > 2) umma_i8_lo_pair(tmem_d, a_lo + 4, b_lo + 4, idesc, 1u);
          if (nk > 3) umma_i8_lo_pair(tmem_d, a_lo + 6, b_lo + 6, idesc, 1u);
        } else {
          umma_i8_lo(tmem_d, a_lo, b_lo, idesc, accumulate);
          if (nk > 1) umma_i8_lo(tmem_d, a_lo + 2, b_lo + 2, idesc, 1u);
          if (nk > 2) umma_i8_lo(tmem_d, a_lo + 4, b_lo + 4, idesc, 1u);
          if (nk > 3) umma_i8_lo(tmem_d, a_lo + 6, b_lo + 6, idesc, 1u);
        }
        accumulate = 1;
        a_lo += a_step;
        b_lo += b_step;
      }
    }
  }
}
__device__ __forceinline__ void tcgen05_commit_if(uint32_t leader, uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
      ::"r"(bar), "r"(leader)
      : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (8-row groups 1024 B apart)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);          // start address >> 4
  d |= (uint64_t)1 << 16;                           // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                 // stride byte offset: 8 rows * 128 B
  d |= (uint64_t)1 << 46;                           // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                           // layout type SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
// 16 lanes x 256 bit, repeated 4x along columns: 16 registers covering 16 TMEM lanes x 32 columns.
// Measured layout (tools/umma_shift_test.cu): thread t, register j holds
//   lane = t/4 + 8*((j>>1)&1),  column = 2*(t%4) + (j&1) + 8*(j>>2)
// i.e. four consecutive threads hold 8 consecutive columns (32 B) of one row: no shared-memory transpose is
// needed, and after one exchange with the neighbouring lane every thread stores 16 contiguous bytes.
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- kernel -------------------------------------------------------------------
struct __align__(16) ColConst {
  int A;        // zp * wsum[o]
  int B;        // w_zp[o]
  float m;      // 1 / (act_scale * w_scale[o])
  float bias;
};

// ---- persistent variant ------------------------------------------------------------------
// One CTA per SM loops over output tiles (static round-robin).  The accumulator is double-buffered in
// TMEM (2 x BN columns), so the epilogue of tile i overlaps the TMA/MMA main loop of tile i+1, the smem
// ring keeps running across tile boundaries, and TMEM allocation / barrier init / descriptor fetch are
// paid once per SM instead of once per tile.
struct TcGeomP {
  int BN, stages, stage_bytes, ncb, tmem_cols;
  int ntn;            // tiles along N
  long long ntiles;   // total tiles
  int stg_off;        // byte offset of the epilogue staging area inside dynamic smem (after the ring)
  int acc_stride;     // TMEM columns between the two accumulators (BN rounded up to 32: loads are 32 wide)
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__global__ void __launch_bounds__(TC_THREADS_P, 1)
qconv_i8_tc_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                              const ConvI8Params p, const TcGeomP g) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ ColConst colc[256];
  __shared__ long long row_pix[TC_EPI_WARPS_P][32];
  __shared__ int row_b[TC_EPI_WARPS_P][32];

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tiles = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int num_kb = p.taps * g.ncb;

  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    }
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"((uint32_t)g.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else if (warp == 1 && lane == 0) {
    for (int s = 0; s < g.stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&tmem_full_bar[a]), 1);
      mbar_init(smem_u32(&tmem_empty_bar[a]), TC_EPI_WARPS_P);   // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      int s = 0;
      uint32_t ph = 0;
      for (long long tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x) {
        const long long m0 = (tile / g.ntn) * TC_BM;
        const int n0 = (int)(tile % g.ntn) * g.BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          const int tap = kb / g.ncb, cb = kb - tap * g.ncb;
          const long long shift = p.taps == 9 ? (long long)(tap / 3) * p.Wp + (tap % 3) : 0;
          mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1);
          const uint32_t a_dst = tiles + (uint32_t)s * g.stage_bytes;
          const uint32_t b_dst = a_dst + TC_BM * TC_BK;
          const uint32_t bar = smem_u32(&full_bar[s]);
          mbar_expect_tx(bar, (uint32_t)g.stage_bytes);
          tma_load_2d(a_dst, &tmA, bar, cb * TC_BK, (int)(m0 + shift));
          tma_load_2d(b_dst, &tmB, bar, tap * p.Cp + cb * TC_BK, n0);
          if (++s == g.stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(g.BN >> 3) << 17) |
                             ((uint32_t)(TC_BM >> 4) << 24);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (long long tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        mbar_wait(smem_u32(&tmem_empty_bar[acc]), (uint32_t)(((it >> 1) & 1) ^ 1));   // epilogue drained it-2
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * g.acc_stride);
        for (int kb = 0; kb < num_kb; ++kb) {
          const int cb = kb % g.ncb;
          int ksteps = (p.Cp - cb * TC_BK + TC_UMMA_K - 1) / TC_UMMA_K;
          if (ksteps > TC_BK / TC_UMMA_K) ksteps = TC_BK / TC_UMMA_K;
          mbar_wait(smem_u32(&full_bar[s]), ph);
          tcgen05_fence_after();
          const uint32_t a_addr = tiles + (uint32_t)s * g.stage_bytes;
          const uint32_t b_addr = a_addr + TC_BM * TC_BK;
          for (int k = 0; k < ksteps; ++k)
            umma_i8(d_tmem, umma_desc_sw128(a_addr + k * TC_UMMA_K), umma_desc_sw128(b_addr + k * TC_UMMA_K), idesc,
                    (kb > 0 || k > 0) ? 1u : 0u);
          tcgen05_commit(smem_u32(&empty_bar[s]));
          if (++s == g.stages) { s = 0; ph ^= 1; }
        }
        tcgen05_commit(smem_u32(&tmem_full_bar[acc]));
      }
    }
  } else {
    // ===== epilogue warps =====
    // eight warps: warp w reads TMEM lanes [32*(w%4), +32); the two warps of a lane quarter take
    // alternate 32-column chunks, so every SM sub-partition has two epilogue warps to interleave
    const int quarter = warp & 3;
    const int ew = warp - 2;                           // 0..7
    const int half = ew >> 2;
    const int zp = *p.act_zp;
    float4* stg = reinterpret_cast<float4*>(smem_raw + (tiles - smem_u32(smem_raw)) + g.stg_off) + ew * (32 * 8);
    const bool vec_ok = (p.O & 3) == 0;
    const int sub = lane >> 3, ch = lane & 7;
    int last_nt = -1;
    int it = 0;
    for (long long tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const long long m0 = (tile / g.ntn) * TC_BM;
      const int nt = (int)(tile % g.ntn);
      const int n0 = nt * g.BN;
      if (nt != last_nt) {                             // warp-uniform across the four epilogue warps
        asm volatile("bar.sync 1, 256;" ::: "memory");  // nobody still reads the old constants
        for (int c = ew * 32 + lane; c < g.BN; c += 32 * TC_EPI_WARPS_P) {
          const int o = n0 + c;
          ColConst cc = {0, 0, 0.f, 0.f};
          if (o < p.O) {
            cc.A = zp * p.wsum[o];
            cc.B = p.w_zp[o];
            cc.m = p.mult[o];
            cc.bias = p.bias ? p.bias[o] : 0.f;
          }
          colc[c] = cc;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        last_nt = nt;
      }
      const long long row = m0 + quarter * 32 + lane;
      long long pix = 0;
      int b = 0;
      const bool valid = conv_row_to_pixel(p, row, pix, b);
      int cs = 0;
      if (valid) cs = (int)conv_window_rowsum(p, row) + zp * (p.taps * p.C);
      __syncwarp();                                    // previous tile's readers of row_pix are done
      row_pix[ew][lane] = valid ? pix : -1;
      row_b[ew][lane] = b;
      __syncwarp();
      mbar_wait(smem_u32(&tmem_full_bar[acc]), (uint32_t)((it >> 1) & 1));
      tcgen05_fence_after();
      const uint32_t t_acc = tmem_base + (uint32_t)(acc * g.acc_stride) + ((uint32_t)(quarter * 32) << 16);
      const int nchunks = (g.BN + 31) >> 5;
      if (half >= nchunks) {                           // nothing to read for this warp: release at once
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&tmem_empty_bar[acc]));
      }
      for (int ci = half; ci < nchunks; ci += 2) {
        const int c0 = ci << 5;
        uint32_t v[32];
        __syncwarp();
        tmem_ld32(t_acc + (uint32_t)c0, v);
        tmem_ld_wait();
        if (ci + 2 >= nchunks) {                       // this warp's last TMEM read of the tile: hand it back
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&tmem_empty_bar[acc]));
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 f;
          ColConst cc = colc[(c0 + 4 * j + 0) & 255];
          f.x = conv_i8_value((int)v[4 * j + 0], cc.A, cc.B, cs, cc.m, cc.bias);
          cc = colc[(c0 + 4 * j + 1) & 255];
          f.y = conv_i8_value((int)v[4 * j + 1], cc.A, cc.B, cs, cc.m, cc.bias);
          cc = colc[(c0 + 4 * j + 2) & 255];
          f.z = conv_i8_value((int)v[4 * j + 2], cc.A, cc.B, cs, cc.m, cc.bias);
          cc = colc[(c0 + 4 * j + 3) & 255];
          f.w = conv_i8_value((int)v[4 * j + 3], cc.A, cc.B, cs, cc.m, cc.bias);
          stg[lane * 8 + (j ^ (lane & 7))] = f;
        }
        __syncwarp();
        const int o = n0 + c0 + 4 * ch;
        const bool col_ok = (c0 + 4 * ch < g.BN) && (o < p.O);
#pragma unroll
        for (int rb = 0; rb < 32; rb += 16) {
          long long pr[4];
          float4 val[4], rs[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = rb + 4 * i + sub;
            pr[i] = row_pix[ew][r];
            val[i] = stg[r * 8 + (ch ^ (r & 7))];
            rs[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.residual && vec_ok && col_ok && pr[i] >= 0)
              rs[i] = *reinterpret_cast<const float4*>(p.residual + pr[i] * p.O + o);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (!col_ok || pr[i] < 0) continue;
            const int r = rb + 4 * i + sub;
            float* dst = p.out + pr[i] * p.O + o;
            if (vec_ok) {
              float4 t = val[i];
              if (p.residual) { t.x = __fadd_rn(t.x, rs[i].x); t.y = __fadd_rn(t.y, rs[i].y); t.z = __fadd_rn(t.z, rs[i].z); t.w = __fadd_rn(t.w, rs[i].w); }
              if (p.temb) {
                const float4 te = *reinterpret_cast<const float4*>(p.temb + (long long)row_b[ew][r] * p.O + o);
                t.x = __fadd_rn(t.x, te.x); t.y = __fadd_rn(t.y, te.y); t.z = __fadd_rn(t.z, te.z); t.w = __fadd_rn(t.w, te.w);
              }
              *reinterpret_cast<float4*>(dst) = t;
            } else {
              const float e[4] = {val[i].x, val[i].y, val[i].z, val[i].w};
              for (int k = 0; k < 4 && o + k < p.O; ++k) {
                float t = e[k];
                if (p.residual) t = __fadd_rn(t, p.residual[pr[i] * p.O + o + k]);
                if (p.temb) t = __fadd_rn(t, p.temb[(long long)row_b[ew][r] * p.O + o + k]);
                dst[k] = t;
              }
            }
          }
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)g.tmem_cols)
                 : "memory");
  }
}

// ---- optional timeline trace (debug, -DATTNDM_TC_TRACE): one CTA records globaltimer at pipeline events ----------
// The buffer and the CTA to record travel in the kernel parameters (TcGeomH::trace / trace_cta): a hook is a
// timer read and one store.  (The first version fetched them with two global LOADS per hook; those queue behind
// the epilogue's stores in the SM's memory pipeline, so every stamp was taken ~1 us late and the hooks themselves
// serialised the warps they were meant to observe.)
// layout: [role 0..15][it 0..31][event 0..3]; [2048 + it] = clock64 at each tile start; [2100 + 4*cta + ev] = per-CTA
// spans; [4096 + 32*cta + it] = MMA issue-complete time of every tile of every CTA
#ifdef ATTNDM_TC_TRACE
#define TC_TRACE(role, it, ev)                                                                        \
  do {                                                                                                \
    if (g.trace != nullptr && blockIdx.x == (unsigned)g.trace_cta && (it) < 32) {                     \
      unsigned long long now_;                                                                        \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now_)::"memory");                              \
      g.trace[((role) * 32 + (it)) * 4 + (ev)] = now_;                                                \
      if ((role) == 1 && (ev) == 0) g.trace[2048 + (it)] = (unsigned long long)clock64();             \
    }                                                                                                 \
  } while (0)
#define TC_SPAN(ev)                                                                                   \
  do {                                                                                                \
    if (g.trace != nullptr) {                                                                         \
      unsigned long long now_;                                                                        \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now_)::"memory");                              \
      g.trace[2100 + 4 * blockIdx.x + (ev)] = now_;                                                   \
    }                                                                                                 \
  } while (0)
#else
#define TC_TRACE(role, it, ev) do { } while (0)
#define TC_SPAN(ev) do { } while (0)
#endif
static unsigned long long* g_tc_trace_host = nullptr;     // set by attndm_debug_set_tc_trace

// ---- halo-reuse variant ------------------------------------------------------------------
// Measured on B200 (tools/umma_shift_test.cu): a K-major SWIZZLE_128B matrix descriptor may start at ANY
// 128-byte row of a TMA-written tile (base_offset = 0) -- the swizzle is keyed on absolute shared-memory
// address bits.  So a 3x3 conv loads ONE halo tile per output tile (128 + 2*(W+2) + 2 rows, once per
// 128-channel block) and feeds all nine taps from it by sliding the A descriptor by (kh*(W+2)+kw)*128 B:
// L2->smem traffic for the activations drops 9x -> ~1.5x.  When the layer's whole weight matrix fits it
// is loaded once per SM and stays resident; otherwise it streams through a ring as before.
// Warps: 0..7 = epilogue, 8 = activation (halo) producer, 9 = MMA issuer, 10 = weight producer, 11 = geometry.
// The role warps carry the HIGHEST warp ids on purpose: the warp scheduler of an SM sub-partition picks the
// eligible warp with the highest id first, and the MMA issuer shares its sub-partition with two epilogue warps
// (a warp may only read the TMEM lane quarter warp % 4).  As warp 1 it lost the issue slot to them whenever they
// had ALU work and the tensor pipe idled between MMAs (85-96 cycles per MMA instead of 64, tools/umma_contend_test.cu).
constexpr int TC_H_EPI0 = 0;                           // first epilogue warp (multiple of 4: quarter = warp % 4)
#ifndef ATTNDM_TC_EPI_WARPS
#define ATTNDM_TC_EPI_WARPS 12
#endif
// Epilogue warps, a multiple of four (one per TMEM lane quarter and group).  The epilogue is latency bound: with two
// epilogue warps per SM sub-partition the schedulers issued 0.3 instructions per cycle (ncu), i.e. ~3 us for the
// ~9000 warp instructions of a tile against 1.3 us of MMAs.  Twelve warps (three groups taking every third tile in
// split mode) leave 128 registers per thread; sixteen (96 registers) spill.
constexpr int TC_H_EPI_WARPS = ATTNDM_TC_EPI_WARPS;
constexpr int TC_H_R0 = TC_H_EPI0 + TC_H_EPI_WARPS;    // first role warp
// Row-geometry buffers: the geometry warp runs this many tiles ahead of the epilogue.  Sixteen on purpose: its
// row-sum gathers are global loads that queue behind the epilogue's stores in the SM's memory pipeline (ncu: the
// epilogue warps' top stall was the wait for this warp), so it takes its lead while the weights are still loading
// and the memory system is idle, and keeps it.
constexpr int TC_H_NGEO = 16;
// (An asynchronous TMA-store epilogue -- per-warp 2 KB staging slots + tensor stores -- was built and measured in round 1/2:
// 76 us against 59 us for direct 128-bit stores on the 128->128 3x3 layer, because one slot per warp is all the shared memory
// left beside resident weights and the warp waits for the TMA unit before every piece.  The code was removed.)
// (measured without effect on the plain conv: reading both accumulator blocks of a warp up front, 62.5 us vs
// 58.8 us, and reading the next block while the current one is processed, 57.1 us vs 56.9 us)
constexpr int TC_H_NRS = 4;                            // row-sum tiles in flight (bulk copies issued by the geometry warp)
constexpr int TC_H_EPI_GROUPS = TC_H_EPI_WARPS / 4;    // warps sharing a quarter split the 32-column chunks
constexpr int TC_THREADS_H = 32 * (TC_H_R0 + 4);
constexpr int TC_H_MAXB = 8;                           // weight ring depth (streamed mode)

__device__ __forceinline__ void tile_geometry(const ConvI8Params& p, long long row, int zp, long long& pix_o,
                                              int& b_o, int& cs_o) {
  long long pix = 0;
  int b = 0;
  pix_o = -1;
  b_o = 0;
  cs_o = 0;
  if (conv_row_to_pixel(p, row, pix, b)) {
    pix_o = pix;
    b_o = b;
    cs_o = (int)conv_window_rowsum(p, row) + zp * (p.taps * p.C);
  }
}

// n / d for n < 2^31 with a precomputed multiplier: q = umulhi(n, m) >> sh, m = ceil(2^(31+s) / d), s = ceil(log2 d), sh = s - 1
struct FastDiv { unsigned m, sh, d; };
__device__ __forceinline__ unsigned fdiv(unsigned n, const FastDiv& f) { return f.d == 1 ? n : (__umulhi(n, f.m) >> f.sh); }

struct TcGeomH {
  int BN, ncb, tmem_cols, acc_stride, ntn;
  int pair;               // 1: CTA pairs (cluster of 2, cta_group::2 MMAs: M = 256 per pair, each CTA holds half of B)
  int na_shift;           // log2(na): na is 1, 2 or 4
  int split;              // 1: the two groups of four epilogue warps take ALTERNATE tiles (all chunks of their quarter)
  int nacc, nacc_shift;   // accumulators in TMEM (4 when they fit in the 512 columns, else 2) and log2 of that
  long long ntiles;
  int hr;             // halo rows per tile (128 for a 1x1 conv)
  int hr_stride;      // bytes per (halo, channel block) in smem, multiple of 1024
  int na;             // halo buffers (ring)
  int b_resident;     // 1: weights loaded once; 0: streamed
  int nb;             // weight ring depth (streamed)
  int a_off, b_off, end_off;   // byte offsets inside the 1024-aligned dynamic smem (end_off: first byte after the weights)
  int rs_off, rs_stride;       // row-sum ring of the geometry warp (TC_H_NRS slots of rs_stride bytes); rs_stride = 0: gather by loads
  unsigned long long* trace;   // debug timeline buffer (trace builds), else nullptr
  int trace_cta;               // which CTA records it (ATTNDM_TRACE_CTA)
  FastDiv d_per, d_wp, d_hw;   // divisions by Hp*Wp, Wp and H*W in the geometry warp
  FastDiv d_cpg;               // division by the channels per GroupNorm group, O / 32 (STATS builds)
  int perm;                    // 1: weight rows arrive channel-permuted inside each unit of 16 (make_map_b_perm): 128-bit epilogue
  int res_prefetch;            // 1: the halo producer pulls each tile's residual rows into L2 (one bulk prefetch per tile)
  int dbg;            // debug experiments (ATTNDM_TC_DBG, bit mask): low two bits 1 = epilogue skips the math/stores,
                      // 2 = skips the TMEM loads too;
                      // 16 = no halo loads; 32 = epilogue arithmetic without loads/stores; 64 = no MMAs; 512 = geometry warp without row-sum loads; 128 / 256 = the MMA warp
                      // skips the halo-full / accumulator-free waits (results are garbage: timing only)
};


// One 32-row x 32-column block of a tile, held in the 16x256b fragment layout: thread (tr = lane/4,
// tq = lane%4) owns rows tr + 8k (k = 0..3; k < 2 from v0, k >= 2 from v1) and the column pairs
// 8i + 2tq + {0,1} (i = 0..3).  Register j of a load: row bit = (j>>1)&1, column group i = j>>2, parity j&1.
// Everything is statically indexed (registers only).
// Per-thread row state of a tile: four fragment rows (tr + 8k), kept as 32-bit element offsets of the output
// row (relative to p.out; the launcher checks that the output has < 2^31 elements), a validity mask, the
// window row-sums and the sample indices.  Residual and time-embedding addresses are derived from these.
struct EpiRows {
  uint32_t off[4];      // pixel * O + n0 + 2*tq   (0 for an invalid row)
  uint32_t te_off[4];   // sample * O + n0 + 2*tq
  int cs[4];
  uint32_t ok;          // bit k: row k is an output pixel
};

// ---- 128-bit variant for full blocks (all 32 columns valid, O % 16 == 0) ----
// The LSU handles one cache line per cycle, so a float2 store of the fragment layout (8 rows x 32 B per warp
// instruction) costs 8 line-cycles for 256 B and the epilogue of a tile is bound by ~4000 of them.  128-bit stores (8
// rows x 64 B per instruction) halve that, but need FOUR consecutive channels per lane where the 16x256b fragment
// gives two pairs eight columns apart.  Rounds 1-2 exchanged pairs with the neighbouring lane (16 SHFL + 48 FSEL of the
// 276 instructions of a block).  Now the WEIGHTS are permuted instead: TMEM column 8i + 2tq + j of a 16-column unit
// holds output channel 4tq + 2i + j, so a lane's two pairs ARE channels 4tq .. 4tq+3.  The permutation costs nothing:
// the weight matrix is loaded through a 5-D tensor map whose box walks the 16 rows of a unit in that order
// (make_map_b_perm); memory layout, packing and the dp4a twin are unchanged.
// The residual elements of one 32-column block: for each column half h and row k the four channels this lane stores.
// All eight loads of a block are issued together, ahead of the block's TMEM load (the first block's ahead of the wait for
// the accumulator), and consumed after it.  (A rolling window of four -- re-load rs[k] as soon as it has been added --
// pins 16 registers instead of 32 but waits for memory four times per half block: 102 us against 71 us on the 128 -> 128
// 3x3 layer at 32x32.)
__device__ __forceinline__ void epi_load_residual_v4(float4 (&rs)[2][4], const float* res, const EpiRows& r, int c0, int tq) {
  const int col4 = 4 * tq;                                      // first of this lane's four channels inside a unit of 16
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int k = 0; k < 4; ++k)
      rs[h][k] = ((r.ok >> k) & 1) ? __ldg(reinterpret_cast<const float4*>(res + (r.off[k] - 2 * tq) + c0 + 16 * h + col4))
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
}

// STATS: also accumulate the GroupNorm statistics of the values stored, in quad order (conv_common.cuh): fp32 over the
// four channels of one pixel, double from there on.  two (warp-uniform): the quarter straddles two samples;
// dst: &gn_out[(first sample * 32) * 2]; m0 / m1: bit k = row k is an output row of the first / second sample.
struct EpiStats {
  double* dst;
  uint32_t m0, m1;
  bool two;
};
// Per-thread partials of one 32-column block, {sum, sumsq} of column half 0 then of column half 1: a = rows of the
// quarter's first sample, b = rows of its second sample (only when EpiStats::two).
struct EpiPend {
  double a[4], b[4];
};

// The reduction over tr (lane xor 4, 8, 16) of a block's four partials as a reduce-scatter: at xor 4 a lane keeps one
// column half and hands the other to its partner, at xor 8 it keeps the sum or the sum of squares, so sixteen lanes each
// end up with ONE finished value (lane bit 2 = column half, bit 3 = sum / sum of squares, bits 0..1 = tq) and add it
// to the (sample, group) accumulator: four shuffled values per block instead of twelve.
__device__ __forceinline__ void epi_stats_flush(const double (&a)[4], double* dst, int col_base, int lane, const FastDiv& d_cpg) {
  const int tq = lane & 3;
  const bool b4 = lane & 4, b8 = lane & 8;
  const int col4 = 4 * tq;
  double kx = b4 ? a[2] : a[0], ky = b4 ? a[3] : a[1];
  const double sx = b4 ? a[0] : a[2], sy = b4 ? a[1] : a[3];
  kx += __shfl_xor_sync(0xffffffffu, sx, 4);
  ky += __shfl_xor_sync(0xffffffffu, sy, 4);
  double k = b8 ? ky : kx;
  const double s = b8 ? kx : ky;
  k += __shfl_xor_sync(0xffffffffu, s, 8);
  k += __shfl_xor_sync(0xffffffffu, k, 16);
  if (lane < 16)
    atomicAdd(dst + 2 * fdiv((unsigned)(col_base + (b4 ? 16 : 0) + col4), d_cpg) + (b8 ? 1 : 0), k);
}

// where the epilogue constants of column c live in `colc` (see epi_block_v4): 16u + 4a + b  ->  16u + 4b + a
__device__ __forceinline__ int colc_slot(int c) { return (c & ~15) | ((c & 3) << 2) | ((c >> 2) & 3); }

template <bool ADD, bool STATS>
__device__ __forceinline__ void epi_block_v4(const uint32_t (&v0)[16], const uint32_t (&v1)[16], const ColConst* colc,
                                             int c0, int tq, const EpiRows& r, float* out, const float4 (&rs)[2][4], bool has_res,
                                             const float* temb, const EpiStats& st, EpiPend& pend) {
  // STATS: the block's partials are left in `pend`; the caller reduces them across the warp (epi_stats_flush) while the
  // next block's accumulator load is in flight
  const int col4 = 4 * tq;
#pragma unroll
  for (int h = 0; h < 2; ++h) {                       // the two units of 16 channels: column groups (0,1) and (2,3)
    const int i0 = 2 * h, i1 = 2 * h + 1;
    // this lane's four channels c0 + 16h + 4tq + {0,1,2,3} (group i0 holds the first two, group i1 the others); the
    // constants of channel 16u + 4tq + j sit in slot 16u + 4j + tq (colc_slot), so each of the four 128-bit loads
    // reads 64 contiguous bytes across tq -- at the natural index the four lanes were 64 B apart: 2-way bank conflicts
    // on the shared-memory port the MMAs saturate
    const int sl = ((c0 + 16 * h) & 255) + tq;
    const ColConst a0c = colc[sl], a1c = colc[sl + 4], b0c = colc[sl + 8], b1c = colc[sl + 12];
    double ds[4], dq[4];                              // STATS: this half's {sum, sumsq} of each row's four channels
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int j0 = (i0 << 2) | ((k & 1) << 1), j1 = (i1 << 2) | ((k & 1) << 1);
      float4 o;
      o.x = conv_i8_value((int)(k < 2 ? v0[j0] : v1[j0]), a0c.A, a0c.B, r.cs[k], a0c.m, a0c.bias);
      o.y = conv_i8_value((int)(k < 2 ? v0[j0 | 1] : v1[j0 | 1]), a1c.A, a1c.B, r.cs[k], a1c.m, a1c.bias);
      o.z = conv_i8_value((int)(k < 2 ? v0[j1] : v1[j1]), b0c.A, b0c.B, r.cs[k], b0c.m, b0c.bias);
      o.w = conv_i8_value((int)(k < 2 ? v0[j1 | 1] : v1[j1 | 1]), b1c.A, b1c.B, r.cs[k], b1c.m, b1c.bias);
      const bool ok = (r.ok >> k) & 1;
      if (ADD) {
        if (has_res) {
          o.x = __fadd_rn(o.x, rs[h][k].x); o.y = __fadd_rn(o.y, rs[h][k].y);
          o.z = __fadd_rn(o.z, rs[h][k].z); o.w = __fadd_rn(o.w, rs[h][k].w);
        }
        if (temb != nullptr) {
          const float4 te = ok ? __ldg(reinterpret_cast<const float4*>(temb + (r.te_off[k] - 2 * tq) + c0 + 16 * h + col4))
                               : make_float4(0.f, 0.f, 0.f, 0.f);
          o.x = __fadd_rn(o.x, te.x); o.y = __fadd_rn(o.y, te.y); o.z = __fadd_rn(o.z, te.z); o.w = __fadd_rn(o.w, te.w);
        }
      }
      if (ok) *reinterpret_cast<float4*>(out + (r.off[k] - 2 * tq) + c0 + 16 * h + col4) = o;
      if (STATS) {
        const float rsum = __fadd_rn(__fadd_rn(o.x, o.y), __fadd_rn(o.z, o.w));
        const float rsq = fmaf(o.w, o.w, fmaf(o.z, o.z, fmaf(o.y, o.y, __fmul_rn(o.x, o.x))));
        ds[k] = (double)rsum;
        dq[k] = (double)rsq;
      }
    }
    if (STATS) {
      // rows of the first sample (m0) / of the second (m1); pairwise, so the double adds are two deep instead of four
      auto pick = [](const double (&d)[4], uint32_t m) {
        return (((m & 1) ? d[0] : 0.0) + ((m & 2) ? d[1] : 0.0)) + (((m & 4) ? d[2] : 0.0) + ((m & 8) ? d[3] : 0.0));
      };
      pend.a[2 * h] = pick(ds, st.m0);
      pend.a[2 * h + 1] = pick(dq, st.m0);
      if (st.two) { pend.b[2 * h] = pick(ds, st.m1); pend.b[2 * h + 1] = pick(dq, st.m1); }
    }
  }
}

// odd channel counts (the 3-channel eps output): scalar loads/stores, same static indexing
__device__ __forceinline__ void epi_block_scalar(const uint32_t (&v0)[16], const uint32_t (&v1)[16], const ColConst* colc,
                                                 int c0, int tq, int BN, int n0, int O, const EpiRows& r, float* out,
                                                 const float* res, const float* temb, bool perm) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int par = 0; par < 2; ++par) {
      // the output channel (relative to n0) held by accumulator column c0 + 8i + 2tq + par: with permuted weight rows
      // (TcGeomH::perm) a unit of 16 columns holds channel 4tq + 2(i & 1) + par at column 8(i & 1) + 2tq + par
      const int cl = perm ? c0 + 16 * (i >> 1) + 4 * tq + 2 * (i & 1) + par : c0 + 8 * i + 2 * tq + par;
      const bool col_ok = (cl < BN) && (n0 + cl < O);
      if (!col_ok) continue;                 // the 3-channel output: 125 of 128 accumulator columns are padding
      const ColConst cc = colc[colc_slot(cl & 255)];
      const int d = cl - 2 * tq;             // r.off / r.te_off already hold n0 + 2tq
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int j = (i << 2) | ((k & 1) << 1) | par;
        float f = conv_i8_value((int)(k < 2 ? v0[j] : v1[j]), cc.A, cc.B, r.cs[k], cc.m, cc.bias);
        if ((r.ok >> k) & 1) {
          if (res) f = __fadd_rn(f, res[r.off[k] + d]);
          if (temb) f = __fadd_rn(f, temb[r.te_off[k] + d]);
          out[r.off[k] + d] = f;
        }
      }
    }
  }
}

// First output pixel at or after GEMM row `row` (the output pixels of consecutive GEMM rows are consecutive: ring rows
// and ring columns of the halo layout are simply skipped), so the residual elements of a tile are ONE contiguous block.
__device__ __forceinline__ unsigned first_pixel_from_row(const ConvI8Params& p, const TcGeomH& g, unsigned row) {
  const unsigned total = (unsigned)(p.B * p.H * p.W);
  if (row >= (unsigned)p.rows) return total;
  if (p.taps == 1) return row;
  const unsigned b = fdiv(row, g.d_per), rem = row - b * (unsigned)(p.Hp * p.Wp);
  const unsigned hp = fdiv(rem, g.d_wp), wp = rem - hp * (unsigned)p.Wp;
  if (hp >= (unsigned)p.H) return (b + 1) * (unsigned)(p.H * p.W);
  if (wp >= (unsigned)p.W) return (b * (unsigned)p.H + hp + 1) * (unsigned)p.W;
  return (b * (unsigned)p.H + hp) * (unsigned)p.W + wp;
}

// ADDS: the epilogue adds a residual and/or a time embedding.  Two instantiations so that each carries only its
// own epilogue code (the kernel's size is felt in the instruction cache).
// Tile index arithmetic in 32 bits (the launcher checks ntiles < 2^31): the role warps run these once per tile on
// their serial path, where a 64-bit division is a ~150-instruction dependent chain (the geometry warp alone spent
// ~1.2 us per tile in them).
__device__ __forceinline__ unsigned tile_mt(const TcGeomH& g, unsigned tile) { return g.ntn == 1 ? tile : tile / (unsigned)g.ntn; }
__device__ __forceinline__ unsigned tile_nt(const TcGeomH& g, unsigned tile) { return g.ntn == 1 ? 0u : tile % (unsigned)g.ntn; }

// PAIR: the CTA-pair (cta_group::2) build of the kernel; it must be launched as clusters of two CTAs, and the single-CTA
// build must not contain cta_group::2 code (the driver rejects its launch without a matching cluster shape).
template <bool ADDS, bool PAIR, bool STATS>
__global__ void __launch_bounds__(TC_THREADS_H, 1)
qconv_i8_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                     const __grid_constant__ CUtensorMap tmB, const ConvI8Params p, const TcGeomH g) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[4], a_empty[4];
  __shared__ __align__(8) uint64_t b_full[TC_H_MAXB], b_empty[TC_H_MAXB];
  __shared__ __align__(8) uint64_t b_res_bar;
  __shared__ __align__(8) uint64_t tmem_full_bar[4], tmem_empty_bar[4];
  __shared__ uint32_t tmem_base_slot;
  __shared__ ColConst colc[256];
  __shared__ __align__(8) uint64_t geo_full[TC_H_NGEO], geo_empty[TC_H_NGEO], rs_bar[2][TC_H_NRS];
  __shared__ int geo_pix[TC_H_NGEO][TC_BM];        // output pixel of each tile row (-1: not an output)
  __shared__ int geo_cs[TC_H_NGEO][TC_BM];         // window row-sum + zp*K (the sample index is pixel / (H*W))

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int nkb = p.taps * g.ncb;
  const int b_tile_bytes = (PAIR ? g.BN / 2 : g.BN) * TC_BK;      // pair: each CTA holds half of the weight rows
  const unsigned ntiles = (unsigned)g.ntiles;
  // Tile sequence of this CTA: tile0, tile0 + gridDim.x, ...  A CTA pair takes two adjacent tiles per step (rank r the
  // tile 2*(pair index + it*pairs) + r), so both CTAs run the same number of iterations; the odd CTA's last tile may lie
  // past the end (all its rows are then invalid: TMA zero-fills, the epilogue stores nothing).
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const unsigned tile0 = PAIR ? 2u * (blockIdx.x >> 1) + rank : blockIdx.x;

  if (warp == TC_H_R0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    }
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                   "r"((uint32_t)g.tmem_cols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                   "r"((uint32_t)g.tmem_cols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  } else if (warp == TC_H_R0 + 1 && lane == 0) {
    for (int i = 0; i < g.na; ++i) { mbar_init(smem_u32(&a_full[i]), 1); mbar_init(smem_u32(&a_empty[i]), 1); }
    for (int i = 0; i < g.nb; ++i) { mbar_init(smem_u32(&b_full[i]), 1); mbar_init(smem_u32(&b_empty[i]), 1); }
    mbar_init(smem_u32(&b_res_bar), 1);
    for (int a = 0; a < g.nacc; ++a) {
      mbar_init(smem_u32(&tmem_full_bar[a]), 1);
      mbar_init(smem_u32(&tmem_empty_bar[a]), (g.split ? 4 : TC_H_EPI_WARPS) * (PAIR ? 2 : 1));   // pair: both CTAs' warps arrive at the leader's
    }
    for (int a = 0; a < TC_H_NRS; ++a) { mbar_init(smem_u32(&rs_bar[0][a]), 1); mbar_init(smem_u32(&rs_bar[1][a]), 1); }
    for (int a = 0; a < TC_H_NGEO; ++a) {
      mbar_init(smem_u32(&geo_full[a]), 1);
      mbar_init(smem_u32(&geo_empty[a]), g.split ? 4 : TC_H_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();                       // the peer's barriers exist before anything signals them
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  if (threadIdx.x == 0) TC_SPAN(0);

  // ===== geometry: per-row output pixel and window row-sum of a tile, prepared ahead of the epilogue.  Run by the
  // geometry warp and, with resident weights, also by the weight producer once its loads are issued (gw = 1 of 2) =====
  auto geometry_role = [&](const int gw, const int ngw) {
    // ===== geometry warp: per-row output pixel / sample / window row-sum for the NEXT tiles, so the
    // epilogue never waits on the nine dependent-latency row-sum loads or the index divisions =====
    // One warp has to keep up with the tile rate, so: 32-bit index arithmetic (the launcher guarantees
    // rows < 2^31), all 36 row-sum loads of a tile in flight together, and the wait for a free buffer only
    // AFTER they were issued -- with four buffers the warp runs up to four tiles ahead of the epilogue.
    pdl_wait();
    const int zp = *p.act_zp;
    const unsigned rows_u = (unsigned)p.rows, per = (unsigned)(p.Hp * p.Wp), hw = (unsigned)(p.H * p.W);
    const int zk = zp * (p.taps * p.C);
    if (g.rs_stride > 0) {
      // Row sums through the TMA unit: the halo tile's row sums are hr consecutive int32 of p.rowsum, so ONE 1-D bulk
      // copy per tile brings them into a small shared-memory ring (this warp issues the copies TC_H_NRS - 1 tiles
      // ahead and is their only reader: no empty barrier).  Gathered with ordinary loads they queued behind the
      // epilogue's stores in the SM's load/store pipeline -- measured ~2.9 us per tile for this warp, which paced
      // the whole kernel (ncu: the epilogue warps' top stall was the wait for this warp's buffer).
      const uint32_t rs_base = base + (uint32_t)(g.rs_off + gw * TC_H_NRS * g.rs_stride);
      const int* rs_gen = reinterpret_cast<const int*>(smem_raw + (base - smem_u32(smem_raw)) + g.rs_off + gw * TC_H_NRS * g.rs_stride);
      auto rs_issue = [&](unsigned tile, int seq) {
        const int slot = seq % TC_H_NRS;
        const unsigned m0 = tile_mt(g, tile) * TC_BM;
        int n_ent = (int)rows_u - (int)m0;
        if (n_ent > g.hr) n_ent = g.hr;
        if (n_ent < 0) n_ent = 0;
        const uint32_t n4 = (uint32_t)n_ent & ~3u;         // bulk copies move multiples of 16 bytes
        const uint32_t bar = smem_u32(&rs_bar[gw][slot]);
        const uint32_t dst = rs_base + (uint32_t)(slot * g.rs_stride);
        asm volatile(
            "{\n\t.reg .pred q, c;\n\t"
            "elect.sync _|q, 0xffffffff;\n\t"
            "setp.ne.and.b32 c, %3, 0, q;\n\t"
            "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %3;\n\t"
            "@c cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%1], [%2], %3, [%0];\n\t}"
            ::"r"(bar), "r"(dst), "l"(p.rowsum + m0), "r"(n4 * 4u)
            : "memory");
        if ((uint32_t)n_ent != n4) {                       // last tile of a row count that is no multiple of 4
          int* gen = const_cast<int*>(rs_gen) + slot * (g.rs_stride >> 2);
          if ((uint32_t)lane < (uint32_t)n_ent - n4) gen[n4 + lane] = __ldg(p.rowsum + m0 + n4 + lane);
          fence_proxy_async_smem();                        // a later bulk copy overwrites these generic-proxy writes
          __syncwarp();
        }
      };
      const unsigned tstride = (unsigned)ngw * gridDim.x;  // this warp takes the tiles gw, gw + ngw, ... of the CTA
      unsigned tile_i = tile0 + (unsigned)gw * gridDim.x;        // issue cursor, TC_H_NRS - 1 of its tiles ahead
      int seq_i = 0;
      for (; seq_i < TC_H_NRS - 1 && tile_i - rank < ntiles; ++seq_i, tile_i += tstride) rs_issue(tile_i, seq_i);
      int it = gw, seq = 0;
      for (unsigned tile = tile0 + (unsigned)gw * gridDim.x; tile - rank < ntiles; tile += tstride, it += ngw, ++seq) {
        if (tile_i - rank < ntiles) { rs_issue(tile_i, seq_i); ++seq_i; tile_i += tstride; }
        const int slot = seq % TC_H_NRS;
        const unsigned m0 = tile_mt(g, tile) * TC_BM;
        int px[4], off[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {                      // index arithmetic before the wait
          const unsigned row = m0 + lane + 32 * j;
          px[j] = -1;
          off[j] = lane + 32 * j;
          if (row < rows_u) {
            if (p.taps == 1) {
              px[j] = (int)row;
            } else {
              const unsigned b = fdiv(row, g.d_per), rem = row - b * per;
              const unsigned hp = fdiv(rem, g.d_wp), wp = rem - hp * (unsigned)p.Wp;
              if (hp < (unsigned)p.H && wp < (unsigned)p.W) px[j] = (int)((b * (unsigned)p.H + hp) * (unsigned)p.W + wp);
            }
          }
        }
        if (lane == 0) TC_TRACE(2 + (gw ^ 1), it, 0);
        mbar_wait(smem_u32(&rs_bar[gw][slot]), (uint32_t)((seq / TC_H_NRS) & 1));
        const int* rs = rs_gen + slot * (g.rs_stride >> 2);
        int cs[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          cs[j] = 0;
          if (px[j] >= 0) {
            const int* w0 = rs + off[j];
            if (p.taps == 1) {
              cs[j] = w0[0] + zk;
            } else {
              const int* w1 = w0 + p.Wp;
              const int* w2 = w1 + p.Wp;
              cs[j] = w0[0] + w0[1] + w0[2] + w1[0] + w1[1] + w1[2] + w2[0] + w2[1] + w2[2] + zk;
            }
          }
        }
        const int buf = it % TC_H_NGEO;
        mbar_wait_relaxed(smem_u32(&geo_empty[buf]), (uint32_t)(((it / TC_H_NGEO) & 1) ^ 1));
        if (lane == 0) TC_TRACE(2 + (gw ^ 1), it, 1);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          geo_pix[buf][lane + 32 * j] = px[j];
          geo_cs[buf][lane + 32 * j] = cs[j];
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&geo_full[buf]));
        if (lane == 0) TC_TRACE(2 + (gw ^ 1), it, 2);
      }
    } else {
      // fallback (row sums not 16-byte aligned): gather with ordinary loads, one tile at a time
      int it = gw;
      for (unsigned tile = tile0 + (unsigned)gw * gridDim.x; tile - rank < ntiles; tile += (unsigned)ngw * gridDim.x, it += ngw) {
        const unsigned m0 = tile_mt(g, tile) * TC_BM;
        int px[4], cs[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const unsigned row = m0 + lane + 32 * j;
          px[j] = -1;
          cs[j] = 0;
          if (row < rows_u) {
            if (p.taps == 1) {
              px[j] = (int)row;
              cs[j] = __ldg(p.rowsum + row) + zk;
            } else {
              const unsigned b = row / per, rem = row - b * per;
              const unsigned hp = rem / (unsigned)p.Wp, wp = rem - hp * (unsigned)p.Wp;
              if (hp < (unsigned)p.H && wp < (unsigned)p.W) {
                px[j] = (int)((b * (unsigned)p.H + hp) * (unsigned)p.W + wp);
                const int32_t* rs = p.rowsum + row;
                cs[j] = __ldg(rs) + __ldg(rs + 1) + __ldg(rs + 2) + __ldg(rs + p.Wp) + __ldg(rs + p.Wp + 1) + __ldg(rs + p.Wp + 2) +
                        __ldg(rs + 2 * p.Wp) + __ldg(rs + 2 * p.Wp + 1) + __ldg(rs + 2 * p.Wp + 2) + zk;
              }
            }
          }
        }
        const int buf = it % TC_H_NGEO;
        mbar_wait_relaxed(smem_u32(&geo_empty[buf]), (uint32_t)(((it / TC_H_NGEO) & 1) ^ 1));
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          geo_pix[buf][lane + 32 * j] = px[j];
          geo_cs[buf][lane + 32 * j] = cs[j];
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&geo_full[buf]));
      }
    }
  };

  // Register file: 512 threads x 128 is all of it.  The four role warps (one warpgroup, the highest warp ids) need few
  // registers; the epilogue's residual block (32), accumulator fragments (32), row state and statistics spilled at 128.
  // The role warpgroup hands back 72 registers per thread, the three epilogue warpgroups take 24 each.
  // (setmaxnreg has to sit INSIDE the branch it governs: ptxas allocates the code after a join for the smaller count.)
  if (warp >= TC_H_R0) {
    if (TC_H_EPI_WARPS == 12) asm volatile("setmaxnreg.dec.sync.aligned.u32 80;" ::: "memory");
    if (warp == TC_H_R0 + 2) {
      {
        // ===== weight producer (weights are static: no need to wait for the previous kernel) =====
        if (g.b_resident) {
          const int ntn_tiles_here = 1;   // resident mode is only chosen when ntn == 1
          (void)ntn_tiles_here;
          if (PAIR) {
            // each CTA loads ITS half of the output channels (rows rank*BN/2 ...); both halves count on the leader's barrier
            if (rank == 0) mbar_expect_tx_elect(smem_u32(&b_res_bar), (uint32_t)(2 * nkb * b_tile_bytes));
            for (int kb = 0; kb < nkb; ++kb) {
              const int tap = kb / g.ncb, cb = kb - tap * g.ncb;
              if (g.perm) tma_load_5d_pair_elect(base + g.b_off + (uint32_t)kb * b_tile_bytes, &tmB, smem_u32(&b_res_bar),
                                                 tap * p.Cp + cb * TC_BK, (int)rank * (g.BN >> 5));
              else tma_load_2d_pair_elect(base + g.b_off + (uint32_t)kb * b_tile_bytes, &tmB, smem_u32(&b_res_bar), tap * p.Cp + cb * TC_BK,
                                          (int)rank * (g.BN >> 1));
            }
          } else {
            mbar_expect_tx_elect(smem_u32(&b_res_bar), (uint32_t)(nkb * b_tile_bytes));
            for (int kb = 0; kb < nkb; ++kb) {
              const int tap = kb / g.ncb, cb = kb - tap * g.ncb;
              if (g.perm) tma_load_5d_elect(base + g.b_off + (uint32_t)kb * b_tile_bytes, &tmB, smem_u32(&b_res_bar), tap * p.Cp + cb * TC_BK, 0);
              else tma_load_2d_elect(base + g.b_off + (uint32_t)kb * b_tile_bytes, &tmB, smem_u32(&b_res_bar), tap * p.Cp + cb * TC_BK, 0);
            }
          }
          geometry_role(1, 2);
        } else {
          int s = 0;
          uint32_t ph = 0;
          for (unsigned tile = tile0; tile - rank < ntiles; tile += gridDim.x) {
            const int n0 = (int)tile_nt(g, tile) * g.BN;
            for (int kb = 0; kb < nkb; ++kb) {
              const int tap = kb / g.ncb, cb = kb - tap * g.ncb;
              mbar_wait_relaxed(smem_u32(&b_empty[s]), ph ^ 1);
              mbar_expect_tx_elect(smem_u32(&b_full[s]), (uint32_t)b_tile_bytes);
              if (g.perm) tma_load_5d_elect(base + g.b_off + (uint32_t)s * b_tile_bytes, &tmB, smem_u32(&b_full[s]), tap * p.Cp + cb * TC_BK, n0 >> 4);
              else tma_load_2d_elect(base + g.b_off + (uint32_t)s * b_tile_bytes, &tmB, smem_u32(&b_full[s]), tap * p.Cp + cb * TC_BK, n0);
              if (++s == g.nb) { s = 0; ph ^= 1; }
            }
          }
        }
      }
    } else if (warp == TC_H_R0) {
      pdl_wait();                                        // the codes come from the previous kernel
      {
        // ===== activation (halo) producer: one halo per (tile, channel block) =====
        const int nfull = g.hr >= 256 ? g.hr / 256 : 1;    // a TMA box holds at most 256 rows: nfull boxes of tmA ...
        const int hr2 = g.hr >= 256 ? g.hr - 256 * nfull : 0;   // ... and the remainder through tmA2
        int it = 0;
        for (unsigned tile = tile0; tile - rank < ntiles; tile += gridDim.x, ++it) {
          const unsigned m0 = tile_mt(g, tile) * TC_BM;
          const int buf = it & (g.na - 1);
          if (lane == 0) TC_TRACE(0, it, 0);
          mbar_wait_relaxed(smem_u32(&a_empty[buf]), (uint32_t)(((it >> g.na_shift) & 1) ^ 1));
          if (lane == 0) TC_TRACE(0, it, 1);
          const uint32_t bar = smem_u32(&a_full[buf]);
          if ((g.dbg & 16) && !PAIR) { mbar_expect_tx_elect(bar, 0u); continue; }       // experiment: no halo loads at all
          if (PAIR) {
            // both CTAs' halos count on the LEADER's barrier (its MMA warp issues for the pair)
            if (rank == 0) mbar_expect_tx_elect(bar, (uint32_t)(2 * g.ncb * g.hr * TC_BK));
            for (int cb = 0; cb < g.ncb; ++cb) {
              const uint32_t dst = base + g.a_off + (uint32_t)(buf * g.ncb + cb) * g.hr_stride;
              for (int bx = 0; bx < nfull; ++bx) tma_load_2d_pair_elect(dst + bx * 256 * TC_BK, &tmA, bar, cb * TC_BK, (int)m0 + 256 * bx);
              if (hr2 > 0) tma_load_2d_pair_elect(dst + nfull * 256 * TC_BK, &tmA2, bar, cb * TC_BK, (int)m0 + 256 * nfull);
            }
          } else {
            mbar_expect_tx_elect(bar, (uint32_t)(g.ncb * g.hr * TC_BK));
            for (int cb = 0; cb < g.ncb; ++cb) {
              const uint32_t dst = base + g.a_off + (uint32_t)(buf * g.ncb + cb) * g.hr_stride;
              for (int bx = 0; bx < nfull; ++bx) tma_load_2d_elect(dst + bx * 256 * TC_BK, &tmA, bar, cb * TC_BK, (int)m0 + 256 * bx);
              if (hr2 > 0) tma_load_2d_elect(dst + nfull * 256 * TC_BK, &tmA2, bar, cb * TC_BK, (int)m0 + 256 * nfull);
            }
          }
          if (ADDS && g.res_prefetch) {
            // this tile's residual rows -> L2, now (the epilogue reaches the tile a few iterations from here): its
            // 128-bit loads then pay L2 latency, and DRAM streams the block instead of answering scattered requests
            const unsigned px0 = first_pixel_from_row(p, g, m0), px1 = first_pixel_from_row(p, g, m0 + TC_BM);
            const uint32_t bytes = (px1 - px0) * (uint32_t)p.O * 4u;
            if (bytes != 0) {
              asm volatile(
                  "{\n\t.reg .pred q;\n\t"
                  "elect.sync _|q, 0xffffffff;\n\t"
                  "@q cp.async.bulk.prefetch.L2.global [%0], %1;\n\t}"
                  ::"l"(p.residual + (size_t)px0 * p.O), "r"(bytes)
                  : "memory");
            }
          }
          // pull the halos of the tiles 2 and 3 iterations ahead into L2
          for (int ahead = (it == 0 ? 1 : 3); ahead <= 3; ++ahead) {
            const unsigned tf = tile + (unsigned)ahead * gridDim.x;
            if (tf - rank < ntiles) {
              const unsigned mf = tile_mt(g, tf) * TC_BM;
              for (int cb = 0; cb < g.ncb; ++cb) {
                for (int bx = 0; bx < nfull; ++bx) tma_prefetch_2d_elect(&tmA, cb * TC_BK, (int)mf + 256 * bx);
                if (hr2 > 0) tma_prefetch_2d_elect(&tmA2, cb * TC_BK, (int)mf + 256 * nfull);
              }
            }
          }
        }
      }
    } else if (warp == TC_H_R0 + 1) {
      pdl_wait();
      if (rank == 0) {                                   // pair mode: the leader CTA issues for both
        // ===== MMA issuer =====
        // The issue loop is scalar code on one warp: anything slow between two tcgen05.mma shows up as idle
        // tensor-pipe time (an integer division per k-block cost ~200 cycles per MMA).  So: nested loops with
        // additive address updates only, and descriptors assembled from a constant high word plus a 14-bit
        // (address >> 4) low field that is simply incremented (+2 per 32-byte K step).
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(g.BN >> 3) << 17) |
                               ((uint32_t)((PAIR ? 2 * TC_BM : TC_BM) >> 4) << 24);
        const uint64_t desc_hi = ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
                                 ((uint64_t)2 << 61);
        const int kdim = p.taps == 9 ? 3 : 1;
        const int ksteps_last = ((p.Cp - (g.ncb - 1) * TC_BK) + TC_UMMA_K - 1) / TC_UMMA_K;   // 1..4
        const uint32_t leader = elect_one();
  #ifdef ATTNDM_TC_TRACE
        unsigned long long* const tr_all = g.trace;        // debug: per-tile issue-complete timestamps of every CTA
  #else
        unsigned long long* const tr_all = nullptr;
  #endif
        int s = 0;
        uint32_t ph = 0;
        int it = 0;
        for (unsigned tile = tile0; tile - rank < ntiles; tile += gridDim.x, ++it) {
          const int acc = it & (g.nacc - 1), buf = it & (g.na - 1);
          if (lane == 0) TC_TRACE(1, it, 0);
          if (!(g.dbg & 256)) mbar_wait(smem_u32(&tmem_empty_bar[acc]), (uint32_t)(((it >> g.nacc_shift) & 1) ^ 1));
          if (lane == 0) TC_TRACE(1, it, 1);
          if (!(g.dbg & 128)) mbar_wait(smem_u32(&a_full[buf]), (uint32_t)((it >> g.na_shift) & 1));
          if (g.b_resident && it == 0) { mbar_wait(smem_u32(&b_res_bar), 0); if (lane == 0) TC_SPAN(1); }
          tcgen05_fence_after();
          if (lane == 0) TC_TRACE(1, it, 2);
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * g.acc_stride);
          const uint32_t a_buf0 = base + g.a_off + (uint32_t)(buf * g.ncb) * g.hr_stride;
          uint32_t b_res = base + g.b_off;               // resident mode: walks through the whole weight block
          uint32_t accumulate = 0;
          if (g.dbg & 64) {
            // experiment: no MMAs at all (the commits below arrive at once): the epilogue's own tile rate
          } else if (g.b_resident) {
            // no waits inside a tile: the elected lane issues the whole tile from a branch (see umma_i8_lo)
            const uint32_t a_lo0 = ((a_buf0 >> 4) & 0x3FFF) | 0x10000u;
            const uint32_t b_lo0 = ((b_res >> 4) & 0x3FFF) | 0x10000u;
            const uint32_t a_step = (uint32_t)g.hr_stride >> 4, b_step = (uint32_t)b_tile_bytes >> 4;
  #ifdef ATTNDM_TC_OLD_ISSUE
            uint32_t b_lo = b_lo0;
            for (int kh = 0; kh < kdim; ++kh) {
              for (int kw = 0; kw < kdim; ++kw) {
                uint32_t a_lo = a_lo0 + (uint32_t)(kh * p.Wp + kw) * (TC_BK >> 4);
                for (int cb = 0; cb < g.ncb; ++cb) {
                  umma_i8_x4_if(leader, d_tmem, a_lo, b_lo, idesc, accumulate,
                                (cb == g.ncb - 1) ? ksteps_last : TC_BK / TC_UMMA_K);
                  accumulate = 1;
                  a_lo += a_step;
                  b_lo += b_step;
                }
              }
            }
  #else
            if (leader) {
              if (PAIR) umma_tile_resident<true>(d_tmem, a_lo0, b_lo0, idesc, kdim, p.Wp * (TC_BK >> 4), g.ncb, a_step, b_step, ksteps_last);
              else        umma_tile_resident<false>(d_tmem, a_lo0, b_lo0, idesc, kdim, p.Wp * (TC_BK >> 4), g.ncb, a_step, b_step, ksteps_last);
            }
            __syncwarp();
            accumulate = 1;
  #endif
          } else
          for (int kh = 0; kh < kdim; ++kh) {
            for (int kw = 0; kw < kdim; ++kw) {
              uint32_t a_addr = a_buf0 + (uint32_t)(kh * p.Wp + kw) * TC_BK;
              for (int cb = 0; cb < g.ncb; ++cb) {
                uint32_t b_addr;
                if (g.b_resident) {
                  b_addr = b_res;
                  b_res += (uint32_t)b_tile_bytes;
                } else {
                  mbar_wait(smem_u32(&b_full[s]), ph);
                  tcgen05_fence_after();
                  b_addr = base + g.b_off + (uint32_t)s * b_tile_bytes;
                }
                const int ksteps = (cb == g.ncb - 1) ? ksteps_last : TC_BK / TC_UMMA_K;
                uint64_t ad = desc_hi | (uint64_t)((a_addr >> 4) & 0x3FFF);
                uint64_t bd = desc_hi | (uint64_t)((b_addr >> 4) & 0x3FFF);
                for (int k = 0; k < ksteps; ++k) {
                  umma_i8_if(leader, d_tmem, ad, bd, idesc, accumulate);
                  accumulate = 1;
                  ad += TC_UMMA_K >> 4;
                  bd += TC_UMMA_K >> 4;
                }
                if (!g.b_resident) {
                  tcgen05_commit_if(leader, smem_u32(&b_empty[s]));
                  if (++s == g.nb) { s = 0; ph ^= 1; }
                }
                a_addr += (uint32_t)g.hr_stride;
              }
            }
          }
          if (PAIR) {
            tcgen05_commit_pair_if(leader, smem_u32(&a_empty[buf]));
            tcgen05_commit_pair_if(leader, smem_u32(&tmem_full_bar[acc]));
          } else {
            tcgen05_commit_if(leader, smem_u32(&a_empty[buf]));
            tcgen05_commit_if(leader, smem_u32(&tmem_full_bar[acc]));
          }
          if (lane == 0) TC_TRACE(1, it, 3);
          if (tr_all != nullptr && lane == 0 && it < 32) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now) :: "memory");
            tr_all[4096 + blockIdx.x * 32 + it] = now;
          }
        }
      }
    } else if (warp == TC_H_R0 + 3) {
      geometry_role(0, g.b_resident ? 2 : 1);
    }
  } else {
    if (TC_H_EPI_WARPS == 12) asm volatile("setmaxnreg.inc.sync.aligned.u32 144;" ::: "memory");
    // ===== epilogue warps (8): quarter = warp % 4, the two warps of a quarter take 32-column chunks round robin.
    // No shared-memory staging: the 16x256b TMEM load shape already hands four consecutive threads 32
    // contiguous bytes of one output row, so results go TMEM -> registers -> global (128-bit per lane after
    // the neighbour-lane exchange of epi_block_v4).
    pdl_wait();
    const int quarter = warp & 3;
    const int ew = warp - TC_H_EPI0;                   // 0..7
    const int half = ew >> 2;                          // which of the quarter's four warps (chunk phase)
    const int zp = *p.act_zp;
    const int tq = lane & 3, tr = lane >> 2;           // fragment coordinates of this thread
    int last_nt = -1;
    // Split mode (one N tile, four accumulators): group j of four warps (warps 4j .. 4j+3) takes the tiles j, j + G,
    // j + 2G, ... of this CTA, each warp all column chunks of its lane quarter.  The two groups then sit in different phases (one
    // waits / loads TMEM while the other stores), which keeps the SM's store path -- the resource that bounds
    // this kernel -- busy; with all eight warps on one tile they stored and idled in lock step.
    const int tile_step = g.split ? TC_H_EPI_GROUPS : 1;
    const int ci0 = g.split ? 0 : half, ci_step = g.split ? 1 : TC_H_EPI_GROUPS;
    int it = g.split ? half : 0;
    for (unsigned tile = tile0 + (unsigned)it * gridDim.x; tile - rank < ntiles || last_nt < 0;
         tile += (unsigned)tile_step * gridDim.x, it += tile_step) {
      const int acc = it & (g.nacc - 1);
      const int nt = (int)tile_nt(g, tile);
      const int n0 = nt * g.BN;
      if (nt != last_nt) {
        asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_H_EPI_WARPS) : "memory");
        for (int c = ew * 32 + lane; c < g.BN; c += 32 * TC_H_EPI_WARPS) {
          const int o = n0 + c;
          ColConst cc = {0, 0, 0.f, 0.f};
          if (o < p.O) {
            cc.A = zp * p.wsum[o];
            cc.B = p.w_zp[o];
            cc.m = p.mult[o];
            cc.bias = p.bias ? p.bias[o] : 0.f;
          }
          colc[colc_slot(c)] = cc;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_H_EPI_WARPS) : "memory");
        last_nt = nt;
        if (tile - rank >= ntiles) break;                   // (split mode, a CTA with a single tile: the odd group only helped with the constants)
      }
      // row geometry of this thread's four fragment rows (tr, tr+8, tr+16, tr+24 of the quarter), prepared
      // by the geometry warp
      const int gb = it % TC_H_NGEO;
      mbar_wait_relaxed(smem_u32(&geo_full[gb]), (uint32_t)((it / TC_H_NGEO) & 1));
      EpiRows rows;
      rows.ok = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int r = quarter * 32 + tr + 8 * k;
        const int pix = geo_pix[gb][r];
        rows.ok |= (pix >= 0 ? 1u : 0u) << k;
        rows.off[k] = (uint32_t)(pix < 0 ? 0 : pix) * (uint32_t)p.O + (uint32_t)(n0 + 2 * tq);
        rows.te_off[k] = (ADDS && p.temb != nullptr && pix >= 0 ? fdiv((uint32_t)pix, g.d_hw) : 0u) * (uint32_t)p.O + (uint32_t)(n0 + 2 * tq);
        rows.cs[k] = geo_cs[gb][r];
      }
      if (g.dbg & 32) rows.ok = 0;                           // experiment: the epilogue math without loads/stores
      EpiStats est = {nullptr, 0u, 0u, false};
      if (STATS) {
        // samples of this quarter's output rows: all rows of the first one go to partial 0, the rest (the next sample:
        // a sample has >= 32 GEMM rows) to partial 1
        int bk[4], mn = 0x7fffffff, mx = -1;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int pix = geo_pix[gb][quarter * 32 + tr + 8 * k];
          bk[k] = pix >= 0 ? (int)fdiv((uint32_t)pix, g.d_hw) : -1;
          if (pix >= 0) { mn = min(mn, bk[k]); mx = max(mx, bk[k]); }
        }
        mn = __reduce_min_sync(0xffffffffu, mn);
        mx = __reduce_max_sync(0xffffffffu, mx);
        est.two = mx > mn;
#pragma unroll
        for (int k = 0; k < 4; ++k) {                        // (no output row in this quarter: bk = -1 everywhere, no bit set)
          est.m0 |= (bk[k] >= 0 && bk[k] == mn ? 1u : 0u) << k;
          est.m1 |= (bk[k] > mn ? 1u : 0u) << k;
        }
        est.dst = p.gn_out + (long long)(mx >= 0 ? mn : 0) * (2 * kGnGroups);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&geo_empty[gb]));
      const int nchunks = (g.BN + 31) >> 5;
      const bool vec4 = g.perm != 0;                         // 128-bit path for full 32-column blocks (permuted weight rows)
      constexpr bool adds = ADDS;
      auto use4_at = [&](int ci) { return vec4 && ci < nchunks && (ci << 5) + 32 <= g.BN && n0 + (ci << 5) + 32 <= p.O; };
      float4 rs4[2][4];
      if (ADDS && p.residual != nullptr && use4_at(ci0))     // first block's residual: issued before the wait below
        epi_load_residual_v4(rs4, p.residual, rows, ci0 << 5, tq);
      mbar_wait_relaxed(smem_u32(&tmem_full_bar[acc]), (uint32_t)((it >> g.nacc_shift) & 1));
      tcgen05_fence_after();
      if (lane == 0) TC_TRACE(4 + ew, it, 0);
      const uint32_t t_acc = tmem_base + (uint32_t)(acc * g.acc_stride) + ((uint32_t)(quarter * 32) << 16);
      if (ci0 >= nchunks) {
        __syncwarp();
        if (lane == 0) { if (PAIR) mbar_arrive_leader(smem_u32(&tmem_empty_bar[acc])); else mbar_arrive(smem_u32(&tmem_empty_bar[acc])); }
      }
      EpiPend pend;
      int pend_c0 = -1;                                      // STATS: the block whose partials wait in `pend`
      auto stats_flush = [&]() {
        epi_stats_flush(pend.a, est.dst, n0 + pend_c0, lane, g.d_cpg);
        if (est.two) epi_stats_flush(pend.b, est.dst + 2 * kGnGroups, n0 + pend_c0, lane, g.d_cpg);
      };
      for (int ci = ci0; ci < nchunks; ci += ci_step) {
        const int c0 = ci << 5;
        uint32_t v0[16], v1[16];
        const bool use4 = use4_at(ci);
        if (ADDS && use4 && p.residual != nullptr && ci != ci0) epi_load_residual_v4(rs4, p.residual, rows, c0, tq);
        __syncwarp();
        if ((g.dbg & 3) < 2) {
          tmem_ld_16x256b_x4(t_acc + (uint32_t)c0, v0);                      // tile rows 32q + 0..15
          tmem_ld_16x256b_x4(t_acc + (16u << 16) + (uint32_t)c0, v1);        // tile rows 32q + 16..31
          if (STATS && pend_c0 >= 0) stats_flush();          // the previous block's statistics, under the load's latency
          tmem_ld_wait();
          if (lane == 0 && ci == ci0) TC_TRACE(4 + ew, it, 1);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) { v0[j] = 0; v1[j] = 0; }
        }
        if (ci + ci_step >= nchunks) {
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) { if (PAIR) mbar_arrive_leader(smem_u32(&tmem_empty_bar[acc])); else mbar_arrive(smem_u32(&tmem_empty_bar[acc])); }
        }
        if ((g.dbg & 3) >= 1) continue;
        if (use4) {
          // two instantiations only (the kernel's code size is felt in the instruction cache): the plain
          // conv, and one variant that checks the residual / time-embedding pointers at run time
          if (adds) epi_block_v4<true, STATS>(v0, v1, colc, c0, tq, rows, p.out, rs4, p.residual != nullptr, p.temb, est, pend);
          else      epi_block_v4<false, STATS>(v0, v1, colc, c0, tq, rows, p.out, rs4, false, nullptr, est, pend);
          if (STATS) pend_c0 = c0;
        } else {
          // ragged last block or O % 4 != 0 (the 3-channel output): scalar, per-column checks
          epi_block_scalar(v0, v1, colc, c0, tq, g.BN, n0, p.O, rows, p.out, p.residual, p.temb, g.perm != 0);
        }
        if (lane == 0) TC_TRACE(4 + ew, it, ci == ci0 ? 2 : 3);
      }
      if (STATS && pend_c0 >= 0) stats_flush();
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();                       // neither CTA leaves (or frees TMEM) while the pair still works
  if (threadIdx.x == 0) TC_SPAN(2);
  if (warp == TC_H_R0) {
    if (PAIR)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)g.tmem_cols) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)g.tmem_cols) : "memory");
  }
}

// ---- host side --------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    // resolved through the runtime so the library has no link-time libcuda dependency
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)f;
  });
  return fn;
}

static int make_map_2d(CUtensorMap* m, const void* base, uint64_t inner, uint64_t outer, uint32_t box_inner,
                       uint32_t box_outer) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("qconv_i8_tc: cuTensorMapEncodeTiled not available"); return ATTNDM_ERR_CUDA; }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {inner};          // bytes between rows (uint8 elements)
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("qconv_i8_tc: cuTensorMapEncodeTiled failed (%d)", (int)r); return ATTNDM_ERR_CUDA; }
  return ATTNDM_OK;
}

// The weight matrix [O][K] (O % 16 == 0) as a 5-D tensor {K, j: 2, tq: 4, i: 2, unit: O / 16} over the same memory, row =
// 16 unit + 4 tq + 2 i + j.  A box {128, 2, 4, 2, n / 16} lands in shared memory as rows 16 u + 8 i + 2 tq + j, i.e. the row
// (= TMEM column) that the 16x256b fragment hands to thread tq as the j-th element of column group i holds output channel
// 4 tq + 2 i + j: every lane of the epilogue owns four consecutive channels (epi_block_v4).
static int make_map_b_perm(CUtensorMap* m, const void* base, uint64_t K, int O, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("qconv_i8_tc: cuTensorMapEncodeTiled not available"); return ATTNDM_ERR_CUDA; }
  cuuint64_t dims[5] = {K, 2, 4, 2, (cuuint64_t)(O / 16)};
  cuuint64_t strides[4] = {K, 4 * K, 2 * K, 16 * K};        // bytes: j, tq, i, unit
  cuuint32_t box[5] = {TC_BK, 2, 4, 2, box_rows / 16};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 5, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("qconv_i8_tc: cuTensorMapEncodeTiled (permuted weights) failed (%d)", (int)r); return ATTNDM_ERR_CUDA; }
  return ATTNDM_OK;
}

static int launch_qconv_i8_tc_persistent(const ConvI8Params& p, cudaStream_t st) {
  TcGeomP g;
  g.BN = p.O <= 256 ? round_up(p.O, 16) : 256;
  const long long mtiles = (p.rows + TC_BM - 1) / TC_BM;
  while (g.BN > 32 && (g.BN / 2) % 16 == 0 && mtiles * cdiv(p.O, g.BN) < 120) g.BN /= 2;
  g.ntn = cdiv(p.O, g.BN);
  g.ntiles = mtiles * g.ntn;
  g.ncb = cdiv(p.Cp, TC_BK);
  g.stage_bytes = TC_BM * TC_BK + g.BN * TC_BK;
  const int grid = (int)(g.ntiles < kNumSMs ? g.ntiles : kNumSMs);
  const long long tiles_per_cta = (g.ntiles + grid - 1) / grid;
  const int num_kb = p.taps * g.ncb;
  int want = (int)(tiles_per_cta > 1 ? 2LL * num_kb : num_kb);       // no point in a ring deeper than the work
  g.stages = (180 * 1024) / g.stage_bytes;
  if (g.stages > TC_MAX_STAGES) g.stages = TC_MAX_STAGES;
  if (g.stages > want) g.stages = want;
  if (g.stages < 2) g.stages = 2;
  g.acc_stride = round_up(g.BN, 32);
  g.tmem_cols = 32;
  while (g.tmem_cols < 2 * g.acc_stride) g.tmem_cols <<= 1;         // two accumulators
  g.stg_off = g.stages * g.stage_bytes;
  CUtensorMap tmA, tmB;
  int rc = make_map_2d(&tmA, p.codes, (uint64_t)p.Cp, (uint64_t)p.rows, TC_BK, TC_BM);
  if (rc) return rc;
  rc = make_map_2d(&tmB, p.qw, (uint64_t)p.taps * p.Cp, (uint64_t)p.O, TC_BK, (uint32_t)g.BN);
  if (rc) return rc;
  const int smem = g.stages * g.stage_bytes + TC_EPI_WARPS_P * 32 * 8 * 16 + 1024;
  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    attr_err = cudaFuncSetAttribute(qconv_i8_tc_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 215 * 1024);
  });
  if (attr_err != cudaSuccess) { set_error("qconv_i8_tc: cannot raise dynamic smem: %s", cudaGetErrorString(attr_err)); return ATTNDM_ERR_CUDA; }
  launch_pdl(qconv_i8_tc_persistent_kernel, dim3(grid), dim3(TC_THREADS_P), smem, st, tmA, tmB, p, g);
  ATTNDM_CUDA_LAUNCH_CHECK("qconv_i8_tc_persistent");
  return ATTNDM_OK;
}

}  // namespace attndm
extern "C" int attndm_debug_set_tc_trace(unsigned long long* buf) {
  attndm::g_tc_trace_host = buf;
  return 0;
}
namespace attndm {

static FastDiv make_fastdiv(unsigned d) {
  FastDiv f;
  f.d = d < 1 ? 1 : d;
  f.m = 0;
  f.sh = 0;
  if (f.d > 1) {
    unsigned sft = 0;
    while ((1ull << sft) < f.d) ++sft;                       // ceil(log2 d) >= 1
    f.m = (unsigned)(((1ull << (31 + sft)) + f.d - 1) / f.d); // ceil(2^(31+s) / d) < 2^32
    f.sh = sft - 1;
  }
  return f;
}

static bool tc_halo_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("ATTNDM_TC_HALO");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

// returns 1 if the halo kernel was launched, 0 if the shape does not fit it, <0 on error
static int launch_qconv_i8_halo(const ConvI8Params& p, cudaStream_t st, bool* stats_fused) {
  TcGeomH g;
  g.BN = p.O <= 256 ? round_up(p.O, 16) : 256;
  const long long mtiles = (p.rows + TC_BM - 1) / TC_BM;
  while (g.BN > 32 && (g.BN / 2) % 16 == 0 && mtiles * cdiv(p.O, g.BN) < 120) g.BN /= 2;
  g.ntn = cdiv(p.O, g.BN);
  g.ntiles = mtiles * g.ntn;
  g.ncb = cdiv(p.Cp, TC_BK);
  g.hr = p.taps == 9 ? TC_BM + 2 * p.Wp + 2 : TC_BM;
  if (g.hr > 1024) return 0;                                          // W <= 446 (the halo is moved in 256-row TMA boxes)
  if ((long long)p.B * p.H * p.W * p.O >= (1LL << 31)) return 0;      // the epilogue keeps 32-bit output offsets
  if (p.rows + TC_BM >= (1LL << 31) || g.ntiles >= (1LL << 31)) return 0;   // 32-bit row and tile indices in the kernel
  g.hr_stride = round_up(g.hr * TC_BK, 1024);
  // row-sum ring of the geometry warp (bulk copies need a 16-byte aligned source; else it gathers with loads)
  static const bool rs_bulk_on = [] { const char* e = getenv("ATTNDM_TC_RS_BULK"); return !(e && e[0] == '0'); }();
  const int rs_bytes = (rs_bulk_on && ((uintptr_t)p.rowsum & 15) == 0) ? 2 * TC_H_NRS * round_up(g.hr * 4, 16) : 0;   // two geometry warps
  // dynamic smem we allow ourselves: 227 KB - 20.5 KB static - alignment slack - the row-sum ring
  const int kBudget = 205 * 1024 - rs_bytes;
  const int nkb = p.taps * g.ncb;
  const int a_buf = g.ncb * g.hr_stride;
  int grid = (int)(g.ntiles < kNumSMs ? g.ntiles : kNumSMs);
  const long long tiles_per_cta = (g.ntiles + grid - 1) / grid;
  g.na = tiles_per_cta > 1 ? 2 : 1;
  int b_tile = g.BN * TC_BK;
  // CTA pairs (cta_group::2): one N tile whose two halves go to the two CTAs of a cluster; the weights then take half
  // the shared memory per CTA and stay resident next to up to four halo buffers.  ATTNDM_TC_PAIR=0 turns it off.
  static const bool pair_on = [] { const char* e = getenv("ATTNDM_TC_PAIR"); return !(e && e[0] == '0'); }();
  g.pair = 0;
  // Chosen where the single-CTA kernel cannot keep the weights resident next to two halo buffers (W >= 64 at 128 -> 128:
  // 63.9 -> 40.8 us at 64x64, 35.9 -> 22.7 us at 128x128); where it can (the CIFAR shapes) the two measure the same and
  // the single-CTA kernel has the shorter prologue.  ATTNDM_TC_PAIR=2 forces pairs wherever they fit.
  static const int pair_mode = [] { const char* e = getenv("ATTNDM_TC_PAIR"); return e ? atoi(e) : 1; }();
  const bool single_resident = g.ntn == 1 && (long long)nkb * b_tile + (long long)g.na * a_buf <= kBudget;
  if (pair_on && g.ntn == 1 && g.BN % 32 == 0 && mtiles >= 2 && (kNumSMs % 2) == 0 &&
      (long long)nkb * (b_tile / 2) + (long long)a_buf <= kBudget && (pair_mode == 2 || !single_resident)) {
    g.pair = 1;
    b_tile /= 2;                                        // each CTA holds BN/2 weight rows
    const long long pairs = (mtiles + 1) / 2;
    grid = (int)(2 * (pairs < kNumSMs / 2 ? pairs : kNumSMs / 2));
    const long long its = (pairs + grid / 2 - 1) / (grid / 2);
    const long long wb = (long long)nkb * b_tile;
    g.na = (wb + 4LL * a_buf <= kBudget && its >= 4) ? 4 : ((wb + 2LL * a_buf <= kBudget && its >= 2) ? 2 : 1);
  }
  // weights resident when they fit next to the halo buffers (and there is a single N tile)
  g.b_resident = (g.ntn == 1 && (long long)nkb * b_tile + (long long)g.na * a_buf <= kBudget) ? 1 : 0;
  if (g.b_resident) {
    g.nb = 1;
  } else {
    long long left = (long long)kBudget - (long long)g.na * a_buf;
    if (left < 2LL * b_tile && g.na == 2) { g.na = 1; left += a_buf; }
    if (left < 2LL * b_tile) return 0;
    g.nb = (int)(left / b_tile);
    if (g.nb > TC_H_MAXB) g.nb = TC_H_MAXB;
    if (g.nb > nkb * 2) g.nb = nkb * 2;
    if (g.nb < 2) g.nb = 2;
  }
  g.na_shift = g.na == 4 ? 2 : (g.na == 2 ? 1 : 0);
  g.a_off = 0;
  g.b_off = g.na * a_buf;
  g.end_off = g.b_off + (g.b_resident ? nkb : g.nb) * b_tile;
  g.rs_off = g.end_off;
  g.rs_stride = rs_bytes / (2 * TC_H_NRS);
  { const char* e = getenv("ATTNDM_TC_DBG"); g.dbg = e ? atoi(e) : 0; }
  g.trace = g_tc_trace_host;
  g.d_per = make_fastdiv((unsigned)(p.Hp * p.Wp));
  g.d_wp = make_fastdiv((unsigned)p.Wp);
  g.d_hw = make_fastdiv((unsigned)(p.H * p.W));
  g.d_cpg = make_fastdiv((unsigned)(p.O >= kGnGroups ? p.O / kGnGroups : 1));
  static const bool res_pf_on = [] { const char* e = getenv("ATTNDM_TC_RES_PREFETCH"); return !(e && e[0] == '0'); }();
  g.res_prefetch = (res_pf_on && p.residual != nullptr && g.ntn == 1 && (p.O & 3) == 0 && ((uintptr_t)p.residual & 15) == 0) ? 1 : 0;
  { const char* e = getenv("ATTNDM_TRACE_CTA"); g.trace_cta = e ? atoi(e) : 0; }
  g.acc_stride = round_up(g.BN, 32);
  // Four accumulators when they fit: the MMA warp may then run up to three tiles ahead of the epilogue, so the
  // latency of the epilogue -> MMA hand-back (and every other per-tile handshake) hides behind queued work.
  { const char* e = getenv("ATTNDM_TC_NACC"); g.nacc = (4 * g.acc_stride <= 512 && !(e && e[0] == '2')) ? 4 : 2; }
  g.nacc_shift = g.nacc == 4 ? 2 : 1;
  { const char* e = getenv("ATTNDM_TC_SPLIT"); g.split = (g.ntn == 1 && g.nacc == 4 && !(e && e[0] == '0')) ? 1 : 0; }
  g.tmem_cols = 32;
  while (g.tmem_cols < g.nacc * g.acc_stride) g.tmem_cols <<= 1;
  const int smem = g.end_off + rs_bytes + 1024;
  CUtensorMap tmA, tmA2, tmB;
  const int hr1 = g.hr > 256 ? 256 : g.hr;
  int rc = make_map_2d(&tmA, p.codes, (uint64_t)p.Cp, (uint64_t)p.rows, TC_BK, (uint32_t)hr1);
  if (rc) return rc;
  tmA2 = tmA;
  if (g.hr > 256 && g.hr % 256 != 0) {
    rc = make_map_2d(&tmA2, p.codes, (uint64_t)p.Cp, (uint64_t)p.rows, TC_BK, (uint32_t)(g.hr % 256));
    if (rc) return rc;
  }
  g.perm = (p.O % 16 == 0) ? 1 : 0;
  if (g.perm) rc = make_map_b_perm(&tmB, p.qw, (uint64_t)p.taps * p.Cp, p.O, (uint32_t)(g.pair ? g.BN / 2 : g.BN));
  else rc = make_map_2d(&tmB, p.qw, (uint64_t)p.taps * p.Cp, (uint64_t)p.O, TC_BK, (uint32_t)(g.pair ? g.BN / 2 : g.BN));
  if (rc) return rc;
  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    auto raise = [](const void* fn) {
      if (attr_err == cudaSuccess) attr_err = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 206 * 1024);
    };
    raise((const void*)qconv_i8_halo_kernel<false, false, false>);
    raise((const void*)qconv_i8_halo_kernel<false, true, false>);
    raise((const void*)qconv_i8_halo_kernel<true, true, false>);
    raise((const void*)qconv_i8_halo_kernel<true, false, false>);
    raise((const void*)qconv_i8_halo_kernel<false, false, true>);
    raise((const void*)qconv_i8_halo_kernel<false, true, true>);
    raise((const void*)qconv_i8_halo_kernel<true, true, true>);
    raise((const void*)qconv_i8_halo_kernel<true, false, true>);
  });
  if (attr_err != cudaSuccess) { set_error("qconv_i8_halo: cannot raise dynamic smem: %s", cudaGetErrorString(attr_err)); return ATTNDM_ERR_CUDA; }
  const bool adds = p.residual != nullptr || p.temb != nullptr;
  // GroupNorm statistics of the output from the epilogue (quad order): every chunk must take the 128-bit path
  static const bool stats_on = [] { const char* e = getenv("ATTNDM_TC_STATS"); return !(e && e[0] == '0'); }();
  const bool stats = stats_on && p.gn_out != nullptr && conv_gn_quad_ok(p) && g.BN % 32 == 0;
#define ATTNDM_HALO_ARGS dim3(grid), dim3(TC_THREADS_H), smem, st
#define ATTNDM_HALO_PARAMS tmA, tmA2, tmB, p, g
#define ATTNDM_HALO_LAUNCH(STATSV)                                                                                  \
  do {                                                                                                              \
    if (g.pair) {                                                                                                   \
      if (adds) launch_pdl_cluster(qconv_i8_halo_kernel<true, true, STATSV>, ATTNDM_HALO_ARGS, 2, ATTNDM_HALO_PARAMS);   \
      else      launch_pdl_cluster(qconv_i8_halo_kernel<false, true, STATSV>, ATTNDM_HALO_ARGS, 2, ATTNDM_HALO_PARAMS);  \
    } else {                                                                                                        \
      if (adds) launch_pdl(qconv_i8_halo_kernel<true, false, STATSV>, ATTNDM_HALO_ARGS, ATTNDM_HALO_PARAMS);        \
      else      launch_pdl(qconv_i8_halo_kernel<false, false, STATSV>, ATTNDM_HALO_ARGS, ATTNDM_HALO_PARAMS);       \
    }                                                                                                               \
  } while (0)
  if (stats) ATTNDM_HALO_LAUNCH(true);
  else ATTNDM_HALO_LAUNCH(false);
  if (stats_fused) *stats_fused = stats;
#undef ATTNDM_HALO_LAUNCH
#undef ATTNDM_HALO_ARGS
#undef ATTNDM_HALO_PARAMS
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("qconv_i8_halo: launch failed: %s", cudaGetErrorString(e)); return ATTNDM_ERR_CUDA; }
  return 1;
}

int launch_qconv_i8_tc(const ConvI8Params& p, cudaStream_t st, bool* stats_fused) {
  if (stats_fused) *stats_fused = false;
  ATTNDM_CHECK_ARG(((uintptr_t)p.codes & 15) == 0 && ((uintptr_t)p.qw & 15) == 0, "qconv_i8_tc: operands must be 16-byte aligned");
  ATTNDM_CHECK_ARG(p.rows + 2LL * p.Wp + 2 + TC_BM < 0x7fffffffLL, "qconv_i8_tc: too many rows for 32-bit TMA coordinates");
  if (tc_halo_enabled()) {
    int rc = launch_qconv_i8_halo(p, st, stats_fused);
    if (rc < 0) return rc;
    if (rc == 1) return ATTNDM_OK;
  }
  return launch_qconv_i8_tc_persistent(p, st);     // shapes the halo kernel does not take (halo > 512 rows, >= 2^31 outputs)
}


// =====================================================================================================
// fp32 GEMM on the tensor cores with fp32-level accuracy ("3xTF32"): out[M][N] = x[M][K] . w[N][K]^T + bias
// Reference op: the lazily created fp32 `channel_proj` 1x1 conv of UpBlock (models/diffusion.py:235-242), the one
// fp32 layer of the step that is large (M = B*H*W = 16384, K = 768, N = 512 on CIFAR: 190 us on the FP32 pipe,
// at 95 % of its peak).
// kind::tf32 keeps 10 mantissa bits, so each operand is split exactly, a = big + small with
//   big = a with its 13 low mantissa bits cleared,  small = a - big  (exact, <= 13 significant bits),
// and the product is accumulated as big*big + big*small + small*big in the fp32 TMEM accumulator; what is
// dropped (small*small and the bits of `small` below tf32) is ~2^-21 relative per product, the size of an fp32
// rounding.  The caller owns the split operands (weights are split once, activations by a streaming pass).
// Kernel: persistent 128 x 128 tiles, TMA (SWIZZLE_128B, 32 floats per row) -> 3-stage ring of
// {A big, A small, B big, B small} -> 12 tcgen05.mma per k-block -> four TMEM accumulators -> 4 epilogue warps.
// =====================================================================================================
constexpr int G32_STAGES = 3;
constexpr int G32_TILE_BYTES = 128 * 128;                 // 128 rows x 32 floats
constexpr int G32_STAGE_BYTES = 4 * G32_TILE_BYTES;
constexpr int G32_NACC = 4;                               // TMEM accumulators per tile (4 x 128 columns)
constexpr int G32_THREADS = 192;                          // producer, MMA issuer, 4 epilogue warps

struct GemmTf32Params {
  const float* bias;
  float* out;
  int M, N, K;
  int ntn, nkb;
  long long ntiles;
};

__global__ void f32_split_tf32_kernel(const float4* __restrict__ x, long long n4, float4* __restrict__ big,
                                      float4* __restrict__ small) {
  pdl_enter();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = ldg_stream(x + i);
    float4 b, s;
    b.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); s.x = __fsub_rn(v.x, b.x);
    b.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); s.y = __fsub_rn(v.y, b.y);
    b.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); s.z = __fsub_rn(v.z, b.z);
    b.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); s.w = __fsub_rn(v.w, b.w);
    big[i] = b;
    small[i] = s;
  }
}

__device__ __forceinline__ void umma_tf32_if(uint32_t leader, uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(leader)
      : "memory");
}

__global__ void __launch_bounds__(G32_THREADS, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap tmAb, const __grid_constant__ CUtensorMap tmAs,
                   const __grid_constant__ CUtensorMap tmBb, const __grid_constant__ CUtensorMap tmBs,
                   const GemmTf32Params p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full[G32_STAGES], empty[G32_STAGES], tmem_full_bar, tmem_empty_bar;
  __shared__ uint32_t tmem_base_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else if (warp == 1 && lane == 0) {
    for (int i = 0; i < G32_STAGES; ++i) { mbar_init(smem_u32(&full[i]), 1); mbar_init(smem_u32(&empty[i]), 1); }
    mbar_init(smem_u32(&tmem_full_bar), 1);
    mbar_init(smem_u32(&tmem_empty_bar), 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  // The tensor core adds into its fp32 accumulator with truncation, so the error of one accumulator grows with
  // the number of MMAs chained into it (measured 5.7e-6 of the output range for K = 768 in ONE accumulator,
  // against 1.1e-6 for sequential fmaf).  The k-blocks are therefore dealt round-robin to G32_NACC accumulators
  // (four times shorter chains) which the epilogue adds with round-to-nearest.
  const int nacc = p.nkb < G32_NACC ? p.nkb : G32_NACC;
  pdl_launch_dependents();
  if (warp == 0) {
    pdl_wait();
    int s = 0;
    uint32_t ph = 0;
    for (long long tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      const int m0 = (int)(tile / p.ntn) * 128, n0 = (int)(tile % p.ntn) * 128;
      for (int kb = 0; kb < p.nkb; ++kb) {
        mbar_wait_relaxed(smem_u32(&empty[s]), ph ^ 1);
        const uint32_t bar = smem_u32(&full[s]);
        const uint32_t dst = base + (uint32_t)s * G32_STAGE_BYTES;
        mbar_expect_tx_elect(bar, (uint32_t)G32_STAGE_BYTES);
        tma_load_2d_elect(dst, &tmAb, bar, kb * 32, m0);
        tma_load_2d_elect(dst + G32_TILE_BYTES, &tmAs, bar, kb * 32, m0);
        tma_load_2d_elect(dst + 2 * G32_TILE_BYTES, &tmBb, bar, kb * 32, n0);
        tma_load_2d_elect(dst + 3 * G32_TILE_BYTES, &tmBs, bar, kb * 32, n0);
        if (++s == G32_STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // c = f32, a = b = tf32, K-major, N = 128, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t desc_hi = ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    const uint32_t leader = elect_one();
    int s = 0, it = 0;
    uint32_t ph = 0;
    for (long long tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      mbar_wait(smem_u32(&tmem_empty_bar), (uint32_t)((it & 1) ^ 1));
      tcgen05_fence_after();
      for (int kb = 0; kb < p.nkb; ++kb) {
        mbar_wait(smem_u32(&full[s]), ph);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)((kb % G32_NACC) * 128);
        uint32_t accumulate = kb >= G32_NACC ? 1u : 0u;
        const uint32_t st = base + (uint32_t)s * G32_STAGE_BYTES;
        const uint64_t ab = desc_hi | (uint64_t)((st >> 4) & 0x3FFF), as = desc_hi | (uint64_t)(((st + G32_TILE_BYTES) >> 4) & 0x3FFF);
        const uint64_t bb = desc_hi | (uint64_t)(((st + 2 * G32_TILE_BYTES) >> 4) & 0x3FFF),
                       bs = desc_hi | (uint64_t)(((st + 3 * G32_TILE_BYTES) >> 4) & 0x3FFF);
#pragma unroll
        for (int k = 0; k < 4; ++k) {                    // 8 floats (32 bytes) of K per MMA
          umma_tf32_if(leader, d_tmem, as + 2 * k, bb + 2 * k, idesc, accumulate);   // small terms first
          umma_tf32_if(leader, d_tmem, ab + 2 * k, bs + 2 * k, idesc, 1u);
          umma_tf32_if(leader, d_tmem, ab + 2 * k, bb + 2 * k, idesc, 1u);
          accumulate = 1;
        }
        tcgen05_commit_if(leader, smem_u32(&empty[s]));
        if (++s == G32_STAGES) { s = 0; ph ^= 1; }
      }
      tcgen05_commit_if(leader, smem_u32(&tmem_full_bar));
    }
  } else {
    pdl_wait();
    const int quarter = warp & 3;
    const int tq = lane & 3, tr = lane >> 2;
    const bool odd = tq & 1;
    const int col4 = odd ? 8 + 2 * (tq - 1) : 2 * tq;
    int it = 0;
    for (long long tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int m0 = (int)(tile / p.ntn) * 128, n0 = (int)(tile % p.ntn) * 128;
      mbar_wait_relaxed(smem_u32(&tmem_full_bar), (uint32_t)(it & 1));
      tcgen05_fence_after();
      const uint32_t t_acc = tmem_base + ((uint32_t)(quarter * 32) << 16);
      for (int ci = 0; ci < 4; ++ci) {
        const int c0 = ci << 5;
        float f0[16], f1[16];
        for (int a = 0; a < nacc; ++a) {
          uint32_t v0[16], v1[16];
          __syncwarp();
          tmem_ld_16x256b_x4(t_acc + (uint32_t)(a * 128 + c0), v0);
          tmem_ld_16x256b_x4(t_acc + (16u << 16) + (uint32_t)(a * 128 + c0), v1);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            f0[j] = a == 0 ? __uint_as_float(v0[j]) : __fadd_rn(f0[j], __uint_as_float(v0[j]));
            f1[j] = a == 0 ? __uint_as_float(v1[j]) : __fadd_rn(f1[j], __uint_as_float(v1[j]));
          }
        }
        if (ci == 3) {
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&tmem_empty_bar));
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int col = n0 + c0 + 16 * h + col4;
          const float4 bi = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int j0 = ((2 * h) << 2) | ((k & 1) << 1), j1 = ((2 * h + 1) << 2) | ((k & 1) << 1);
            const float lo0 = k < 2 ? f0[j0] : f1[j0], lo1 = k < 2 ? f0[j0 | 1] : f1[j0 | 1];
            const float hi0 = k < 2 ? f0[j1] : f1[j1], hi1 = k < 2 ? f0[j1 | 1] : f1[j1 | 1];
            const float s0 = odd ? lo0 : hi0, s1 = odd ? lo1 : hi1;
            const float g0 = __shfl_xor_sync(0xffffffffu, s0, 1), g1 = __shfl_xor_sync(0xffffffffu, s1, 1);
            float4 o = odd ? make_float4(g0, g1, hi0, hi1) : make_float4(lo0, lo1, g0, g1);
            o.x = __fadd_rn(o.x, bi.x); o.y = __fadd_rn(o.y, bi.y); o.z = __fadd_rn(o.z, bi.z); o.w = __fadd_rn(o.w, bi.w);
            const int row = m0 + quarter * 32 + tr + 8 * k;
            if (row < p.M) *reinterpret_cast<float4*>(p.out + (long long)row * p.N + col) = o;
          }
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

static int make_map_2d_f32(CUtensorMap* m, const void* base, uint64_t inner, uint64_t outer) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("gemm_tf32: cuTensorMapEncodeTiled not available"); return ATTNDM_ERR_CUDA; }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {inner * 4};
  cuuint32_t box[2] = {32, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("gemm_tf32: cuTensorMapEncodeTiled failed (%d)", (int)r); return ATTNDM_ERR_CUDA; }
  return ATTNDM_OK;
}

int conv_f32_tc_fits(long long rows, int C, int O) {
  return rows >= 1 && rows < (1LL << 31) - 256 && C >= 64 && (C & 3) == 0 && (O & 127) == 0 && O >= 128 ? 1 : 0;
}

int launch_split_tf32(const float* x, long long n, float* big, float* small, cudaStream_t st) {
  ATTNDM_CHECK_ARG(x && big && small && n > 0 && (n & 3) == 0, "split_tf32: n must be a positive multiple of 4");
  ATTNDM_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)big & 15) == 0 && ((uintptr_t)small & 15) == 0,
                   "split_tf32: operands must be 16-byte aligned");
  const long long n4 = n / 4;
  const long long blocks = cdiv(n4, 256);
  launch_pdl(f32_split_tf32_kernel, dim3((unsigned)(blocks < 4 * kNumSMs ? blocks : 4 * kNumSMs)), dim3(256), 0, st,
             reinterpret_cast<const float4*>(x), n4, reinterpret_cast<float4*>(big), reinterpret_cast<float4*>(small));
  ATTNDM_CUDA_LAUNCH_CHECK("split_tf32");
  return ATTNDM_OK;
}

int launch_gemm_tf32x3(const float* a_big, const float* a_small, long long rows, int C, const float* w_big,
                       const float* w_small, int O, const float* bias, float* out, cudaStream_t st) {
  if (!conv_f32_tc_fits(rows, C, O)) { set_error("gemm_tf32x3: shape %lld x %d -> %d not supported", rows, C, O); return ATTNDM_ERR_UNSUPPORTED; }
  ATTNDM_CHECK_ARG(((uintptr_t)a_big & 15) == 0 && ((uintptr_t)a_small & 15) == 0 && ((uintptr_t)w_big & 15) == 0 &&
                   ((uintptr_t)w_small & 15) == 0 && ((uintptr_t)out & 15) == 0 && (!bias || ((uintptr_t)bias & 15) == 0),
                   "gemm_tf32x3: operands must be 16-byte aligned");
  CUtensorMap tmAb, tmAs, tmBb, tmBs;
  int rc = make_map_2d_f32(&tmAb, a_big, (uint64_t)C, (uint64_t)rows); if (rc) return rc;
  rc = make_map_2d_f32(&tmAs, a_small, (uint64_t)C, (uint64_t)rows); if (rc) return rc;
  rc = make_map_2d_f32(&tmBb, w_big, (uint64_t)C, (uint64_t)O); if (rc) return rc;
  rc = make_map_2d_f32(&tmBs, w_small, (uint64_t)C, (uint64_t)O); if (rc) return rc;
  GemmTf32Params p;
  p.bias = bias; p.out = out; p.M = (int)rows; p.N = O; p.K = C;
  p.ntn = O / 128; p.nkb = cdiv(C, 32);
  p.ntiles = (long long)cdiv(rows, 128) * p.ntn;
  const int smem = G32_STAGES * G32_STAGE_BYTES + 1024;
  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [smem] { attr_err = cudaFuncSetAttribute(gemm_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); });
  if (attr_err != cudaSuccess) { set_error("gemm_tf32x3: cannot raise dynamic smem: %s", cudaGetErrorString(attr_err)); return ATTNDM_ERR_CUDA; }
  const int grid = (int)(p.ntiles < kNumSMs ? p.ntiles : kNumSMs);
  launch_pdl(gemm_tf32x3_kernel, dim3(grid), dim3(G32_THREADS), smem, st, tmAb, tmAs, tmBb, tmBs, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("gemm_tf32x3: launch failed: %s", cudaGetErrorString(e)); return ATTNDM_ERR_CUDA; }
  return ATTNDM_OK;
}

}  // namespace attndm
