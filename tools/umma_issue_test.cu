// Micro-benchmark (GPU box): issue cost of different ways to write the tcgen05.mma issue loop.
// One CTA per SM, warp 1 issues `iters` x 36 MMAs (kind::i8, M=128, N = 16 so the tensor pipe needs only ~39
// cycles per MMA and the ISSUE rate shows), operands walk like the conv kernel's nine taps.
//   form 0: C-level loop, 64-bit descriptors built per MMA, elect.sync inside every asm (umma_rate_test)
//   form 1: the conv kernel's block of four MMAs under a per-lane leader predicate (one asm, additive low words)
//   form 2: one elected thread branches around the whole loop (no predicates on the MMAs)
//   form 3: like 1, but one predicate for all four MMAs
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/umma_issue_test tools/umma_issue_test.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t l;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(l));
  return l;
}
__device__ __forceinline__ void x4_if(uint32_t leader, uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate, int nk) {
  asm volatile(
      "{\n\t.reg .pred q, p, tr, k2, k3, k4;\n\t.reg .b64 ad, bd;\n\t.reg .b32 al, bl;\n\t"
      "setp.ne.b32 q, %5, 0;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.eq.u32 tr, 0, 0;\n\t"
      "setp.gt.and.s32 k2, %6, 1, q;\n\tsetp.gt.and.s32 k3, %6, 2, q;\n\tsetp.gt.and.s32 k4, %6, 3, q;\n\t"
      "mov.b64 ad, {%1, %7};\n\tmov.b64 bd, {%2, %7};\n\t@q tcgen05.mma.cta_group::1.kind::i8 [%0], ad, bd, %3, p;\n\t"
      "add.u32 al, %1, 2;\n\tadd.u32 bl, %2, 2;\n\tmov.b64 ad, {al, %7};\n\tmov.b64 bd, {bl, %7};\n\t@k2 tcgen05.mma.cta_group::1.kind::i8 [%0], ad, bd, %3, tr;\n\t"
      "add.u32 al, %1, 4;\n\tadd.u32 bl, %2, 4;\n\tmov.b64 ad, {al, %7};\n\tmov.b64 bd, {bl, %7};\n\t@k3 tcgen05.mma.cta_group::1.kind::i8 [%0], ad, bd, %3, tr;\n\t"
      "add.u32 al, %1, 6;\n\tadd.u32 bl, %2, 6;\n\tmov.b64 ad, {al, %7};\n\tmov.b64 bd, {bl, %7};\n\t@k4 tcgen05.mma.cta_group::1.kind::i8 [%0], ad, bd, %3, tr;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(leader), "r"(nk), "r"(0x40004040u) : "memory");
}
__device__ __forceinline__ void x4_one_pred(uint32_t leader, uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred q, p, tr;\n\t.reg .b64 ad, bd;\n\t.reg .b32 al, bl;\n\t"
      "setp.ne.b32 q, %5, 0;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.eq.u32 tr, 0, 0;\n\t"
      "mov.b64 ad, {%1, %6};\n\tmov.b64 bd, {%2, %6};\n\t@q tcgen05.mma.cta_group::1.kind::i8 [%0], ad, bd, %3, p;\n\t"
      "add.u32 al, %1, 2;\n\tadd.u32 bl, %2, 2;\n\tmov.b64 ad, {al, %6};\n\tmov.b64 bd, {bl, %6};\n\t@q tcgen05.mma.cta_group::1.kind::i8 [%0], ad, bd, %3, tr;\n\t"
      "add.u32 al, %1, 4;\n\tadd.u32 bl, %2, 4;\n\tmov.b64 ad, {al, %6};\n\tmov.b64 bd, {bl, %6};\n\t@q tcgen05.mma.cta_group::1.kind::i8 [%0], ad, bd, %3, tr;\n\t"
      "add.u32 al, %1, 6;\n\tadd.u32 bl, %2, 6;\n\tmov.b64 ad, {al, %6};\n\tmov.b64 bd, {bl, %6};\n\t@q tcgen05.mma.cta_group::1.kind::i8 [%0], ad, bd, %3, tr;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(leader), "r"(0x40004040u) : "memory");
}
__device__ __forceinline__ void mma_plain(uint32_t tmem_d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(ad), "l"(bd), "r"(idesc), "r"(accumulate) : "memory");
}
__global__ void __launch_bounds__(128) issue(int form, int N, int iters, int Wp, long long* out) {
  extern __shared__ uint8_t raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(raw + (base - smem_u32(raw)))[i] = 0x01010101u;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 32) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = slot;
  if (warp == 1) {
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t desc_hi = ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    const uint32_t a_buf0 = base, b_base = base + 48 * 1024;
    const uint32_t b_tile = (uint32_t)N * 128u;
    const uint32_t leader = elect_one();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      uint32_t accumulate = 0;
      if (form == 0) {
        for (int kb = 0; kb < 9; ++kb) {
          const uint32_t a0 = a_buf0 + (uint32_t)((kb / 3) * Wp + kb % 3) * 128, b0 = b_base + kb * b_tile;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ad = desc_hi | (uint64_t)(((a0 + k * 32) >> 4) & 0x3FFF), bd = desc_hi | (uint64_t)(((b0 + k * 32) >> 4) & 0x3FFF);
            asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t@q tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tm), "l"(ad), "l"(bd), "r"(idesc), "r"(accumulate) : "memory");
            accumulate = 1;
          }
        }
      } else if (form == 1 || form == 3) {
        const uint32_t a_lo0 = ((a_buf0 >> 4) & 0x3FFF) | 0x10000u;
        uint32_t b_lo = ((b_base >> 4) & 0x3FFF) | 0x10000u;
        for (int kh = 0; kh < 3; ++kh)
          for (int kw = 0; kw < 3; ++kw) {
            const uint32_t a_lo = a_lo0 + (uint32_t)(kh * Wp + kw) * 8;
            if (form == 1) x4_if(leader, tm, a_lo, b_lo, idesc, accumulate, 4);
            else x4_one_pred(leader, tm, a_lo, b_lo, idesc, accumulate);
            accumulate = 1;
            b_lo += b_tile >> 4;
          }
      } else {
        if (leader) {
          const uint32_t a_lo0 = (a_buf0 >> 4) & 0x3FFF;
          uint32_t b_lo = (b_base >> 4) & 0x3FFF;
          for (int kh = 0; kh < 3; ++kh)
            for (int kw = 0; kw < 3; ++kw) {
              const uint32_t a_lo = a_lo0 + (uint32_t)(kh * Wp + kw) * 8;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                mma_plain(tm, desc_hi | (uint64_t)(a_lo + 2 * k), desc_hi | (uint64_t)(b_lo + 2 * k), idesc, accumulate);
                accumulate = 1;
              }
              b_lo += b_tile >> 4;
            }
        }
        __syncwarp();
      }
    }
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(256u) : "memory");
}
int main() {
  long long* d; cudaMalloc(&d, 148 * 8);
  const int smem = 200 * 1024 + 2048, iters = 500;
  cudaFuncSetAttribute(issue, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int N : {16, 128})
    for (int form = 0; form < 4; ++form) {
      for (int rep = 0; rep < 2; ++rep) issue<<<148, 128, smem>>>(form, N, iters, 34, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("form %d: CUDA error %s\n", form, cudaGetErrorString(e)); return 1; }
      long long h[148]; cudaMemcpy(h, d, 148 * 8, cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < 148; ++i) if (h[i] > mx) mx = h[i];
      printf("N=%3d form %d: %6.1f cycles per MMA\n", N, form, (double)mx / (iters * 36.0));
    }
  return 0;
}
