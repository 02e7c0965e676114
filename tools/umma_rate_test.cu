// Micro-benchmark (GPU box): issue-to-completion rate of back-to-back tcgen05.mma instructions from one
// CTA, for kind::i8 / kind::f8f6f4 / kind::f16, N in {16,128,256}, A start aligned vs shifted by one row.
// Operands are whatever is in shared memory (values do not matter for timing).  Prints cycles per MMA and
// the implied MAC/clk/SM.  One CTA; run on all SMs concurrently with gridDim = 148 for the chip number.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/umma_rate_test tools/umma_rate_test.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <int KIND>   // 0 = i8, 1 = f8f6f4 (e4m3), 2 = f16 (bf16)
__global__ void __launch_bounds__(128) rate(int N, int shift_rows, int iters, long long* out, int walk) {
  extern __shared__ uint8_t raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(raw + (base - smem_u32(raw)))[i] = 0x01010101u * (i & 3);
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
    // instruction descriptor
    uint32_t idesc;
    if (KIND == 0) idesc = (2u << 4) | (1u << 7) | (1u << 10);          // s32 accum, s8 x s8
    else if (KIND == 1) idesc = (1u << 4) | (0u << 7) | (0u << 10);     // f32 accum, e4m3 x e4m3
    else idesc = (1u << 4) | (1u << 7) | (1u << 10);                    // f32 accum, bf16 x bf16
    idesc |= ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a00 = base + (uint32_t)shift_rows * 128, b00 = base + 48 * 1024;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      // walk != 0: every k-block uses a different weight tile (36 x N*128 B region) and a different A row
      // shift, like the nine taps of the conv kernel
      const int kb = walk ? (it % 9) : 0;
      const uint32_t a0 = a00 + (walk ? (uint32_t)((kb / 3) * 34 + kb % 3) * 128 : 0u);
      const uint32_t b0 = b00 + (walk ? (uint32_t)kb * (uint32_t)(N * 128) : 0u);
      for (int k = 0; k < 4; ++k) {
        const uint64_t ad = desc_sw128(a0 + k * 32), bd = desc_sw128(b0 + k * 32);
        const uint32_t acc = (it | k) ? 1u : 0u;
        if (KIND == 0)
          asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t@q tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tm), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
        else if (KIND == 1)
          asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t@q tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(tm), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
        else
          asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tm), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
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
  const int smem = 200 * 1024 + 2048, iters = 2000;
  cudaFuncSetAttribute(rate<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(rate<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(rate<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* names[3] = {"i8 (K=32)", "f8f6f4 e4m3 (K=32)", "f16 bf16 (K=16)"};
  for (int grid : {1, 148}) {
    for (int kind = 0; kind < 3; ++kind) {
      for (int N : {16, 128}) {
        for (int sh : {0, 2}) {
          const int walk = sh == 2;
          if (kind == 0) rate<0><<<grid, 128, smem>>>(N, 0, iters, d, walk);
          if (kind == 1) rate<1><<<grid, 128, smem>>>(N, 0, iters, d, walk);
          if (kind == 2) rate<2><<<grid, 128, smem>>>(N, 0, iters, d, walk);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("%s N=%d shift=%d: CUDA error %s\n", names[kind], N, sh, cudaGetErrorString(e)); return 1; }
          long long h[148]; cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
          long long mx = 0; for (int i = 0; i < grid; ++i) if (h[i] > mx) mx = h[i];
          double cyc = (double)mx / (iters * 4);
          int K = kind == 2 ? 16 : 32;
          printf("grid=%3d %-20s N=%3d walk=%d: %7.1f cyc/MMA  -> %7.0f MAC/clk/SM\n", grid, names[kind], N, sh, cyc, 128.0 * N * K / cyc);
        }
      }
    }
  }
  return 0;
}
