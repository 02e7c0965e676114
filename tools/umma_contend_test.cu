// Micro-benchmark (GPU box): how much do other warps of the same CTA slow a saturated tcgen05.mma stream?
// Warp 1 issues back-to-back kind::i8 128x128x32 MMAs (A and B from shared memory, walking over nine taps
// like the conv kernel); warps 4..11 run one kind of interference until it finishes.  Prints cycles per
// MMA (clock64), ns per MMA (globaltimer) and the implied SM clock.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/umma_contend_test tools/umma_contend_test.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}

enum { M_NONE, M_ALU, M_STG, M_LDS, M_TMEM, M_SHFL, M_LDG, M_STG64, M_I2F, M_IMAD, M_FFMA, M_FSEL, M_I2F_OFF, M_NMODES };
static const char* mode_names[M_NMODES] = {"none", "alu (imad/ffma)", "global stores 128-bit", "shared loads", "tcgen05.ld",
                                           "shuffles", "global loads 128-bit", "global stores 64-bit strided", "i2f", "imad only", "ffma only", "fsel/iadd mix", "i2f, not on the issuer SMSP"};

__global__ void __launch_bounds__(384) contend(int mode, int iters, int nwarps_active, long long* out, float* scratch,
                                               size_t scratch_per_cta) {
  extern __shared__ uint8_t raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  __shared__ volatile int done_flag;
  __shared__ float lut[1024];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(raw + (base - smem_u32(raw)))[i] = 0x01010101u * (i & 3);
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) lut[i] = (float)i;
  if (threadIdx.x == 0) done_flag = 0;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
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
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a00 = base, b00 = base + 48 * 1024;
    unsigned long long g0, g1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int kb = it % 9;
      const uint32_t a0 = a00 + (uint32_t)((kb / 3) * 34 + kb % 3) * 128;
      const uint32_t b0 = b00 + (uint32_t)kb * (uint32_t)(128 * 128);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t ad = desc_sw128(a0 + k * 32), bd = desc_sw128(b0 + k * 32);
        const uint32_t acc = (it | k) ? 1u : 0u;
        asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t@q tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tm), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
      }
    }
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    long long t1 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    if (lane == 0) { out[2 * blockIdx.x] = t1 - t0; out[2 * blockIdx.x + 1] = (long long)(g1 - g0); done_flag = 1; }
  } else if (warp >= 4 && warp < 4 + nwarps_active && mode != M_NONE) {
    const int w = warp - 4, t = w * 32 + lane;
    float* mine = scratch + (size_t)blockIdx.x * scratch_per_cta;
    const size_t nvec = scratch_per_cta / 4;     // float4 elements
    float facc = 0.f;
    int iacc = lane;
    size_t pos = t;
    while (!done_flag) {
      if (mode == M_ALU) {
#pragma unroll
        for (int i = 0; i < 64; ++i) { iacc = iacc * 3 + i; facc = fmaf(facc, 1.0001f, (float)(iacc & 1)); }
      } else if (mode == M_IMAD) {
        int b = iacc ^ 5, c = iacc + 9, d = iacc * 7;
#pragma unroll
        for (int i = 0; i < 32; ++i) { iacc = iacc * 3 + b; b = b * 5 + c; c = c * 7 + d; d = d * 9 + iacc; }
        iacc += b + c + d;
      } else if (mode == M_FFMA) {
        float b = facc + 1.f, c = facc + 2.f, d = facc + 3.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) { facc = fmaf(facc, 1.0001f, b); b = fmaf(b, 1.0002f, c); c = fmaf(c, 1.0003f, d); d = fmaf(d, 1.0004f, facc); }
        facc += b + c + d;
      } else if (mode == M_FSEL) {
        float b = facc + 1.f; int c = iacc + 1;
#pragma unroll
        for (int i = 0; i < 32; ++i) { facc = (c & 1) ? facc : b; c += iacc; b = (c & 2) ? b : facc; iacc += c; }
        facc += b; iacc += c;
      } else if (mode == M_I2F_OFF) {
        if ((warp & 3) != 1) {
#pragma unroll
          for (int i = 0; i < 64; ++i) { iacc += i; facc += (float)iacc; }
        } else {
          __nanosleep(1000);
        }
      } else if (mode == M_I2F) {
#pragma unroll
        for (int i = 0; i < 64; ++i) { iacc += i; facc += (float)iacc; }
      } else if (mode == M_STG) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          reinterpret_cast<float4*>(mine)[pos] = make_float4(facc, 1.f, 2.f, 3.f);
          pos += 256; if (pos >= nvec) pos = t;
        }
      } else if (mode == M_STG64) {
        // the fragment-layout pattern: 8 rows x 32 bytes per warp instruction, rows 512 B apart
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const size_t row = (pos >> 5) * 8 + (lane >> 2);
          reinterpret_cast<float2*>(mine)[row * 64 + (lane & 3) + 4 * (i & 3)] = make_float2(facc, 1.f);
          if ((i & 3) == 3) { pos += 256; if (pos * 8 + 2048 >= nvec * 2 / 64 * 8) pos = t; }
        }
      } else if (mode == M_LDG) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(mine) + pos);
          facc += v.x + v.w;
          pos += 256; if (pos >= nvec) pos = t;
        }
      } else if (mode == M_LDS) {
#pragma unroll
        for (int i = 0; i < 64; ++i) facc += lut[(iacc + i * 4 + (lane & 3)) & 1023];
        iacc += 7;
      } else if (mode == M_SHFL) {
#pragma unroll
        for (int i = 0; i < 64; ++i) facc += __shfl_xor_sync(0xffffffffu, facc, 1);
      } else if (mode == M_TMEM) {
        uint32_t v[16];
        const uint32_t addr = tm + 256u + ((uint32_t)((warp & 3) * 32) << 16);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                       : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                         "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                       : "r"(addr + (uint32_t)(i * 32)));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          iacc += (int)v[0] + (int)v[15];
        }
      }
    }
    if (facc == 123.456f || iacc == 0x7fffffff) mine[t] = facc + iacc;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}

int main() {
  long long* d; cudaMalloc(&d, 148 * 16);
  const size_t per_cta = 4u << 20;            // floats per CTA (16 MB): streams through HBM, not L2
  float* scratch; cudaMalloc(&scratch, 148 * per_cta * sizeof(float));
  cudaMemset(scratch, 0, 148 * per_cta * sizeof(float));
  const int smem = 200 * 1024 + 2048, iters = 4000;
  cudaFuncSetAttribute(contend, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep)
    for (int mode = 0; mode < M_NMODES; ++mode) {
      for (int nw : {8, 2}) {
        if (mode == M_NONE && nw != 8) continue;
        contend<<<148, 384, smem>>>(mode, iters, nw, d, scratch, per_cta);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d: CUDA error %s\n", mode, cudaGetErrorString(e)); return 1; }
        long long h[296]; cudaMemcpy(h, d, 148 * 16, cudaMemcpyDeviceToHost);
        long long mc = 0, mn = 0;
        for (int i = 0; i < 148; ++i) { if (h[2 * i] > mc) mc = h[2 * i]; if (h[2 * i + 1] > mn) mn = h[2 * i + 1]; }
        const double cyc = (double)mc / (iters * 4), ns = (double)mn / (iters * 4);
        if (rep == 1)
          printf("%-30s warps=%d: %6.1f cyc/MMA  %6.1f ns/MMA  (%.0f MHz)\n", mode_names[mode], nw, cyc, ns, cyc / ns * 1e3);
      }
    }
  return 0;
}
