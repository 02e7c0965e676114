// Micro-benchmark (GPU box): per-SM global store throughput.  Each CTA streams 128-bit stores over its own
// region (region size chosen L2-resident or HBM-streaming); prints bytes/clk/SM and GB/s for several grid sizes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/store_rate_test tools/store_rate_test.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__global__ void __launch_bounds__(256) st(float4* base, size_t vec_per_cta, int reps, long long* out, int mode) {
  float4* mine = base + (size_t)blockIdx.x * vec_per_cta;
  unsigned long long g0, g1;
  __syncthreads();
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0) :: "memory");
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r)
    for (size_t i = threadIdx.x; i < vec_per_cta; i += 256 * 4) {
      // mode 0: each warp instruction writes 512 contiguous bytes; mode 1: 8 rows x 64 B (rows 512 B apart)
      if (mode == 0) {
#pragma unroll
        for (int u = 0; u < 4; ++u) if (i + u * 256 < vec_per_cta) mine[i + u * 256] = make_float4(1.f, 2.f, 3.f, (float)r);
      } else {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const size_t blk = (i - threadIdx.x) + warp * 128;         // 128 float4 = 8 rows x 8 columns x 2 halves
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const size_t idx = blk + (size_t)(lane >> 2) * 32 + (lane & 3) + 4 * u + (warp & 1) * 16;
          if (idx < vec_per_cta) mine[idx] = make_float4(1.f, 2.f, 3.f, (float)r);
        }
      }
    }
  __syncthreads();
  long long t1 = clock64();
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1) :: "memory");
  if (threadIdx.x == 0) { out[2 * blockIdx.x] = t1 - t0; out[2 * blockIdx.x + 1] = (long long)(g1 - g0); }
}
int main() {
  long long* d; cudaMalloc(&d, 148 * 16);
  const size_t total = (size_t)2 << 30;
  float4* buf; cudaMalloc(&buf, total);
  for (int mode = 0; mode < 2; ++mode)
    for (int grid : {1, 8, 37, 74, 148})
      for (size_t kb_per_cta : {(size_t)256, (size_t)8192}) {
        const size_t vec = kb_per_cta * 1024 / 16;
        const int reps = kb_per_cta == 256 ? 64 : 2;
        st<<<grid, 256>>>(buf, vec, reps, d, mode);
        st<<<grid, 256>>>(buf, vec, reps, d, mode);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
        long long h[296]; cudaMemcpy(h, d, grid * 16, cudaMemcpyDeviceToHost);
        long long mc = 0, mn = 0;
        for (int i = 0; i < grid; ++i) { if (h[2 * i] > mc) mc = h[2 * i]; if (h[2 * i + 1] > mn) mn = h[2 * i + 1]; }
        const double bytes = (double)kb_per_cta * 1024 * reps;
        printf("mode=%d grid=%3d region=%5zu KB/CTA: %6.1f B/clk/SM  %7.1f GB/s/SM  total %7.0f GB/s  (%.0f MHz)\n", mode, grid,
               kb_per_cta, bytes / mc, bytes / mn, bytes * grid / mn, (double)mc / mn * 1e3);
      }
  return 0;
}
