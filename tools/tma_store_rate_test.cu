// Micro-benchmark (GPU box): per-SM throughput of small TMA tensor stores (cp.async.bulk.tensor shared -> global).
// Eight warps each own a staging slot in shared memory and repeatedly store a {COLS x 32 rows} fp32 box into a
// [npix][128] fp32 output (row pitch 512 B), waiting only for the previous store's shared-memory read
// (cp.async.bulk.wait_group.read) -- the pattern a conv epilogue would use.  Prints bytes/clk/SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/tma_store_rate_test tools/tma_store_rate_test.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int COLS>
__global__ void __launch_bounds__(256) k(const __grid_constant__ CUtensorMap tm, int rows_per_cta, int reps, long long* out) {
  extern __shared__ __align__(128) uint8_t stage[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* mine = reinterpret_cast<float*>(stage + warp * (COLS * 32 * 4));
  for (int i = lane; i < COLS * 32; i += 32) mine[i] = (float)i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  long long t0 = clock64();
  unsigned long long g0, g1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0) :: "memory");
  const int row0 = blockIdx.x * rows_per_cta;
  for (int r = 0; r < reps; ++r)
    for (int row = warp * 32; row < rows_per_cta; row += 8 * 32)
      for (int c = 0; c < 128; c += COLS) {
        if (lane == 0) {
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                       ::"l"(&tm), "r"(c), "r"(row0 + row), "r"(smem_u32(mine)) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        __syncwarp();
      }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  __syncthreads();
  long long t1 = clock64();
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1) :: "memory");
  if (threadIdx.x == 0) { out[2 * blockIdx.x] = t1 - t0; out[2 * blockIdx.x + 1] = (long long)(g1 - g0); }
}
int main() {
  const int rows_per_cta = 16384;          // 16384 rows x 512 B = 8 MB per CTA
  float* buf; cudaMalloc(&buf, (size_t)148 * rows_per_cta * 512);
  long long* d; cudaMalloc(&d, 148 * 16);
  for (int cols : {16, 32, 64}) {
    CUtensorMap tm;
    cuuint64_t dims[2] = {128, (cuuint64_t)148 * rows_per_cta};
    cuuint64_t strides[1] = {512};
    cuuint32_t box[2] = {(cuuint32_t)cols, 32};
    cuuint32_t es[2] = {1, 1};
    CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, buf, dims, strides, box, es,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                        CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    for (int grid : {1, 148}) {
      const int reps = 2, smem = 8 * cols * 32 * 4;
      for (int rep = 0; rep < 2; ++rep) {
        if (cols == 16) k<16><<<grid, 256, smem>>>(tm, rows_per_cta, reps, d);
        if (cols == 32) k<32><<<grid, 256, smem>>>(tm, rows_per_cta, reps, d);
        if (cols == 64) k<64><<<grid, 256, smem>>>(tm, rows_per_cta, reps, d);
      }
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
      long long h[296]; cudaMemcpy(h, d, grid * 16, cudaMemcpyDeviceToHost);
      long long mc = 0, mn = 0;
      for (int i = 0; i < grid; ++i) { if (h[2 * i] > mc) mc = h[2 * i]; if (h[2 * i + 1] > mn) mn = h[2 * i + 1]; }
      const double bytes = (double)rows_per_cta * 512 * reps;
      printf("box %2d cols x 32 rows (%d B) grid=%3d: %6.1f B/clk/SM  %6.1f GB/s/SM  total %6.0f GB/s  %.0f clk per store per SM\n", cols,
             cols * 128, grid, bytes / mc, bytes / mn, bytes * grid / mn, (double)mc / (bytes / (cols * 128)));
    }
  }
  return 0;
}
