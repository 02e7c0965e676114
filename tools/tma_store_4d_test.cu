// Functional check (GPU box): 4-D TMA tensor store with a box taller than the x extent, negative start
// coordinates and clipping -- the addressing the conv epilogue relies on.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/tma_store_4d_test tools/tma_store_4d_test.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap tm, int c, int x, int y, int b) {
  __shared__ __align__(128) float stage[32 * 16];
  for (int i = threadIdx.x; i < 512; i += 32) stage[i] = 1000.f * (i / 16) + (i % 16);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (threadIdx.x == 0) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(&tm), "r"(smem_u32(stage)), "r"(c), "r"(x), "r"(y), "r"(b) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}
int main() {
  const int O = 128, W = 16, H = 16, B = 4;
  float* buf; cudaMalloc(&buf, (size_t)B * H * W * O * 4); cudaMemset(buf, 0, (size_t)B * H * W * O * 4);
  CUtensorMap tm;
  cuuint64_t dims[4] = {O, W, H, B};
  cuuint64_t strides[3] = {(cuuint64_t)O * 4, (cuuint64_t)W * O * 4, (cuuint64_t)H * W * O * 4};
  cuuint32_t box[4] = {16, 32, 1, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, buf, dims, strides, box, es,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                      CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d\n", (int)r);
  if (r != CUDA_SUCCESS) return 1;
  const int cases[3][4] = {{32, 5, 3, 1}, {48, -10, 7, 2}, {0, 0, 15, 3}};
  for (auto& cs : cases) {
    cudaMemset(buf, 0, (size_t)B * H * W * O * 4);
    k<<<1, 32>>>(tm, cs[0], cs[1], cs[2], cs[3]);
    cudaError_t e = cudaDeviceSynchronize();
    printf("case c=%d x=%d y=%d b=%d: %s\n", cs[0], cs[1], cs[2], cs[3], cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<float> h((size_t)B * H * W * O);
    cudaMemcpy(h.data(), buf, h.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0, written = 0;
    for (int b = 0; b < B; ++b) for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) for (int c = 0; c < O; ++c) {
      const float v = h[(((size_t)b * H + y) * W + x) * O + c];
      const int i = x - cs[1], j = c - cs[0];
      const bool in = b == cs[3] && y == cs[2] && i >= 0 && i < 32 && j >= 0 && j < 16;
      const float want = in ? 1000.f * i + j : 0.f;
      if (v != want) ++bad;
      if (v != 0.f) ++written;
    }
    printf("  mismatches %d, elements written %d\n", bad, written);
  }
  return 0;
}
