// Micro-experiment (GPU box): can a tcgen05 K-major SWIZZLE_128B matrix descriptor start at a row that is
// NOT a multiple of 8 (i.e. not 1024-byte aligned) when the tile was written by TMA with the swizzle
// keyed on absolute shared-memory addresses?  This decides whether a 3x3 conv can load ONE halo tile
// and feed all nine taps from it by sliding the A descriptor (start += shift*128 B).
// For each shift we run D[128 x 16] = A[rows shift..shift+127][0..127] * B^T with (a) base_offset = 0 and
// (b) base_offset = (start_addr >> 7) & 7 and compare with the CPU.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/umma_shift_test tools/umma_shift_test.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 2; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done; long long t0 = clock64();
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (!done && clock64() - t0 > 2000000000LL) return false;
  } while (!done);
  return true;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}

constexpr int ROWS = 256, KB = 128, N = 16;

__global__ void __launch_bounds__(128) k(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                         int shift, int use_base_off, int* out, int* status) {
  extern __shared__ uint8_t raw[];
  __shared__ __align__(8) uint64_t bar_full, bar_mma;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t a_s = base, b_s = base + ROWS * KB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(32u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 32) {
    mbar_init(smem_u32(&bar_full), 1);
    mbar_init(smem_u32(&bar_mma), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_slot;
  bool ok = true;
  if (threadIdx.x == 0) {
    mbar_expect_tx(smem_u32(&bar_full), ROWS * KB + N * KB);
    tma_load_2d(a_s, &tmA, smem_u32(&bar_full), 0, 0);                 // rows 0..127
    tma_load_2d(a_s + 128 * KB, &tmA, smem_u32(&bar_full), 0, 128);    // rows 128..255, contiguous in smem
    tma_load_2d(b_s, &tmB, smem_u32(&bar_full), 0, 0);
    ok = mbar_wait(smem_u32(&bar_full), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a0 = a_s + (uint32_t)shift * KB;
    for (int kk = 0; kk < 4; ++kk) {
      const uint32_t aa = a0 + kk * 32, bb = b_s + kk * 32;
      const uint32_t bo = use_base_off ? ((aa >> 7) & 7) : 0;
      uint64_t ad = desc_sw128(aa, bo), bd = desc_sw128(bb, 0);
      uint32_t accum = kk > 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tm), "l"(ad), "l"(bd), "r"(idesc), "r"(accum) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_mma)) : "memory");
  }
  __syncwarp();
  bool ok2 = mbar_wait(smem_u32(&bar_mma), 0);
  if (!ok || !ok2) { if (threadIdx.x == 0) *status = 1; }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t v[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                 "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(tm + ((uint32_t)(warp * 32) << 16)) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * 16 + j] = (int)v[j];
  // layout probe: .16x256b.x2 (16 lanes x 16 columns -> 8 regs/thread), lane offsets 0 and 16 of this warp's window
  for (int half = 0; half < 2; ++half) {
    uint32_t w[8];
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                 : "r"(tm + ((uint32_t)(warp * 32 + half * 16) << 16)) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) out[128 * 16 + ((warp * 2 + half) * 32 + lane) * 8 + j] = (int)w[j];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(32u) : "memory");
}

typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* f = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
  EncFn enc = (EncFn)f;
  std::vector<int8_t> hA((ROWS + 128) * KB), hB(N * KB);
  srand(1);
  for (auto& x : hA) x = (int8_t)(rand() % 255 - 127);
  for (auto& x : hB) x = (int8_t)(rand() % 255 - 127);
  int8_t *dA, *dB; int *dO, *dS;
  CK(cudaMalloc(&dA, hA.size())); CK(cudaMalloc(&dB, hB.size())); CK(cudaMalloc(&dO, 128 * 16 * 4 * 2)); CK(cudaMalloc(&dS, 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
  CUtensorMap tA, tB;
  cuuint64_t dimsA[2] = {KB, (cuuint64_t)(ROWS + 128)}, strA[1] = {KB}; cuuint32_t boxA[2] = {KB, 128}, es[2] = {1, 1};
  cuuint64_t dimsB[2] = {KB, N}; cuuint32_t boxB[2] = {KB, N};
  if (enc(&tA, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, dA, dimsA, strA, boxA, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) ||
      enc(&tB, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, dB, dimsB, strA, boxB, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) {
    printf("encode failed\n"); return 2;
  }
  const int smem = ROWS * KB + N * KB + 2048;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int shifts[] = {0, 8, 16, 1, 2, 3, 7, 9, 34, 35, 36, 68, 69, 70, 127};
  std::vector<int> hO(128 * 16 * 2);
  for (int s : shifts) {
    for (int bo = 0; bo < 2; ++bo) {
      CK(cudaMemset(dO, 0xff, 128 * 16 * 4 * 2)); CK(cudaMemset(dS, 0, 4));
      k<<<1, 128, smem>>>(tA, tB, s, bo, dO, dS);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("shift %3d base_off=%d : CUDA error %s\n", s, bo, cudaGetErrorString(e)); return 3; }
      int st; CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost));
      int bad = 0;
      for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
        int ref = 0;
        for (int kk = 0; kk < KB; ++kk) ref += (int)hA[(m + s) * KB + kk] * (int)hB[n * KB + kk];
        if (ref != hO[m * 16 + n]) ++bad;
      }
      printf("shift %3d base_off=%d : %s (%d / %d wrong)%s\n", s, bo, bad ? "MISMATCH" : "ok", bad, 128 * N, st ? " [barrier timeout]" : "");
      if (s == 0 && bo == 0) {
        // deduce the 16x256b.x2 fragment layout: for warp 0, half 0: which (row, col) does (thread t, reg j) hold?
        int okmap = 1;
        for (int wh = 0; wh < 8 && okmap; ++wh)
          for (int t = 0; t < 32; ++t)
            for (int j = 0; j < 8; ++j) {
              int val = hO[128 * 16 + (wh * 32 + t) * 8 + j];
              int row = (wh / 2) * 32 + (wh % 2) * 16 + t / 4 + ((j >> 1) & 1) * 8;
              int col = (t % 4) * 2 + (j & 1) + (j >> 2) * 8;
              if (hO[row * 16 + col] != val) okmap = 0;
            }
        printf("16x256b.x2 layout hypothesis [row = t/4 + 8*((j>>1)&1), col = 2*(t%%4) + (j&1) + 8*(j>>2)] : %s\n", okmap ? "CONFIRMED" : "WRONG");
        if (!okmap) {
          for (int t = 0; t < 8; ++t) for (int j = 0; j < 8; ++j) {
            int val = hO[128 * 16 + t * 8 + j], fr = -1, fc = -1;
            for (int r = 0; r < 32; ++r) for (int c = 0; c < 16; ++c) if (hO[r * 16 + c] == val) { fr = r; fc = c; }
            printf("  t=%d j=%d -> row %d col %d\n", t, j, fr, fc);
          }
        }
      }
    }
  }
  return 0;
}
