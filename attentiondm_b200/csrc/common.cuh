// Shared helpers for the attndm_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <utility>

#include "../../include/attndm_b200.h"

namespace attndm {

void set_error(const char* fmt, ...);

#define ATTNDM_CHECK_ARG(cond, ...)            \
  do {                                         \
    if (!(cond)) {                             \
      ::attndm::set_error(__VA_ARGS__);        \
      return ATTNDM_ERR_ARG;                   \
    }                                          \
  } while (0)

#define ATTNDM_CUDA_LAUNCH_CHECK(name)                                              \
  do {                                                                              \
    cudaError_t e__ = cudaGetLastError();                                           \
    if (e__ != cudaSuccess) {                                                       \
      ::attndm::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));  \
      return ATTNDM_ERR_CUDA;                                                       \
    }                                                                               \
  } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

constexpr int kNumSMs = 148;       // B200
constexpr int kGnGroups = 32;      // every GroupNorm in the reference uses 32 groups

// ---- exact (non-contracted) fp32 arithmetic of the reference quantizer -----
// utils/quant_util.py:273-279: mul, sub, round-half-even, clamp, add, divide,
// each a separately rounded fp32 op.  The _rn intrinsics are never fused.
__device__ __forceinline__ float quant_code(float v, float s, float zp, float lo, float hi) {
  float q = rintf(__fsub_rn(__fmul_rn(s, v), zp));
  return fminf(fmaxf(q, lo), hi);
}
__device__ __forceinline__ float dequant(float code, float s, float zp) {
  return __fdiv_rn(__fadd_rn(code, zp), s);
}

// SiLU, x / (1 + exp(-x)) (models/diffusion.py:122, torch.nn.functional.silu).  ONE set of definitions shared by
// every kernel, so fused and unfused paths agree bit for bit.
//  * silu_acc: the exponential by range reduction + a degree-6 polynomial on the FMA pipe (<= 1 ulp, no
//    special-function unit) and a correctly rounded quotient (rcp.approx seed + one Newton step + a residual
//    correction).  It agrees with torch's CPU kernel (Sleef expf, 1 ulp, + IEEE divide) to the last bit for
//    99.6 % of inputs (<= 2 ulp always; tests/test_gpu_parity.py).  ~22 FMA-pipe instructions + one MUFU: used
//    wherever the SiLU output is consumed as fp32 (calibration branch, time_embed).
//  * silu_sfu: ex2.approx + rcp.approx, relative error <= ~5e-7, five instructions.
//  * silu_quant_t: the int8 hot path.  The consumer of SiLU there is always the activation quantizer,
//    code = clamp(rne(s * y - zp)).  Three builds, measured on the B200 (profiles/parity_r02.json, CIFAR-10 UNet,
//    batch 8, 4 steps, 53 M activations; "flips" = codes that differ from the CPU reference's on identical layer
//    inputs, next to the same count for torch's own CUDA kernels against its CPU kernels):
//      default            SFU form                          60 flips (torch CUDA: 62)   285 ms per DDIM-100 pass
//      -DATTNDM_SILU_GUARD   SFU first, silu_acc when t lands within kSiluGuard of a rounding boundary (codes
//                            identical to the accurate form, asserted)   34 (35)        302 ms
//      -DATTNDM_SILU_ACCURATE  silu_acc everywhere            34 (35)                   310 ms
//    The flips are dominated by GroupNorm statistics (torch's CPU kernel is the odd one out: ours and torch's CUDA
//    kernels flip the SAME elements), so the SFU form is no worse than torch-CUDA-vs-torch-CPU and 6 % faster:
//    it is the default; the other two stay as A/B builds (attentiondm_b200/build.py VARIANTS).
constexpr float kSiluGuard = 2.5e-4f;

__device__ __forceinline__ float silu_sfu(float v) {
  return __fdividef(v, __fadd_rn(1.0f, __expf(-v)));
}
__device__ __forceinline__ float silu_acc(float v) {
  // e = exp(a), a = -v clamped to [-87, 80]: beyond that the quotient is v (e -> 0) or |v| * 2e-35 (far below
  // any quantization step), and every intermediate stays a normal number
  const float a = fminf(fmaxf(-v, -87.0f), 80.0f);
  const float n = rintf(a * 1.44269504088896341f);
  float r = fmaf(n, -0.693359375f, a);                 // ln2 split hi/lo (Cody-Waite)
  r = fmaf(n, 2.12194440e-4f, r);
  float p = 1.9875691500e-4f;
  p = fmaf(p, r, 1.3981999507e-3f);
  p = fmaf(p, r, 8.3334519073e-3f);
  p = fmaf(p, r, 4.1665795894e-2f);
  p = fmaf(p, r, 1.6666665459e-1f);
  p = fmaf(p, r, 5.0000001201e-1f);
  p = fmaf(p, r * r, r);
  p = __fadd_rn(p, 1.0f);
  const float e = p * __int_as_float(((int)n + 127) << 23);      // n in [-126, 116]: a normal power of two
  const float d = __fadd_rn(1.0f, e);
  float r0;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(d));
  r0 = fmaf(fmaf(-d, r0, 1.0f), r0, r0);                         // Newton: |r0 - 1/d| < 1 ulp
  const float q = v * r0;
  return fmaf(fmaf(-q, d, v), r0, q);                            // residual correction -> correctly rounded v / d
}
__device__ __forceinline__ float silu_f(float v) { return silu_acc(v); }

// t = s * silu(u) - zp, the argument of the quantizer's round (utils/quant_util.py:273), see above
__device__ __forceinline__ float silu_quant_t(float u, float s, float zp) {
#if defined(ATTNDM_SILU_ACCURATE)
  return __fsub_rn(__fmul_rn(s, silu_acc(u)), zp);
#elif defined(ATTNDM_SILU_GUARD)
  const float sy = __fmul_rn(s, silu_sfu(u));
  float t = __fsub_rn(sy, zp);
  const float w = __fadd_rn(t, 12582912.0f);                      // 1.5 * 2^23: w - 1.5 * 2^23 == rne(t) for |t| < 2^22
  const float d = fabsf(fabsf(__fsub_rn(t, __fsub_rn(w, 12582912.0f))) - 0.5f);
  // the guard is > 2x the worst-case SFU error of t for |s * y| < 400; beyond that always the accurate form
  if (d < kSiluGuard || !(fabsf(sy) < 400.0f)) t = __fsub_rn(__fmul_rn(s, silu_acc(u)), zp);
  return t;
#else
  return __fsub_rn(__fmul_rn(s, silu_sfu(u)), zp);
#endif
}
// the quantizer's round + clamp on t (same codes as quant_code)
__device__ __forceinline__ float quant_code_t(float t, float lo, float hi) {
  return fminf(fmaxf(rintf(t), lo), hi);
}


// MixedPrecisionAttention.quantize_tensor (utils/attention_quant_utils.py:30-38), shared by attention_kernel and the
// fused one-position attention of rowprog.cu
__device__ __forceinline__ float attn_fake_quant(float x, float s, float zp, float qmax) {
  // clamp(round(x / scale) + zero_point, 0, qmax); (x_q - zero_point) * scale
  float q = __fadd_rn(rintf(__fdiv_rn(x, s)), zp);
  q = fminf(fmaxf(q, 0.f), qmax);
  return __fmul_rn(__fsub_rn(q, zp), s);
}

// GroupNorm(32) finalisation and application, shared by every kernel that normalises (so that the fused
// and the stand-alone paths agree bit for bit): mean / rstd from double sums, each step separately rounded.
__device__ __forceinline__ void gn_mean_rstd(double s, double ss, double inv_n, float eps, float& mean, float& rstd) {
  const double m = __dmul_rn(s, inv_n);
  double var = __dsub_rn(__dmul_rn(ss, inv_n), __dmul_rn(m, m));
  if (var < 0.0) var = 0.0;
  mean = (float)m;
  // correctly rounded fp32 reciprocal square root of the (double-accumulated) variance: <= 1 ulp from the
  // double-precision 1/sqrt it replaces, without the ~100-instruction software double sqrt + divide that every
  // warp of the small-map kernels paid per sample
  rstd = __frsqrt_rn((float)__dadd_rn(var, (double)eps));
}
// silu(groupnorm(v)) with a = rstd*gamma, b = beta - mean*a  (models/diffusion.py:121-122)
__device__ __forceinline__ float gn_apply(float v, float mean, float rstd, float gamma, float beta) {
  const float a = __fmul_rn(rstd, gamma);
  return fmaf(v, a, fmaf(-mean, a, beta));
}
__device__ __forceinline__ float gn_silu_apply(float v, float mean, float rstd, float gamma, float beta) {
  return silu_f(gn_apply(v, mean, rstd, gamma, beta));
}

// (c0, c1) = (a*b0 + c0, a*b1 + c1), each an IEEE fma, in one packed instruction
__device__ __forceinline__ void fma2(float& c0, float& c1, float a, float b0, float b1) {
  asm("{\n\t.reg .b64 ra, rb, rc;\n\t"
      "mov.b64 ra, {%2, %2};\n\t"
      "mov.b64 rb, {%3, %4};\n\t"
      "mov.b64 rc, {%0, %1};\n\t"
      "fma.rn.f32x2 rc, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rc;\n\t}"
      : "+f"(c0), "+f"(c1)
      : "f"(a), "f"(b0), "f"(b1));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- programmatic dependent launch (PDL) ------------------------------------------------------
// Every kernel of the sampling graph starts with pdl_enter(): it lets the NEXT kernel of the stream be
// launched early (its CTAs are scheduled and run their own prologue while this grid drains) and then
// waits until the PREVIOUS grid has completed and its writes are visible.  Nothing that a predecessor
// writes may be touched before pdl_wait().
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() { pdl_launch_dependents(); pdl_wait(); }

bool pdl_enabled();   // ATTNDM_PDL=0 in the environment switches the launch attribute off

template <typename... ExpTypes, typename... ActTypes>
inline cudaError_t launch_pdl(void (*kernel)(ExpTypes...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              ActTypes&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr = {};
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<ExpTypes>(args)...);
}

// the same with a thread-block cluster of `cluster` CTAs along x (1: no cluster attribute)
template <typename... ExpTypes, typename... ActTypes>
inline cudaError_t launch_pdl_cluster(void (*kernel)(ExpTypes...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                      int cluster, ActTypes&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2] = {};
  int n = 0;
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = (unsigned)cluster;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<ExpTypes>(args)...);
}

// streaming 128-bit global accesses (activations are touched once per kernel)
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

}  // namespace attndm
