// HBM-bound quantizer / collector kernels.
//   act_quant      : utils/quant_util.py:260-282 (+ fused GroupNorm/SiLU producer)
//   gn_stats       : GroupNorm(32) statistics, double accumulation
//   minmax_c       : utils/quant_util.py:187-191
//   group_ranges   : utils/quant_util.py:193-205, 403-437
//   calib_mix      : utils/quant_util.py:207-224, 54-66
//   kth_value      : utils/quant_util.py:440-450
// Mapping used by the row kernels: one warp owns one NHWC pixel row at a time and
// its lanes walk the row in float4 steps (fully coalesced 512 B per request), so
// the per-pixel code sum is a warp shuffle reduction and never an atomic.  Warps
// own contiguous row ranges so per-sample state (GroupNorm mean/rstd, held one
// group per lane) is refreshed only when the sample index changes.
#include <stdlib.h>
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;
namespace attndm {

// ---------------------------------------------------------------------------
// act_quant
// ---------------------------------------------------------------------------
struct ActQuantParams {
  const float* x;
  int B, H, W, C, Cp;
  const float* scale;
  const float* zp;
  float qlo, qhi;
  const double* gn_stats;
  const float* gamma;
  const float* beta;
  float eps;
  int8_t* codes;
  int32_t* rowsum;
  int halo;
  float* y;
  long long rows;          // rows of the code layout (halo or plain)
  long long rows_per_warp;
  const float* x2;         // CAT kernels: x = [B][H/2][W/2][C1] (read through a nearest x2 upsample), x2 = [B][H][W][C - C1]
  int C1;
  // DUAL kernels: a second, producer-less quantizer of the same input (the shortcut conv of a ResidualBlock reads what
  // conv1 reads behind GroupNorm+SiLU): same bit width, its own tables, outputs and layout
  const float* scale2;
  const float* zp2;
  int8_t* codes2;
  int32_t* rowsum2;
  int halo2;
};

template <int PRE>
__device__ __forceinline__ float pre_op(float v, float a, float b) {
  if (PRE == ATTNDM_PRE_GN_SILU) return silu_f(fmaf(v, a, b));
  if (PRE == ATTNDM_PRE_SILU) return silu_f(v);
  return v;
}
// t = s * pre_op(v) - zp, the argument of the quantizer's round; a SiLU producer goes through silu_quant_t
// (SFU evaluation, refined next to a rounding boundary: same codes as the accurate form, common.cuh)
template <int PRE>
__device__ __forceinline__ float quant_t(float v, float a, float b, float s, float zp) {
  if (PRE == ATTNDM_PRE_GN_SILU) return silu_quant_t(fmaf(v, a, b), s, zp);
  if (PRE == ATTNDM_PRE_SILU) return silu_quant_t(v, s, zp);
  return __fsub_rn(__fmul_rn(s, v), zp);
}

__device__ __forceinline__ void gn_refresh(const double* stats, int b, int lane, double inv_n,
                                           float eps, float& mean, float& rstd) {
  // lane g owns group g
  double s = stats[((long long)b * kGnGroups + lane) * 2 + 0];
  double ss = stats[((long long)b * kGnGroups + lane) * 2 + 1];
  gn_mean_rstd(s, ss, inv_n, eps, mean, rstd);
}

template <int PRE, bool QUANT, bool VEC>
__global__ void __launch_bounds__(256) act_quant_kernel(ActQuantParams p) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  long long r0 = warp * p.rows_per_warp;
  long long r1 = r0 + p.rows_per_warp;
  if (r1 > p.rows) r1 = p.rows;
  const int Hp = p.halo ? p.H + 2 : p.H, Wp = p.halo ? p.W + 2 : p.W;
  const int cpg = p.C / kGnGroups;
  const double inv_n = (PRE == ATTNDM_PRE_GN_SILU) ? 1.0 / ((double)p.H * p.W * cpg) : 0.0;
  int cur_b = -1;
  float mean = 0.f, rstd = 0.f;
  for (long long r = r0; r < r1; ++r) {
    int b = (int)(r / ((long long)Hp * Wp));
    int rem = (int)(r - (long long)b * Hp * Wp);
    int hp = rem / Wp, wp = rem - hp * Wp;
    bool interior = true;
    int h = hp, w = wp;
    if (p.halo) {
      interior = (hp >= 1 && hp <= p.H && wp >= 1 && wp <= p.W);
      h = hp - 1;
      w = wp - 1;
    }
    const long long pix = ((long long)b * p.H + h) * p.W + w;
    if (PRE == ATTNDM_PRE_GN_SILU && interior && b != cur_b) {
      gn_refresh(p.gn_stats, b, lane, inv_n, p.eps, mean, rstd);
      cur_b = b;
    }
    int acc = 0;
    if (VEC) {
      const int Q = p.C >> 2;
      // uniform trip count: the GroupNorm shuffles below need every lane present
      for (int q0 = 0; q0 < Q; q0 += 32) {
        const bool act = (q0 + lane) < Q;
        const int c = act ? ((q0 + lane) << 2) : 0;
        float4 s4 = make_float4(1.f, 1.f, 1.f, 1.f), z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (QUANT) {
          s4 = *reinterpret_cast<const float4*>(p.scale + c);
          z4 = *reinterpret_cast<const float4*>(p.zp + c);
        }
        float4 cd;
        if (interior) {
          float4 v = ldg_stream(reinterpret_cast<const float4*>(p.x + pix * p.C + c));
          float4 ga = make_float4(0.f, 0.f, 0.f, 0.f), gb = ga;
          if (PRE == ATTNDM_PRE_GN_SILU) {
            // group of each of the 4 channels (cpg may be < 4 for tiny models)
            float4 g4 = *reinterpret_cast<const float4*>(p.gamma + c);
            float4 b4 = *reinterpret_cast<const float4*>(p.beta + c);
            int g0 = c / cpg, g1 = (c + 1) / cpg, g2 = (c + 2) / cpg, g3 = (c + 3) / cpg;
            float m0 = __shfl_sync(0xffffffffu, mean, g0), r0_ = __shfl_sync(0xffffffffu, rstd, g0);
            float m1 = m0, r1_ = r0_, m2 = m0, r2_ = r0_, m3 = m0, r3_ = r0_;
            if (cpg & 3) {   // a float4 may straddle groups; warp-uniform branch (cpg is kernel-wide)
              m1 = __shfl_sync(0xffffffffu, mean, g1); r1_ = __shfl_sync(0xffffffffu, rstd, g1);
              m2 = __shfl_sync(0xffffffffu, mean, g2); r2_ = __shfl_sync(0xffffffffu, rstd, g2);
              m3 = __shfl_sync(0xffffffffu, mean, g3); r3_ = __shfl_sync(0xffffffffu, rstd, g3);
            }
            ga.x = r0_ * g4.x; gb.x = fmaf(-m0, ga.x, b4.x);
            ga.y = r1_ * g4.y; gb.y = fmaf(-m1, ga.y, b4.y);
            ga.z = r2_ * g4.z; gb.z = fmaf(-m2, ga.z, b4.z);
            ga.w = r3_ * g4.w; gb.w = fmaf(-m3, ga.w, b4.w);
          }
          if (QUANT) {
            cd.x = quant_code_t(quant_t<PRE>(v.x, ga.x, gb.x, s4.x, z4.x), p.qlo, p.qhi);
            cd.y = quant_code_t(quant_t<PRE>(v.y, ga.y, gb.y, s4.y, z4.y), p.qlo, p.qhi);
            cd.z = quant_code_t(quant_t<PRE>(v.z, ga.z, gb.z, s4.z, z4.z), p.qlo, p.qhi);
            cd.w = quant_code_t(quant_t<PRE>(v.w, ga.w, gb.w, s4.w, z4.w), p.qlo, p.qhi);
            if (p.y && act) {
              float4 o;
              o.x = dequant(cd.x, s4.x, z4.x);
              o.y = dequant(cd.y, s4.y, z4.y);
              o.z = dequant(cd.z, s4.z, z4.z);
              o.w = dequant(cd.w, s4.w, z4.w);
              *reinterpret_cast<float4*>(p.y + pix * p.C + c) = o;
            }
          } else {
            v.x = pre_op<PRE>(v.x, ga.x, gb.x);
            v.y = pre_op<PRE>(v.y, ga.y, gb.y);
            v.z = pre_op<PRE>(v.z, ga.z, gb.z);
            v.w = pre_op<PRE>(v.w, ga.w, gb.w);
            if (p.y && act) *reinterpret_cast<float4*>(p.y + pix * p.C + c) = v;
            cd = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        } else {
          // halo ring: the code of 0.0, i.e. -zero_point (asserted in range by the host)
          cd.x = fminf(fmaxf(-z4.x, p.qlo), p.qhi);
          cd.y = fminf(fmaxf(-z4.y, p.qlo), p.qhi);
          cd.z = fminf(fmaxf(-z4.z, p.qlo), p.qhi);
          cd.w = fminf(fmaxf(-z4.w, p.qlo), p.qhi);
        }
        if (QUANT && p.codes && act) {
          int ix = (int)cd.x, iy = (int)cd.y, iz = (int)cd.z, iw = (int)cd.w;
          acc += ix + iy + iz + iw;
          char4 c4 = make_char4((signed char)ix, (signed char)iy, (signed char)iz, (signed char)iw);
          *reinterpret_cast<char4*>(p.codes + r * p.Cp + c) = c4;
        }
      }
    } else {
      for (int c = lane; c < p.Cp; c += 32) {
        float cd = 0.f;
        if (c < p.C) {
          float s = QUANT ? p.scale[c] : 1.f, z = QUANT ? p.zp[c] : 0.f;
          if (interior) {
            float v = p.x[pix * p.C + c];
            // C % 4 != 0 only occurs for the image latent, which has no GroupNorm (host-checked)
            constexpr int PS = PRE == ATTNDM_PRE_GN_SILU ? ATTNDM_PRE_NONE : PRE;
            if (QUANT) {
              cd = quant_code_t(quant_t<PS>(v, 0.f, 0.f, s, z), p.qlo, p.qhi);
              if (p.y) p.y[pix * p.C + c] = dequant(cd, s, z);
            } else if (p.y) {
              p.y[pix * p.C + c] = pre_op<PS>(v, 0.f, 0.f);
            }
          } else {
            cd = fminf(fmaxf(-z, p.qlo), p.qhi);
          }
        }
        if (QUANT && p.codes) {
          acc += (int)cd;
          p.codes[r * p.Cp + c] = (int8_t)(int)cd;
        }
      }
    }
    if (QUANT && p.rowsum) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) p.rowsum[r] = acc;
    }
  }
}

// Fast path for the shapes that carry the traffic (C = 128 or 256, i.e. NQ = 1 or 2 float4 per lane):
// per-lane quantizer / GroupNorm constants live in registers across rows, the row index is advanced
// incrementally (no 64-bit division per row) and the next row's load is issued before the current row
// is processed (two 512-B requests in flight per warp).
template <int PRE, bool QUANT, int NQ>
__global__ void __launch_bounds__(256, NQ == 1 ? 3 : 2) act_quant_fast_kernel(ActQuantParams p) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long r0 = warp * p.rows_per_warp;
  long long r1 = r0 + p.rows_per_warp;
  if (r1 > p.rows) r1 = p.rows;
  if (r0 >= r1) return;
  const int Hp = p.halo ? p.H + 2 : p.H, Wp = p.halo ? p.W + 2 : p.W;
  const int cpg = p.C / kGnGroups;
  const double inv_n = (PRE == ATTNDM_PRE_GN_SILU) ? 1.0 / ((double)p.H * p.W * cpg) : 0.0;
  const long long per = (long long)Hp * Wp;
  int b = (int)(r0 / per);
  int rem = (int)(r0 - (long long)b * per);
  int hp = rem / Wp, wp = rem - hp * Wp;
  float4 s4[NQ], z4[NQ], ga[NQ], gb[NQ];       // per-lane quantizer and GroupNorm constants, in registers across rows
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    const int c = (i * 32 + lane) << 2;
    s4[i] = make_float4(1.f, 1.f, 1.f, 1.f);
    z4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (QUANT) {
      s4[i] = *reinterpret_cast<const float4*>(p.scale + c);
      z4[i] = *reinterpret_cast<const float4*>(p.zp + c);
    }
    ga[i] = gb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  int cur_b = -1;
  auto interior_of = [&](int hh, int ww) { return !p.halo || (hh >= 1 && hh <= p.H && ww >= 1 && ww <= p.W); };
  auto pix_of = [&](int bb, int hh, int ww) {
    return p.halo ? (bb * p.H + (hh - 1)) * p.W + (ww - 1) : (bb * p.H + hh) * p.W + ww;
  };
  // R rows per iteration: all their loads are issued before any arithmetic (R x NQ independent 512-byte
  // requests in flight per warp), and the per-row bookkeeping (coordinates, code-sum reduction, row-sum store)
  // is paid once per group.
  constexpr int R = 4;
  for (long long r = r0; r < r1; r += R) {
    bool in[R];
    unsigned pix[R];                   // element offset of the pixel row (the launcher checks it fits 32 bits)
    int bb[R];
    float4 v[R][NQ];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      in[k] = (r + k < r1) && interior_of(hp, wp);
      pix[k] = in[k] ? (unsigned)pix_of(b, hp, wp) * (unsigned)p.C : 0u;
      bb[k] = b;
      if (in[k]) {
#pragma unroll
        for (int i = 0; i < NQ; ++i)
          v[k][i] = ldg_stream(reinterpret_cast<const float4*>(p.x + pix[k] + ((i * 32 + lane) << 2)));
      }
      if (++wp == Wp) { wp = 0; if (++hp == Hp) { hp = 0; ++b; } }
    }
    int sums[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      if (PRE == ATTNDM_PRE_GN_SILU && in[k] && bb[k] != cur_b) {        // warp-uniform
        float mean, rstd;
        gn_refresh(p.gn_stats, bb[k], lane, inv_n, p.eps, mean, rstd);
        cur_b = bb[k];
#pragma unroll
        for (int i = 0; i < NQ; ++i) {                        // once per sample: gamma / beta come from L1/L2
          const int c = (i * 32 + lane) << 2;
          const int gi = c / cpg;                              // cpg % 4 == 0 on this path: one group per float4
          const float4 g4 = *reinterpret_cast<const float4*>(p.gamma + c);
          const float4 be4 = *reinterpret_cast<const float4*>(p.beta + c);
          const float m = __shfl_sync(0xffffffffu, mean, gi), rs = __shfl_sync(0xffffffffu, rstd, gi);
          ga[i].x = rs * g4.x; gb[i].x = fmaf(-m, ga[i].x, be4.x);
          ga[i].y = rs * g4.y; gb[i].y = fmaf(-m, ga[i].y, be4.y);
          ga[i].z = rs * g4.z; gb[i].z = fmaf(-m, ga[i].z, be4.z);
          ga[i].w = rs * g4.w; gb[i].w = fmaf(-m, ga[i].w, be4.w);
        }
      }
      int acc = 0;
      if (r + k < r1) {
#pragma unroll
        for (int i = 0; i < NQ; ++i) {
          const int c = (i * 32 + lane) << 2;
          float4 cd;
          if (in[k]) {
            float4 t = v[k][i];
            if (QUANT) {
              cd.x = quant_code_t(quant_t<PRE>(t.x, ga[i].x, gb[i].x, s4[i].x, z4[i].x), p.qlo, p.qhi);
              cd.y = quant_code_t(quant_t<PRE>(t.y, ga[i].y, gb[i].y, s4[i].y, z4[i].y), p.qlo, p.qhi);
              cd.z = quant_code_t(quant_t<PRE>(t.z, ga[i].z, gb[i].z, s4[i].z, z4[i].z), p.qlo, p.qhi);
              cd.w = quant_code_t(quant_t<PRE>(t.w, ga[i].w, gb[i].w, s4[i].w, z4[i].w), p.qlo, p.qhi);
              if (p.y) {
                float4 o;
                o.x = dequant(cd.x, s4[i].x, z4[i].x);
                o.y = dequant(cd.y, s4[i].y, z4[i].y);
                o.z = dequant(cd.z, s4[i].z, z4[i].z);
                o.w = dequant(cd.w, s4[i].w, z4[i].w);
                *reinterpret_cast<float4*>(p.y + pix[k] + c) = o;
              }
            } else {
              t.x = pre_op<PRE>(t.x, ga[i].x, gb[i].x);
              t.y = pre_op<PRE>(t.y, ga[i].y, gb[i].y);
              t.z = pre_op<PRE>(t.z, ga[i].z, gb[i].z);
              t.w = pre_op<PRE>(t.w, ga[i].w, gb[i].w);
              if (p.y) *reinterpret_cast<float4*>(p.y + pix[k] + c) = t;
              cd = make_float4(0.f, 0.f, 0.f, 0.f);
            }
          } else {                                             // halo ring: the code of 0.0
            cd.x = fminf(fmaxf(-z4[i].x, p.qlo), p.qhi);
            cd.y = fminf(fmaxf(-z4[i].y, p.qlo), p.qhi);
            cd.z = fminf(fmaxf(-z4[i].z, p.qlo), p.qhi);
            cd.w = fminf(fmaxf(-z4[i].w, p.qlo), p.qhi);
          }
          if (QUANT && p.codes) {
            const int ix = (int)cd.x, iy = (int)cd.y, iz = (int)cd.z, iw = (int)cd.w;
            acc += ix + iy + iz + iw;
            *reinterpret_cast<char4*>(p.codes + (r + k) * p.Cp + c) =
                make_char4((signed char)ix, (signed char)iy, (signed char)iz, (signed char)iw);
          }
        }
      }
      sums[k] = acc;
    }
    if (QUANT && p.rowsum) {
      int mine = 0;
#pragma unroll
      for (int k = 0; k < R; ++k) {
        const int t = __reduce_add_sync(0xffffffffu, sums[k]);
        if (lane == k) mine = t;
      }
      if (lane < R && r + lane < r1) p.rowsum[r + lane] = mine;
    }
  }
}

// ---------------------------------------------------------------------------
// The int8 hot path at the large feature maps (codes + row sums only, C = 128 or 256, W % 4 == 0):
// warps walk INTERIOR image rows, four pixels per step with all loads issued first; the halo ring is
// written by the warp that owns the adjacent image row, so the inner loop has no border logic.
// Quantizer: clamp before round (bounds are integers, so it equals round-then-clamp, NaN -> lo as before)
// and one F2I.RNI instead of rint + cvt.  ~15 instructions per element instead of ~25.
// ---------------------------------------------------------------------------
__device__ __forceinline__ int quant_code_i(float t, float lo, float hi) {
  return __float2int_rn(fminf(fmaxf(t, lo), hi));
}

// CAT: the input is the never-materialised concat of an UpBlock (models/diffusion.py:225-229 + torch.cat): channels
// [0, C1) are p.x [B][H/2][W/2][C1] seen through a nearest-neighbour x2 upsample, channels [C1, C) are p.x2
// [B][H][W][C - C1].  C1 % 128 == 0, so each of a lane's NQ float4 slots lies entirely in one part.
// DUAL: every element is read ONCE and quantized twice -- by the main quantizer (PRE) and by a second one without a
// producer (ActQuantParams::scale2 ...).
template <int PRE, int NQ, bool A8, bool CAT, bool DUAL = false>
__global__ void __launch_bounds__(256, DUAL ? 2 : (NQ == 1 ? 4 : NQ == 2 ? 3 : 2)) act_quant_rows_kernel(ActQuantParams p) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const int warp = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int nimg = p.B * p.H;                           // interior image rows
  const int ir0 = warp * (int)p.rows_per_warp;
  int ir1 = ir0 + (int)p.rows_per_warp;
  if (ir1 > nimg) ir1 = nimg;
  if (ir0 >= ir1) return;
  const int W = p.W, C = p.C, Cp = p.Cp;
  const int Hp = p.halo ? p.H + 2 : p.H, Wp = p.halo ? W + 2 : W;
  const int cpg = C / kGnGroups;
  const double inv_n = (PRE == ATTNDM_PRE_GN_SILU) ? 1.0 / ((double)p.H * W * cpg) : 0.0;
  float4 s4[NQ], z4[NQ], ga[NQ], gb[NQ];
  int padw[NQ], padsum = 0;                              // the ring code (packed) and its row sum
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    const int c = (i * 32 + lane) << 2;
    s4[i] = *reinterpret_cast<const float4*>(p.scale + c);
    z4[i] = *reinterpret_cast<const float4*>(p.zp + c);
    const int a = (int)fminf(fmaxf(-z4[i].x, p.qlo), p.qhi), b2 = (int)fminf(fmaxf(-z4[i].y, p.qlo), p.qhi);
    const int c2 = (int)fminf(fmaxf(-z4[i].z, p.qlo), p.qhi), d = (int)fminf(fmaxf(-z4[i].w, p.qlo), p.qhi);
    padw[i] = (a & 0xff) | ((b2 & 0xff) << 8) | ((c2 & 0xff) << 16) | ((d & 0xff) << 24);
    padsum += a + b2 + c2 + d;
    ga[i] = gb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  padsum = __reduce_add_sync(0xffffffffu, padsum);
  auto ring_row = [&](long long row) {                   // one ring pixel: its codes and row sum
#pragma unroll
    for (int i = 0; i < NQ; ++i) *reinterpret_cast<int*>(p.codes + row * Cp + ((i * 32 + lane) << 2)) = padw[i];
    if (lane == 0) p.rowsum[row] = padsum;
  };
  float4 s4b[DUAL ? NQ : 1], z4b[DUAL ? NQ : 1];
  int padwb[DUAL ? NQ : 1], padsumb = 0;
  if (DUAL) {
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      const int c = (i * 32 + lane) << 2;
      s4b[i] = *reinterpret_cast<const float4*>(p.scale2 + c);
      z4b[i] = *reinterpret_cast<const float4*>(p.zp2 + c);
      const int a = (int)fminf(fmaxf(-z4b[i].x, p.qlo), p.qhi), b2 = (int)fminf(fmaxf(-z4b[i].y, p.qlo), p.qhi);
      const int c2 = (int)fminf(fmaxf(-z4b[i].z, p.qlo), p.qhi), d = (int)fminf(fmaxf(-z4b[i].w, p.qlo), p.qhi);
      padwb[i] = (a & 0xff) | ((b2 & 0xff) << 8) | ((c2 & 0xff) << 16) | ((d & 0xff) << 24);
      padsumb += a + b2 + c2 + d;
    }
    padsumb = __reduce_add_sync(0xffffffffu, padsumb);
  }
  auto ring_row2 = [&](long long row) {
#pragma unroll
    for (int i = 0; i < (DUAL ? NQ : 1); ++i) *reinterpret_cast<int*>(p.codes2 + row * Cp + ((i * 32 + lane) << 2)) = padwb[i];
    if (lane == 0) p.rowsum2[row] = padsumb;
  };
  int cur_b = -1;
  constexpr int R = NQ <= 2 ? 4 : 2;                    // pixels per step (register budget)
  for (int ir = ir0; ir < ir1; ++ir) {
    const int b = ir / p.H, h = ir - b * p.H;
    if (PRE == ATTNDM_PRE_GN_SILU && b != cur_b) {       // warp-uniform, once per sample
      float mean, rstd;
      gn_refresh(p.gn_stats, b, lane, inv_n, p.eps, mean, rstd);
      cur_b = b;
#pragma unroll
      for (int i = 0; i < NQ; ++i) {
        const int c = (i * 32 + lane) << 2;
        const int gi = c / cpg;                            // cpg % 4 == 0 on this path: one group per float4
        const float4 g4 = *reinterpret_cast<const float4*>(p.gamma + c);
        const float4 be4 = *reinterpret_cast<const float4*>(p.beta + c);
        const float m = __shfl_sync(0xffffffffu, mean, gi), rs = __shfl_sync(0xffffffffu, rstd, gi);
        ga[i].x = rs * g4.x; gb[i].x = fmaf(-m, ga[i].x, be4.x);
        ga[i].y = rs * g4.y; gb[i].y = fmaf(-m, ga[i].y, be4.y);
        ga[i].z = rs * g4.z; gb[i].z = fmaf(-m, ga[i].z, be4.z);
        ga[i].w = rs * g4.w; gb[i].w = fmaf(-m, ga[i].w, be4.w);
      }
    }
    const float* xrow = p.x + (long long)ir * W * C + (lane << 2);
    const int C2 = C - p.C1, Wa = W >> 1;
    const float* arow = CAT ? p.x + ((long long)(b * (p.H >> 1) + (h >> 1)) * Wa) * p.C1 + (lane << 2) : nullptr;
    const float* brow = CAT ? p.x2 + (long long)ir * W * C2 + (lane << 2) - p.C1 : nullptr;
    const long long rbase = p.halo ? ((long long)b * Hp + h + 1) * Wp + 1 : (long long)ir * W;   // code row of pixel w = 0
    int8_t* crow = p.codes + rbase * Cp + (lane << 2);
    const long long rbase2 = !DUAL ? 0 : (p.halo2 ? ((long long)b * (p.H + 2) + h + 1) * (W + 2) + 1 : (long long)ir * W);
    int8_t* crow2 = DUAL ? p.codes2 + rbase2 * Cp + (lane << 2) : nullptr;
    for (int w0 = 0; w0 < W; w0 += R) {
      float4 v[R][NQ];
#pragma unroll
      for (int k = 0; k < R; ++k)
#pragma unroll
        for (int i = 0; i < NQ; ++i)
          if (!CAT) {
            v[k][i] = ldg_stream(reinterpret_cast<const float4*>(xrow + (long long)(w0 + k) * C + i * 128));
          } else if (i * 128 < p.C1) {                          // upsampled part: pixels 2j and 2j + 1 read the same element
            if ((k & 1) == 0) v[k][i] = __ldg(reinterpret_cast<const float4*>(arow + (long long)((w0 + k) >> 1) * p.C1 + i * 128));
            else v[k][i] = v[k - 1][i];
          } else {
            v[k][i] = ldg_stream(reinterpret_cast<const float4*>(brow + (long long)(w0 + k) * C2 + i * 128));
          }
      int sums[R], sums2[R];
#pragma unroll
      for (int k = 0; k < R; ++k) {
        int acc = 0, acc2 = 0;
#pragma unroll
        for (int i = 0; i < NQ; ++i) {
          if (DUAL) {                                          // the second quantizer: s2 * x - zp2, no producer
            const float4 x4 = v[k][i];
            const float ux = quant_t<ATTNDM_PRE_NONE>(x4.x, 0.f, 0.f, s4b[i].x, z4b[i].x);
            const float uy = quant_t<ATTNDM_PRE_NONE>(x4.y, 0.f, 0.f, s4b[i].y, z4b[i].y);
            const float uz = quant_t<ATTNDM_PRE_NONE>(x4.z, 0.f, 0.f, s4b[i].z, z4b[i].z);
            const float uw = quant_t<ATTNDM_PRE_NONE>(x4.w, 0.f, 0.f, s4b[i].w, z4b[i].w);
            int word2;
            if (A8) {
              const int ix = __float2int_rn(ux), iy = __float2int_rn(uy), iz = __float2int_rn(uz), iw = __float2int_rn(uw);
              int hi2;
              asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, 0;" : "=r"(hi2) : "r"(iw), "r"(iz));
              asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(word2) : "r"(iy), "r"(ix), "r"(hi2));
              acc2 = __dp4a(word2, 0x01010101, acc2);
            } else {
              const int ix = quant_code_i(ux, p.qlo, p.qhi), iy = quant_code_i(uy, p.qlo, p.qhi);
              const int iz = quant_code_i(uz, p.qlo, p.qhi), iw = quant_code_i(uw, p.qlo, p.qhi);
              acc2 += ix + iy + iz + iw;
              word2 = (ix & 0xff) | ((iy & 0xff) << 8) | ((iz & 0xff) << 16) | (iw << 24);
            }
            *reinterpret_cast<int*>(crow2 + (long long)(w0 + k) * Cp + i * 128) = word2;
          }
          float4 t = v[k][i];                                  // -> the quantizer's pre-round values s * pre(x) - zp
          t.x = quant_t<PRE>(t.x, ga[i].x, gb[i].x, s4[i].x, z4[i].x);
          t.y = quant_t<PRE>(t.y, ga[i].y, gb[i].y, s4[i].y, z4[i].y);
          t.z = quant_t<PRE>(t.z, ga[i].z, gb[i].z, s4[i].z, z4[i].z);
          t.w = quant_t<PRE>(t.w, ga[i].w, gb[i].w, s4[i].w, z4[i].w);
          int word;
          if (A8) {
            // 8-bit codes: round to nearest even, then ONE saturating pack per two codes (cvt.pack.sat.s8.s32 clamps
            // to [-128, 127], which is the quantizer's clamp), and the code sum as one dp4a with a vector of ones
            const int ix = __float2int_rn(t.x), iy = __float2int_rn(t.y);
            const int iz = __float2int_rn(t.z), iw = __float2int_rn(t.w);
            int hi2;
            asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, 0;" : "=r"(hi2) : "r"(iw), "r"(iz));
            asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(word) : "r"(iy), "r"(ix), "r"(hi2));
            acc = __dp4a(word, 0x01010101, acc);
          } else {
            const int ix = quant_code_i(t.x, p.qlo, p.qhi), iy = quant_code_i(t.y, p.qlo, p.qhi);
            const int iz = quant_code_i(t.z, p.qlo, p.qhi), iw = quant_code_i(t.w, p.qlo, p.qhi);
            acc += ix + iy + iz + iw;
            word = (ix & 0xff) | ((iy & 0xff) << 8) | ((iz & 0xff) << 16) | (iw << 24);
          }
          *reinterpret_cast<int*>(crow + (long long)(w0 + k) * Cp + i * 128) = word;
        }
        sums[k] = acc;
        sums2[k] = acc2;
      }
      int mine = 0, mine2 = 0;
#pragma unroll
      for (int k = 0; k < R; ++k) {
        const int t = __reduce_add_sync(0xffffffffu, sums[k]);
        if (lane == k) mine = t;
        if (DUAL) {
          const int t2 = __reduce_add_sync(0xffffffffu, sums2[k]);
          if (lane == k) mine2 = t2;
        }
      }
      if (lane < R) p.rowsum[rbase + w0 + lane] = mine;
      if (DUAL && lane < R) p.rowsum2[rbase2 + w0 + lane] = mine2;
    }
    if (p.halo) {                                        // the ring next to this image row
      ring_row(rbase - 1);
      ring_row(rbase + W);
      if (h == 0)
        for (int w = 0; w < Wp; ++w) ring_row(rbase - 1 - Wp + w);
      if (h == p.H - 1)
        for (int w = 0; w < Wp; ++w) ring_row(rbase - 1 + Wp + w);
    }
    if (DUAL && p.halo2) {
      const int Wp2 = W + 2;
      ring_row2(rbase2 - 1);
      ring_row2(rbase2 + W);
      if (h == 0)
        for (int w = 0; w < Wp2; ++w) ring_row2(rbase2 - 1 - Wp2 + w);
      if (h == p.H - 1)
        for (int w = 0; w < Wp2; ++w) ring_row2(rbase2 - 1 + Wp2 + w);
    }
  }
}

// ---------------------------------------------------------------------------
// Narrow inputs (C <= 16: the 3-channel latent of init_conv): one THREAD per code row, one 16-byte store.
// The warp-per-row kernels keep 29 of 32 lanes idle on such rows (61 us for 0.8 M elements).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) act_quant_narrow_kernel(ActQuantParams p) {
  pdl_enter();
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= p.rows) return;
  const int Hp = p.halo ? p.H + 2 : p.H, Wp = p.halo ? p.W + 2 : p.W;
  const long long per = (long long)Hp * Wp;
  const int b = (int)(r / per);
  const int rem = (int)(r - (long long)b * per);
  const int hp = rem / Wp, wp = rem - hp * Wp;
  bool interior = true;
  int h = hp, w = wp;
  if (p.halo) { interior = hp >= 1 && hp <= p.H && wp >= 1 && wp <= p.W; h = hp - 1; w = wp - 1; }
  const float* xr = p.x + (((long long)b * p.H + h) * p.W + w) * p.C;
  int q[16];
  int acc = 0;
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    q[c] = 0;
    if (c < p.C) {
      const float s = p.scale[c], z = p.zp[c];
      q[c] = interior ? quant_code_i(__fsub_rn(__fmul_rn(s, xr[c]), z), p.qlo, p.qhi) : (int)fminf(fmaxf(-z, p.qlo), p.qhi);
      acc += q[c];
    }
  }
  uint4 o;
  o.x = (q[0] & 0xff) | ((q[1] & 0xff) << 8) | ((q[2] & 0xff) << 16) | (q[3] << 24);
  o.y = (q[4] & 0xff) | ((q[5] & 0xff) << 8) | ((q[6] & 0xff) << 16) | (q[7] << 24);
  o.z = (q[8] & 0xff) | ((q[9] & 0xff) << 8) | ((q[10] & 0xff) << 16) | (q[11] << 24);
  o.w = (q[12] & 0xff) | ((q[13] & 0xff) << 8) | ((q[14] & 0xff) << 16) | (q[15] << 24);
  *reinterpret_cast<uint4*>(p.codes + r * 16) = o;
  p.rowsum[r] = acc;
}

template <int PRE, bool QUANT>
static void launch_act_quant(const ActQuantParams& p, int blocks, cudaStream_t st) {
  // the GN shuffle in the scalar path needs all lanes converged per channel step;
  // it is only used for C % 4 != 0 (the 3-channel latent), which never has a GN.
  const bool gn_ok = ((PRE != ATTNDM_PRE_GN_SILU) || ((p.C / kGnGroups) % 4 == 0)) &&
                     (long long)p.B * p.H * p.W * p.C < (1LL << 31);          // 32-bit element offsets in the fast kernel
  if (p.C == 128 && gn_ok)
    launch_pdl(act_quant_fast_kernel<PRE, QUANT, 1>, dim3(blocks), dim3(256), 0, st, p);
  else if (p.C == 256 && gn_ok)
    launch_pdl(act_quant_fast_kernel<PRE, QUANT, 2>, dim3(blocks), dim3(256), 0, st, p);
  else if ((p.C & 3) == 0)
    launch_pdl(act_quant_kernel<PRE, QUANT, true>, dim3(blocks), dim3(256), 0, st, p);
  else
    launch_pdl(act_quant_kernel<PRE, QUANT, false>, dim3(blocks), dim3(256), 0, st, p);
}

static int act_quant_impl(const float* x, int B, int H, int W, int C, const float* scale,
                          const float* zp, int a_bit, int pre, const double* gn_stats,
                          const float* gamma, const float* beta, float eps, int8_t* codes,
                          int32_t* rowsum, int rows_layout, float* y, bool quant,
                          cudaStream_t st) {
  ATTNDM_CHECK_ARG(x && B > 0 && H > 0 && W > 0 && C > 0, "act_quant: bad shape");
  ATTNDM_CHECK_ARG(!quant || (scale && zp && a_bit >= 2 && a_bit <= 8), "act_quant: bad quant params");
  ATTNDM_CHECK_ARG(pre != ATTNDM_PRE_GN_SILU || (gn_stats && gamma && beta && C % kGnGroups == 0 && C % 4 == 0),
                   "act_quant: GroupNorm pre-op needs stats/gamma/beta and C %% 32 == 0");
  ATTNDM_CHECK_ARG(rows_layout == ATTNDM_ROWS_PLAIN || rows_layout == ATTNDM_ROWS_HALO, "act_quant: bad layout");
  ActQuantParams p = {};
  p.x = x; p.B = B; p.H = H; p.W = W; p.C = C; p.Cp = round_up(C, 16);
  p.scale = scale; p.zp = zp;
  p.qlo = quant ? -(float)(1 << (a_bit - 1)) : 0.f;
  p.qhi = quant ? (float)((1 << (a_bit - 1)) - 1) : 0.f;
  p.gn_stats = gn_stats; p.gamma = gamma; p.beta = beta; p.eps = eps;
  p.codes = codes; p.rowsum = rowsum; p.halo = (rows_layout == ATTNDM_ROWS_HALO && codes) ? 1 : 0;
  p.y = y;
  p.x2 = nullptr; p.C1 = 0;
  p.rows = p.halo ? (long long)B * (H + 2) * (W + 2) : (long long)B * H * W;
  // ~8 warps per block; aim at <= 16 resident blocks per SM worth of warps, rows contiguous per warp
  // one wave: the fast kernels keep 3 (C = 128) or 2 (C = 256) blocks of 8 warps resident per SM
  long long max_warps = (long long)kNumSMs * 8 * (C == 128 ? 3 : C == 256 ? 2 : 8);
  long long warps = p.rows < max_warps ? p.rows : max_warps;
  p.rows_per_warp = (p.rows + warps - 1) / warps;
  warps = (p.rows + p.rows_per_warp - 1) / p.rows_per_warp;
  int blocks = cdiv(warps, 8);
  if (quant && codes && rowsum && !y && pre == ATTNDM_PRE_NONE && C <= 16 && (((uintptr_t)codes) & 15) == 0) {
    launch_pdl(act_quant_narrow_kernel, dim3(cdiv(p.rows, 256)), dim3(256), 0, st, p);
    ATTNDM_CUDA_LAUNCH_CHECK("act_quant");
    return ATTNDM_OK;
  }
  // int8 hot path at the large maps: codes + row sums only
  if (quant && codes && rowsum && !y && (C == 128 || C == 256 || C == 384 || C == 512) && (W & 3) == 0 &&
      (pre != ATTNDM_PRE_GN_SILU || (C / kGnGroups) % 4 == 0) && (long long)B * H * W * C < (1LL << 31) &&
      (((uintptr_t)x | (uintptr_t)codes | (uintptr_t)scale | (uintptr_t)zp) & 15) == 0) {
    const int nimg = B * H;
    int w2 = kNumSMs * 8 * (C == 128 ? 4 : C == 256 ? 3 : 2);
    if (w2 > nimg) w2 = nimg;
    p.rows_per_warp = (nimg + w2 - 1) / w2;
    w2 = (nimg + (int)p.rows_per_warp - 1) / (int)p.rows_per_warp;
    const int nb = cdiv(w2, 8);
#define ATTNDM_AQ_ROWS_N(PREV, NQV)                                                                                 \
    do {                                                                                                              \
      if (a_bit == 8) launch_pdl(act_quant_rows_kernel<PREV, NQV, true, false>, dim3(nb), dim3(256), 0, st, p);      \
      else launch_pdl(act_quant_rows_kernel<PREV, NQV, false, false>, dim3(nb), dim3(256), 0, st, p);                \
    } while (0)
#define ATTNDM_AQ_ROWS(PREV)                                                                                         \
    do {                                                                                                              \
      if (C == 128) ATTNDM_AQ_ROWS_N(PREV, 1);                                                                        \
      else if (C == 256) ATTNDM_AQ_ROWS_N(PREV, 2);                                                                   \
      else if (C == 384) ATTNDM_AQ_ROWS_N(PREV, 3);                                                                   \
      else ATTNDM_AQ_ROWS_N(PREV, 4);                                                                                 \
    } while (0)
    if (pre == ATTNDM_PRE_GN_SILU) ATTNDM_AQ_ROWS(ATTNDM_PRE_GN_SILU);
    else if (pre == ATTNDM_PRE_SILU) ATTNDM_AQ_ROWS(ATTNDM_PRE_SILU);
    else ATTNDM_AQ_ROWS(ATTNDM_PRE_NONE);
#undef ATTNDM_AQ_ROWS_N
#undef ATTNDM_AQ_ROWS
    ATTNDM_CUDA_LAUNCH_CHECK("act_quant");
    return ATTNDM_OK;
  }
  if (!quant) {
    if (pre == ATTNDM_PRE_GN_SILU) launch_act_quant<ATTNDM_PRE_GN_SILU, false>(p, blocks, st);
    else if (pre == ATTNDM_PRE_SILU) launch_act_quant<ATTNDM_PRE_SILU, false>(p, blocks, st);
    else launch_act_quant<ATTNDM_PRE_NONE, false>(p, blocks, st);
  } else {
    if (pre == ATTNDM_PRE_GN_SILU) launch_act_quant<ATTNDM_PRE_GN_SILU, true>(p, blocks, st);
    else if (pre == ATTNDM_PRE_SILU) launch_act_quant<ATTNDM_PRE_SILU, true>(p, blocks, st);
    else launch_act_quant<ATTNDM_PRE_NONE, true>(p, blocks, st);
  }
  ATTNDM_CUDA_LAUNCH_CHECK("act_quant");
  return ATTNDM_OK;
}

// ---------------------------------------------------------------------------
// Fused GroupNorm(32) + SiLU + quantize, one CTA per sample, the sample's [HW][C] tile resident in
// shared memory (one HBM read, no separate statistics kernel).  Used for every feature map whose
// per-sample tile fits in 64 KB (up to 8x8x256 fp32): ~75 of the 97 GroupNorms of the CIFAR model.
// ---------------------------------------------------------------------------
struct GnActParams {
  const float* x;
  int B, H, W, C, Cp;
  const float* scale;
  const float* zp;
  float qlo, qhi;
  const float* gamma;
  const float* beta;
  float eps;
  int8_t* codes;
  int32_t* rowsum;
  int halo;
  float* y;
  int quant;
};

__global__ void __launch_bounds__(256) gn_act_quant_sample_kernel(GnActParams p) {
  pdl_enter();
  extern __shared__ float4 tile4[];                 // [HW][C/4]
  __shared__ float s_mean[kGnGroups], s_rstd[kGnGroups];
  const int b = blockIdx.x;
  const int HW = p.H * p.W, Q = p.C >> 2, cpg = p.C / kGnGroups;
  const int n4 = HW * Q;
  const float4* src = reinterpret_cast<const float4*>(p.x + (long long)b * HW * p.C);
  // phase 1: stream the sample's tile into shared memory (coalesced float4)
  for (int e = threadIdx.x; e < n4; e += blockDim.x) tile4[e] = ldg_stream(src + e);
  __syncthreads();
  // phase 1b: one warp per group at a time, double accumulation from shared memory, shuffle reduction
  {
    const int lane_ = threadIdx.x & 31, wid_ = threadIdx.x >> 5, nw_ = blockDim.x >> 5;
    const float* tile = reinterpret_cast<const float*>(tile4);
    const int per_group = HW * cpg;
    const double inv_n = 1.0 / (double)per_group;
    if (nw_ == 8) {
      // the warp's four groups side by side: four independent chains of double adds and 64-bit shuffles instead of one
      // after the other (same arithmetic per group; this phase was most of the kernel's 7 us on a 2x2 map)
      double a[4] = {0.0, 0.0, 0.0, 0.0}, q[4] = {0.0, 0.0, 0.0, 0.0};
      for (int i = lane_; i < per_group; i += 32) {
        const int px = i / cpg, k = i - px * cpg;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const double v = (double)tile[px * p.C + (wid_ + 8 * u) * cpg + k];
          a[u] += v;
          q[u] += v * v;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          a[u] += __shfl_xor_sync(0xffffffffu, a[u], o);
          q[u] += __shfl_xor_sync(0xffffffffu, q[u], o);
        }
      if (lane_ < 4) {
        const double as = lane_ == 0 ? a[0] : lane_ == 1 ? a[1] : lane_ == 2 ? a[2] : a[3];
        const double qs = lane_ == 0 ? q[0] : lane_ == 1 ? q[1] : lane_ == 2 ? q[2] : q[3];
        gn_mean_rstd(as, qs, inv_n, p.eps, s_mean[wid_ + 8 * lane_], s_rstd[wid_ + 8 * lane_]);
      }
    } else
    for (int g = wid_; g < kGnGroups; g += nw_) {
      double a = 0.0, q = 0.0;
      for (int i = lane_; i < per_group; i += 32) {
        const int px = i / cpg, k = i - px * cpg;
        const double v = (double)tile[px * p.C + g * cpg + k];
        a += v;
        q += v * v;
      }
      a = warp_sum_d(a);
      q = warp_sum_d(q);
      if (lane_ == 0) gn_mean_rstd(a, q, inv_n, p.eps, s_mean[g], s_rstd[g]);
    }
  }
  __syncthreads();
  // phase 2: apply, one warp per output row of this sample (halo rows included)
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int Hp = p.halo ? p.H + 2 : p.H, Wp = p.halo ? p.W + 2 : p.W;
  const long long row0 = (long long)b * Hp * Wp;
  for (int rr = wid; rr < Hp * Wp; rr += nw) {
    const int hp = rr / Wp, wp = rr - hp * Wp;
    bool interior = true;
    int h = hp, w = wp;
    if (p.halo) { interior = (hp >= 1 && hp <= p.H && wp >= 1 && wp <= p.W); h = hp - 1; w = wp - 1; }
    const int px = h * p.W + w;
    const long long r = row0 + rr;
    int acc = 0;
    for (int q0 = 0; q0 < Q; q0 += 32) {
      const int q = q0 + lane;
      if (q >= Q) break;
      const int c = q << 2;
      float4 s4 = make_float4(1.f, 1.f, 1.f, 1.f), z4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.quant) {
        s4 = *reinterpret_cast<const float4*>(p.scale + c);
        z4 = *reinterpret_cast<const float4*>(p.zp + c);
      }
      float4 cd;
      if (interior) {
        float4 v = tile4[px * Q + q];
        const float4 g4 = *reinterpret_cast<const float4*>(p.gamma + c);
        const float4 b4 = *reinterpret_cast<const float4*>(p.beta + c);
        const int g0 = c / cpg, g1 = (c + 1) / cpg, g2 = (c + 2) / cpg, g3 = (c + 3) / cpg;
        v.x = gn_apply(v.x, s_mean[g0], s_rstd[g0], g4.x, b4.x);      // SiLU follows below
        v.y = gn_apply(v.y, s_mean[g1], s_rstd[g1], g4.y, b4.y);
        v.z = gn_apply(v.z, s_mean[g2], s_rstd[g2], g4.z, b4.z);
        v.w = gn_apply(v.w, s_mean[g3], s_rstd[g3], g4.w, b4.w);
        const long long pix = (long long)b * HW + px;
        if (p.quant) {
          cd.x = quant_code_t(silu_quant_t(v.x, s4.x, z4.x), p.qlo, p.qhi);
          cd.y = quant_code_t(silu_quant_t(v.y, s4.y, z4.y), p.qlo, p.qhi);
          cd.z = quant_code_t(silu_quant_t(v.z, s4.z, z4.z), p.qlo, p.qhi);
          cd.w = quant_code_t(silu_quant_t(v.w, s4.w, z4.w), p.qlo, p.qhi);
          if (p.y) {
            float4 o;
            o.x = dequant(cd.x, s4.x, z4.x); o.y = dequant(cd.y, s4.y, z4.y);
            o.z = dequant(cd.z, s4.z, z4.z); o.w = dequant(cd.w, s4.w, z4.w);
            *reinterpret_cast<float4*>(p.y + pix * p.C + c) = o;
          }
        } else {
          v.x = silu_f(v.x); v.y = silu_f(v.y); v.z = silu_f(v.z); v.w = silu_f(v.w);
          if (p.y) *reinterpret_cast<float4*>(p.y + pix * p.C + c) = v;
          cd = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      } else {
        cd.x = fminf(fmaxf(-z4.x, p.qlo), p.qhi); cd.y = fminf(fmaxf(-z4.y, p.qlo), p.qhi);
        cd.z = fminf(fmaxf(-z4.z, p.qlo), p.qhi); cd.w = fminf(fmaxf(-z4.w, p.qlo), p.qhi);
      }
      if (p.quant && p.codes) {
        const int ix = (int)cd.x, iy = (int)cd.y, iz = (int)cd.z, iw = (int)cd.w;
        acc += ix + iy + iz + iw;
        *reinterpret_cast<char4*>(p.codes + r * p.Cp + c) =
            make_char4((signed char)ix, (signed char)iy, (signed char)iz, (signed char)iw);
      }
    }
    if (p.quant && p.rowsum) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) p.rowsum[r] = acc;
    }
  }
}



// ---------------------------------------------------------------------------
// GroupNorm statistics
// ---------------------------------------------------------------------------
// block = (C/4 channel quads) x P pixel lanes, fixed quad per thread; grid = (splits, B)
// reverse != 0: blocks walk the tensor from its END.  The producer (a conv epilogue) wrote it front to back and it is
// larger than L2 at the big maps, so the back is what L2 still holds; the consumer of the statistics
// (GroupNorm+SiLU+quantize) then walks front to back again and meets what THIS kernel touched last.
// cpg / g0 / mult: the tensor may be one PART of a GroupNorm input (the upsampled half or the skip half of an
// UpBlock's concat, which is never materialised): its channels fall into groups of cpg channels starting at group g0
// of the stats row, and every sum is scaled by mult (4 for the half that nearest-neighbour upsampling repeats 2x2).
// QUAD: the "quad order" of csrc/conv_common.cuh -- per pixel and aligned group of four channels the fp32 sum
// (x0 + x1) + (x2 + x3) and the fp32 fma chain of the squares, double from there on: the same statistics the tcgen05
// conv epilogue accumulates, for convs computed by another kernel (cpg % 4 == 0 there).
template <int UNROLL, bool QUAD = false>
__global__ void gn_stats_kernel(const float* __restrict__ x, int HW, int C, int P, int rows_per_block,
                                double* __restrict__ stats, int reverse, int cpg, int g0, double mult) {
  pdl_enter();
  __shared__ double s_sum[kGnGroups], s_sq[kGnGroups];
  const int b = reverse ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y;
  const int bx = reverse ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x;
  const int Q = C >> 2;
  if (threadIdx.x < kGnGroups) { s_sum[threadIdx.x] = 0.0; s_sq[threadIdx.x] = 0.0; }
  __syncthreads();
  const int q = threadIdx.x % Q, pl = threadIdx.x / Q;
  int r0 = bx * rows_per_block, r1 = min(HW, r0 + rows_per_block);
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0, q0 = 0, q1 = 0, q2 = 0, q3 = 0;
  if (pl < P) {
    const float* base = x + ((long long)b * HW) * C + (q << 2);
    int r = r0 + pl;
    for (; r + (UNROLL - 1) * P < r1; r += UNROLL * P) {   // UNROLL independent loads in flight, accumulated in row order
      float4 v[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) v[u] = ldg_stream(reinterpret_cast<const float4*>(base + (long long)(r + u * P) * C));
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        if (QUAD) {
          a0 += (double)__fadd_rn(__fadd_rn(v[u].x, v[u].y), __fadd_rn(v[u].z, v[u].w));
          q0 += (double)fmaf(v[u].w, v[u].w, fmaf(v[u].z, v[u].z, fmaf(v[u].y, v[u].y, __fmul_rn(v[u].x, v[u].x))));
        } else {
          a0 += v[u].x; q0 += (double)v[u].x * v[u].x;
          a1 += v[u].y; q1 += (double)v[u].y * v[u].y;
          a2 += v[u].z; q2 += (double)v[u].z * v[u].z;
          a3 += v[u].w; q3 += (double)v[u].w * v[u].w;
        }
      }
    }
    for (; r < r1; r += P) {
      float4 v = ldg_stream(reinterpret_cast<const float4*>(base + (long long)r * C));
      if (QUAD) {
        a0 += (double)__fadd_rn(__fadd_rn(v.x, v.y), __fadd_rn(v.z, v.w));
        q0 += (double)fmaf(v.w, v.w, fmaf(v.z, v.z, fmaf(v.y, v.y, __fmul_rn(v.x, v.x))));
      } else {
        a0 += v.x; q0 += (double)v.x * v.x;
        a1 += v.y; q1 += (double)v.y * v.y;
        a2 += v.z; q2 += (double)v.z * v.z;
        a3 += v.w; q3 += (double)v.w * v.w;
      }
    }
    const int c = q << 2;
    if ((cpg & 3) == 0) {   // all four channels of the quad are in one group
      int g = c / cpg;
      atomicAdd(&s_sum[g], (a0 + a1) + (a2 + a3));
      atomicAdd(&s_sq[g], (q0 + q1) + (q2 + q3));
    } else {
      atomicAdd(&s_sum[c / cpg], a0); atomicAdd(&s_sq[c / cpg], q0);
      atomicAdd(&s_sum[(c + 1) / cpg], a1); atomicAdd(&s_sq[(c + 1) / cpg], q1);
      atomicAdd(&s_sum[(c + 2) / cpg], a2); atomicAdd(&s_sq[(c + 2) / cpg], q2);
      atomicAdd(&s_sum[(c + 3) / cpg], a3); atomicAdd(&s_sq[(c + 3) / cpg], q3);
    }
  }
  __syncthreads();
  if (threadIdx.x < C / cpg) {                  // scaling by a power of two is exact
    atomicAdd(&stats[((long long)b * kGnGroups + g0 + threadIdx.x) * 2 + 0], mult * s_sum[threadIdx.x]);
    atomicAdd(&stats[((long long)b * kGnGroups + g0 + threadIdx.x) * 2 + 1], mult * s_sq[threadIdx.x]);
  }
}

// ---------------------------------------------------------------------------
// per-channel min / max
// ---------------------------------------------------------------------------
constexpr int kMinMaxBlocks = 2 * kNumSMs;

__device__ __forceinline__ void atomic_min_f(float* addr, float v) {
  if (v >= 0.f) atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_f(float* addr, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// pass 1: block partials -> workspace[blk][2][C]
__global__ void minmax_partial_kernel(const float* __restrict__ x, long long rows, int C, int P,
                                      long long rows_per_block, float* __restrict__ ws) {
  extern __shared__ float sm[];   // [2][C]
  float* smin = sm;
  float* smax = sm + C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) { smin[c] = INFINITY; smax[c] = -INFINITY; }
  __syncthreads();
  long long r0 = blockIdx.x * rows_per_block, r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  if ((C & 3) == 0) {
    const int Q = C >> 2;
    const int q = threadIdx.x % Q, pl = threadIdx.x / Q;
    if (pl < P && r0 + pl < r1) {
      float4 mn = make_float4(INFINITY, INFINITY, INFINITY, INFINITY);
      float4 mx = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      for (long long r = r0 + pl; r < r1; r += P) {
        float4 v = ldg_stream(reinterpret_cast<const float4*>(x + r * C + (q << 2)));
        mn.x = fminf(mn.x, v.x); mx.x = fmaxf(mx.x, v.x);
        mn.y = fminf(mn.y, v.y); mx.y = fmaxf(mx.y, v.y);
        mn.z = fminf(mn.z, v.z); mx.z = fmaxf(mx.z, v.z);
        mn.w = fminf(mn.w, v.w); mx.w = fmaxf(mx.w, v.w);
      }
      const int c = q << 2;
      atomic_min_f(&smin[c], mn.x); atomic_max_f(&smax[c], mx.x);
      atomic_min_f(&smin[c + 1], mn.y); atomic_max_f(&smax[c + 1], mx.y);
      atomic_min_f(&smin[c + 2], mn.z); atomic_max_f(&smax[c + 2], mx.z);
      atomic_min_f(&smin[c + 3], mn.w); atomic_max_f(&smax[c + 3], mx.w);
    }
  } else {
    const int c = threadIdx.x % C, pl = threadIdx.x / C;
    if (pl < P && r0 + pl < r1) {
      float mn = INFINITY, mx = -INFINITY;
      for (long long r = r0 + pl; r < r1; r += P) {
        float v = x[r * C + c];
        mn = fminf(mn, v); mx = fmaxf(mx, v);
      }
      atomic_min_f(&smin[c], mn); atomic_max_f(&smax[c], mx);
    }
  }
  __syncthreads();
  float* o = ws + (long long)blockIdx.x * 2 * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) { o[c] = smin[c]; o[C + c] = smax[c]; }
}

__global__ void minmax_final_kernel(const float* __restrict__ ws, int nblk, int C,
                                    float* __restrict__ min_c, float* __restrict__ max_c) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float mn = INFINITY, mx = -INFINITY;
  for (int b = 0; b < nblk; ++b) {
    mn = fminf(mn, ws[(long long)b * 2 * C + c]);
    mx = fmaxf(mx, ws[(long long)b * 2 * C + C + c]);
  }
  min_c[c] = mn;
  max_c[c] = mx;
}

// ---------------------------------------------------------------------------
// group ranges (floor + GroupWise_Quantizaion x2), one block
// ---------------------------------------------------------------------------
__device__ float block_reduce(float v, bool is_max, float* red) {
  v = is_max ? warp_max(v) : warp_min(v);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  int nw = blockDim.x >> 5;
  float r = (threadIdx.x < nw) ? red[threadIdx.x] : (is_max ? -INFINITY : INFINITY);
  if (w == 0) {
    r = is_max ? warp_max(r) : warp_min(r);
    if (lane == 0) red[0] = r;
  }
  __syncthreads();
  r = red[0];
  __syncthreads();
  return r;
}

// one GroupWise pass on vec[C] (already floored, in smem); writes vals[G], xq[C]
__device__ void group_wise_device(const float* vec, int C, int G, bool mode_max, float* edges,
                                  float* vals, int* mark, float* red, float* xq_out) {
  float lmx = -INFINITY, lmn = INFINITY;
  for (int c = threadIdx.x; c < C; c += blockDim.x) { lmx = fmaxf(lmx, vec[c]); lmn = fminf(lmn, vec[c]); }
  float rmax = block_reduce(lmx, true, red);
  float rmin = block_reduce(lmn, false, red);
  float div = __fsub_rn(rmax, rmin);
  if (threadIdx.x == 0) {
    edges[0] = rmin;
    for (int m = 0; m < G; ++m)
      edges[m + 1] = __fadd_rn(rmin, __fdiv_rn(__fmul_rn(div, (float)(m + 1)), (float)G));
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float v = vec[c];
    int mk = 0;
    for (int m = 0; m < G; ++m)
      if (v >= edges[m] && v <= edges[m + 1]) mk = m + 1;   // later bins win ties
    mark[c] = mk;
  }
  __syncthreads();
  for (int m = 0; m < G; ++m) {
    float l = mode_max ? -INFINITY : INFINITY;
    int cnt = 0;
    for (int c = threadIdx.x; c < C; c += blockDim.x)
      if (mark[c] == m + 1) { l = mode_max ? fmaxf(l, vec[c]) : fminf(l, vec[c]); cnt = 1; }
    float r = block_reduce(l, mode_max, red);
    int any = __syncthreads_or(cnt);
    if (threadIdx.x == 0) vals[m] = any ? r : edges[m + 1];   // empty bin -> its upper edge
    __syncthreads();
  }
  if (xq_out)
    for (int c = threadIdx.x; c < C; c += blockDim.x) xq_out[c] = mark[c] ? vals[mark[c] - 1] : 0.f;
  __syncthreads();
}

__global__ void group_ranges_kernel(const float* __restrict__ min_c, const float* __restrict__ max_c, int C,
                                    int G, float init_min, float init_max, float* __restrict__ gr,
                                    float* __restrict__ xq_min, float* __restrict__ xq_max) {
  extern __shared__ float sm[];
  float* vec = sm;                    // [C]
  int* mark = reinterpret_cast<int*>(sm + C);   // [C]
  __shared__ float edges[65], vals[64], red[32];
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float v = min_c[c];
    vec[c] = (v > init_min) ? init_min : v;       // utils/quant_util.py:193-194
  }
  __syncthreads();
  group_wise_device(vec, C, G, false, edges, vals, mark, red, xq_min);
  if (threadIdx.x < G) gr[threadIdx.x * 2 + 0] = vals[threadIdx.x];
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float v = max_c[c];
    vec[c] = (v < init_max) ? init_max : v;       // utils/quant_util.py:195-196
  }
  __syncthreads();
  group_wise_device(vec, C, G, true, edges, vals, mark, red, xq_max);
  if (threadIdx.x < G) gr[threadIdx.x * 2 + 1] = vals[threadIdx.x];
}

// ---------------------------------------------------------------------------
// calibration mix
// ---------------------------------------------------------------------------
constexpr int kMaxGroups = 16;

__global__ void __launch_bounds__(256) calib_mix_kernel(const float* __restrict__ x, long long rows, int C, int G,
                                                        const float* __restrict__ gr, const float* __restrict__ sw,
                                                        int a_bit, float* __restrict__ y, double* lp_sum, float lp_p,
                                                        long long rows_per_warp) {
  __shared__ float s_s[kMaxGroups], s_z[kMaxGroups];
  __shared__ int s_same[kMaxGroups];      // group g has the same (scale, zero point) as group g - 1: its branch output is reused
  __shared__ double s_lp[8];
  if (threadIdx.x < G) {
    float lo = gr[threadIdx.x * 2], hi = gr[threadIdx.x * 2 + 1];
    float n = (float)((1 << a_bit) - 1);
    float s = __fdiv_rn(n, __fsub_rn(hi, lo));                       // quant_utils.py:119-123
    s_s[threadIdx.x] = s;
    s_z[threadIdx.x] = __fadd_rn(rintf(__fmul_rn(s, lo)), (float)(1 << (a_bit - 1)));
  }
  __syncthreads();
  // With the [-4, 6] floor most (often all) groups share one range: the G branches are then the
  // SAME fake-quant of x and differ only in their mixing weight -- one quantize/de-quantize (an IEEE divide) per
  // distinct range instead of G.  Bit-identical: the reused value is the value the branch would recompute.
  if (threadIdx.x < G)
    s_same[threadIdx.x] = threadIdx.x > 0 && s_s[threadIdx.x] == s_s[threadIdx.x - 1] && s_z[threadIdx.x] == s_z[threadIdx.x - 1];
  __syncthreads();
  const float qlo = -(float)(1 << (a_bit - 1)), qhi = (float)((1 << (a_bit - 1)) - 1);
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  long long r0 = warp * rows_per_warp, r1 = r0 + rows_per_warp;
  if (r1 > rows) r1 = rows;
  double lp = 0.0;
  const bool vec = (C & 3) == 0;
  for (long long r = r0; r < r1; ++r) {
    if (vec) {
      for (int q = lane; q < (C >> 2); q += 32) {
        const int c = q << 2;
        float4 v = ldg_stream(reinterpret_cast<const float4*>(x + r * C + c));
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
        for (int g = 0; g < G; ++g) {
          float4 w4 = *reinterpret_cast<const float4*>(sw + (long long)g * C + c);
          if (!s_same[g]) {
            const float s = s_s[g], z = s_z[g];
            d0 = dequant(quant_code(v.x, s, z, qlo, qhi), s, z);
            d1 = dequant(quant_code(v.y, s, z, qlo, qhi), s, z);
            d2 = dequant(quant_code(v.z, s, z, qlo, qhi), s, z);
            d3 = dequant(quant_code(v.w, s, z, qlo, qhi), s, z);
          }
          float t0 = __fmul_rn(d0, w4.x);
          float t1 = __fmul_rn(d1, w4.y);
          float t2 = __fmul_rn(d2, w4.z);
          float t3 = __fmul_rn(d3, w4.w);
          if (g == 0) { acc.x = t0; acc.y = t1; acc.z = t2; acc.w = t3; }
          else {
            acc.x = __fadd_rn(acc.x, t0); acc.y = __fadd_rn(acc.y, t1);
            acc.z = __fadd_rn(acc.z, t2); acc.w = __fadd_rn(acc.w, t3);
          }
        }
        *reinterpret_cast<float4*>(y + r * C + c) = acc;
        if (lp_sum) {
          lp += pow((double)fabsf(acc.x - v.x), (double)lp_p) + pow((double)fabsf(acc.y - v.y), (double)lp_p) +
                pow((double)fabsf(acc.z - v.z), (double)lp_p) + pow((double)fabsf(acc.w - v.w), (double)lp_p);
        }
      }
    } else {
      for (int c = lane; c < C; c += 32) {
        float v = x[r * C + c];
        float acc = 0.f, d = 0.f;
        for (int g = 0; g < G; ++g) {
          if (!s_same[g]) d = dequant(quant_code(v, s_s[g], s_z[g], qlo, qhi), s_s[g], s_z[g]);
          float t = __fmul_rn(d, sw[(long long)g * C + c]);
          acc = (g == 0) ? t : __fadd_rn(acc, t);
        }
        y[r * C + c] = acc;
        if (lp_sum) lp += pow((double)fabsf(acc - v), (double)lp_p);
      }
    }
  }
  if (lp_sum) {
    lp = warp_sum_d(lp);
    if (lane == 0) s_lp[threadIdx.x >> 5] = lp;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0;
      for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += s_lp[i];
      atomicAdd(lp_sum, t);
    }
  }
}

// ---------------------------------------------------------------------------
// k-th smallest (radix select, 4 x 8 bits)
// workspace: hist[4][256] | state: {prefix, k_lo, k_hi(unused), pad...}
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t float_key(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void kth_init_kernel(uint32_t* ws, long long k) {
  for (int i = threadIdx.x; i < 4 * 256 + 8; i += blockDim.x) ws[i] = 0;
  __syncthreads();
  if (threadIdx.x == 0) {
    ws[4 * 256 + 1] = (uint32_t)(k & 0xffffffffu);
    ws[4 * 256 + 2] = (uint32_t)((unsigned long long)k >> 32);
  }
}

__global__ void kth_hist_kernel(const float* __restrict__ x, long long n, int pass, uint32_t* ws) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;     // blockDim.x == 256
  __syncthreads();
  const uint32_t prefix = ws[4 * 256 + 0];
  const int shift = 24 - 8 * pass;
  const uint32_t mask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    uint32_t key = float_key(x[i]);
    if ((key & mask) == (prefix & mask)) atomicAdd(&h[(key >> shift) & 255u], 1u);
  }
  __syncthreads();
  if (h[threadIdx.x]) atomicAdd(&ws[pass * 256 + threadIdx.x], h[threadIdx.x]);
}

__global__ void kth_pick_kernel(int pass, uint32_t* ws, float* out) {
  if (threadIdx.x != 0) return;
  unsigned long long k = (unsigned long long)ws[4 * 256 + 1] | ((unsigned long long)ws[4 * 256 + 2] << 32);
  const int shift = 24 - 8 * pass;
  unsigned long long cum = 0;
  int bin = 255;
  for (int b = 0; b < 256; ++b) {
    unsigned long long c = ws[pass * 256 + b];
    if (k < cum + c) { bin = b; break; }
    cum += c;
  }
  k -= cum;
  uint32_t prefix = ws[4 * 256 + 0] | ((uint32_t)bin << shift);
  ws[4 * 256 + 0] = prefix;
  ws[4 * 256 + 1] = (uint32_t)(k & 0xffffffffu);
  ws[4 * 256 + 2] = (uint32_t)(k >> 32);
  if (pass == 3) {
    uint32_t u = (prefix & 0x80000000u) ? (prefix & 0x7fffffffu) : ~prefix;
    *out = __uint_as_float(u);
  }
}

// ---------------------------------------------------------------------------
// weights
// ---------------------------------------------------------------------------
__global__ void weight_clamp_pack_kernel(const float* __restrict__ w, int O, int C, int taps,
                                         const float* __restrict__ lo, const float* __restrict__ hi,
                                         float* __restrict__ w_eff) {
  long long n = (long long)O * C * taps;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int tap = (int)((i / C) % taps);
    int o = (int)(i / ((long long)C * taps));
    float v = w[((long long)o * C + c) * taps + tap];
    float l = lo[o], h = hi[o];
    // 0.5 * ((-w + lo).abs() - (w - hi).abs() + lo + hi), utils/quant_util.py:293-300
    float a = fabsf(__fadd_rn(-v, l));
    float b = fabsf(__fsub_rn(v, h));
    float t = __fadd_rn(__fadd_rn(__fsub_rn(a, b), l), h);
    w_eff[i] = __fmul_rn(0.5f, t);
  }
}

__global__ void weight_to_i8_kernel(const float* __restrict__ w_eff, int O, int C, int taps,
                                    const float* __restrict__ ws, const float* __restrict__ wz, int w_bit,
                                    int8_t* __restrict__ qw, int Cp, int32_t* __restrict__ wsum,
                                    int32_t* __restrict__ wzp_out, int* on_grid) {
  // One block per output channel.  Codes are only defined up to the shift (q, zp) -> (q - d, zp + d):
  // if rounding the zero point one way pushed the code range to [-127, 128] it is slid back into the
  // signed range instead of declaring the channel off-grid.
  const int o = blockIdx.x;
  const float s = ws[o], z = wz[o];
  const int qlo = -(1 << (w_bit - 1)), qhi = (1 << (w_bit - 1)) - 1;
  bool ok = isfinite(s) && s > 0.f && isfinite(z);
  const int K = taps * Cp;
  __shared__ int s_red[32];
  __shared__ int s_shift;
  int mn = 1 << 30, mx = -(1 << 30);
  for (int i = threadIdx.x; i < K; i += blockDim.x) {
    int tap = i / Cp, c = i - tap * Cp;
    if (c < C) {
      float back = __fsub_rn(__fmul_rn(s, w_eff[((long long)o * taps + tap) * C + c]), z);
      float qf = rintf(back);
      // on the grid if within 1e-3 of a quantization step (the reference clamp identity perturbs
      // on-grid weights by ~1 ulp)
      if (!(fabsf(back - qf) <= 1e-3f) || fabsf(qf) > 1e6f) ok = false;
      int q = ok ? (int)qf : 0;
      mn = min(mn, q);
      mx = max(mx, q);
    }
  }
  for (int off = 16; off > 0; off >>= 1) {
    mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, off));
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (lane == 0) s_red[wid] = mn;
  __syncthreads();
  if (threadIdx.x == 0) { int t = s_red[0]; for (int i = 1; i < nw; ++i) t = min(t, s_red[i]); s_shift = t; }
  __syncthreads();
  mn = s_shift;
  __syncthreads();
  if (lane == 0) s_red[wid] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = s_red[0];
    for (int i = 1; i < nw; ++i) t = max(t, s_red[i]);
    int shift = 0;
    if (t > qhi) shift = t - qhi;          // slide down
    else if (mn < qlo) shift = mn - qlo;   // slide up (negative shift)
    s_shift = shift;
    if (t - shift > qhi || mn - shift < qlo) s_shift = 1 << 20;   // range wider than the grid: off-grid
  }
  __syncthreads();
  const int shift = s_shift;
  if (shift == (1 << 20)) ok = false;
  // The conv epilogue forms acc + zp*wsum + w_zp*cs in int32 (conv_common.cuh): |acc|, |zp*wsum| <= K*2^14 and
  // |cs| <= K*2^8 for a_bit <= 8, so the channel's zero point must satisfy K*2^15 + |w_zp|*K*2^8 < 2^31.  A
  // same-sign or narrow-range channel (w_zp = 2^(w-1) + round(s*lo), arbitrarily large) is declared off-grid
  // and the layer takes the fp32 kernel instead of overflowing silently.
  if (ok) {
    const long long wz_final = (long long)z + shift;
    const long long mag = wz_final < 0 ? -wz_final : wz_final;
    if ((long long)K * 32768 + mag * (long long)K * 256 >= 2147483648LL) ok = false;
  }
  int acc = 0;
  for (int i = threadIdx.x; i < K; i += blockDim.x) {
    int tap = i / Cp, c = i - tap * Cp;
    int q = 0;
    if (c < C && ok) {
      float qf = rintf(__fsub_rn(__fmul_rn(s, w_eff[((long long)o * taps + tap) * C + c]), z));
      q = (int)qf - shift;
      q = max(qlo, min(qhi, q));
    }
    qw[(long long)o * K + i] = (int8_t)q;
    acc += q;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  __syncthreads();
  if (lane == 0) s_red[wid] = acc;
  if (!ok) atomicAnd(on_grid, 0);
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int i = 0; i < nw; ++i) t += s_red[i];
    wsum[o] = t;
    wzp_out[o] = (ok && shift != (1 << 20)) ? (int)z + shift : 0;
  }
}

__global__ void set_int_kernel(int* p, int v) { *p = v; }

}  // namespace attndm

using namespace attndm;

extern "C" {

int attndm_act_quant(const float* x, int B, int H, int W, int C, const float* scale, const float* zp,
                     int a_bit, int pre_op, const double* gn_stats, const float* gn_gamma,
                     const float* gn_beta, float gn_eps, int8_t* codes, int32_t* rowsum,
                     int rows_layout, float* y_f32, void* stream) {
  ATTNDM_CHECK_ARG(codes || y_f32, "act_quant: no output requested");
  if (a_bit == 0) {   // quantizer off: y = pre_op(x) (the calibration branch needs the un-quantized producer)
    ATTNDM_CHECK_ARG(y_f32 && !codes, "act_quant: a_bit == 0 only produces y_f32");
    return act_quant_impl(x, B, H, W, C, nullptr, nullptr, 0, pre_op, gn_stats, gn_gamma, gn_beta, gn_eps, nullptr,
                          nullptr, ATTNDM_ROWS_PLAIN, y_f32, false, (cudaStream_t)stream);
  }
  return act_quant_impl(x, B, H, W, C, scale, zp, a_bit, pre_op, gn_stats, gn_gamma, gn_beta, gn_eps, codes,
                        rowsum, rows_layout, y_f32, true, (cudaStream_t)stream);
}

int attndm_act_quant_cat_fits(int H, int W, int C1, int C2) {
  const int C = C1 + C2;
  return (C1 > 0 && C2 > 0 && C1 % 128 == 0 && (C == 256 || C == 384 || C == 512) && (W & 3) == 0 && (H & 1) == 0 &&
          (C / kGnGroups) % 4 == 0 && C1 % (C / kGnGroups) == 0) ? 1 : 0;
}

static int act_quant_cat_impl(const float* xa, int C1, const float* xb, int C2, int B, int H, int W, const float* scale,
                              const float* zp, int a_bit, int pre_op, const double* gn_stats, const float* gn_gamma,
                              const float* gn_beta, float gn_eps, int8_t* codes, int32_t* rowsum, int rows_layout,
                              const float* scale2, const float* zp2, int8_t* codes2, int32_t* rowsum2, int rows_layout2,
                              void* stream) {
  const bool dual = scale2 != nullptr;
  ATTNDM_CHECK_ARG(xa && xb && B > 0 && H > 0 && W > 0 && scale && zp && codes && rowsum && a_bit >= 2 && a_bit <= 8,
                   "act_quant_cat: bad args");
  ATTNDM_CHECK_ARG(rows_layout == ATTNDM_ROWS_PLAIN || rows_layout == ATTNDM_ROWS_HALO, "act_quant_cat: bad layout");
  ATTNDM_CHECK_ARG(pre_op == ATTNDM_PRE_NONE || pre_op == ATTNDM_PRE_SILU || (pre_op == ATTNDM_PRE_GN_SILU && gn_stats && gn_gamma && gn_beta),
                   "act_quant_cat: bad pre-op");
  ATTNDM_CHECK_ARG(!dual || (zp2 && codes2 && rowsum2 && codes2 != codes && rowsum2 != rowsum &&
                             (rows_layout2 == ATTNDM_ROWS_PLAIN || rows_layout2 == ATTNDM_ROWS_HALO)),
                   "act_quant_cat2: bad second quantizer");
  const int C = C1 + C2;
  if (!attndm_act_quant_cat_fits(H, W, C1, C2) || (long long)B * H * W * C >= (1LL << 31) ||
      ((((uintptr_t)xa | (uintptr_t)xb | (uintptr_t)codes | (uintptr_t)scale | (uintptr_t)zp | (uintptr_t)scale2 |
         (uintptr_t)zp2 | (uintptr_t)codes2) & 15) != 0)) {
    set_error("act_quant_cat: shape %dx%dx(%d+%d) not supported (see attndm_act_quant_cat_fits)", H, W, C1, C2);
    return ATTNDM_ERR_UNSUPPORTED;
  }
  ActQuantParams p = {};
  p.x = xa; p.x2 = xb; p.C1 = C1;
  p.B = B; p.H = H; p.W = W; p.C = C; p.Cp = round_up(C, 16);
  p.scale = scale; p.zp = zp;
  p.qlo = -(float)(1 << (a_bit - 1));
  p.qhi = (float)((1 << (a_bit - 1)) - 1);
  p.gn_stats = gn_stats; p.gamma = gn_gamma; p.beta = gn_beta; p.eps = gn_eps;
  p.codes = codes; p.rowsum = rowsum; p.halo = rows_layout == ATTNDM_ROWS_HALO ? 1 : 0;
  p.y = nullptr;
  p.rows = p.halo ? (long long)B * (H + 2) * (W + 2) : (long long)B * H * W;
  p.scale2 = scale2; p.zp2 = zp2; p.codes2 = codes2; p.rowsum2 = rowsum2;
  p.halo2 = rows_layout2 == ATTNDM_ROWS_HALO ? 1 : 0;
  const int nimg = B * H;
  int w2 = kNumSMs * 8 * (dual ? 2 : (C == 256 ? 3 : 2));
  if (w2 > nimg) w2 = nimg;
  p.rows_per_warp = (nimg + w2 - 1) / w2;
  w2 = (nimg + (int)p.rows_per_warp - 1) / (int)p.rows_per_warp;
  const int nb = cdiv(w2, 8);
  cudaStream_t st = (cudaStream_t)stream;
#define ATTNDM_AQ_CAT_N(PREV, NQV, DUALV)                                                                           \
  do {                                                                                                              \
    if (a_bit == 8) launch_pdl(act_quant_rows_kernel<PREV, NQV, true, true, DUALV>, dim3(nb), dim3(256), 0, st, p); \
    else launch_pdl(act_quant_rows_kernel<PREV, NQV, false, true, DUALV>, dim3(nb), dim3(256), 0, st, p);           \
  } while (0)
#define ATTNDM_AQ_CAT(PREV, DUALV)                                                                                  \
  do {                                                                                                              \
    if (C == 256) ATTNDM_AQ_CAT_N(PREV, 2, DUALV);                                                                  \
    else if (C == 384) ATTNDM_AQ_CAT_N(PREV, 3, DUALV);                                                             \
    else ATTNDM_AQ_CAT_N(PREV, 4, DUALV);                                                                           \
  } while (0)
  if (dual) {
    // the pair a ResidualBlock asks for: conv1 behind GroupNorm+SiLU and the shortcut conv on the raw input
    ATTNDM_CHECK_ARG(pre_op == ATTNDM_PRE_GN_SILU, "act_quant_cat2: the main quantizer must be the GroupNorm+SiLU one");
    ATTNDM_AQ_CAT(ATTNDM_PRE_GN_SILU, true);
  } else if (pre_op == ATTNDM_PRE_GN_SILU) ATTNDM_AQ_CAT(ATTNDM_PRE_GN_SILU, false);
  else if (pre_op == ATTNDM_PRE_SILU) ATTNDM_AQ_CAT(ATTNDM_PRE_SILU, false);
  else ATTNDM_AQ_CAT(ATTNDM_PRE_NONE, false);
#undef ATTNDM_AQ_CAT_N
#undef ATTNDM_AQ_CAT
  ATTNDM_CUDA_LAUNCH_CHECK("act_quant_cat");
  return ATTNDM_OK;
}

int attndm_act_quant_cat(const float* xa, int C1, const float* xb, int C2, int B, int H, int W, const float* scale,
                         const float* zp, int a_bit, int pre_op, const double* gn_stats, const float* gn_gamma,
                         const float* gn_beta, float gn_eps, int8_t* codes, int32_t* rowsum, int rows_layout,
                         void* stream) {
  return act_quant_cat_impl(xa, C1, xb, C2, B, H, W, scale, zp, a_bit, pre_op, gn_stats, gn_gamma, gn_beta, gn_eps, codes,
                            rowsum, rows_layout, nullptr, nullptr, nullptr, nullptr, ATTNDM_ROWS_PLAIN, stream);
}

int attndm_act_quant_cat2(const float* xa, int C1, const float* xb, int C2, int B, int H, int W, const float* scale,
                          const float* zp, int a_bit, const double* gn_stats, const float* gn_gamma,
                          const float* gn_beta, float gn_eps, int8_t* codes, int32_t* rowsum, int rows_layout,
                          const float* scale2, const float* zp2, int8_t* codes2, int32_t* rowsum2, int rows_layout2,
                          void* stream) {
  ATTNDM_CHECK_ARG(scale2 != nullptr, "act_quant_cat2: null second scale");
  return act_quant_cat_impl(xa, C1, xb, C2, B, H, W, scale, zp, a_bit, ATTNDM_PRE_GN_SILU, gn_stats, gn_gamma, gn_beta,
                            gn_eps, codes, rowsum, rows_layout, scale2, zp2, codes2, rowsum2, rows_layout2, stream);
}

int attndm_gn_act_quant_fits(int H, int W, int C) {
  static int max_hw = -1;
  if (max_hw < 0) {
    const char* e = getenv("ATTNDM_GN_FUSED_MAX_HW");
    max_hw = e ? atoi(e) : 4;     // above 2x2 the statistics kernel + the row kernel are faster (measured: 9 us vs 25 us at 8x8x128)
  }
  return (C % kGnGroups == 0) && (C % 4 == 0) && ((long long)H * W * C * 4 <= 64 * 1024) && H * W <= max_hw ? 1 : 0;
}

int attndm_gn_act_quant(const float* x, int B, int H, int W, int C, const float* gamma, const float* beta,
                        float eps, const float* scale, const float* zp, int a_bit, int8_t* codes,
                        int32_t* rowsum, int rows_layout, float* y_f32, void* stream) {
  ATTNDM_CHECK_ARG(x && gamma && beta && B > 0 && H > 0 && W > 0 && C > 0, "gn_act_quant: bad args");
  ATTNDM_CHECK_ARG(codes || y_f32, "gn_act_quant: no output requested");
  ATTNDM_CHECK_ARG(a_bit == 0 || (scale && zp && a_bit >= 2 && a_bit <= 8), "gn_act_quant: bad quant params");
  ATTNDM_CHECK_ARG(a_bit != 0 || (y_f32 && !codes), "gn_act_quant: a_bit == 0 only produces y_f32");
  if (!attndm_gn_act_quant_fits(H, W, C)) {
    set_error("gn_act_quant: per-sample tile %dx%dx%d does not fit in shared memory", H, W, C);
    return ATTNDM_ERR_UNSUPPORTED;
  }
  GnActParams p;
  p.x = x; p.B = B; p.H = H; p.W = W; p.C = C; p.Cp = round_up(C, 16);
  p.scale = scale; p.zp = zp; p.quant = a_bit != 0;
  p.qlo = p.quant ? -(float)(1 << (a_bit - 1)) : 0.f;
  p.qhi = p.quant ? (float)((1 << (a_bit - 1)) - 1) : 0.f;
  p.gamma = gamma; p.beta = beta; p.eps = eps; p.codes = codes; p.rowsum = rowsum;
  p.halo = (rows_layout == ATTNDM_ROWS_HALO && codes) ? 1 : 0;
  p.y = y_f32;
  const size_t smem = (size_t)H * W * C * sizeof(float);
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(gn_act_quant_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         64 * 1024);
    if (e != cudaSuccess) { set_error("gn_act_quant: smem attr: %s", cudaGetErrorString(e)); return ATTNDM_ERR_CUDA; }
    smem_set = 64 * 1024;
  }
  launch_pdl(gn_act_quant_sample_kernel, dim3(B), dim3(256), smem, (cudaStream_t)stream, p);
  ATTNDM_CUDA_LAUNCH_CHECK("gn_act_quant");
  return ATTNDM_OK;
}

int attndm_gn_silu(const float* x, int B, int H, int W, int C, const double* gn_stats, const float* gamma,
                   const float* beta, float eps, float* y, void* stream) {
  ATTNDM_CHECK_ARG(y, "gn_silu: y is NULL");
  return act_quant_impl(x, B, H, W, C, nullptr, nullptr, 0, ATTNDM_PRE_GN_SILU, gn_stats, gamma, beta, eps,
                        nullptr, nullptr, ATTNDM_ROWS_PLAIN, y, false, (cudaStream_t)stream);
}

static int gn_stats_impl(const float* x, int B, int HW, int C, int cpg, int g0, double mult, double* stats, cudaStream_t st,
                         bool quad = false) {
  const int Q = C / 4;
  int P = Q >= 256 ? 1 : 256 / Q;
  if (P > HW) P = HW;
  int threads = round_up(Q * P, 32);
  static const int tune_ctas = [] { const char* e = getenv("ATTNDM_GN_CTAS_PER_SM"); return e ? atoi(e) : 8; }();
  static const int tune_unroll = [] { const char* e = getenv("ATTNDM_GN_UNROLL"); return e ? atoi(e) : 4; }();
  static const int reverse = [] { const char* e = getenv("ATTNDM_GN_REVERSE"); return e ? atoi(e) : 1; }();
  int splits = cdiv(tune_ctas * kNumSMs, B);
  int max_splits = cdiv(HW, P * 16);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int rows_per_block = cdiv(HW, splits);
  splits = cdiv(HW, rows_per_block);
  dim3 grid(splits, B);
  // (the partial sums of a sample's `splits` blocks meet in double-precision atomics: their order can change the last
  // bit of a double, far below the fp32 mean / rstd the consumers form from them)
  if (quad)
    launch_pdl(gn_stats_kernel<4, true>, dim3(grid), dim3(threads), 0, st, x, HW, C, P, rows_per_block, stats, reverse, cpg, g0, mult);
  else if (tune_unroll >= 8)
    launch_pdl(gn_stats_kernel<8>, dim3(grid), dim3(threads), 0, st, x, HW, C, P, rows_per_block, stats, reverse, cpg, g0, mult);
  else
    launch_pdl(gn_stats_kernel<4>, dim3(grid), dim3(threads), 0, st, x, HW, C, P, rows_per_block, stats, reverse, cpg, g0, mult);
  ATTNDM_CUDA_LAUNCH_CHECK("gn_stats");
  return ATTNDM_OK;
}

int attndm_gn_stats(const float* x, int B, int H, int W, int C, double* stats, void* stream) {
  ATTNDM_CHECK_ARG(x && stats && B > 0 && H > 0 && W > 0, "gn_stats: bad args");
  ATTNDM_CHECK_ARG(C % kGnGroups == 0 && C % 4 == 0 && C <= 4096, "gn_stats: C must be a multiple of 32, <= 4096");
  return gn_stats_impl(x, B, H * W, C, C / kGnGroups, 0, 1.0, stats, (cudaStream_t)stream);
}

extern "C++" {
namespace attndm {
// GroupNorm statistics of a conv output in quad order, from the stored output (the twin of the tcgen05 epilogue's
// accumulation: same fp32 quad sums, double above; conv_common.cuh)
int launch_gn_stats_quad(const float* out, int B, int HW, int C, double* stats, cudaStream_t st) {
  ATTNDM_CHECK_ARG(out && stats && C % 128 == 0, "gn_stats_quad: C must be a multiple of 128");
  return gn_stats_impl(out, B, HW, C, C / kGnGroups, 0, 1.0, stats, st, true);
}
}  // namespace attndm
}  // extern "C++"

int attndm_gn_stats_cat(const float* xa, int Ha, int Wa, int C1, const float* xb, int H, int W, int C2, int B,
                        double* stats, void* stream) {
  ATTNDM_CHECK_ARG(xa && xb && stats && B > 0 && Ha > 0 && Wa > 0 && H == 2 * Ha && W == 2 * Wa,
                   "gn_stats_cat: the first part must be half the size of the second");
  const int C = C1 + C2;
  ATTNDM_CHECK_ARG(C1 > 0 && C2 > 0 && C % kGnGroups == 0 && C <= 4096, "gn_stats_cat: C1 + C2 must be a multiple of 32, <= 4096");
  const int cpg = C / kGnGroups;
  ATTNDM_CHECK_ARG(cpg % 4 == 0 && C1 % cpg == 0, "gn_stats_cat: the parts must meet on a group boundary (groups of a multiple of 4 channels)");
  // nearest-neighbour x2 repeats every element of the first part 2x2 times: its sums are 4x the low-resolution sums
  int rc = gn_stats_impl(xa, B, Ha * Wa, C1, cpg, 0, 4.0, stats, (cudaStream_t)stream);
  if (rc) return rc;
  return gn_stats_impl(xb, B, H * W, C2, cpg, C1 / cpg, 1.0, stats, (cudaStream_t)stream);
}

int attndm_minmax_workspace_blocks(void) { return kMinMaxBlocks; }

int attndm_minmax_c(const float* x, long long rows, int C, float* min_c, float* max_c, float* workspace,
                    void* stream) {
  ATTNDM_CHECK_ARG(x && min_c && max_c && workspace && rows > 0 && C > 0 && C <= 4096, "minmax_c: bad args");
  int P, threads;
  if ((C & 3) == 0) {
    int Q = C / 4;
    P = Q >= 256 ? 1 : 256 / Q;
    threads = round_up(Q * P, 32);
  } else {
    P = C >= 256 ? 1 : 256 / C;
    threads = round_up(C * P, 32);
  }
  long long rpb = (rows + kMinMaxBlocks - 1) / kMinMaxBlocks;
  if (rpb < P) rpb = P;
  int nblk = (int)((rows + rpb - 1) / rpb);
  size_t smem = 2 * (size_t)C * sizeof(float);
  minmax_partial_kernel<<<nblk, threads, smem, (cudaStream_t)stream>>>(x, rows, C, P, rpb, workspace);
  ATTNDM_CUDA_LAUNCH_CHECK("minmax_partial");
  minmax_final_kernel<<<cdiv(C, 128), 128, 0, (cudaStream_t)stream>>>(workspace, nblk, C, min_c, max_c);
  ATTNDM_CUDA_LAUNCH_CHECK("minmax_final");
  return ATTNDM_OK;
}

int attndm_group_ranges(const float* min_c, const float* max_c, int C, int G, float init_min, float init_max,
                        float* groups_range_t, float* xq_min, float* xq_max, void* stream) {
  ATTNDM_CHECK_ARG(min_c && max_c && groups_range_t && C > 0 && C <= 4096 && G >= 1 && G <= 64,
                   "group_ranges: bad args");
  size_t smem = 2 * (size_t)C * sizeof(float);
  group_ranges_kernel<<<1, 256, smem, (cudaStream_t)stream>>>(min_c, max_c, C, G, init_min, init_max,
                                                             groups_range_t, xq_min, xq_max);
  ATTNDM_CUDA_LAUNCH_CHECK("group_ranges");
  return ATTNDM_OK;
}

int attndm_calib_mix(const float* x, long long rows, int C, int G, const float* groups_range_t, const float* sw,
                     int a_bit, float* y, double* lp_sum, float lp_p, void* stream) {
  ATTNDM_CHECK_ARG(x && y && groups_range_t && sw && rows > 0 && C > 0, "calib_mix: bad args");
  ATTNDM_CHECK_ARG(G >= 1 && G <= kMaxGroups && a_bit >= 2 && a_bit <= 8, "calib_mix: G <= 16, 2 <= a_bit <= 8");
  // one wave: the fast kernels keep 3 (C = 128) or 2 (C = 256) blocks of 8 warps resident per SM
  long long max_warps = (long long)kNumSMs * 8 * (C == 128 ? 3 : C == 256 ? 2 : 8);
  long long warps = rows < max_warps ? rows : max_warps;
  long long rpw = (rows + warps - 1) / warps;
  warps = (rows + rpw - 1) / rpw;
  calib_mix_kernel<<<cdiv(warps, 8), 256, 0, (cudaStream_t)stream>>>(x, rows, C, G, groups_range_t, sw, a_bit, y,
                                                                    lp_sum, lp_p, rpw);
  ATTNDM_CUDA_LAUNCH_CHECK("calib_mix");
  return ATTNDM_OK;
}

int attndm_kth_value(const float* x, long long n, long long k, float* out, uint32_t* workspace, void* stream) {
  ATTNDM_CHECK_ARG(x && out && workspace && n > 0 && k >= 0 && k < n, "kth_value: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  kth_init_kernel<<<1, 256, 0, st>>>(workspace, k);
  int blocks = cdiv(n, 256 * 8);
  if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
  for (int pass = 0; pass < 4; ++pass) {
    kth_hist_kernel<<<blocks, 256, 0, st>>>(x, n, pass, workspace);
    kth_pick_kernel<<<1, 32, 0, st>>>(pass, workspace, out);
  }
  ATTNDM_CUDA_LAUNCH_CHECK("kth_value");
  return ATTNDM_OK;
}

int attndm_weight_clamp_pack(const float* w, int O, int C, int KH, int KW, const float* lo, const float* hi,
                             float* w_eff, void* stream) {
  ATTNDM_CHECK_ARG(w && lo && hi && w_eff && O > 0 && C > 0 && KH > 0 && KW > 0, "weight_clamp_pack: bad args");
  long long n = (long long)O * C * KH * KW;
  int blocks = cdiv(n, 256);
  if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
  weight_clamp_pack_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w, O, C, KH * KW, lo, hi, w_eff);
  ATTNDM_CUDA_LAUNCH_CHECK("weight_clamp_pack");
  return ATTNDM_OK;
}

int attndm_weight_to_i8(const float* w_eff, int O, int C, int taps, const float* w_scale, const float* w_zp,
                        int w_bit, int8_t* qw, int Cp, int32_t* wsum, int32_t* w_zp_i32, int* on_grid,
                        void* stream) {
  ATTNDM_CHECK_ARG(w_eff && w_scale && w_zp && qw && wsum && w_zp_i32 && on_grid, "weight_to_i8: null pointer");
  ATTNDM_CHECK_ARG(Cp >= C && Cp % 16 == 0 && w_bit >= 2 && w_bit <= 8, "weight_to_i8: bad Cp / w_bit");
  set_int_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(on_grid, 1);
  weight_to_i8_kernel<<<O, 256, 0, (cudaStream_t)stream>>>(w_eff, O, C, taps, w_scale, w_zp, w_bit, qw, Cp, wsum,
                                                          w_zp_i32, on_grid);
  ATTNDM_CUDA_LAUNCH_CHECK("weight_to_i8");
  return ATTNDM_OK;
}

}  // extern "C"
