// Parameter block and epilogue arithmetic shared by the two int8 convolution
// kernels (conv_simt.cu, conv_tc.cu), so that both produce bit-identical output.
#pragma once
#include "common.cuh"

namespace attndm {

struct ConvI8Params {
  const int8_t* codes;     // [rows][Cp]
  const int32_t* rowsum;   // [rows]
  int B, H, W, C, Cp;
  int Hp, Wp;              // row geometry: H+2, W+2 for taps == 9 (halo rows), H, W for taps == 1
  long long rows;
  const int8_t* qw;        // [O][taps*Cp]
  const int32_t* wsum;     // [O]
  const int32_t* w_zp;     // [O]
  int O, taps;
  const float* mult;       // [O]
  const int32_t* act_zp;   // [1]
  const float* bias;       // [O] or NULL
  const float* residual;   // [B*H*W][O] or NULL
  const float* temb;       // [B][O] or NULL
  float* out;              // [B*H*W][O]
  double* gn_out;          // [B][32][2] or NULL: GroupNorm {sum, sumsq} of `out`, accumulated (quad order, below)
};

// code-row index -> output pixel.  For 3x3 the GEMM row r is the halo-layout row of
// the window's top-left corner, so output pixel (h, w) = (hp, wp) and it is valid for
// hp < H, wp < W.
__device__ __forceinline__ bool conv_row_to_pixel(const ConvI8Params& p, long long row, long long& pix, int& b) {
  if (row >= p.rows) return false;
  if (p.taps == 1) {
    pix = row;
    b = (int)(row / ((long long)p.H * p.W));
    return true;
  }
  const long long per = (long long)p.Hp * p.Wp;
  b = (int)(row / per);
  const int rem = (int)(row - (long long)b * per);
  const int hp = rem / p.Wp, wp = rem - hp * p.Wp;
  if (hp >= p.H || wp >= p.W) return false;
  pix = ((long long)b * p.H + hp) * p.W + wp;
  return true;
}

// sum of the per-pixel code sums over the receptive field of GEMM row `row`
__device__ __forceinline__ long long conv_window_rowsum(const ConvI8Params& p, long long row) {
  if (p.taps == 1) return p.rowsum[row];
  long long s = 0;
#pragma unroll
  for (int kh = 0; kh < 3; ++kh)
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) s += p.rowsum[row + (long long)kh * p.Wp + kw];
  return s;
}

// Exact integer I = acc + A + B*cs with A = zp*wsum[o], B = w_zp[o], cs = window rowsum + zp*taps*C,
// then one fused multiply-add: out = float(I) * mult[o] + bias[o].
// Range: on the integer path 0.0 is representable, so |zp| <= 2^(a-1) <= 128; |code|, |q| <= 128 give
// |acc|, |A| <= K*2^14 and |cs| <= K*2^8.  |w_zp| is NOT bounded by the bit width (a channel whose weights do
// not straddle 0 has w_zp = 2^(w-1) + round(s*lo)): attndm_weight_to_i8 declares a channel off-grid unless
// K*2^15 + |w_zp|*K*2^8 < 2^31, and the launcher checks K = taps*Cp < 2^15, so I always fits in int32.
// Both conv kernels call this, so they agree bit for bit.
__device__ __forceinline__ float conv_i8_value(int acc, int A, int B, int cs, float m, float bias) {
  return fmaf((float)(acc + A + B * cs), m, bias);
}

__device__ __forceinline__ float conv_epilogue_add(const float* residual, const float* temb, float v,
                                                   long long pix, int b, int o, int O) {
  if (residual) v = __fadd_rn(v, residual[pix * O + o]);     // ResidualBlock: x + h
  if (temb) v = __fadd_rn(v, temb[(long long)b * O + o]);    // block: x + time_mlp(t_emb)
  return v;
}

// ---- GroupNorm statistics of a conv output in "quad" order -------------------------------------------------------
// The statistics a conv hands to the GroupNorm behind it are DEFINED as follows, so that any kernel can reproduce them:
//   per output pixel and aligned group of four channels (x0..x3), fp32:
//       s = (x0 + x1) + (x2 + x3),   q = fma(x3,x3, fma(x2,x2, fma(x1,x1, x0*x0)))
//   per (sample, GroupNorm group): the sum of these s and q in DOUBLE precision, in whatever order the kernel meets
//   them (thread partials, shuffles, atomics): reordering moves the last bits of a double, far below the fp32 mean /
//   rstd formed from the totals -- as in attndm_gn_stats.
// Nothing in the fp32 part depends on how rows fall into tiles, so a sample's statistics do not depend on its position
// in the batch (sharding a batch over GPUs gives the same images, tests/test_gpu_parity.py).
// The tcgen05 epilogue holds the output as 128-row tiles of GEMM rows, a 32-row quarter per warp, thread (tr = lane / 4,
// tq = lane % 4) owning rows tr + 8k (k = 0..3) x channels 4tq .. 4tq+3 of every 16-channel unit; a quarter that
// straddles two samples keeps two partials.
// Supported when O % 128 == 0 (a thread's four channels share a group) and a sample has >= 32 GEMM rows.
inline bool conv_gn_quad_ok(const ConvI8Params& p) {
  return p.O % 128 == 0 && (long long)p.Hp * p.Wp >= 32 && p.rows + 128 < (1LL << 31);
}
int launch_gn_stats_quad(const float* out, int B, int HW, int C, double* stats, cudaStream_t st);   // from the stored output (quant_kernels.cu)

int launch_qconv_i8_simt(const ConvI8Params& p, cudaStream_t st);
// *stats_fused = true when the kernel that ran also accumulated p.gn_out in its epilogue
int launch_qconv_i8_tc(const ConvI8Params& p, cudaStream_t st, bool* stats_fused);
int conv_f32_tc_fits(long long rows, int C, int O);
int launch_split_tf32(const float* x, long long n, float* big, float* small, cudaStream_t st);
int launch_gemm_tf32x3(const float* a_big, const float* a_small, long long rows, int C, const float* w_big,
                       const float* w_small, int O, const float* bias, float* out, cudaStream_t st);

}  // namespace attndm
