// CUDA-core convolutions (any shape): the dp4a int8 implicit GEMM that is the
// correctness twin of the tcgen05 kernel in conv_tc.cu, and the fp32 implicit
// GEMM used where the operand is not on one integer grid (calibration branch,
// trained alpha_activ, off-grid weights).  Reference op: F.conv2d at
// utils/quant_util.py:385 (3x3/s1/p1 and 1x1, groups=1).
#include "common.cuh"
#include "conv_common.cuh"

namespace attndm {

// ---------------------------------------------------------------------------
// int8 implicit GEMM, 64 rows x 64 outputs per CTA, dp4a
// ---------------------------------------------------------------------------
constexpr int BM = 64, BN = 64, KCB = 64;   // KCB bytes of K per smem chunk

__global__ void __launch_bounds__(256) qconv_i8_simt_kernel(ConvI8Params p) {
  pdl_enter();
  __shared__ int As[BM][KCB / 4 + 1];
  __shared__ int Bs[BN][KCB / 4 + 1];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int K = p.taps * p.Cp;
  int acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0;

  const int lrow = tid >> 2, lq = tid & 3;   // loader: row 0..63, 16-byte quarter 0..3
  for (int tap = 0; tap < p.taps; ++tap) {
    const long long shift = p.taps == 9 ? (long long)(tap / 3) * p.Wp + (tap % 3) : 0;
    for (int kc = 0; kc < p.Cp; kc += KCB) {
      uint4 av = make_uint4(0, 0, 0, 0), bv = make_uint4(0, 0, 0, 0);
      const int kb = kc + lq * 16;
      long long ar = m0 + lrow + shift;
      if (kb < p.Cp && ar < p.rows) av = *reinterpret_cast<const uint4*>(p.codes + ar * p.Cp + kb);
      if (kb < p.Cp && n0 + lrow < p.O)
        bv = *reinterpret_cast<const uint4*>(p.qw + (long long)(n0 + lrow) * K + (long long)tap * p.Cp + kb);
      __syncthreads();
      As[lrow][lq * 4 + 0] = av.x; As[lrow][lq * 4 + 1] = av.y; As[lrow][lq * 4 + 2] = av.z; As[lrow][lq * 4 + 3] = av.w;
      Bs[lrow][lq * 4 + 0] = bv.x; Bs[lrow][lq * 4 + 1] = bv.y; Bs[lrow][lq * 4 + 2] = bv.z; Bs[lrow][lq * 4 + 3] = bv.w;
      __syncthreads();
#pragma unroll
      for (int k4 = 0; k4 < KCB / 4; ++k4) {
        int a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[ty * 4 + i][k4];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Bs[tx * 4 + j][k4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = __dp4a(a[i], b[j], acc[i][j]);
      }
    }
  }
  // epilogue
  const int zp = *p.act_zp;
  const int ktot = p.taps * p.C;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long row = m0 + ty * 4 + i;
    long long pix;
    int b;
    if (!conv_row_to_pixel(p, row, pix, b)) continue;
    const int cs = (int)conv_window_rowsum(p, row) + zp * ktot;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int o = n0 + tx * 4 + j;
      if (o >= p.O) continue;
      float v = conv_i8_value(acc[i][j], zp * p.wsum[o], p.w_zp[o], cs, p.mult[o], p.bias ? p.bias[o] : 0.f);
      p.out[pix * p.O + o] = conv_epilogue_add(p.residual, p.temb, v, pix, b, o, p.O);
    }
  }
}

// ---------------------------------------------------------------------------
// fp32 implicit GEMM on NHWC, 64 pixels x 64 outputs per CTA
// ---------------------------------------------------------------------------
struct ConvF32Params {
  const float* x;
  int B, H, W, C;
  const float* w;     // [O][taps][C]
  int O, taps;
  const float* bias;
  const float* residual;
  const float* temb;
  float* out;
  long long rows;     // B*H*W
};

constexpr int KCF = 16;

// TM x TM outputs per thread, (16 TM) x (16 TM) per CTA: TM = 4 (64 x 64 tiles; small problems, many CTAs) or
// TM = 8 (128 x 128; the large channel_proj GEMMs, four 128-bit smem loads per 32 packed FMAs).
template <int TM>
__global__ void __launch_bounds__(256) conv_f32_simt_kernel(ConvF32Params p) {
  pdl_enter();
  constexpr int TB = 16 * TM;                          // tile edge
  constexpr int NL = TM / 4;                           // float4 loads per operand per thread and chunk
  __shared__ __align__(16) float As[KCF][TB + 4];     // [k][pixel]: a thread's pixels are 128-bit loads
  __shared__ __align__(16) float Bs[KCF][TB + 4];     // [k][output]
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const long long m0 = (long long)blockIdx.x * TB;
  const int n0 = blockIdx.y * TB;
  float acc[TM][TM];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TM; ++j) acc[i][j] = 0.f;
  const int lrow = tid >> 2, lq = tid & 3;             // loader: rows lrow + 64 r, channels 4 lq .. 4 lq + 3 of the chunk
  int lb[NL], lh[NL], lw[NL];
  bool lvalid[NL];
#pragma unroll
  for (int r = 0; r < NL; ++r) {
    const long long lpix = m0 + lrow + 64 * r;
    lvalid[r] = lpix < p.rows;
    lb[r] = lh[r] = lw[r] = 0;
    if (lvalid[r]) {
      lb[r] = (int)(lpix / ((long long)p.H * p.W));
      const int rem = (int)(lpix - (long long)lb[r] * p.H * p.W);
      lh[r] = rem / p.W;
      lw[r] = rem - lh[r] * p.W;
    }
  }
  const bool vec = (p.C & 3) == 0;
  // (tap, 16-channel chunk) pairs are walked as one sequence; the loads of chunk i+1 are issued before the FMAs of
  // chunk i (register double buffering), which is what matters for the long-K, few-CTA shapes (the 1024-wide
  // time_embed linears were pure load latency: 64 chunks x ~0.7 us)
  const int nkc = (p.C + KCF - 1) / KCF, nchunks = p.taps * nkc;
  float av[NL][4], bv[NL][4];
  auto load_chunk = [&](int idx) {
    const int tap = idx / nkc, kc = (idx - tap * nkc) * KCF;
    const int dh = p.taps == 9 ? tap / 3 - 1 : 0, dw = p.taps == 9 ? tap % 3 - 1 : 0;
    const int k = kc + lq * 4;
#pragma unroll
    for (int r = 0; r < NL; ++r) {
      const int hh = lh[r] + dh, ww = lw[r] + dw;
      const bool inb = lvalid[r] && hh >= 0 && hh < p.H && ww >= 0 && ww < p.W;
      const float* arow = p.x + (((long long)lb[r] * p.H + hh) * p.W + ww) * p.C;
      const int orow = n0 + lrow + 64 * r;
      const float* brow = p.w + ((long long)orow * p.taps + tap) * p.C;
      const bool bvalid = orow < p.O;
#pragma unroll
      for (int e = 0; e < 4; ++e) { av[r][e] = 0.f; bv[r][e] = 0.f; }
      if (vec) {
        if (inb && k < p.C) { float4 t = *reinterpret_cast<const float4*>(arow + k); av[r][0] = t.x; av[r][1] = t.y; av[r][2] = t.z; av[r][3] = t.w; }
        if (bvalid && k < p.C) { float4 t = *reinterpret_cast<const float4*>(brow + k); bv[r][0] = t.x; bv[r][1] = t.y; bv[r][2] = t.z; bv[r][3] = t.w; }
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (inb && k + e < p.C) av[r][e] = arow[k + e];
          if (bvalid && k + e < p.C) bv[r][e] = brow[k + e];
        }
      }
    }
  };
  load_chunk(0);
  for (int idx = 0; idx < nchunks; ++idx) {
    __syncthreads();
#pragma unroll
    for (int r = 0; r < NL; ++r)
#pragma unroll
      for (int e = 0; e < 4; ++e) { As[lq * 4 + e][lrow + 64 * r] = av[r][e]; Bs[lq * 4 + e][lrow + 64 * r] = bv[r][e]; }
    __syncthreads();
    if (idx + 1 < nchunks) load_chunk(idx + 1);
#pragma unroll
    for (int kk = 0; kk < KCF; ++kk) {
      float a[TM], b[TM];
#pragma unroll
      for (int q = 0; q < NL; ++q) {
        const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * TM + 4 * q]);
        const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * TM + 4 * q]);
        a[4 * q] = a4.x; a[4 * q + 1] = a4.y; a[4 * q + 2] = a4.z; a[4 * q + 3] = a4.w;
        b[4 * q] = b4.x; b[4 * q + 1] = b4.y; b[4 * q + 2] = b4.z; b[4 * q + 3] = b4.w;
      }
      // packed fp32 FMAs (fma.rn.f32x2, sm_100): two IEEE-rounded fmas per instruction -- same results as
      // fmaf, twice the FMA rate of the 3-register scalar form.  Per output the order over k is unchanged.
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TM; j += 2) fma2(acc[i][j], acc[i][j + 1], a[i], b[j], b[j + 1]);
    }
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const long long pix = m0 + ty * TM + i;
    if (pix >= p.rows) continue;
    const int b = (int)(pix / ((long long)p.H * p.W));
#pragma unroll
    for (int j = 0; j < TM; ++j) {
      const int o = n0 + tx * TM + j;
      if (o >= p.O) continue;
      float v = acc[i][j] + (p.bias ? p.bias[o] : 0.f);
      p.out[pix * p.O + o] = conv_epilogue_add(p.residual, p.temb, v, pix, b, o, p.O);
    }
  }
}

int launch_qconv_i8_simt(const ConvI8Params& p, cudaStream_t st) {
  dim3 grid(cdiv(p.rows, BM), cdiv(p.O, BN));
  launch_pdl(qconv_i8_simt_kernel, dim3(grid), dim3(256), 0, st, p);
  ATTNDM_CUDA_LAUNCH_CHECK("qconv_i8_simt");
  return ATTNDM_OK;
}

}  // namespace attndm

using namespace attndm;

extern "C" int attndm_conv_f32(const float* x, int B, int H, int W, int C, const float* w_eff, int O, int taps,
                               const float* bias, const float* residual, const float* temb, float* out,
                               void* stream) {
  ATTNDM_CHECK_ARG(x && w_eff && out && B > 0 && H > 0 && W > 0 && C > 0 && O > 0, "conv_f32: bad args");
  ATTNDM_CHECK_ARG(taps == 1 || taps == 9, "conv_f32: only 1x1 and 3x3/s1/p1 are on the hot path");
  ConvF32Params p;
  p.x = x; p.B = B; p.H = H; p.W = W; p.C = C; p.w = w_eff; p.O = O; p.taps = taps;
  p.bias = bias; p.residual = residual; p.temb = temb; p.out = out;
  p.rows = (long long)B * H * W;
  // large GEMMs (>= one wave of 128 x 128 tiles) take the 8 x 8 register tile
  if ((long long)cdiv(p.rows, 128) * cdiv(O, 128) >= kNumSMs) {
    dim3 grid(cdiv(p.rows, 128), cdiv(O, 128));
    launch_pdl(conv_f32_simt_kernel<8>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, p);
  } else {
    dim3 grid(cdiv(p.rows, 64), cdiv(O, 64));
    launch_pdl(conv_f32_simt_kernel<4>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, p);
  }
  ATTNDM_CUDA_LAUNCH_CHECK("conv_f32");
  return ATTNDM_OK;
}

/* fp32 1x1 conv (plain GEMM + bias) on the tensor cores, fp32-level accuracy by operand splitting (conv_tc.cu) */
extern "C" int attndm_conv_f32_tc_fits(long long rows, int C, int O) { return conv_f32_tc_fits(rows, C, O); }
extern "C" int attndm_split_tf32(const float* x, long long n, float* big, float* small, void* stream) {
  return launch_split_tf32(x, n, big, small, (cudaStream_t)stream);
}
extern "C" int attndm_gemm_tf32x3(const float* a_big, const float* a_small, long long rows, int C, const float* w_big,
                                  const float* w_small, int O, const float* bias, float* out, void* stream) {
  ATTNDM_CHECK_ARG(a_big && a_small && w_big && w_small && out && rows > 0 && C > 0 && O > 0, "gemm_tf32x3: bad args");
  return launch_gemm_tf32x3(a_big, a_small, rows, C, w_big, w_small, O, bias, out, (cudaStream_t)stream);
}
