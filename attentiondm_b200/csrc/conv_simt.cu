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

__global__ void __launch_bounds__(256) conv_f32_simt_kernel(ConvF32Params p) {
  pdl_enter();
  __shared__ float As[BM][KCF + 1];
  __shared__ float Bs[BN][KCF + 1];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int lrow = tid >> 2, lq = tid & 3;
  // decode the loader's pixel once
  const long long lpix = m0 + lrow;
  int lb = 0, lh = 0, lw = 0;
  const bool lvalid = lpix < p.rows;
  if (lvalid) {
    lb = (int)(lpix / ((long long)p.H * p.W));
    int rem = (int)(lpix - (long long)lb * p.H * p.W);
    lh = rem / p.W;
    lw = rem - lh * p.W;
  }
  const bool vec = (p.C & 3) == 0;
  // (tap, 16-channel chunk) pairs are walked as one sequence; the loads of chunk i+1 are issued before the FMAs of
  // chunk i (register double buffering), which is what matters for the long-K, few-CTA shapes (the 1024-wide
  // time_embed linears were pure load latency: 64 chunks x ~0.7 us)
  const int nkc = (p.C + KCF - 1) / KCF, nchunks = p.taps * nkc;
  float av[4], bv[4];
  auto load_chunk = [&](int idx) {
    const int tap = idx / nkc, kc = (idx - tap * nkc) * KCF;
    const int dh = p.taps == 9 ? tap / 3 - 1 : 0, dw = p.taps == 9 ? tap % 3 - 1 : 0;
    const int hh = lh + dh, ww = lw + dw;
    const bool inb = lvalid && hh >= 0 && hh < p.H && ww >= 0 && ww < p.W;
    const float* arow = p.x + (((long long)lb * p.H + hh) * p.W + ww) * p.C;
    const float* brow = p.w + ((long long)(n0 + lrow) * p.taps + tap) * p.C;
    const bool bvalid = n0 + lrow < p.O;
    const int k = kc + lq * 4;
#pragma unroll
    for (int e = 0; e < 4; ++e) { av[e] = 0.f; bv[e] = 0.f; }
    if (vec) {
      if (inb && k < p.C) { float4 t = *reinterpret_cast<const float4*>(arow + k); av[0] = t.x; av[1] = t.y; av[2] = t.z; av[3] = t.w; }
      if (bvalid && k < p.C) { float4 t = *reinterpret_cast<const float4*>(brow + k); bv[0] = t.x; bv[1] = t.y; bv[2] = t.z; bv[3] = t.w; }
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (inb && k + e < p.C) av[e] = arow[k + e];
        if (bvalid && k + e < p.C) bv[e] = brow[k + e];
      }
    }
  };
  load_chunk(0);
  for (int idx = 0; idx < nchunks; ++idx) {
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 4; ++e) { As[lrow][lq * 4 + e] = av[e]; Bs[lrow][lq * 4 + e] = bv[e]; }
    __syncthreads();
    if (idx + 1 < nchunks) load_chunk(idx + 1);
#pragma unroll
    for (int kk = 0; kk < KCF; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[ty * 4 + i][kk];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[tx * 4 + j][kk];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long pix = m0 + ty * 4 + i;
    if (pix >= p.rows) continue;
    const int b = (int)(pix / ((long long)p.H * p.W));
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int o = n0 + tx * 4 + j;
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
  dim3 grid(cdiv(p.rows, BM), cdiv(O, BN));
  launch_pdl(conv_f32_simt_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, p);
  ATTNDM_CUDA_LAUNCH_CHECK("conv_f32");
  return ATTNDM_OK;
}
