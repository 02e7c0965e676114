// Attention core, UNet glue and the sampler update.
#include "common.cuh"

namespace attndm {

// ---------------------------------------------------------------------------
// attention: one warp per (sample, head, query row)
// models/self_attention.py:141-144; utils/attention_quant_utils.py:65-107
// ---------------------------------------------------------------------------
struct AttnParams {
  const float* q;
  const float* k;
  const float* v;
  float* out;
  int B, N, d, dv, heads;
  float scale, softmax_scale;
  attndm_attn_quant qk_q, p_q;
};

// R query rows per warp (consecutive rows of one sample and head): every K row and every V element the warp loads
// serves R dot products, and the probabilities are read from shared memory four keys at a time.  With one row per warp
// the kernel streamed all of V (N x dv floats, 1 MB at N = 1024, dv = 256) from L1/L2 for EVERY query.  The arithmetic of
// a row -- the order of every sum -- is the one-row kernel's, so results do not depend on R.
template <int R>
__global__ void __launch_bounds__(128) attention_kernel(AttnParams p) {
  pdl_enter();
  extern __shared__ __align__(16) float sm[];         // [4 warps][R][N]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float* sc = sm + (long long)w * R * p.N;
  const long long gw = (long long)blockIdx.x * 4 + w;            // group of R rows
  const int ngrp = p.N / R;                                      // (the launcher guarantees N % R == 0)
  const long long total = (long long)p.B * p.heads * ngrp;
  if (gw >= total) return;
  const int i0 = (int)(gw % ngrp) * R;
  const int head = (int)((gw / ngrp) % p.heads);
  const int b = (int)(gw / ((long long)ngrp * p.heads));
  const int dq = p.d / p.heads, dvh = p.dv / p.heads;
  const float* qrow = p.q + ((long long)b * p.N + i0) * p.d + head * dq;     // q: head-major channels; row r at + r * d
  const float* kb = p.k + (long long)b * p.N * p.d;                          // k: channel = e*heads + head
  float mx[R];
#pragma unroll
  for (int r = 0; r < R; ++r) mx[r] = -INFINITY;
  for (int j = lane; j < p.N; j += 32) {
    const float* krow = kb + (long long)j * p.d;
    float acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.f;
    for (int e = 0; e < dq; ++e) {
      const float kv = krow[e * p.heads + head];
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] = fmaf(qrow[r * p.d + e], kv, acc[r]);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float s = __fmul_rn(acc[r], p.scale);
      if (p.qk_q.bits > 0) s = attn_fake_quant(s, p.qk_q.scale, p.qk_q.zero_point, (float)((1 << p.qk_q.bits) - 1));
      s = __fmul_rn(s, p.softmax_scale);
      sc[r * p.N + j] = s;
      mx[r] = fmaxf(mx[r], s);
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const float m = warp_max(mx[r]);
    float sum = 0.f;
    for (int j = lane; j < p.N; j += 32) {
      float e = expf(sc[r * p.N + j] - m);
      sc[r * p.N + j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    for (int j = lane; j < p.N; j += 32) {
      float pr = __fdiv_rn(sc[r * p.N + j], sum);
      if (p.p_q.bits > 0) pr = attn_fake_quant(pr, p.p_q.scale, p.p_q.zero_point, (float)((1 << p.p_q.bits) - 1));
      sc[r * p.N + j] = pr;
    }
  }
  __syncwarp();
  const float* vb = p.v + (long long)b * p.N * p.dv + head * dvh;
  float* orow = p.out + ((long long)b * p.N + i0) * p.dv + head * dvh;
  const int n4 = (R > 1 && (p.N & 3) == 0) ? p.N : 0;            // keys taken four at a time (128-bit reads of the probabilities)
  for (int c = lane; c < dvh; c += 32) {
    float acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.f;
    int j = 0;
    // sixteen keys per step, all sixteen V loads issued before the first FMA, where the score rows leave room for enough
    // warps per SM (N <= 512: 236 -> 130 us at N = 256, heads 8, batch 32; at N = 1024 the same loop measured 3.25 ms
    // against 3.0 ms for four loads per step)
    const int n16 = p.N <= 512 ? n4 : 0;
    for (; j + 16 <= n16; j += 16) {
      float vv[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) vv[u] = __ldg(vb + (long long)(j + u) * p.dv + c);
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4)
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float4 pr = *reinterpret_cast<const float4*>(sc + r * p.N + j + 4 * q4);
          acc[r] = fmaf(pr.x, vv[4 * q4], acc[r]);
          acc[r] = fmaf(pr.y, vv[4 * q4 + 1], acc[r]);
          acc[r] = fmaf(pr.z, vv[4 * q4 + 2], acc[r]);
          acc[r] = fmaf(pr.w, vv[4 * q4 + 3], acc[r]);
        }
    }
    for (; j < n4; j += 4) {
      const float v0 = vb[(long long)j * p.dv + c], v1 = vb[(long long)(j + 1) * p.dv + c];
      const float v2 = vb[(long long)(j + 2) * p.dv + c], v3 = vb[(long long)(j + 3) * p.dv + c];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float4 pr = *reinterpret_cast<const float4*>(sc + r * p.N + j);
        acc[r] = fmaf(pr.x, v0, acc[r]);
        acc[r] = fmaf(pr.y, v1, acc[r]);
        acc[r] = fmaf(pr.z, v2, acc[r]);
        acc[r] = fmaf(pr.w, v3, acc[r]);
      }
    }
    for (; j < p.N; ++j) {
      const float vv = vb[(long long)j * p.dv + c];
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] = fmaf(sc[r * p.N + j], vv, acc[r]);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) orow[(long long)r * p.dv + c] = acc[r];
  }
}

__global__ void scale_add_kernel(const float* __restrict__ a, const float* __restrict__ x, const float* __restrict__ gamma,
                                 float* __restrict__ out, long long n) {
  pdl_enter();
  const float g = *gamma;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __fadd_rn(__fmul_rn(g, a[i]), x[i]);
}

// ---------------------------------------------------------------------------
// UNet glue (NHWC)
// ---------------------------------------------------------------------------
__global__ void maxpool2_kernel(const float* __restrict__ x, int B, int H, int W, int C, float* __restrict__ y) {
  pdl_enter();
  const int Ho = H / 2, Wo = W / 2;
  const long long n = (long long)B * Ho * Wo * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    long long t = i / C;
    int wo = (int)(t % Wo);
    t /= Wo;
    int ho = (int)(t % Ho);
    int b = (int)(t / Ho);
    const float* p00 = x + (((long long)b * H + 2 * ho) * W + 2 * wo) * C + c;
    float m = fmaxf(fmaxf(p00[0], p00[C]), fmaxf(p00[(long long)W * C], p00[(long long)W * C + C]));
    y[i] = m;
  }
}

template <int V>   // V = 4: float4 over channels (Cx % 4 == 0 && Cs % 4 == 0), V = 1: scalar
__global__ void upsample_concat_kernel(const float* __restrict__ x, int B, int H, int W, int Cx,
                                       const float* __restrict__ skip, int Hs, int Ws, int Cs,
                                       float* __restrict__ out) {
  pdl_enter();
  const int Ct = Cx + Cs, Cv = Ct / V;
  const long long n = (long long)B * Hs * Ws * Cv;
  // nearest x2 (src = dst/2) followed, when sizes differ, by nearest resize 2H x 2W -> Hs x Ws
  // (src = min(floor(dst * in/out), in-1) in fp32, as ATen's nearest kernel computes it)
  const bool same = (2 * H == Hs) && (2 * W == Ws);
  const float sh = (float)(2 * H) / (float)Hs, sw = (float)(2 * W) / (float)Ws;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cv) * V;
    long long t = i / Cv;
    const int ws = (int)(t % Ws);
    t /= Ws;
    const int hs = (int)(t % Hs);
    const int b = (int)(t / Hs);
    const float* src;
    if (c < Cx) {
      int uh = hs, uw = ws;
      if (!same) {
        uh = min((int)floorf(__fmul_rn((float)hs, sh)), 2 * H - 1);
        uw = min((int)floorf(__fmul_rn((float)ws, sw)), 2 * W - 1);
      }
      src = x + (((long long)b * H + (uh >> 1)) * W + (uw >> 1)) * Cx + c;
    } else {
      src = skip + (((long long)b * Hs + hs) * Ws + ws) * Cs + (c - Cx);
    }
    if (V == 4) *reinterpret_cast<float4*>(out + i * 4) = *reinterpret_cast<const float4*>(src);
    else out[i] = *src;
  }
}

__global__ void timestep_embedding_kernel(const float* __restrict__ t, int B, int dim, float* __restrict__ emb) {
  pdl_enter();
  // The reference evaluates exp / sin / cos in fp32 on fp32 arguments.  |t * f| reaches ~1e3, so one
  // ulp of f moves sin/cos by ~1e-4 -- enough to flip 8-bit codes in the time MLP.  We therefore
  // evaluate each transcendental in double on the SAME fp32 argument and round once: that is the
  // correctly rounded fp32 result, which is what an accurate fp32 libm (the CPU reference) returns.
  const int half = dim / 2;
  const float coef = -(float)(log(10000.0) / (double)(half - 1));
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B * half; i += gridDim.x * blockDim.x) {
    int b = i / half, e = i - b * half;
    float f = (float)exp((double)__fmul_rn((float)e, coef));
    float a = __fmul_rn(t[b], f);
    emb[(long long)b * dim + e] = (float)sin((double)a);
    emb[(long long)b * dim + half + e] = (float)cos((double)a);
    if ((dim & 1) && e == 0) emb[(long long)b * dim + dim - 1] = 0.f;
  }
}

// ---------------------------------------------------------------------------
// sampler
// ---------------------------------------------------------------------------
__device__ __forceinline__ void ddim_update_1(float x, float e, float nz, bool has_noise, float s1mat, float sat, float satn,
                                              float c1, float c2, float& r, float& x0) {
  // x0_t = (xt - et * (1 - at).sqrt()) / at.sqrt()
  x0 = __fdiv_rn(__fsub_rn(x, __fmul_rn(e, s1mat)), sat);
  // xt_next = at_next.sqrt() * x0_t + c1 * randn + c2 * et
  r = __fmul_rn(satn, x0);
  if (has_noise) r = __fadd_rn(r, __fmul_rn(c1, nz));
  r = __fadd_rn(r, __fmul_rn(c2, e));
}

// VEC = 4: 128-bit accesses (n % 4 == 0, 16-byte aligned operands); VEC = 1: scalar.  12 B/element of traffic
// (x_t, eps in; x_next out) + the optional x0 / history outputs.
template <int VEC>
__global__ void ddim_step_kernel(const float* __restrict__ xt, const float* __restrict__ eps,
                                 const float* __restrict__ coef, const float* __restrict__ noise,
                                 float* __restrict__ x_next, float* __restrict__ x0_out, long long n,
                                 float* __restrict__ hist_x, float* __restrict__ hist_x0,
                                 const int* __restrict__ step_after, int T) {
  pdl_enter();
  const float s1mat = coef[0], sat = coef[1], satn = coef[2], c1 = coef[3], c2 = coef[4];
  if (hist_x != nullptr) {                       // history slot of this step (the counter was advanced already)
    const int k = (*step_after + T - 1) % T;
    hist_x += (long long)k * n;
    hist_x0 += (long long)k * n;
  }
  const bool has_noise = noise != nullptr;
  if (VEC == 4) {
    const long long n4 = n >> 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
      const float4 x = reinterpret_cast<const float4*>(xt)[i], e = reinterpret_cast<const float4*>(eps)[i];
      const float4 z = has_noise ? reinterpret_cast<const float4*>(noise)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 r, x0;
      ddim_update_1(x.x, e.x, z.x, has_noise, s1mat, sat, satn, c1, c2, r.x, x0.x);
      ddim_update_1(x.y, e.y, z.y, has_noise, s1mat, sat, satn, c1, c2, r.y, x0.y);
      ddim_update_1(x.z, e.z, z.z, has_noise, s1mat, sat, satn, c1, c2, r.z, x0.z);
      ddim_update_1(x.w, e.w, z.w, has_noise, s1mat, sat, satn, c1, c2, r.w, x0.w);
      reinterpret_cast<float4*>(x_next)[i] = r;
      if (x0_out) reinterpret_cast<float4*>(x0_out)[i] = x0;
      if (hist_x != nullptr) {
        reinterpret_cast<float4*>(hist_x)[i] = r;
        reinterpret_cast<float4*>(hist_x0)[i] = x0;
      }
    }
    return;
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float r, x0;
    ddim_update_1(xt[i], eps[i], has_noise ? noise[i] : 0.f, has_noise, s1mat, sat, satn, c1, c2, r, x0);
    x_next[i] = r;
    if (x0_out) x0_out[i] = x0;
    if (hist_x != nullptr) {
      hist_x[i] = r;
      hist_x0[i] = x0;
    }
  }
}

__global__ void stage_copy_kernel(const float* __restrict__ table, long long n, const int* __restrict__ step,
                                  float* __restrict__ dst) {
  pdl_enter();
  const float* src = table + (long long)(*step) * n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = src[i];
}
// the current step's rows of the staged table fanned out to [B][width] tensors (attndm_bcast_rows)
__global__ void bcast_rows_kernel(const float* __restrict__ cur, const int32_t* __restrict__ desc, int B, float* __restrict__ dst) {
  pdl_enter();
  const int off_dst = desc[3 * blockIdx.y], off_src = desc[3 * blockIdx.y + 1], width = desc[3 * blockIdx.y + 2];
  const int n = B * width;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int c = i % width;
    dst[off_dst + i] = cur[off_src + c];
  }
}
__global__ void stage_advance_kernel(int* step, int T) {
  pdl_enter();
  int s = *step + 1;
  *step = s >= T ? 0 : s;
}


// ---- the callers either side of the sampler -------------------------------------------------
// ddpm_steps, functions/denoising.py:119-151.  coef (device float[6]) = {(1/at).sqrt(), (1/at - 1).sqrt(),
// atm1.sqrt() * beta_t, (1 - beta_t).sqrt() * (1 - atm1), 1 - at, mask * exp(0.5 * log(beta_t))} in the reference's fp32
// op order (host side); every elementwise operation below is rounded separately like the eager ops it replaces.
__global__ void ddpm_step_kernel(const float* __restrict__ xt, const float* __restrict__ eps, const float* __restrict__ coef,
                                 const float* __restrict__ noise, float* __restrict__ x_next, float* __restrict__ x0_out,
                                 long long n) {
  pdl_enter();
  const float ca = coef[0], cb = coef[1], cc = coef[2], cd = coef[3], ce = coef[4], cf = coef[5];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float x = xt[i], e = eps[i];
    float x0 = __fsub_rn(__fmul_rn(ca, x), __fmul_rn(cb, e));            // x0_from_e (:137)
    x0 = fminf(fmaxf(x0, -1.f), 1.f);                                    // torch.clamp(-1, 1) (:138)
    const float mean = __fdiv_rn(__fadd_rn(__fmul_rn(cc, x0), __fmul_rn(cd, x)), ce);   // (:140-142)
    x_next[i] = __fadd_rn(mean, __fmul_rn(cf, noise[i]));                // mean + mask * exp(0.5 logvar) * noise (:149)
    if (x0_out) x0_out[i] = x0;
  }
}

// noise_estimation_loss, functions/denoising.py:52-54: x = x0 * a.sqrt() + e * (1.0 - a).sqrt(), coef = {a.sqrt(), (1 - a).sqrt()}
__global__ void noise_mix_kernel(const float* __restrict__ x0, const float* __restrict__ e, const float* __restrict__ coef,
                                 float* __restrict__ x, long long n) {
  pdl_enter();
  const float sa = coef[0], s1a = coef[1];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    x[i] = __fadd_rn(__fmul_rn(x0[i], sa), __fmul_rn(e[i], s1a));
}

// (e - output).square().sum(dim=(1,2,3)) per sample (:58-60), accumulated in double: out[b]
__global__ void sq_err_kernel(const float* __restrict__ a, const float* __restrict__ b, long long per, double* __restrict__ out) {
  pdl_enter();
  __shared__ double part[8];
  const float* pa = a + (long long)blockIdx.x * per;
  const float* pb = b + (long long)blockIdx.x * per;
  double acc = 0.0;
  for (long long i = threadIdx.x; i < per; i += blockDim.x) {
    const float d = __fsub_rn(pa[i], pb[i]);
    acc += (double)__fmul_rn(d, d);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += part[w];
    out[blockIdx.x] = t;
  }
}

// The entropy regulariser of generalized_steps_loss (functions/denoising.py:83-100) for ONE layer and timestep:
//   s = softmax(alpha_t, dim over the G groups), H = cal_entropy(s) = -(1/G) sum_g sum_c s log s, term = H / (G*C);
// writes grad[g][c] = weight * d term / d alpha_t[g][c]
//        = -(weight / (G*G*C)) * s[g][c] * ((log s[g][c] + 1) - sum_g' s[g'][c] (log s[g'][c] + 1))
// and adds weight * term to *value (double).  One thread per channel; G <= 32.
__global__ void alpha_entropy_grad_kernel(const float* __restrict__ alpha_t, int G, int C, float weight, float* __restrict__ grad,
                                          double* __restrict__ value) {
  pdl_enter();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  double h = 0.0;
  if (c < C) {
    double mx = -1e300;
    for (int g = 0; g < G; ++g) mx = fmax(mx, (double)alpha_t[(long long)g * C + c]);
    double den = 0.0;
    for (int g = 0; g < G; ++g) den += exp((double)alpha_t[(long long)g * C + c] - mx);
    const double lden = log(den);
    double m1 = 0.0;                                   // sum_g s (log s + 1)
    for (int g = 0; g < G; ++g) {
      const double ls = (double)alpha_t[(long long)g * C + c] - mx - lden, sg = exp(ls);
      m1 += sg * (ls + 1.0);
      h -= sg * ls;
    }
    const double k = -(double)weight / ((double)G * G * C);
    for (int g = 0; g < G; ++g) {
      const double ls = (double)alpha_t[(long long)g * C + c] - mx - lden, sg = exp(ls);
      grad[(long long)g * C + c] = (float)(k * sg * ((ls + 1.0) - m1));
    }
  }
  if (value != nullptr) {
    for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
    if ((threadIdx.x & 31) == 0 && h != 0.0) atomicAdd(value, (double)weight * h / ((double)G * G * C));
  }
}

static inline int ew_blocks(long long n) {
  long long b = (n + 255) / 256;
  long long cap = (long long)kNumSMs * 32;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace attndm

using namespace attndm;

static void launch_ddim(const float* xt, const float* eps, const float* coef, const float* noise, float* x_next, float* x0_out,
                        long long n, float* hist_x, float* hist_x0, const int* step_after, int T, cudaStream_t st) {
  auto al = [](const void* p) { return p == nullptr || ((uintptr_t)p & 15) == 0; };
  const bool v4 = (n & 3) == 0 && al(xt) && al(eps) && al(noise) && al(x_next) && al(x0_out) && al(hist_x) && al(hist_x0);
  if (v4)
    launch_pdl(ddim_step_kernel<4>, dim3(ew_blocks(n >> 2)), dim3(256), 0, st, xt, eps, coef, noise, x_next, x0_out, n, hist_x,
               hist_x0, step_after, T);
  else
    launch_pdl(ddim_step_kernel<1>, dim3(ew_blocks(n)), dim3(256), 0, st, xt, eps, coef, noise, x_next, x0_out, n, hist_x,
               hist_x0, step_after, T);
}

extern "C" {

int attndm_attention(const float* q, const float* k, const float* v, float* out, int B, int N, int d, int dv,
                     float scale, int heads, float softmax_scale, attndm_attn_quant qk_q, attndm_attn_quant p_q,
                     void* stream) {
  ATTNDM_CHECK_ARG(q && k && v && out && B > 0 && N > 0 && d > 0 && dv > 0, "attention: bad args");
  ATTNDM_CHECK_ARG(heads >= 1 && d % heads == 0 && dv % heads == 0, "attention: heads must divide d and dv");
  ATTNDM_CHECK_ARG(N <= 8192, "attention: N <= 8192 (per-warp score row lives in shared memory)");
  ATTNDM_CHECK_ARG(qk_q.bits >= 0 && qk_q.bits <= 8 && p_q.bits >= 0 && p_q.bits <= 8, "attention: bad quant bits");
  AttnParams p;
  p.q = q; p.k = k; p.v = v; p.out = out; p.B = B; p.N = N; p.d = d; p.dv = dv; p.heads = heads;
  p.scale = scale; p.softmax_scale = softmax_scale; p.qk_q = qk_q; p.p_q = p_q;
  // rows per warp: four where the score rows of a CTA (4 warps x R x N floats) fit in 96 KB and there are enough row
  // groups to fill the chip, else two, else one
  int R = 1;
  for (int r = 4; r >= 2; r >>= 1)
    if (N % r == 0 && 4 * (size_t)r * N * sizeof(float) <= 96 * 1024 && (long long)B * heads * (N / r) >= 4LL * 2 * kNumSMs) { R = r; break; }
  const long long total = (long long)B * heads * (N / R);
  const size_t smem = 4 * (size_t)R * N * sizeof(float);
  auto kern = R == 4 ? attention_kernel<4> : (R == 2 ? attention_kernel<2> : attention_kernel<1>);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("attention: smem attr: %s", cudaGetErrorString(e)); return ATTNDM_ERR_CUDA; }
  }
  launch_pdl(kern, dim3((unsigned)cdiv(total, 4)), dim3(128), smem, (cudaStream_t)stream, p);
  ATTNDM_CUDA_LAUNCH_CHECK("attention");
  return ATTNDM_OK;
}

int attndm_scale_add(const float* a, const float* x, const float* gamma, float* out, long long n, void* stream) {
  ATTNDM_CHECK_ARG(a && x && gamma && out && n > 0, "scale_add: bad args");
  launch_pdl(scale_add_kernel, dim3(ew_blocks(n)), dim3(256), 0, (cudaStream_t)stream, a, x, gamma, out, n);
  ATTNDM_CUDA_LAUNCH_CHECK("scale_add");
  return ATTNDM_OK;
}

int attndm_maxpool2(const float* x, int B, int H, int W, int C, float* y, void* stream) {
  ATTNDM_CHECK_ARG(x && y && B > 0 && H >= 2 && W >= 2 && C > 0, "maxpool2: bad args");
  long long n = (long long)B * (H / 2) * (W / 2) * C;
  launch_pdl(maxpool2_kernel, dim3(ew_blocks(n)), dim3(256), 0, (cudaStream_t)stream, x, B, H, W, C, y);
  ATTNDM_CUDA_LAUNCH_CHECK("maxpool2");
  return ATTNDM_OK;
}

int attndm_upsample_concat(const float* x, int B, int H, int W, int Cx, const float* skip, int Hs, int Ws, int Cs,
                           float* out, void* stream) {
  ATTNDM_CHECK_ARG(x && out && B > 0 && H > 0 && W > 0 && Cx > 0 && Hs > 0 && Ws > 0 && Cs >= 0, "upsample_concat: bad args");
  ATTNDM_CHECK_ARG(Cs == 0 || skip, "upsample_concat: skip is NULL");
  long long n = (long long)B * Hs * Ws * (Cx + Cs);
  if ((Cx & 3) == 0 && (Cs & 3) == 0)
    launch_pdl(upsample_concat_kernel<4>, dim3(ew_blocks(n / 4)), dim3(256), 0, (cudaStream_t)stream, x, B, H, W, Cx, skip, Hs, Ws, Cs, out);
  else
    launch_pdl(upsample_concat_kernel<1>, dim3(ew_blocks(n)), dim3(256), 0, (cudaStream_t)stream, x, B, H, W, Cx, skip, Hs, Ws, Cs, out);
  ATTNDM_CUDA_LAUNCH_CHECK("upsample_concat");
  return ATTNDM_OK;
}

int attndm_timestep_embedding(const float* t, int B, int dim, float* emb, void* stream) {
  ATTNDM_CHECK_ARG(t && emb && B > 0 && dim >= 4, "timestep_embedding: bad args");
  launch_pdl(timestep_embedding_kernel, dim3(cdiv((long long)B * (dim / 2), 256)), dim3(256), 0, (cudaStream_t)stream, t, B, dim, emb);
  ATTNDM_CUDA_LAUNCH_CHECK("timestep_embedding");
  return ATTNDM_OK;
}

int attndm_ddim_step(const float* xt, const float* eps, const float* coef, const float* noise, float* x_next,
                     float* x0_out, long long n, void* stream) {
  ATTNDM_CHECK_ARG(xt && eps && coef && x_next && n > 0, "ddim_step: bad args");
  launch_ddim(xt, eps, coef, noise, x_next, x0_out, n, nullptr, nullptr, nullptr, 1, (cudaStream_t)stream);
  ATTNDM_CUDA_LAUNCH_CHECK("ddim_step");
  return ATTNDM_OK;
}

int attndm_ddim_step_hist(const float* xt, const float* eps, const float* coef, const float* noise, float* x_next,
                          float* x0_out, long long n, float* hist_x, float* hist_x0, const int* step_after, int T,
                          void* stream) {
  ATTNDM_CHECK_ARG(xt && eps && coef && x_next && n > 0, "ddim_step_hist: bad args");
  ATTNDM_CHECK_ARG(hist_x && hist_x0 && step_after && T > 0, "ddim_step_hist: history rings, step counter and T are required");
  launch_ddim(xt, eps, coef, noise, x_next, x0_out, n, hist_x, hist_x0, step_after, T, (cudaStream_t)stream);
  ATTNDM_CUDA_LAUNCH_CHECK("ddim_step_hist");
  return ATTNDM_OK;
}

int attndm_ddpm_step(const float* xt, const float* eps, const float* coef, const float* noise, float* x_next, float* x0_out,
                     long long n, void* stream) {
  ATTNDM_CHECK_ARG(xt && eps && coef && noise && x_next && n > 0, "ddpm_step: bad args");
  launch_pdl(ddpm_step_kernel, dim3(ew_blocks(n)), dim3(256), 0, (cudaStream_t)stream, xt, eps, coef, noise, x_next, x0_out, n);
  ATTNDM_CUDA_LAUNCH_CHECK("ddpm_step");
  return ATTNDM_OK;
}

int attndm_noise_mix(const float* x0, const float* e, const float* coef, float* x, long long n, void* stream) {
  ATTNDM_CHECK_ARG(x0 && e && coef && x && n > 0, "noise_mix: bad args");
  launch_pdl(noise_mix_kernel, dim3(ew_blocks(n)), dim3(256), 0, (cudaStream_t)stream, x0, e, coef, x, n);
  ATTNDM_CUDA_LAUNCH_CHECK("noise_mix");
  return ATTNDM_OK;
}

int attndm_sq_err(const float* a, const float* b, int B, long long per, double* out, void* stream) {
  ATTNDM_CHECK_ARG(a && b && out && B > 0 && per > 0, "sq_err: bad args");
  launch_pdl(sq_err_kernel, dim3(B), dim3(256), 0, (cudaStream_t)stream, a, b, per, out);
  ATTNDM_CUDA_LAUNCH_CHECK("sq_err");
  return ATTNDM_OK;
}

int attndm_alpha_entropy_grad(const float* alpha_t, int G, int C, float weight, float* grad, double* value, void* stream) {
  ATTNDM_CHECK_ARG(alpha_t && grad && G > 0 && G <= 32 && C > 0, "alpha_entropy_grad: bad args (1 <= G <= 32)");
  launch_pdl(alpha_entropy_grad_kernel, dim3((C + 127) / 128), dim3(128), 0, (cudaStream_t)stream, alpha_t, G, C, weight, grad, value);
  ATTNDM_CUDA_LAUNCH_CHECK("alpha_entropy_grad");
  return ATTNDM_OK;
}

int attndm_bcast_rows(const float* cur, const int32_t* desc, int n, int B, int max_width, float* dst, void* stream) {
  ATTNDM_CHECK_ARG(cur && desc && dst && n > 0 && B > 0 && max_width > 0 && (long long)B * max_width < (1LL << 31), "bcast_rows: bad args");
  launch_pdl(bcast_rows_kernel, dim3((unsigned)cdiv((long long)B * max_width, 256), (unsigned)n), dim3(256), 0, (cudaStream_t)stream,
             cur, desc, B, dst);
  ATTNDM_CUDA_LAUNCH_CHECK("bcast_rows");
  return ATTNDM_OK;
}

int attndm_stage_tables(const float* table, long long n, int T, int* step, int advance, float* dst, void* stream) {
  ATTNDM_CHECK_ARG(table && step && dst && n > 0 && T > 0, "stage_tables: bad args");
  launch_pdl(stage_copy_kernel, dim3(ew_blocks(n)), dim3(256), 0, (cudaStream_t)stream, table, n, step, dst);
  if (advance) launch_pdl(stage_advance_kernel, dim3(1), dim3(1), 0, (cudaStream_t)stream, step, T);
  ATTNDM_CUDA_LAUNCH_CHECK("stage_tables");
  return ATTNDM_OK;
}

}  // extern "C"
