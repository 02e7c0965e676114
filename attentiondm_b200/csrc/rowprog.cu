// Fused per-sample layer programs for 1x1 feature maps (include/attndm_b200.h: attndm_rowprog).
//
// On a 1x1 map everything the UNet does is row-local (one row = one sample): GroupNorm(32) reduces
// over the channels of one sample (models/diffusion.py:119-127), a 3x3/pad-1 conv is its centre tap
// (utils/quant_util.py:383-385), attention over a single position is the identity on V
// (models/self_attention.py:132-144).  ~150 of the 198 QConv2d of the CIFAR model run there, as ~330
// launches of a few microseconds each.  Here one CTA owns NS samples, keeps their fp32 activations in a
// shared-memory arena and interprets a host-built op list.
//
// The work per op is tiny, so the kernel is built around latency, not throughput:
//   * the GEMM of a CONV op is [O x C] x [C x NS<=8] on int8: legacy warp-level tensor-core MMAs
//     (mma.sync m16n8k32, s8 x s8 -> s32; dp4a measured ~8x slower on this part).  The weights are
//     stored in MMA-fragment order, so a warp loads its A fragments with ONE coalesced 128-bit load per
//     32 input channels, straight from L2 into registers -- and it does so for the NEXT conv right after
//     the current conv's MMAs, so that the load latency hides behind the epilogue and the next conv's
//     GroupNorm/quantize phases (weights are static);
//   * the per-channel parameter vectors of the next conv (scale / zero-point / multiplier row of the
//     staged table, GroupNorm affine, bias, weight sums) are fetched the same way into a double-buffered
//     shared-memory block;
//   * op descriptors are staged one op ahead.
// Every op reproduces the arithmetic of the stand-alone kernel it replaces bit for bit (same helper
// functions, same summation order), which is what the tests check.
#include <mutex>

#include "common.cuh"
#include "conv_common.cuh"

namespace attndm {

constexpr int RP_THREADS = 512;            // 16 warps
constexpr int RP_WARPS = RP_THREADS / 32;
constexpr int RP_OCH = 256;                // FCONV: output channels per pass; the two thread halves split the samples
constexpr int RP_SPLIT = RP_THREADS / RP_OCH;
constexpr int RP_MAX_CONV_O = 512;         // CONV: two 16-channel MMA tiles per warp at most
constexpr int RP_MAXT = RP_MAX_CONV_O / 16 / RP_WARPS;
constexpr int RP_KCH = 8;                  // k32 steps per register chunk of A fragments (256 input channels)
constexpr int RP_CODE_PAD = 16;            // bytes added to a code row so that the 8 samples of a B fragment hit different banks

__device__ __forceinline__ void rp_sync() { __syncthreads(); }

// D(16x8,s32) += A(16x32,s8,row) * B(32x8,s8,col): one instruction covers 16 output channels x 8 samples x
// 32 input channels.
__device__ __forceinline__ void rp_mma_s8(int (&d)[4], const uint4& a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
      : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}
// Bulk asynchronous global -> shared copies (TMA, cp.async.bulk) tracked by mbarriers: ONE instruction moves a
// warp's whole 4 KB block of A fragments (or a conv's parameter block), holds no register scoreboard, and is
// waited for only where the data is consumed.  (Per-thread 16-byte cp.async cost ~8 issue cycles each --
// 5600 of them per conv took longer than the conv's MMAs.)
__device__ __forceinline__ uint32_t rp_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void rp_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void rp_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  long long t0 = clock64();
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && clock64() - t0 > 4000000000LL) __trap();     // a lost arrival fails the launch instead of hanging
  } while (!done);
}
__device__ __forceinline__ void rp_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void rp_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void rp_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ uint4 rp_ldg128(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// parameter block of one CONV in shared memory (floats):
//   [row: scale Cq | zp Cq | mult Oq | act_zp 4]  (a verbatim copy of the layer's staged-table row)
//   [stat: gamma C | beta C | bias O | wsum O | w_zp O]  (host-built, static)
__host__ __device__ __forceinline__ int rp_q4(int v) { return (v + 3) & ~3; }
__host__ __device__ __forceinline__ int rp_row_floats(int C, int O) { return 2 * rp_q4(C) + rp_q4(O) + 4; }
__host__ __device__ __forceinline__ int rp_stat_floats(int C, int O) { return 2 * C + 3 * O; }

// GroupNorm sums of one (sample, group) of a 1x1 map by ONE thread, in exactly the order of the 32-lane
// reduction of gn_act_quant_sample_kernel: lane i holds element i (0 for i >= CPG), then lane i += lane i^o
// for o = 16, 8, 4, 2, 1 -- the steps with o >= CPG add zeros, the others form this in-place tree.
template <int CPG>
__device__ __forceinline__ void rp_gn_pair(const float* x, double inv_n, float eps, float& mean, float& rstd) {
  double a[CPG], q[CPG];
#pragma unroll
  for (int i = 0; i < CPG; ++i) {
    const double v = (double)x[i];
    a[i] = v;              // (the reduction's "0.0 + v" is v: nine double adds less per pair on a slow FP64 pipe)
    q[i] = v * v;
  }
#pragma unroll
  for (int o = CPG >> 1; o > 0; o >>= 1)
#pragma unroll
    for (int i = 0; i < o; ++i) {
      a[i] += a[i + o];
      q[i] += q[i + o];
    }
  gn_mean_rstd(a[0], q[0], inv_n, eps, mean, rstd);
}

struct RpExt {
  const void* p[4];
};

// optional timeline (debug builds, -DATTNDM_RP_TRACE: `python -m attentiondm_b200.build --variant=rp_trace`, tools/rowprog_trace.py):
// thread 0 of CTA (0,0) stamps globaltimer at phase boundaries of every op.  Compiled out otherwise -- a hook is a
// dependent global load of the buffer pointer by every thread, seven of them per conv.
__device__ unsigned long long* g_rp_trace = nullptr;    // [op][8]
#ifdef ATTNDM_RP_TRACE
__device__ __forceinline__ void rp_trace(int opi, int ev) {
  unsigned long long* t = g_rp_trace;
  if (t != nullptr && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && opi < 256) {
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    t[opi * 8 + ev] = now;
  }
}
#else
__device__ __forceinline__ void rp_trace(int, int) {}
#endif

template <int NS>
__global__ void __launch_bounds__(RP_THREADS, 1)
rowprog_kernel(const attndm_rowop* __restrict__ ops, const int32_t* __restrict__ prog_start, int B, int arena_floats,
               int cp_max, int pbuf_floats, const float* __restrict__ cur_base, long long cur_cta_stride, const RpExt ext) {
  extern __shared__ __align__(16) uint8_t rp_smem[];
  __shared__ __align__(16) attndm_rowop s_op[2];       // the current and the next op, staged from global memory
  __shared__ float s_mean[NS * kGnGroups], s_rstd[NS * kGnGroups];
  __shared__ int s_rowsum[8];
  __shared__ float s_prob[8 * 8];                      // [NS][heads <= 8]
  __shared__ __align__(8) uint64_t s_wbar[RP_WARPS];   // per warp: its staged A fragments have landed
  __shared__ __align__(8) uint64_t s_pbar[2];          // per parameter buffer
  pdl_launch_dependents();

  uint8_t* const wstage = rp_smem;                                               // [16 warps][8 k32 steps][512 B]
  float* const params = reinterpret_cast<float*>(rp_smem + RP_WARPS * RP_KCH * 512);   // [2][pbuf_floats]
  float* const arena = params + 2 * (size_t)pbuf_floats;                         // [arena_floats]
  int8_t* const codes = reinterpret_cast<int8_t*>(arena + arena_floats);        // [NS][C + RP_CODE_PAD]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int s0 = blockIdx.x * NS;
  // the table row this CTA's samples are quantized with: the staged current-step row, or (cur_cta_stride != 0) row
  // blockIdx.x of a whole [T][width] table -- the time path of ALL sampler steps in one launch (engine.py)
  const float* const cur = cur_base + (long long)blockIdx.x * cur_cta_stride;
  const attndm_rowop* prog = ops + prog_start[blockIdx.y];
  constexpr int OPW = (int)(sizeof(attndm_rowop) / 4);
  if (tid < OPW) reinterpret_cast<int*>(&s_op[0])[tid] = reinterpret_cast<const int*>(prog)[tid];
  if (tid == 0) {
    for (int i = 0; i < RP_WARPS; ++i) rp_mbar_init(rp_smem_u32(&s_wbar[i]), 1);
    rp_mbar_init(rp_smem_u32(&s_pbar[0]), 1);
    rp_mbar_init(rp_smem_u32(&s_pbar[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  pdl_wait();                          // inputs (activations, staged table) come from earlier kernels

  const int hh = tid / RP_OCH, ot = tid - hh * RP_OCH;    // FCONV mapping: sample half, output channel
  constexpr int NST = NS / RP_SPLIT;
  const int grp = lane >> 2, tig = lane & 3;              // MMA fragment coordinates

  uint32_t nconv = 0;                  // CONV ops seen so far: parity selects the parameter buffer
  uint8_t* const wmine = wstage + (size_t)warp * (RP_KCH * 512) + lane * 16;    // this lane's slots of the staged fragments
  uint32_t wuse = 0;                   // how many times this warp's fragment barrier has completed (parity)

  // out[n][o] = (sum over c, sequential fmaf, of x[n][c] * wt[c][o]) + bias[o]: the arithmetic of conv_f32_simt_kernel on a
  // 1x1 map.  Used by the FCONV op (the un-quantized channel_proj) and by CONV ops whose layer takes the fp32 path.
  // Weight-bandwidth bound per CTA (every CTA streams the whole [C][O] matrix for its few samples): each thread owns four
  // consecutive output channels and keeps 16 independent 128-bit loads in flight.  bias: global or shared, or NULL.
  auto fconv_rows = [&](const float* wt, const float* bias, const float* x0, int ld, int doff, int dld, int C, int O) {
    const float* xs = x0 + hh * NST * ld;
    if (O <= RP_OCH) {
      // narrow outputs (the q/k/v/out convs of an attention block on the fp32 path: O <= 256): ONE output channel per
      // thread, 32 loads in flight each -- with four channels per thread only O/4 threads per sample half had work
      // and the bytes in flight per CTA bounded the stream (12 us per 256 -> 256 conv)
      const int o = ot;
      if (o < O) {
        float acc[NST];
#pragma unroll
        for (int n = 0; n < NST; ++n) acc[n] = 0.f;
        const float* wp = wt + o;
        int c = 0;
        for (; c + 32 <= C; c += 32) {
          float w[32];
#pragma unroll
          for (int u = 0; u < 32; ++u) w[u] = __ldg(wp + (long long)(c + u) * O);
#pragma unroll
          for (int u = 0; u < 32; ++u)
#pragma unroll
            for (int n = 0; n < NST; ++n) acc[n] = fmaf(xs[n * ld + c + u], w[u], acc[n]);
        }
        for (; c < C; ++c) {
          const float w = __ldg(wp + (long long)c * O);
#pragma unroll
          for (int n = 0; n < NST; ++n) acc[n] = fmaf(xs[n * ld + c], w, acc[n]);
        }
        const float bz = bias ? bias[o] : 0.f;
#pragma unroll
        for (int n = 0; n < NST; ++n) arena[doff + (hh * NST + n) * dld + o] = acc[n] + bz;
      }
      return;
    }
    for (int o = 4 * ot; o < O; o += 4 * RP_OCH) {
      float acc[NST][4];
#pragma unroll
      for (int n = 0; n < NST; ++n)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[n][e] = 0.f;
      const float* wp = wt + o;
      int c = 0;
      for (; c + 16 <= C; c += 16) {
        float4 w[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) w[u] = __ldg(reinterpret_cast<const float4*>(wp + (long long)(c + u) * O));
#pragma unroll
        for (int u = 0; u < 16; ++u)
#pragma unroll
          for (int n = 0; n < NST; ++n) {
            const float xv = xs[n * ld + c + u];
            fma2(acc[n][0], acc[n][1], xv, w[u].x, w[u].y);      // packed fp32 FMAs: same results as fmaf
            fma2(acc[n][2], acc[n][3], xv, w[u].z, w[u].w);
          }
      }
      for (; c < C; ++c) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(wp + (long long)c * O));
#pragma unroll
        for (int n = 0; n < NST; ++n) {
          const float xv = xs[n * ld + c];
          fma2(acc[n][0], acc[n][1], xv, w.x, w.y);
          fma2(acc[n][2], acc[n][3], xv, w.z, w.w);
        }
      }
      float4 bz = make_float4(0.f, 0.f, 0.f, 0.f);
      if (bias) bz = *reinterpret_cast<const float4*>(bias + o);
#pragma unroll
      for (int n = 0; n < NST; ++n) {
        float* d = arena + doff + (hh * NST + n) * dld + o;
        d[0] = acc[n][0] + bz.x; d[1] = acc[n][1] + bz.y; d[2] = acc[n][2] + bz.z; d[3] = acc[n][3] + bz.w;
      }
    }
  };

  for (int opi = 0;; ++opi) {
    const attndm_rowop& op = s_op[opi & 1];
    const int type = op.type;
    if (type == ATTNDM_ROWOP_END) break;
    int next_word = 0;                   // the next op: loaded now, parked in shared memory at the end of this op
    if (tid < OPW) next_word = reinterpret_cast<const int*>(prog + opi + 1)[tid];
    const int C = op.C, O = op.O;
    // Prefetch of the NEXT conv (described by the nx_* fields of a CONV op and of a program's first op) with
    // bulk copies: its first 256 input channels of A fragments (tile = warp) into this warp's staging block
    // (one copy per warp), its parameter block into the other parameter buffer (two copies by one thread).
    // Nobody waits for them until that conv's phase A2 / phase B.  (The staging block and the parameter buffer
    // were last READ through the generic proxy and those reads have retired -- their values fed the MMAs /
    // the epilogue before this point -- so no proxy fence is needed before the async writes; an explicit
    // fence.proxy.async here cost ~0.4 us per conv.)
    const bool has_next = (type == ATTNDM_ROWOP_CONV || opi == 0) && op.nx_qw != nullptr;
    const int nC = op.nx_C, nO = op.nx_O;
#define RP_PREFETCH_ISSUE(parity)                                                                        \
    if (has_next) {                                                                                        \
      const int nk32n = (nC + 31) >> 5, ntilen = (nO + 15) >> 4;                                           \
      __syncwarp();                                                                                        \
      if (warp < ntilen && lane == 0) {                                                                    \
        const uint32_t nb = (uint32_t)(nk32n < RP_KCH ? nk32n : RP_KCH) * 512u;                            \
        rp_mbar_expect_tx(rp_smem_u32(&s_wbar[warp]), nb);                                                 \
        rp_bulk_g2s(rp_smem_u32(wstage + (size_t)warp * (RP_KCH * 512)),                                   \
                    reinterpret_cast<const uint8_t*>(op.nx_qw) + (size_t)warp * nk32n * 512, nb,           \
                    rp_smem_u32(&s_wbar[warp]));                                                           \
      }                                                                                                    \
      if (tid == RP_THREADS - 32) {                                                                        \
        const uint32_t nrowb = (uint32_t)rp_row_floats(nC, nO) * 4u, nstatb = (uint32_t)rp_stat_floats(nC, nO) * 4u; \
        float* dstp = params + (size_t)((parity) & 1) * pbuf_floats;                                       \
        const uint32_t bar = rp_smem_u32(&s_pbar[(parity) & 1]);                                           \
        rp_mbar_expect_tx(bar, nrowb + nstatb);                                                            \
        rp_bulk_g2s(rp_smem_u32(dstp), cur + op.nx_tab_off, nrowb, bar);                                   \
        rp_bulk_g2s(rp_smem_u32(dstp) + nrowb, op.nx_stat, nstatb, bar);                                   \
      }                                                                                                    \
    }
    rp_trace(opi, 0);
    if (type != ATTNDM_ROWOP_CONV) { RP_PREFETCH_ISSUE(nconv); }  // (only a program's first op qualifies)
    switch (type) {
      case ATTNDM_ROWOP_LOAD: {
        const float* g = reinterpret_cast<const float*>(op.g0_ext >= 0 ? ext.p[op.g0_ext & 3] : op.g0);
        const int Q = C >> 2, gld = op.g_ld, doff = op.dst_off, dld = op.dst_ld;
        for (int n = 0; n < NS; ++n) {
          for (int q = tid; q < Q; q += RP_THREADS) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (s0 + n < B) v = *reinterpret_cast<const float4*>(g + (long long)(s0 + n) * gld + 4 * q);
            *reinterpret_cast<float4*>(arena + doff + n * dld + 4 * q) = v;
          }
        }
        break;
      }
      case ATTNDM_ROWOP_LOAD_POOL: {     // nn.MaxPool2d(2) of a 2x2 map, same expression as maxpool2_kernel
        const float* g = reinterpret_cast<const float*>(op.g0_ext >= 0 ? ext.p[op.g0_ext & 3] : op.g0);
        const int doff = op.dst_off, dld = op.dst_ld;
        for (int n = 0; n < NS; ++n) {
          for (int c = tid; c < C; c += RP_THREADS) {
            float v = 0.f;
            if (s0 + n < B) {
              const float* p00 = g + (long long)(s0 + n) * 4 * C + c;
              v = fmaxf(fmaxf(p00[0], p00[C]), fmaxf(p00[2 * C], p00[3 * C]));
            }
            arena[doff + n * dld + c] = v;
          }
        }
        break;
      }
      case ATTNDM_ROWOP_STORE: {
        float* g = reinterpret_cast<float*>(const_cast<void*>(op.g0_ext >= 0 ? ext.p[op.g0_ext & 3] : op.g0));
        const int Q = C >> 2, gld = op.g_ld, soff = op.src_off, sld = op.src_ld;
        for (int n = 0; n < NS; ++n) {
          if (s0 + n >= B) break;
          for (int q = tid; q < Q; q += RP_THREADS)
            *reinterpret_cast<float4*>(g + (long long)(s0 + n) * gld + 4 * q) =
                *reinterpret_cast<const float4*>(arena + soff + n * sld + 4 * q);
        }
        break;
      }
      case ATTNDM_ROWOP_COPY: {
        const int Q = C >> 2, soff = op.src_off, sld = op.src_ld, doff = op.dst_off, dld = op.dst_ld;
        for (int n = 0; n < NS; ++n)
          for (int q = tid; q < Q; q += RP_THREADS)
            *reinterpret_cast<float4*>(arena + doff + n * dld + 4 * q) =
                *reinterpret_cast<const float4*>(arena + soff + n * sld + 4 * q);
        break;
      }
      case ATTNDM_ROWOP_SCALE_ADD: {       // models/self_attention.py:151, same expression as scale_add_kernel
        const float g = *reinterpret_cast<const float*>(op.g0);
        const int soff = op.src_off, sld = op.src_ld, doff = op.dst_off, dld = op.dst_ld, aoff = op.add0_off, ald = op.add0_ld;
        for (int n = 0; n < NS; ++n)
          for (int c = tid; c < C; c += RP_THREADS)
            arena[doff + n * dld + c] = __fadd_rn(__fmul_rn(g, arena[soff + n * sld + c]), arena[aoff + n * ald + c]);
        break;
      }
      case ATTNDM_ROWOP_ATTN1: {           // attention_kernel with N = 1: C = d (q/k channels), O = dv
        // op.g0 (optional): 8 floats {heads, softmax_scale, qk scale, qk zero point, qk bits, p scale, p zero point,
        // p bits} -- MixedPrecisionAttention (utils/attention_quant_utils.py:51-107); NULL: one head, no quantizers
        const float* ap = reinterpret_cast<const float*>(op.g0);
        const int heads = ap ? (int)ap[0] : 1;
        if (tid < NS * heads) {
          const int n = tid / heads, head = tid - n * heads, dq = C / heads;
          const float* qr = arena + op.src_off + n * op.src_ld + head * dq;       // q: head-major channels
          const float* kr = arena + op.add0_off + n * op.add0_ld;                 // k: channel = e * heads + head
          float acc = 0.f;
          for (int e = 0; e < dq; ++e) acc = fmaf(qr[e], kr[e * heads + head], acc);
          float s = __fmul_rn(acc, op.fparam);
          if (ap && (int)ap[4] > 0) s = attn_fake_quant(s, ap[2], ap[3], (float)((1 << (int)ap[4]) - 1));
          s = __fmul_rn(s, ap ? ap[1] : 1.0f);
          const float mx = fmaxf(-INFINITY, s);
          const float e1 = expf(s - mx);
          float sum = e1;
#pragma unroll
          for (int k = 0; k < 5; ++k) sum += 0.f;      // the warp reduction over 31 empty lanes
          float pr = __fdiv_rn(e1, sum);
          if (ap && (int)ap[7] > 0) pr = attn_fake_quant(pr, ap[5], ap[6], (float)((1 << (int)ap[7]) - 1));
          s_prob[tid] = pr;
        }
        rp_sync();
        const int doff = op.dst_off, dld = op.dst_ld, aoff = op.aux_off, ald = op.aux_ld, dvh = O / heads;
        for (int n = 0; n < NS; ++n)
          for (int c = tid; c < O; c += RP_THREADS)
            arena[doff + n * dld + c] = fmaf(s_prob[n * heads + c / dvh], arena[aoff + n * ald + c], 0.f);
        break;
      }
      case ATTNDM_ROWOP_FCONV: {           // conv_f32_simt_kernel: sequential fmaf over c, then + bias
        fconv_rows(reinterpret_cast<const float*>(op.g0), reinterpret_cast<const float*>(op.g1), arena + op.src_off, op.src_ld,
                   op.dst_off, op.dst_ld, C, O);
        break;
      }
      case ATTNDM_ROWOP_CONV: {
        const float* prm = params + (size_t)(nconv & 1) * pbuf_floats;      // staged by the previous conv's prefetch
        const int Cq = rp_q4(C), Oq = rp_q4(O);
        const float* scale = prm;
        const float* zpv = prm + Cq;
        const float* mult = prm + 2 * Cq;
        const float* stat = prm + rp_row_floats(C, O);
        const float* gamma = stat;
        const float* beta = stat + C;
        const float* bias = stat + 2 * C;
        const int32_t* wsum = reinterpret_cast<const int32_t*>(stat + 2 * C + O);
        const int32_t* wzp = reinterpret_cast<const int32_t*>(stat + 2 * C + 2 * O);
        const int pre = op.pre;
        // rsv0 != NULL: the layer takes the fp32 path on every step (non-uniform alpha, off-grid weights; QConv2d.forward_fused):
        // same quantizer, then the fp32 conv of its fake-quantized input (op.qw = fp32 weights [C][O], op.aux = scratch row)
        const bool f32 = op.rsv0 != nullptr;
        const float qlo = -(float)(1 << (op.a_bit - 1)), qhi = (float)((1 << (op.a_bit - 1)) - 1);
        const float* src = arena + op.src_off;
        const int sld = op.src_ld;
        const int crow_bytes = C + RP_CODE_PAD;
        if (tid < NS) s_rowsum[tid] = 0;
        rp_trace(opi, 1);
        // ---- phase A1: GroupNorm statistics, the summation tree of gn_act_quant_sample_kernel ----
        if (pre == ATTNDM_PRE_GN_SILU) {
          const int cpg = C / kGnGroups;
          const double inv_n = 1.0 / (double)cpg;
          const float eps = op.fparam;
          if (cpg <= 32 && (cpg & (cpg - 1)) == 0) {
            for (int pair = tid; pair < NS * kGnGroups; pair += RP_THREADS) {     // one thread per (sample, group)
              const int n = pair / kGnGroups, g = pair - n * kGnGroups;
              const float* x = src + n * sld + g * cpg;
              float mean, rstd;
              switch (cpg) {
                case 1: rp_gn_pair<1>(x, inv_n, eps, mean, rstd); break;
                case 2: rp_gn_pair<2>(x, inv_n, eps, mean, rstd); break;
                case 4: rp_gn_pair<4>(x, inv_n, eps, mean, rstd); break;
                case 8: rp_gn_pair<8>(x, inv_n, eps, mean, rstd); break;
                case 16: rp_gn_pair<16>(x, inv_n, eps, mean, rstd); break;
                default: rp_gn_pair<32>(x, inv_n, eps, mean, rstd); break;
              }
              s_mean[pair] = mean;
              s_rstd[pair] = rstd;
            }
          } else {
            for (int pair = warp; pair < NS * kGnGroups; pair += RP_WARPS) {
              const int n = pair / kGnGroups, g = pair - n * kGnGroups;
              double a = 0.0, q = 0.0;
              for (int i = lane; i < cpg; i += 32) {
                const double v = (double)src[n * sld + g * cpg + i];
                a += v;
                q += v * v;
              }
              a = warp_sum_d(a);
              q = warp_sum_d(q);
              if (lane == 0) gn_mean_rstd(a, q, inv_n, eps, s_mean[pair], s_rstd[pair]);
            }
          }
        }
        rp_mbar_wait(rp_smem_u32(&s_pbar[nconv & 1]), (nconv >> 1) & 1);    // this conv's parameter block has landed
        rp_sync();
        const int zp = *reinterpret_cast<const int*>(prm + 2 * Cq + Oq);     // (nothing of `prm` may be read before the wait)
        rp_trace(opi, 2);
        // ---- phase A2: producer op + quantize (utils/quant_util.py:260-282), codes and per-sample code sums ----
        {
          const int Q = C >> 1, cpg = C / kGnGroups;        // items of two channels: short dependent chains
          const bool warp_rows = (C & 63) == 0;           // a warp's 32 items belong to one sample
          const bool pow2 = (Q & (Q - 1)) == 0;
          const int qsh = 31 - __clz(Q);
          for (int i = tid; i < NS * Q; i += RP_THREADS) {
            int n, q;
            if (pow2) { n = i >> qsh; q = i & (Q - 1); }
            else { n = i / Q; q = i - n * Q; }
            const int c = q << 1;
            float2 v = *reinterpret_cast<const float2*>(src + n * sld + c);
            if (pre == ATTNDM_PRE_GN_SILU) {
              const float2 g2 = *reinterpret_cast<const float2*>(gamma + c);
              const float2 b2 = *reinterpret_cast<const float2*>(beta + c);
              const float* mn = s_mean + n * kGnGroups;
              const float* rs = s_rstd + n * kGnGroups;
              const int g0 = c / cpg, g1 = (c + 1) / cpg;
              v.x = gn_apply(v.x, mn[g0], rs[g0], g2.x, b2.x);
              v.y = gn_apply(v.y, mn[g1], rs[g1], g2.y, b2.y);
            }
            const float2 s2 = *reinterpret_cast<const float2*>(scale + c);
            const float2 z2 = *reinterpret_cast<const float2*>(zpv + c);
            // same pre-round values as the stand-alone quantizer kernels (quant_t in quant_kernels.cu)
            const float tx = pre == ATTNDM_PRE_NONE ? __fsub_rn(__fmul_rn(s2.x, v.x), z2.x) : silu_quant_t(v.x, s2.x, z2.x);
            const float ty = pre == ATTNDM_PRE_NONE ? __fsub_rn(__fmul_rn(s2.y, v.y), z2.y) : silu_quant_t(v.y, s2.y, z2.y);
            const int ix = (int)quant_code_t(tx, qlo, qhi), iy = (int)quant_code_t(ty, qlo, qhi);
            if (f32)          // fp32 path: the fake-quantized activation itself (dequant of the same code, common.cuh)
              *reinterpret_cast<float2*>(arena + op.aux_off + n * op.aux_ld + c) =
                  make_float2(dequant((float)ix, s2.x, z2.x), dequant((float)iy, s2.y, z2.y));
            *reinterpret_cast<char2*>(codes + n * crow_bytes + c) = make_char2((signed char)ix, (signed char)iy);
            int part = ix + iy;
            if (warp_rows) {
              part = __reduce_add_sync(0xffffffffu, part);
              if (lane == 0) atomicAdd(&s_rowsum[n], part);
            } else {
              atomicAdd(&s_rowsum[n], part);
            }
          }
        }
        rp_sync();
        rp_trace(opi, 3);
        // ---- phase B: the GEMM on the tensor cores.  warp = one 16-channel tile (two for O > 256); the A fragments
        //      of the first 256 input channels of tile `warp` are already in registers (prefetched by the
        //      previous conv); anything beyond is loaded here ----
        if (f32) {
          // the fragment prefetch an earlier op issued for this conv (it cannot know the path) is consumed unread
          if (warp < ((O + 15) >> 4)) { rp_mbar_wait(rp_smem_u32(&s_wbar[warp]), wuse & 1); ++wuse; }
          fconv_rows(reinterpret_cast<const float*>(op.qw), bias, arena + op.aux_off, op.aux_ld, op.dst_off, op.dst_ld, C, O);
          rp_trace(opi, 4);
          RP_PREFETCH_ISSUE(nconv + 1);
          ++nconv;
          rp_trace(opi, 5);
          break;
        }
        int acc[RP_MAXT][4], acc2[RP_MAXT][4];
        float tev[RP_MAXT][4];             // the time-embedding terms of phase C, fetched from global memory ahead of the MMAs
#pragma unroll
        for (int t = 0; t < RP_MAXT; ++t)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            acc[t][j] = 0;
            acc2[t][j] = 0;
            tev[t][j] = 0.f;
            if (op.g1 != nullptr) {
              const int o = (warp + t * RP_WARPS) * 16 + grp + 8 * (j >> 1), n = 2 * tig + (j & 1);
              if (n < NS && o < O && s0 + n < B) tev[t][j] = __ldg(reinterpret_cast<const float*>(op.g1) + (long long)(s0 + n) * O + o);
            }
          }
        {
          const int nk32 = (C + 31) >> 5, ntile = (O + 15) >> 4;
          const uint8_t* crow = reinterpret_cast<const uint8_t*>(codes) + grp * crow_bytes + tig * 4;
          const bool has_n = grp < NS;
#pragma unroll
          for (int t = 0; t < RP_MAXT; ++t) {
            const int tile = warp + t * RP_WARPS;
            if (tile < ntile) {
              const uint8_t* w = reinterpret_cast<const uint8_t*>(op.qw) + ((size_t)tile * nk32 * 32 + lane) * 16;
              for (int k0 = 0; k0 < nk32; k0 += RP_KCH) {
                uint4 wf[RP_KCH];
                if (t > 0 || k0 > 0) {
#pragma unroll
                  for (int j = 0; j < RP_KCH; ++j)
                    if (k0 + j < nk32) wf[j] = rp_ldg128(w + (size_t)(k0 + j) * 512);
                } else {
                  rp_mbar_wait(rp_smem_u32(&s_wbar[warp]), wuse & 1);       // prefetched by the previous conv
                  ++wuse;
                  rp_trace(opi, 7);
#pragma unroll
                  for (int j = 0; j < RP_KCH; ++j)
                    if (j < nk32) wf[j] = *reinterpret_cast<const uint4*>(wmine + j * 512);
                }
                // all B fragments of the chunk first, then the MMAs back to back (one load-to-use latency per chunk
                // instead of one per MMA)
                uint32_t bf0[RP_KCH], bf1[RP_KCH];
#pragma unroll
                for (int j = 0; j < RP_KCH; ++j) {
                  bf0[j] = 0;
                  bf1[j] = 0;
                  if (has_n && k0 + j < nk32) {
                    bf0[j] = *reinterpret_cast<const uint32_t*>(crow + (k0 + j) * 32);
                    bf1[j] = *reinterpret_cast<const uint32_t*>(crow + (k0 + j) * 32 + 16);
                  }
                }
#pragma unroll
                for (int j = 0; j < RP_KCH; ++j) {
                  if (k0 + j < nk32) {
                    if (j & 1) rp_mma_s8(acc2[t], wf[j], bf0[j], bf1[j]);     // two independent accumulation chains
                    else rp_mma_s8(acc[t], wf[j], bf0[j], bf1[j]);
                  }
                }
              }
            }
          }
        }
        rp_trace(opi, 4);
        RP_PREFETCH_ISSUE(nconv + 1);    // the next conv's weights and parameters: in flight from here on
        // ---- phase C: exact integer finish + fused adds (conv_common.cuh), written to the arena.
        //      accumulator j of a tile: output channel 16*tile + grp + 8*(j>>1), sample 2*tig + (j&1) ----
        {
          const float* temb = reinterpret_cast<const float*>(op.g1);
          const int aoff = op.add0_off, ald = op.add0_ld, doff = op.dst_off, dld = op.dst_ld;
          const int ntile = (O + 15) >> 4;
#pragma unroll
          for (int t = 0; t < RP_MAXT; ++t) {
            const int tile = warp + t * RP_WARPS;
            if (tile < ntile) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int o = tile * 16 + grp + 8 * (j >> 1), n = 2 * tig + (j & 1);
                if (n < NS && o < O) {
                  float v = conv_i8_value(acc[t][j] + acc2[t][j], zp * wsum[o], wzp[o], s_rowsum[n] + zp * C, mult[o], bias[o]);
                  if (aoff >= 0) v = __fadd_rn(v, arena[aoff + n * ald + o]);
                  if (temb && s0 + n < B) v = __fadd_rn(v, tev[t][j]);
                  arena[doff + n * dld + o] = v;
                }
              }
            }
          }
        }
        ++nconv;
        rp_trace(opi, 5);
        break;
      }
      default:
        __trap();
    }
    if (tid < OPW) reinterpret_cast<int*>(&s_op[(opi + 1) & 1])[tid] = next_word;
    rp_sync();
    rp_trace(opi, 6);
  }
#undef RP_PREFETCH_ISSUE
}

// int8 weights [O][Cp] -> MMA-fragment order [tile][k32][lane][16 B]: lane (grp, tig) holds
//   a0 = W[16 tile + grp][32 j + 4 tig ..+3], a1 = row + 8, a2 = a0's row at k + 16, a3 = row + 8 at k + 16
// (rows >= O and columns >= Cp are zero).
__global__ void rowprog_pack_weights_kernel(const int8_t* __restrict__ qw, int O, int Cp, int8_t* __restrict__ out) {
  const int nk32 = (Cp + 31) >> 5, ntile = (O + 15) >> 4;
  const long long n = (long long)ntile * nk32 * 32 * 4;         // one 4-byte register per thread iteration
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int reg = (int)(i & 3), lane = (int)((i >> 2) & 31);
    const long long tj = i >> 7;
    const int j = (int)(tj % nk32), tile = (int)(tj / nk32);
    const int grp = lane >> 2, tig = lane & 3;
    const int row = tile * 16 + grp + 8 * (reg & 1), col = j * 32 + 16 * (reg >> 1) + 4 * tig;
    uint32_t v = 0;
    if (row < O) {
#pragma unroll
      for (int b = 0; b < 4; ++b)
        if (col + b < Cp) v |= (uint32_t)(uint8_t)qw[(long long)row * Cp + col + b] << (8 * b);
    }
    reinterpret_cast<uint32_t*>(out)[i] = v;
  }
}

static int rp_smem_bytes(int ns, int arena_floats, int cp_max, int pbuf_floats) {
  return RP_WARPS * RP_KCH * 512 + 2 * pbuf_floats * 4 + arena_floats * 4 + ns * (cp_max + RP_CODE_PAD) + 16;
}

}  // namespace attndm

using namespace attndm;

extern "C" {

int attndm_debug_set_rp_trace(unsigned long long* buf) {
  cudaError_t e = cudaMemcpyToSymbol(attndm::g_rp_trace, &buf, sizeof(buf));
  return e == cudaSuccess ? 0 : -2;
}

int attndm_rowprog_smem_bytes(int ns, int arena_floats, int cp_max, int pbuf_floats) {
  return rp_smem_bytes(ns, arena_floats, cp_max, pbuf_floats);
}

long long attndm_rowprog_packed_weight_bytes(int O, int Cp) {
  return (long long)((O + 15) / 16) * ((Cp + 31) / 32) * 512;
}

int attndm_rowprog(const attndm_rowop* ops, const int32_t* prog_start, int nprog, int B, int ns, int arena_floats,
                   int cp_max, int pbuf_floats, const float* cur, long long cur_cta_stride, const void* const* ext_ptrs,
                   int n_ext, void* stream) {
  ATTNDM_CHECK_ARG(ops && prog_start && nprog > 0 && B > 0, "rowprog: bad args");
  ATTNDM_CHECK_ARG(ns == 2 || ns == 4 || ns == 8, "rowprog: ns must be 2, 4 or 8");
  ATTNDM_CHECK_ARG(arena_floats > 0 && (arena_floats & 3) == 0 && cp_max > 0 && (cp_max & 15) == 0, "rowprog: bad arena / cp_max");
  ATTNDM_CHECK_ARG(pbuf_floats >= 4 && (pbuf_floats & 3) == 0, "rowprog: bad pbuf_floats");
  ATTNDM_CHECK_ARG(n_ext >= 0 && n_ext <= 4 && (n_ext == 0 || ext_ptrs), "rowprog: at most 4 external pointers");
  ATTNDM_CHECK_ARG(cur && cur_cta_stride >= 0 && (cur_cta_stride & 3) == 0, "rowprog: bad table / per-CTA table stride");
  RpExt ext = {{nullptr, nullptr, nullptr, nullptr}};
  for (int i = 0; i < n_ext; ++i) ext.p[i] = ext_ptrs[i];
  const int smem = rp_smem_bytes(ns, arena_floats, cp_max, pbuf_floats);
  ATTNDM_CHECK_ARG(smem <= 220 * 1024, "rowprog: program needs %d bytes of shared memory (max 220 KB)", smem);
  dim3 grid(cdiv(B, ns), nprog);
  static std::once_flag once2, once4, once8;
  static cudaError_t attr_err = cudaSuccess;
  auto raise = [](void (*k)(const attndm_rowop*, const int32_t*, int, int, int, int, const float*, long long, const RpExt)) {
    cudaError_t r = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (r != cudaSuccess) attr_err = r;
  };
  if (ns == 2) { std::call_once(once2, raise, rowprog_kernel<2>); }
  else if (ns == 4) { std::call_once(once4, raise, rowprog_kernel<4>); }
  else { std::call_once(once8, raise, rowprog_kernel<8>); }
  if (attr_err != cudaSuccess) { set_error("rowprog: cannot raise dynamic smem: %s", cudaGetErrorString(attr_err)); return ATTNDM_ERR_CUDA; }
  if (ns == 2) launch_pdl(rowprog_kernel<2>, grid, dim3(RP_THREADS), smem, (cudaStream_t)stream, ops, prog_start, B, arena_floats, cp_max, pbuf_floats, cur, cur_cta_stride, ext);
  else if (ns == 4) launch_pdl(rowprog_kernel<4>, grid, dim3(RP_THREADS), smem, (cudaStream_t)stream, ops, prog_start, B, arena_floats, cp_max, pbuf_floats, cur, cur_cta_stride, ext);
  else launch_pdl(rowprog_kernel<8>, grid, dim3(RP_THREADS), smem, (cudaStream_t)stream, ops, prog_start, B, arena_floats, cp_max, pbuf_floats, cur, cur_cta_stride, ext);
  ATTNDM_CUDA_LAUNCH_CHECK("rowprog");
  return ATTNDM_OK;
}

int attndm_rowprog_pack_weights(const int8_t* qw, int O, int Cp, int8_t* out, void* stream) {
  ATTNDM_CHECK_ARG(qw && out && O > 0 && Cp > 0 && (Cp & 15) == 0, "rowprog_pack_weights: bad args");
  const long long n = attndm_rowprog_packed_weight_bytes(O, Cp) / 4;
  rowprog_pack_weights_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(qw, O, Cp, out);
  ATTNDM_CUDA_LAUNCH_CHECK("rowprog_pack_weights");
  return ATTNDM_OK;
}

}  // extern "C"
