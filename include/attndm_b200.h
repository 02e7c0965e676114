/*
 * attndm_b200.h -- C-ABI of the B200-native hot path of PTQ-AttnDM.
 *
 * The reference (aqilmarwan/attentionDM) is pure Python/PyTorch and has no FFI;
 * its "operator API" is the nn.Module surface.  Each entry point below replaces
 * the eager op chain of one reference method; the citation is the reference
 * file:line whose arithmetic the kernel reproduces.  attentiondm_b200/_ffi.py is
 * the ctypes binding; INTEGRATION.md shows the same binding as a patch to the
 * reference's own modules.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (raw cudaMalloc /
 *     torch storage); the library allocates nothing and keeps no global state
 *     except a mutex-guarded TMA-descriptor cache;
 *   - `stream` is a cudaStream_t passed as void*; kernels are stream-ordered and
 *     never synchronise;
 *   - fp32 activations are NHWC ("channels last"): x[((b*H + h)*W + w)*C + c];
 *   - int8 activation codes come in two row layouts, both with row pitch
 *     Cp = round_up(C, 16) bytes (pad channels are 0):
 *        ATTNDM_ROWS_PLAIN : row = (b*H + h)*W + w
 *        ATTNDM_ROWS_HALO  : row = (b*(H+2) + h+1)*(W+2) + w+1, with a one-pixel
 *                            ring holding the code of 0.0 (= -zero_point), so a
 *                            3x3/pad-1 convolution is nine row-shifted GEMMs;
 *     each code tensor travels with rowsum[row] = sum_c code (int32, same rows);
 *   - return value: 0 = ok, negative = error (attndm_last_error() has the text).
 *     No exception crosses the boundary and there is no CPU fallback.
 */
#ifndef ATTNDM_B200_H
#define ATTNDM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ATTNDM_OK 0
#define ATTNDM_ERR_ARG (-1)
#define ATTNDM_ERR_CUDA (-2)
#define ATTNDM_ERR_UNSUPPORTED (-3)

#define ATTNDM_ROWS_PLAIN 0
#define ATTNDM_ROWS_HALO 1

#define ATTNDM_PRE_NONE 0      /* quantize x as is                               */
#define ATTNDM_PRE_SILU 1      /* quantize silu(x)            (time_mlp)         */
#define ATTNDM_PRE_GN_SILU 2   /* quantize silu(groupnorm(x)) (ResidualBlock)    */

#define ATTNDM_CONV_SIMT 0     /* CUDA-core dp4a implicit GEMM (any shape)       */
#define ATTNDM_CONV_TCGEN05 1  /* tcgen05 kind::i8 + TMA + TMEM implicit GEMM    */

const char* attndm_last_error(void);
int attndm_version(void);
/* 1 if the current device is sm_100 (tcgen05 usable), 0 otherwise, <0 on error */
int attndm_device_supported(void);

/* ---- activation quantizers (HBM-bound) ---------------------------------- */

/* QModule._quantize_activation, inference branch: utils/quant_util.py:260-282
 * with scale/zero_point from asymmetric_linear_quantization_params,
 * utils/quantization_utils/quant_utils.py:109-133 (computed by the host on the
 * tiny [C] vectors and passed in).  Optional pre-op fuses the producer:
 * GroupNorm(32, eps)+SiLU of ResidualBlock.forward (models/diffusion.py:119-127)
 * or the SiLU of time_mlp (models/diffusion.py:157-161).
 *   code = clamp(rne(scale[c]*v - zp[c]), -2^(a-1), 2^(a-1)-1)
 *   y    = (code + zp[c]) / scale[c]
 * Outputs (any may be NULL): codes/rowsum in `rows_layout`, y_f32 NHWC.
 * a_bit == 0 switches the quantizer off: y_f32 = pre_op(x), scale/zp ignored.
 * gn_stats: per (b, group) {sum, sumsq} as double[B*32*2] from attndm_gn_stats. */
int attndm_act_quant(const float* x, int B, int H, int W, int C,
                     const float* scale, const float* zp, int a_bit,
                     int pre_op, const double* gn_stats, const float* gn_gamma,
                     const float* gn_beta, float gn_eps,
                     int8_t* codes, int32_t* rowsum, int rows_layout,
                     float* y_f32, void* stream);

/* GroupNorm statistics (32 groups) of an NHWC tensor: stats[b][g] = {sum, sumsq}
 * accumulated in double.  `stats` must be zeroed by the caller (it is an
 * accumulation target so that a conv epilogue may also feed it).
 * models/diffusion.py:36-37 (Normalize), 91,94 (GroupNorm eps 1e-6). */
int attndm_gn_stats(const float* x, int B, int H, int W, int C, double* stats, void* stream);

/* The input of UpBlock.res1 is cat([upsample_x2(x), skip], dim=1) (models/diffusion.py:225-229,244); that concat
 * is pure data movement, so the hot path never materialises it: its GroupNorm statistics and its two quantizers
 * (res1.conv1 behind GroupNorm+SiLU, models/diffusion.py:119-122, and res1.nin_shortcut, :131-134) read the two parts
 * in place.  xa: [B][H/2][W/2][C1] (seen through the nearest-neighbour x2 upsample), xb: [B][H][W][C2].
 * attndm_gn_stats_cat accumulates the statistics of the (C1 + C2)-channel concat into stats (caller-zeroed): the sums
 * of the upsampled part are exactly 4x its low-resolution sums.  attndm_act_quant_cat = attndm_act_quant on the concat
 * (codes + row sums only); shapes outside attndm_act_quant_cat_fits return ATTNDM_ERR_UNSUPPORTED (the caller then
 * materialises the concat with attndm_upsample_concat). */
int attndm_gn_stats_cat(const float* xa, int Ha, int Wa, int C1, const float* xb, int H, int W, int C2, int B,
                        double* stats, void* stream);
int attndm_act_quant_cat_fits(int H, int W, int C1, int C2);
int attndm_act_quant_cat(const float* xa, int C1, const float* xb, int C2, int B, int H, int W,
                         const float* scale, const float* zp, int a_bit, int pre_op, const double* gn_stats,
                         const float* gn_gamma, const float* gn_beta, float gn_eps, int8_t* codes,
                         int32_t* rowsum, int rows_layout, void* stream);
/* Both quantizers of UpBlock.res1 in ONE pass over the concat: (codes, rowsum) = the GroupNorm+SiLU quantizer of
 * res1.conv1 (models/diffusion.py:119-122), (codes2, rowsum2) = the producer-less quantizer of res1.nin_shortcut /
 * conv_shortcut (:131-134) with its own tables and layout; same bit width.  Bit-identical to two attndm_act_quant_cat
 * calls (pre_op = ATTNDM_PRE_GN_SILU, then ATTNDM_PRE_NONE); the 200 MB input is read once instead of twice. */
int attndm_act_quant_cat2(const float* xa, int C1, const float* xb, int C2, int B, int H, int W,
                          const float* scale, const float* zp, int a_bit, const double* gn_stats,
                          const float* gn_gamma, const float* gn_beta, float gn_eps, int8_t* codes,
                          int32_t* rowsum, int rows_layout, const float* scale2, const float* zp2,
                          int8_t* codes2, int32_t* rowsum2, int rows_layout2, void* stream);

/* GroupNorm(32)+SiLU+quantize in ONE kernel (statistics computed in-kernel, one CTA per sample with
 * the sample's [H*W][C] tile resident in shared memory): same outputs as attndm_gn_stats followed by
 * attndm_act_quant(pre_op = ATTNDM_PRE_GN_SILU).  Only for tiles that fit (attndm_gn_act_quant_fits);
 * returns ATTNDM_ERR_UNSUPPORTED otherwise.  a_bit == 0: y_f32 = silu(groupnorm(x)).
 * models/diffusion.py:119-127 + utils/quant_util.py:260-282. */
int attndm_gn_act_quant_fits(int H, int W, int C);
int attndm_gn_act_quant(const float* x, int B, int H, int W, int C, const float* gamma,
                        const float* beta, float eps, const float* scale, const float* zp, int a_bit,
                        int8_t* codes, int32_t* rowsum, int rows_layout, float* y_f32, void* stream);

/* silu(groupnorm(x)) -> fp32 (used by the calibration branch, which needs the
 * un-quantized activation).  models/diffusion.py:121-122,125-126. */
int attndm_gn_silu(const float* x, int B, int H, int W, int C, const double* gn_stats,
                   const float* gamma, const float* beta, float eps, float* y, void* stream);

/* ---- calibration collectors --------------------------------------------- */

/* Per-channel min / max over (B,H,W): utils/quant_util.py:187-191.
 * workspace: float[2 * C * attndm_minmax_workspace_blocks()]. */
int attndm_minmax_workspace_blocks(void);
int attndm_minmax_c(const float* x, long long rows, int C, float* min_c, float* max_c,
                    float* workspace, void* stream);

/* Floor to the init range then GroupWise_Quantizaion of both vectors, on device:
 * utils/quant_util.py:193-205 and 403-437 (equal-width bins evaluated as
 * min + (div*(m+1))/G in fp32, later bins win ties, empty bin -> upper edge,
 * unassigned channel -> 0).  Writes groups_range_t[G][2] = {gmin[g], gmax[g]} and
 * the per-channel snapped vectors xq_min/xq_max[C] (activation_range_min1/max1). */
int attndm_group_ranges(const float* min_c, const float* max_c, int C, int G,
                        float init_min, float init_max, float* groups_range_t,
                        float* xq_min, float* xq_max, void* stream);

/* Calibration output y = sum_g softmax(alpha)[g,c] * FQ(x; gmin[g], gmax[g], a_bit):
 * utils/quant_util.py:207-224 with Quant.forward :54-66.  sw = softmax weights
 * [G][C]; groups_range_t as written by attndm_group_ranges.  Optionally also
 * accumulates sum |y-x|^p into lp_sum (double, caller-zeroed) for the
 * first-calibrate search, utils/quant_util.py:37-44,237-254. */
int attndm_calib_mix(const float* x, long long rows, int C, int G, const float* groups_range_t,
                     const float* sw, int a_bit, float* y, double* lp_sum, float lp_p,
                     void* stream);

/* find_scale_by_percentile_min/max: utils/quant_util.py:440-450.  k-th smallest
 * of x[n] (k = int(n*(1-p)) or int(n*p)), exact, by 4-pass radix select.
 * workspace: uint32[4*256 + 8]. */
int attndm_kth_value(const float* x, long long n, long long k, float* out,
                     uint32_t* workspace, void* stream);

/* ---- weights -------------------------------------------------------------- */

/* QModule._quantize_weight clamp, utils/quant_util.py:284-303, applied once:
 * w_eff[o][tap][c] (tap-major, channels innermost) from w[o][c][kh][kw]. */
int attndm_weight_clamp_pack(const float* w, int O, int C, int KH, int KW, const float* lo,
                             const float* hi, float* w_eff, void* stream);

/* If every clamped weight sits on the per-out-channel w_bit grid of
 * AsymmetricQuantFunction (utils/quantization_utils/quant_utils.py:136-162)
 * emit int8 codes qw[o][tap*Cp + c], wsum[o] = sum qw, the integer zero points
 * w_zp_i32[o] that go with those codes, and set *on_grid = 1; otherwise
 * *on_grid = 0.  w_scale/w_zp[O] are the grid parameters computed by the host
 * (codes are defined up to the shift (q, zp) -> (q-d, zp+d); the kernel slides a
 * channel whose codes came out as [-127, 128] back into the signed range). */
int attndm_weight_to_i8(const float* w_eff, int O, int C, int taps, const float* w_scale,
                        const float* w_zp, int w_bit, int8_t* qw, int Cp, int32_t* wsum,
                        int32_t* w_zp_i32, int* on_grid, void* stream);

/* ---- convolutions --------------------------------------------------------- */

/* QConv2d.forward -> F.conv2d (utils/quant_util.py:383-385) for 3x3/s1/p1 and 1x1
 * on integer codes.  out[pix][o] = float(I) * mult[o] + bias[o] (+ residual[pix][o])
 * (+ temb[b][o]) with the exact integer
 *   I = sum code*qw + zp*wsum[o] + w_zp[o]*(sum_window rowsum + zp*taps*C).
 * codes/rowsum: HALO rows for taps = 9, PLAIN rows for taps = 1.
 * mult[o] = 1/(act_scale*w_scale[o]); act_zp: int32[1] on device.
 * residual (NHWC fp32, same shape as out) and temb ([B][O]) may be NULL:
 * they fuse ResidualBlock's `x + h` (models/diffusion.py:136) and the block's
 * `x + time_mlp(t_emb)` (models/diffusion.py:175-177).
 * gn_stats_out (double[B*32*2], caller-zeroed) may be NULL: accumulates the
 * GroupNorm statistics of `out` for the next layer.
 * impl: ATTNDM_CONV_SIMT or ATTNDM_CONV_TCGEN05. */
int attndm_qconv_i8(const int8_t* codes, const int32_t* rowsum, int B, int H, int W, int C,
                    const int8_t* qw, const int32_t* wsum, const int32_t* w_zp, int O, int taps,
                    const float* mult, const int32_t* act_zp, const float* bias,
                    const float* residual, const float* temb, float* out,
                    double* gn_stats_out, int impl, void* stream);

/* fp32 convolution on NHWC (3x3/s1/p1 or 1x1) with the same fused epilogue; used
 * by the calibration branch (its input is a G-way mix, not on one integer grid),
 * by layers whose per-channel scales differ (trained alpha_activ) and by weights
 * that are not on the integer grid.  w_eff[o][tap][c]. */
int attndm_conv_f32(const float* x, int B, int H, int W, int C, const float* w_eff, int O,
                    int taps, const float* bias, const float* residual, const float* temb,
                    float* out, void* stream);

/* The same fp32 1x1 conv (a GEMM, out[rows][O] = x[rows][C] . w[O][C]^T + bias) on the tensor cores with
 * fp32-level accuracy: each operand is split exactly into a tf32 "big" part and an fp32 remainder
 * (attndm_split_tf32: big = x with the 13 low mantissa bits cleared, small = x - big) and the product is
 * accumulated as big*big + big*small + small*big by tcgen05.mma kind::tf32 (~2^-21 relative per product).
 * Used for the reference's lazily created fp32 `channel_proj` (models/diffusion.py:235-242) on spatial maps;
 * 1x1 maps keep attndm_conv_f32 (the fused programs reproduce its summation order bit for bit).
 * attndm_conv_f32_tc_fits: rows < 2^31, C % 4 == 0, C >= 64, O % 128 == 0. */
int attndm_conv_f32_tc_fits(long long rows, int C, int O);
int attndm_split_tf32(const float* x, long long n, float* big, float* small, void* stream);
int attndm_gemm_tf32x3(const float* a_big, const float* a_small, long long rows, int C, const float* w_big,
                       const float* w_small, int O, const float* bias, float* out, void* stream);

/* ---- attention ------------------------------------------------------------ */

/* EnhancedQSelfAttention core (models/self_attention.py:132-144):
 *   out = softmax(q k^T * scale) v   per sample; q,k: [B][N][d], v,out: [B][N][dv].
 * heads > 1 selects MixedPrecisionAttention's head split and optional
 * fake-quant of logits / probabilities (utils/attention_quant_utils.py:30-38,
 * 65-107): qk_q / p_q = {scale, zero_point, bits} on the host, bits 0 = off;
 * softmax_scale multiplies the logits before softmax (:91). */
typedef struct {
  float scale;
  float zero_point;
  int bits;
} attndm_attn_quant;
int attndm_attention(const float* q, const float* k, const float* v, float* out, int B, int N,
                     int d, int dv, float scale, int heads, float softmax_scale,
                     attndm_attn_quant qk_q, attndm_attn_quant p_q, void* stream);

/* out = gamma[0] * a + x  (models/self_attention.py:151) */
int attndm_scale_add(const float* a, const float* x, const float* gamma, float* out,
                     long long n, void* stream);

/* ---- UNet glue ------------------------------------------------------------ */

/* nn.MaxPool2d(2) on NHWC (models/diffusion.py:143,183). */
int attndm_maxpool2(const float* x, int B, int H, int W, int C, float* y, void* stream);

/* UpBlock.forward resize + concat (models/diffusion.py:225-244): nearest x2, then
 * nearest-resize to the skip's (Hs, Ws), written to out[..., 0:Cx]; skip copied to
 * out[..., Cx:Cx+Cs].  out is NHWC [B][Hs][Ws][Cx+Cs]. */
int attndm_upsample_concat(const float* x, int B, int H, int W, int Cx, const float* skip,
                           int Hs, int Ws, int Cs, float* out, void* stream);

/* get_timestep_embedding (models/diffusion.py:11-29): emb[b] = [sin(t f), cos(t f)]. */
int attndm_timestep_embedding(const float* t, int B, int dim, float* emb, void* stream);

/* ---- sampler -------------------------------------------------------------- */

/* One generalized (DDIM) update, functions/denoising.py:33-39.
 * coef (device float[5]) = {sqrt(1-at), sqrt(at), sqrt(at_next), c1, c2};
 * noise may be NULL when c1 == 0 (eta = 0).  x0_out may be NULL. */
int attndm_ddim_step(const float* xt, const float* eps, const float* coef, const float* noise,
                     float* x_next, float* x0_out, long long n, void* stream);

/* The same update that ALSO records x_next and x0 of this step in device-side history rings
 * hist_x / hist_x0 [T][n] -- what generalized_steps returns as its per-step lists (functions/denoising.py:34,40
 * appends x0_t.to('cpu') and xt_next.to('cpu') every step).  The slot is (*step_after - 1) mod T, where
 * step_after is the device counter attndm_stage_tables has already advanced for this step, so that one captured
 * CUDA graph fills a different slot at every replay; the host then copies whole chunks of the rings. */
int attndm_ddim_step_hist(const float* xt, const float* eps, const float* coef, const float* noise,
                          float* x_next, float* x0_out, long long n, float* hist_x, float* hist_x0,
                          const int* step_after, int T, void* stream);

/* ---- the callers either side of the sampler ---------- */

/* One ancestral (DDPM) update, functions/denoising.py:137-149 (ddpm_steps; the ablation driver's sampler,
 * ablation_study_attention_quantization.py:300-340).  coef (device float[6]) = {(1/at).sqrt(), (1/at - 1).sqrt(),
 * atm1.sqrt() * beta_t, (1 - beta_t).sqrt() * (1 - atm1), 1 - at, mask * exp(0.5 * log(beta_t))}; x0_out (the
 * clamped x0 prediction) may be NULL. */
int attndm_ddpm_step(const float* xt, const float* eps, const float* coef, const float* noise,
                     float* x_next, float* x0_out, long long n, void* stream);

/* x = x0 * a.sqrt() + e * (1 - a).sqrt(), the forward-noised input of noise_estimation_loss
 * (functions/denoising.py:52-54); coef (device float[2]) = {a.sqrt(), (1 - a).sqrt()}. */
int attndm_noise_mix(const float* x0, const float* e, const float* coef, float* x, long long n, void* stream);

/* out[b] = sum over a sample of (a - b)^2 (double), the per-sample term of noise_estimation_loss
 * (functions/denoising.py:58-60); a, b: [B][per]. */
int attndm_sq_err(const float* a, const float* b, int B, long long per, double* out, void* stream);

/* Gradient of the entropy regulariser of generalized_steps_loss (functions/denoising.py:83-100) for one layer and
 * timestep: term = cal_entropy(softmax(alpha_t, over the G groups)) / (G*C), alpha_t [G][C];
 * grad[g][c] = weight * d term / d alpha_t[g][c]; *value (device double, may be NULL) += weight * term.
 * It is the ONLY non-zero gradient the attention alphas receive: torch.round in every downstream quantizer
 * (utils/quant_util.py:271) has derivative zero and conv_out is itself a QConv2d, so the noise-estimation loss
 * contributes exactly 0 (measured on the reference: oracle/make_golden_calib.py, tests/golden/tiny_calib.npz). */
int attndm_alpha_entropy_grad(const float* alpha_t, int G, int C, float weight, float* grad, double* value,
                              void* stream);

/* Copy row `*step` of a [T][n] table into `dst` and (if advance) increment *step
 * modulo T: lets one captured CUDA graph serve every denoising step, mirroring
 * the per-module index_seq counter (utils/quant_util.py:228-229,281). */
int attndm_stage_tables(const float* table, long long n, int T, int* step, int advance,
                        float* dst, void* stream);


/* ---- fused per-sample layer programs (1x1 feature maps) ------------------- */

/* On a 1x1 feature map every op of the UNet is row-local: GroupNorm reduces over the channels of ONE
 * sample, a 3x3/pad-1 conv is its centre tap, attention over a single position is the identity on V.
 * So a run of blocks at 1x1 (11 DownBlock/UpBlock/middle blocks, ~150 of the 198 QConv2d of the CIFAR
 * model: models/diffusion.py:119-136,170-190,224-252, models/self_attention.py:127-151) is executed by
 * ONE kernel: each CTA owns `ns` samples, keeps their activations in shared memory and interprets a
 * list of ops.  The tiny GEMMs run as warp-level int8 tensor-core MMAs whose weight fragments (and the
 * per-channel parameter vectors) of the NEXT conv are fetched while the current one finishes.
 * Every op reproduces the arithmetic of the stand-alone kernel it replaces bit for bit
 * (attndm_gn_act_quant, attndm_act_quant, attndm_qconv_i8 with taps = 1, attndm_conv_f32,
 * attndm_attention with N = 1, attndm_scale_add, attndm_maxpool2), which is how it is tested.
 *
 * Activation buffers live in a per-CTA fp32 arena; a buffer reference is (offset, leading dimension):
 * element (sample n, channel c) = arena[off + n*ld + c].  Offsets and lds are multiples of 4.
 *
 * A CONV op reads its parameters from shared memory, where the PREVIOUS conv of the program (or the
 * program's first op, which must not be a CONV) has put them: the nx_* fields of those ops describe the
 * next CONV -- nx_qw its weights (attndm_rowprog_pack_weights order), nx_tab_off its row of the staged
 * table ([scale Cq | zero point Cq | mult Oq | act_zp 4], Cq = round_up(C, 4)), nx_stat its static block
 * (floats: gamma[C] beta[C] bias[O], then int32 wsum[O] w_zp[O]; gamma/beta unused without GroupNorm,
 * bias zero when the layer has none). */
#define ATTNDM_ROWOP_END 0
#define ATTNDM_ROWOP_LOAD 1        /* dst[n][0..C) = g0[(s0+n)*g_ld + c]                                   */
#define ATTNDM_ROWOP_LOAD_POOL 2   /* dst[n][c] = max over the 2x2 pixels of g0 viewed as [B][2][2][C]      */
#define ATTNDM_ROWOP_STORE 3       /* g0[(s0+n)*g_ld + c] = src[n][c]                                       */
#define ATTNDM_ROWOP_COPY 4        /* dst = src                                                             */
#define ATTNDM_ROWOP_CONV 5        /* dst = qconv1x1(quant(pre(src))) + bias (+ add0) (+ g1[(s0+n)*O + o]) */
#define ATTNDM_ROWOP_FCONV 6       /* dst = fp32 1x1 conv: g0 = w^T [C][O], g1 = bias[O]                    */
#define ATTNDM_ROWOP_ATTN1 7       /* dst = softmax(q.k * scale) v for one position: src = q, add0 = k, aux = v */
#define ATTNDM_ROWOP_SCALE_ADD 8   /* dst = g0[0] * src + add0                                              */

typedef struct {
  int32_t type;
  int32_t C, O;                    /* input / output channels                                               */
  int32_t src_off, src_ld;
  int32_t dst_off, dst_ld;
  int32_t add0_off, add0_ld;       /* -1 = none                                                             */
  int32_t aux_off, aux_ld;
  int32_t pre, a_bit;              /* CONV: ATTNDM_PRE_*, activation bits                                   */
  int32_t tab_off;                 /* CONV: float offset of this layer's row in the staged table            */
  int32_t nx_C, nx_O, nx_tab_off;  /* the next CONV of the program (see above)                              */
  int32_t g_ld;                    /* LOAD/STORE: row pitch of g0 in floats                                 */
  float fparam;                    /* CONV: GroupNorm eps; ATTN1: logit scale                               */
  int32_t g0_ext;                  /* >= 0: g0 is taken from ext[g0_ext] of the launch instead (LOAD/STORE) */
  const void* g0;                  /* LOAD/STORE/FCONV/SCALE_ADD global pointer                             */
  const void* g1;                  /* CONV: temb [B][O] or NULL; FCONV: bias                                */
  const void* qw;                  /* CONV: int8 weights in attndm_rowprog_pack_weights order               */
  const void* stat;                /* CONV: this layer's static block (informational)                       */
  const void* nx_qw;               /* next CONV: weights, or NULL when there is none                        */
  const void* nx_stat;             /* next CONV: static block                                               */
  const void* rsv0;
  const void* rsv1;
} attndm_rowop;

/* ops: device array of attndm_rowop; prog_start: device int32[nprog], index of each program's first op
 * (a program ends at ATTNDM_ROWOP_END).  Grid = ceil(B/ns) x nprog CTAs; ns in {2,4,8}.
 * arena_floats: per-CTA arena size; cp_max: largest input channel count of any CONV (multiple of 16);
 * pbuf_floats: largest parameter block over the CONV ops: (2*Cq + Oq + 4) + (2*C + 3*O) floats.
 * cur: the staged per-step table (attndm_stage_tables); CTA x reads its rows at cur + x * cur_cta_stride floats
 * (0: every sample uses the staged row; the row stride of the whole [T][width] table with ns samples per step: the
 * time path of all T sampler steps in one launch, attentiondm_b200/engine.py).
 * ext: host array of n_ext (<= 4) device pointers bound at launch time, for tensors whose address is
 * only known per call (the program's input and output activations). */
int attndm_rowprog(const attndm_rowop* ops, const int32_t* prog_start, int nprog, int B, int ns,
                   int arena_floats, int cp_max, int pbuf_floats, const float* cur, long long cur_cta_stride,
                   const void* const* ext, int n_ext, void* stream);
/* dst[off_dst[j] + b * width[j] + c] = cur[off_src[j] + c] for b < B, c < width[j], j < n: the per-step rows of a staged
 * table fanned out to [B][width] tensors.  The sampler's time path (timestep embedding -> time_embed Linears ->
 * every block's time_mlp, models/diffusion.py:157-161,273-277,347-351) depends on the step alone, not on the sample:
 * the engine evaluates it once per pass for all T steps, keeps the results as table columns, and this kernel hands the
 * current step's rows to the consumers that expect [B][O].  desc: device int32[n][3] = {off_dst, off_src, width}. */
int attndm_bcast_rows(const float* cur, const int32_t* desc, int n, int B, int max_width, float* dst, void* stream);
/* shared memory the kernel would need (bytes), for the host-side planner */
int attndm_rowprog_smem_bytes(int ns, int arena_floats, int cp_max, int pbuf_floats);
/* int8 weights [O][Cp] -> MMA-fragment order [ceil(O/16)][ceil(Cp/32)][32 lanes][16 B] (zero padded);
 * `out` needs attndm_rowprog_packed_weight_bytes(O, Cp) bytes. */
long long attndm_rowprog_packed_weight_bytes(int O, int Cp);
int attndm_rowprog_pack_weights(const int8_t* qw, int O, int Cp, int8_t* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ATTNDM_B200_H */
