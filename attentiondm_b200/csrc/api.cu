// C-ABI plumbing: error text, version/device probes and the int8 conv dispatch.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "conv_common.cuh"

namespace attndm {

static thread_local char g_err[512] = "";

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("ATTNDM_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

}  // namespace attndm

using namespace attndm;

extern "C" {

const char* attndm_last_error(void) { return g_err; }

int attndm_version(void) { return 100; }

int attndm_device_supported(void) {
  int dev = 0;
  cudaDeviceProp prop;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    set_error("device_supported: no CUDA device");
    return ATTNDM_ERR_CUDA;
  }
  return prop.major == 10 ? 1 : 0;
}

int attndm_qconv_i8(const int8_t* codes, const int32_t* rowsum, int B, int H, int W, int C, const int8_t* qw,
                    const int32_t* wsum, const int32_t* w_zp, int O, int taps, const float* mult,
                    const int32_t* act_zp, const float* bias, const float* residual, const float* temb, float* out,
                    double* gn_stats_out, int impl, void* stream) {
  ATTNDM_CHECK_ARG(codes && rowsum && qw && wsum && w_zp && mult && act_zp && out, "qconv_i8: null pointer");
  ATTNDM_CHECK_ARG(B > 0 && H > 0 && W > 0 && C > 0 && O > 0, "qconv_i8: bad shape");
  ATTNDM_CHECK_ARG(taps == 1 || taps == 9, "qconv_i8: only 1x1 and 3x3/s1/p1 are on the hot path");
  ATTNDM_CHECK_ARG((long long)taps * round_up(C, 16) < 32768, "qconv_i8: K = taps*C must stay below 2^15 (int32 epilogue)");
  ConvI8Params p;
  p.codes = codes; p.rowsum = rowsum; p.B = B; p.H = H; p.W = W; p.C = C; p.Cp = round_up(C, 16);
  p.Hp = taps == 9 ? H + 2 : H;
  p.Wp = taps == 9 ? W + 2 : W;
  p.rows = (long long)B * p.Hp * p.Wp;
  p.qw = qw; p.wsum = wsum; p.w_zp = w_zp; p.O = O; p.taps = taps; p.mult = mult; p.act_zp = act_zp;
  p.bias = bias; p.residual = residual; p.temb = temb; p.out = out;
  if (gn_stats_out && O % 32 != 0) { set_error("qconv_i8: gn_stats_out needs O %% 32 == 0"); return ATTNDM_ERR_ARG; }
  // The statistics of one shape always follow ONE summation order, whichever kernel computes the conv: quad order
  // (conv_common.cuh) where the tcgen05 epilogue can produce it, else the order of attndm_gn_stats.
  const bool quad_order = gn_stats_out != nullptr && conv_gn_quad_ok(p);
  p.gn_out = quad_order ? gn_stats_out : nullptr;
  bool fused = false;
  int rc;
  if (impl == ATTNDM_CONV_TCGEN05) rc = launch_qconv_i8_tc(p, (cudaStream_t)stream, &fused);
  else if (impl == ATTNDM_CONV_SIMT) rc = launch_qconv_i8_simt(p, (cudaStream_t)stream);
  else { set_error("qconv_i8: unknown impl %d", impl); return ATTNDM_ERR_ARG; }
  if (rc) return rc;
  if (gn_stats_out && !fused)
    return quad_order ? launch_gn_stats_quad(out, B, H * W, O, gn_stats_out, (cudaStream_t)stream) : attndm_gn_stats(out, B, H, W, O, gn_stats_out, stream);
  return ATTNDM_OK;
}

}  // extern "C"
