"""Tensor-level wrappers over the C-ABI (attentiondm_b200/_ffi.py).

All activations here are NHWC fp32 CUDA tensors of shape [B, H, W, C]; the
nn.Module layer converts from/to the reference's logical NCHW at its boundary
(a channels_last tensor IS this layout, so the conversion is a view).
torch is used for device memory and streams only.
"""
from __future__ import annotations

import os

from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import _ffi as F_
from ._ffi import (CONV_SIMT, CONV_TCGEN05, PRE_GN_SILU, PRE_NONE, PRE_SILU, ROWS_HALO, ROWS_PLAIN, AttnQuant,
                   call, ptr, stream)

GN_GROUPS = 32

# which int8 conv kernel the modules use; tests flip it to cross-check the two
DEFAULT_CONV_IMPL = CONV_TCGEN05


def _chk(x: torch.Tensor, what="tensor"):
    if not (isinstance(x, torch.Tensor) and x.is_cuda):
        raise RuntimeError(f"attentiondm_b200: {what} must be a CUDA tensor (no CPU fallback on this path)")
    if x.dtype != torch.float32:
        raise RuntimeError(f"attentiondm_b200: {what} must be float32, got {x.dtype}")
    if not x.is_contiguous():
        raise RuntimeError(f"attentiondm_b200: {what} must be contiguous NHWC")


def to_nhwc(x: torch.Tensor) -> torch.Tensor:
    """logical NCHW -> contiguous [B,H,W,C] (a view when x is channels_last)."""
    if x.dim() != 4:
        raise RuntimeError("attentiondm_b200: expected a 4-D NCHW tensor")
    if not x.is_cuda:
        raise RuntimeError("attentiondm_b200: CUDA tensor required; this path has no CPU fallback")
    y = x.permute(0, 2, 3, 1)
    if y.dtype != torch.float32:
        y = y.float()
    return y.contiguous()


def to_nchw(y: torch.Tensor) -> torch.Tensor:
    """[B,H,W,C] -> logical NCHW view (channels_last memory)."""
    return y.permute(0, 3, 1, 2)


def cp_of(c: int) -> int:
    return (c + 15) // 16 * 16


@dataclass
class GnArgs:
    stats: Optional[torch.Tensor]   # double [B, 32, 2] sums; None = compute in the fused per-sample kernel
    gamma: torch.Tensor
    beta: torch.Tensor
    eps: float


class CatView:
    """The input of UpBlock.res1 -- cat([upsample_x2(x), skip], channel) (models/diffusion.py:225-229,244) -- WITHOUT the
    copy: `lo` is x [B, H/2, W/2, C1], `skip` is [B, H, W, C2].  gn_stats / act_quant read the two parts in place
    (attndm_gn_stats_cat, attndm_act_quant_cat); anything else calls materialize(), which runs the concat kernel once."""

    def __init__(self, lo: torch.Tensor, skip: torch.Tensor):
        _chk(lo, "CatView x")
        _chk(skip, "CatView skip")
        self.lo, self.skip = lo, skip
        B, H, W, C2 = skip.shape
        self.shape = torch.Size((B, H, W, lo.shape[-1] + C2))
        self.device = skip.device
        self._full = None

    @staticmethod
    def fits(lo: torch.Tensor, skip: torch.Tensor) -> bool:
        """The shapes the in-place kernels take, and only where the per-sample fused GroupNorm kernel is not used."""
        if os.environ.get("ATTNDM_CAT", "1") == "0":
            return False
        B, H, W, C1 = lo.shape
        Bs, Hs, Ws, C2 = skip.shape
        return (B == Bs and Hs == 2 * H and Ws == 2 * W and not gn_fits_fused(Hs, Ws, C1 + C2)
                and bool(F_.lib().attndm_act_quant_cat_fits(Hs, Ws, C1, C2)) and B * Hs * Ws * (C1 + C2) < 2 ** 31)

    def materialize(self) -> torch.Tensor:
        if self._full is None:
            self._full = upsample_concat(self.lo, self.skip)
        return self._full

    # -- the two quantizers of UpBlock.res1 (conv1 behind GroupNorm+SiLU, the shortcut conv on the raw concat) in one
    #    pass over the input (attndm_act_quant_cat2); act_quant() picks the prepared codes up by their table pointers
    def prepare_pair(self, main, second, gn: "GnArgs"):
        """main / second: (scale, zp, a_bit, halo) of the GroupNorm+SiLU quantizer and of the producer-less one."""
        (s1, z1, a1, h1), (s2, z2, a2, h2) = main, second
        if a1 != a2 or gn is None or gn.stats is None or os.environ.get("ATTNDM_CAT_DUAL", "1") == "0":
            return False
        B, H, W, Cc = self.shape
        Cp = cp_of(Cc)
        out = []
        for halo in (h1, h2):
            rows = B * (H + 2) * (W + 2) if halo else B * H * W
            out.append((torch.empty(rows, Cp, dtype=torch.int8, device=self.device),
                        torch.empty(rows, dtype=torch.int32, device=self.device)))
        call("attndm_act_quant_cat2", ptr(self.lo), self.lo.shape[-1], ptr(self.skip), self.skip.shape[-1], B, H, W,
             ptr(s1), ptr(z1), int(a1), ptr(gn.stats), ptr(gn.gamma), ptr(gn.beta), float(gn.eps),
             ptr(out[0][0]), ptr(out[0][1]), ROWS_HALO if h1 else ROWS_PLAIN,
             ptr(s2), ptr(z2), ptr(out[1][0]), ptr(out[1][1]), ROWS_HALO if h2 else ROWS_PLAIN, stream())
        self._prepared = {(s1.data_ptr(), z1.data_ptr(), int(a1), PRE_GN_SILU, bool(h1)): out[0],
                          (s2.data_ptr(), z2.data_ptr(), int(a2), PRE_NONE, bool(h2)): out[1]}
        return True

    def take_prepared(self, scale, zp, a_bit, pre, halo):
        prep = getattr(self, "_prepared", None)
        if not prep:
            return None
        return prep.pop((scale.data_ptr(), zp.data_ptr(), int(a_bit), int(pre), bool(halo)), None)

    def cpu(self):
        return self.materialize().cpu()


def gn_fits_fused(H: int, W: int, Cc: int) -> bool:
    return bool(F_.lib().attndm_gn_act_quant_fits(H, W, Cc))


# One zero-filled [n, B, 32, 2] double buffer per UNet forward, handed out slice by slice, instead of
# one memset per GroupNorm (97 of them on the CIFAR model).
_gn_pool = None
_gn_pool_next = 0


def gn_pool_begin(n: int, B: int, device):
    global _gn_pool, _gn_pool_next
    _gn_pool = torch.zeros(n, B, GN_GROUPS, 2, dtype=torch.float64, device=device)
    _gn_pool_next = 0


def gn_pool_end():
    global _gn_pool
    _gn_pool = None


def gn_stats_buffer(B: int, device) -> torch.Tensor:
    """A zeroed [B, 32, 2] double accumulation target (a slice of the per-forward pool when one is open)."""
    global _gn_pool_next
    if _gn_pool is not None and _gn_pool_next < _gn_pool.shape[0] and _gn_pool.shape[1] == B:
        out = _gn_pool[_gn_pool_next]
        _gn_pool_next += 1
        return out
    return torch.zeros(B, GN_GROUPS, 2, dtype=torch.float64, device=device)


def conv_stats_enabled() -> bool:
    """GroupNorm statistics of a conv output from the conv's own epilogue (ATTNDM_CONV_STATS=0: separate pass)."""
    return os.environ.get("ATTNDM_CONV_STATS", "1") != "0"


def gn_stats(x, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    cat = x if isinstance(x, CatView) else None
    if cat is None:
        _chk(x, "gn_stats input")
    B, H, W, Cc = x.shape
    if out is None:
        out = gn_stats_buffer(B, x.device)
    else:
        out.zero_()
    if cat is not None:
        call("attndm_gn_stats_cat", ptr(cat.lo), H // 2, W // 2, cat.lo.shape[-1], ptr(cat.skip), H, W,
             cat.skip.shape[-1], B, ptr(out), stream())
        return out
    call("attndm_gn_stats", ptr(x), B, H, W, Cc, ptr(out), stream())
    return out


def gn_silu(x, gn: GnArgs) -> torch.Tensor:
    if isinstance(x, CatView):
        x = x.materialize()
    _chk(x, "gn_silu input")
    B, H, W, Cc = x.shape
    y = torch.empty_like(x)
    if gn.stats is None and not gn_fits_fused(H, W, Cc):
        gn = GnArgs(stats=gn_stats(x), gamma=gn.gamma, beta=gn.beta, eps=gn.eps)    # deferred statistics
    if gn.stats is None:
        call("attndm_gn_act_quant", ptr(x), B, H, W, Cc, ptr(gn.gamma), ptr(gn.beta), float(gn.eps), None, None, 0,
             None, None, ROWS_PLAIN, ptr(y), stream())
        return y
    call("attndm_gn_silu", ptr(x), B, H, W, Cc, ptr(gn.stats), ptr(gn.gamma), ptr(gn.beta), float(gn.eps), ptr(y),
         stream())
    return y


def act_quant(x: torch.Tensor, scale: torch.Tensor, zp: torch.Tensor, a_bit: int, pre: int = PRE_NONE,
              gn: Optional[GnArgs] = None, want_codes: bool = True, halo: bool = False, want_f32: bool = False):
    """Returns (codes int8 [rows, Cp] | None, rowsum int32 [rows] | None, y fp32 NHWC | None)."""
    if isinstance(x, CatView):
        hit = x.take_prepared(scale, zp, a_bit, pre, halo) if want_codes and not want_f32 else None
        if hit is not None:                                  # both quantizers of this concat came out of one pass
            return hit[0], hit[1], None
        if want_codes and not want_f32 and (pre != PRE_GN_SILU or gn.stats is not None):
            B, H, W, Cc = x.shape
            rows = B * (H + 2) * (W + 2) if halo else B * H * W
            codes = torch.empty(rows, cp_of(Cc), dtype=torch.int8, device=x.device)
            rowsum = torch.empty(rows, dtype=torch.int32, device=x.device)
            call("attndm_act_quant_cat", ptr(x.lo), x.lo.shape[-1], ptr(x.skip), x.skip.shape[-1], B, H, W, ptr(scale),
                 ptr(zp), int(a_bit), int(pre), ptr(gn.stats) if gn else None, ptr(gn.gamma) if gn else None,
                 ptr(gn.beta) if gn else None, float(gn.eps) if gn else 0.0, ptr(codes), ptr(rowsum),
                 ROWS_HALO if halo else ROWS_PLAIN, stream())
            return codes, rowsum, None
        x = x.materialize()
    _chk(x, "act_quant input")
    B, H, W, Cc = x.shape
    Cp = cp_of(Cc)
    codes = rowsum = y = None
    if want_codes:
        rows = B * (H + 2) * (W + 2) if halo else B * H * W
        alloc = torch.zeros if Cp != Cc else torch.empty
        codes = alloc(rows, Cp, dtype=torch.int8, device=x.device)
        rowsum = torch.empty(rows, dtype=torch.int32, device=x.device)
    if want_f32:
        y = torch.empty_like(x)
    if pre == PRE_GN_SILU and gn.stats is None and not gn_fits_fused(H, W, Cc):
        # statistics were deferred but this consumer cannot use the fused per-sample kernel: compute them now
        gn = GnArgs(stats=gn_stats(x), gamma=gn.gamma, beta=gn.beta, eps=gn.eps)
    if pre == PRE_GN_SILU and gn.stats is None:
        call("attndm_gn_act_quant", ptr(x), B, H, W, Cc, ptr(gn.gamma), ptr(gn.beta), float(gn.eps), ptr(scale),
             ptr(zp), int(a_bit), ptr(codes), ptr(rowsum), ROWS_HALO if halo else ROWS_PLAIN, ptr(y), stream())
        return codes, rowsum, y
    call("attndm_act_quant", ptr(x), B, H, W, Cc, ptr(scale), ptr(zp), int(a_bit), int(pre),
         ptr(gn.stats) if gn else None, ptr(gn.gamma) if gn else None, ptr(gn.beta) if gn else None,
         float(gn.eps) if gn else 0.0, ptr(codes), ptr(rowsum), ROWS_HALO if halo else ROWS_PLAIN, ptr(y), stream())
    return codes, rowsum, y


def minmax_c(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    _chk(x, "minmax_c input")
    Cc = x.shape[-1]
    rows = x.numel() // Cc
    nblk = F_.lib().attndm_minmax_workspace_blocks()
    ws = torch.empty(nblk * 2 * Cc, dtype=torch.float32, device=x.device)
    mn = torch.empty(Cc, dtype=torch.float32, device=x.device)
    mx = torch.empty(Cc, dtype=torch.float32, device=x.device)
    call("attndm_minmax_c", ptr(x), rows, Cc, ptr(mn), ptr(mx), ptr(ws), stream())
    return mn, mx


def group_ranges(min_c, max_c, G: int, init_min: float, init_max: float, out_gr_t: torch.Tensor):
    """Writes out_gr_t [G,2] in place; returns (xq_min, xq_max) [C]."""
    Cc = min_c.numel()
    if not (out_gr_t.is_cuda and out_gr_t.is_contiguous() and out_gr_t.shape == (G, 2)):
        raise RuntimeError("group_ranges: out_gr_t must be a contiguous CUDA [G,2] tensor")
    xq_min = torch.empty_like(min_c)
    xq_max = torch.empty_like(max_c)
    call("attndm_group_ranges", ptr(min_c), ptr(max_c), Cc, G, float(init_min), float(init_max), ptr(out_gr_t),
         ptr(xq_min), ptr(xq_max), stream())
    return xq_min, xq_max


def calib_mix(x: torch.Tensor, gr_t: torch.Tensor, sw: torch.Tensor, a_bit: int, lp_p: Optional[float] = None):
    """Returns y (and the lp sum as a 1-element double tensor when lp_p is given)."""
    _chk(x, "calib_mix input")
    Cc = x.shape[-1]
    rows = x.numel() // Cc
    G = gr_t.shape[0]
    y = torch.empty_like(x)
    lp = torch.zeros(1, dtype=torch.float64, device=x.device) if lp_p is not None else None
    gr_c, sw_c = gr_t.contiguous(), sw.contiguous()      # named: a temporary could be freed (and its block reused) before the call
    call("attndm_calib_mix", ptr(x), rows, Cc, G, ptr(gr_c), ptr(sw_c), int(a_bit), ptr(y),
         ptr(lp), float(lp_p or 0.0), stream())
    return (y, lp) if lp_p is not None else y


def kth_value(x: torch.Tensor, k: int) -> torch.Tensor:
    xf = x.reshape(-1)
    _chk(xf, "kth_value input")
    ws = torch.empty(4 * 256 + 8, dtype=torch.int32, device=x.device)
    out = torch.empty(1, dtype=torch.float32, device=x.device)
    call("attndm_kth_value", ptr(xf), xf.numel(), int(k), ptr(out), ptr(ws), stream())
    return out


def weight_clamp_pack(w: torch.Tensor, lo: torch.Tensor, hi: torch.Tensor) -> torch.Tensor:
    """w [O,C,kh,kw] -> clamped w_eff [O, taps, C]."""
    O, Cc, KH, KW = w.shape
    w = w.detach().float().contiguous()
    w_eff = torch.empty(O, KH * KW, Cc, dtype=torch.float32, device=w.device)
    lo_f, hi_f = lo.float().contiguous(), hi.float().contiguous()     # keep both conversions alive across the call
    call("attndm_weight_clamp_pack", ptr(w), O, Cc, KH, KW, ptr(lo_f), ptr(hi_f), ptr(w_eff), stream())
    return w_eff


@dataclass
class I8Pack:
    qw: torch.Tensor        # int8 [O, taps*Cp]
    wsum: torch.Tensor      # int32 [O]
    w_zp: torch.Tensor      # int32 [O]
    w_scale: torch.Tensor   # float32 [O]
    on_grid: bool


def weight_grid(w_eff: torch.Tensor, w_bit: int, slack: int = 0):
    """(scale[O], zero_point[O]) of AsymmetricQuantFunction's grid (utils/quantization_utils/
    quant_utils.py:109-133) as recovered from the per-out-channel min/max of on-grid weights.
    `slack` = number of grid steps by which the attained extremes fall short of the full 2^b - 1 span
    (a channel whose original min/max tie at a rounding boundary never attains its top code)."""
    flat = w_eff.reshape(w_eff.shape[0], -1)
    lo, hi = flat.min(1)[0], flat.max(1)[0]
    n = 2 ** w_bit - 1 - slack
    w_scale = n / (hi - lo)
    w_zp = (w_scale * lo).round() + 2 ** (w_bit - 1)
    return w_scale, w_zp


def _weight_to_i8_on(w_eff, w_bit, grid) -> I8Pack:
    O, taps, Cc = w_eff.shape
    Cp = cp_of(Cc)
    w_scale, w_zp = grid
    qw = torch.empty(O, taps * Cp, dtype=torch.int8, device=w_eff.device)
    wsum = torch.empty(O, dtype=torch.int32, device=w_eff.device)
    flag = torch.empty(1, dtype=torch.int32, device=w_eff.device)
    wzp_i = torch.empty(O, dtype=torch.int32, device=w_eff.device)
    ws_c, wz_c = w_scale.float().contiguous(), w_zp.float().contiguous()
    call("attndm_weight_to_i8", ptr(w_eff), O, Cc, taps, ptr(ws_c), ptr(wz_c),
         int(w_bit), ptr(qw), Cp, ptr(wsum), ptr(wzp_i), ptr(flag), stream())
    on_grid = bool(flag.item() == 1) and bool(torch.isfinite(w_scale).all().item())
    return I8Pack(qw=qw, wsum=wsum, w_zp=wzp_i, w_scale=w_scale, on_grid=on_grid)


def weight_to_i8(w_eff: torch.Tensor, w_bit: int, grid=None) -> I8Pack:
    """Integer codes of w_eff [O, taps, C].  With grid=None the per-channel grid is recovered from the
    weights themselves: per channel, the first span in {2^b-1, 2^b-2, 2^b-3} steps that puts every
    weight of that channel on an integer."""
    w_eff = w_eff.contiguous()
    if grid is not None:
        return _weight_to_i8_on(w_eff, w_bit, grid)
    pack = _weight_to_i8_on(w_eff, w_bit, weight_grid(w_eff, w_bit, 0))
    if pack.on_grid:
        return pack
    # rare: some channels do not attain their extreme codes -> choose the span per channel
    O = w_eff.shape[0]
    flat = w_eff.reshape(O, -1)
    best_s, best_z = weight_grid(w_eff, w_bit, 0)
    done = torch.zeros(O, dtype=torch.bool, device=w_eff.device)
    for slack in (0, 1, 2):
        s_k, z_k = weight_grid(w_eff, w_bit, slack)
        back = s_k[:, None] * flat - z_k[:, None]
        ok = ((back - back.round()).abs().amax(1) <= 1e-3) & torch.isfinite(s_k)
        take = ok & ~done
        best_s = torch.where(take, s_k, best_s)
        best_z = torch.where(take, z_k, best_z)
        done |= ok
    return _weight_to_i8_on(w_eff, w_bit, (best_s, best_z))


def qconv_i8(codes, rowsum, B: int, H: int, W: int, Cc: int, pack: I8Pack, taps: int, mult, act_zp, bias,
             residual=None, temb=None, impl: Optional[int] = None, out: Optional[torch.Tensor] = None,
             gn_stats_out: Optional[torch.Tensor] = None):
    """gn_stats_out: zeroed double [B, 32, 2]; receives the GroupNorm {sum, sumsq} of the output (from the conv's own
    epilogue where the kernel supports it, include/attndm_b200.h)."""
    O = pack.qw.shape[0]
    if out is None:
        out = torch.empty(B, H, W, O, dtype=torch.float32, device=codes.device)
    call("attndm_qconv_i8", ptr(codes), ptr(rowsum), B, H, W, Cc, ptr(pack.qw), ptr(pack.wsum), ptr(pack.w_zp), O,
         taps, ptr(mult), ptr(act_zp), ptr(bias), ptr(residual), ptr(temb), ptr(out), ptr(gn_stats_out),
         DEFAULT_CONV_IMPL if impl is None else impl, stream())
    return out


def conv_f32_tc_fits(rows: int, Cc: int, O: int) -> bool:
    if os.environ.get("ATTNDM_F32_TC", "1") == "0":
        return False
    return bool(F_.lib().attndm_conv_f32_tc_fits(int(rows), int(Cc), int(O)))


def split_tf32(x: torch.Tensor):
    """x = big + small exactly: big has the 13 low mantissa bits cleared (a tf32 value), small is the remainder."""
    big, small = torch.empty_like(x), torch.empty_like(x)
    call("attndm_split_tf32", ptr(x), x.numel(), ptr(big), ptr(small), stream())
    return big, small


def conv1x1_f32_tc(x: torch.Tensor, w_split, bias) -> torch.Tensor:
    """fp32 1x1 conv on the tensor cores (3xTF32).  x NHWC fp32; w_split = split_tf32 of the [O, C] weight."""
    _chk(x, "conv1x1_f32_tc input")
    B, H, W, Cc = x.shape
    w_big, w_small = w_split
    O = w_big.shape[0]
    a_big, a_small = split_tf32(x)
    out = torch.empty(B, H, W, O, dtype=torch.float32, device=x.device)
    call("attndm_gemm_tf32x3", ptr(a_big), ptr(a_small), B * H * W, Cc, ptr(w_big), ptr(w_small), O, ptr(bias), ptr(out),
         stream())
    return out


def conv_f32(x: torch.Tensor, w_eff: torch.Tensor, bias, residual=None, temb=None):
    _chk(x, "conv_f32 input")
    B, H, W, Cc = x.shape
    O, taps, _ = w_eff.shape
    out = torch.empty(B, H, W, O, dtype=torch.float32, device=x.device)
    call("attndm_conv_f32", ptr(x), B, H, W, Cc, ptr(w_eff), O, taps, ptr(bias), ptr(residual), ptr(temb), ptr(out),
         stream())
    return out


def attention(q, k, v, scale: float, heads: int = 1, softmax_scale: float = 1.0, qk_q=None, p_q=None):
    """q,k [B,N,d]; v [B,N,dv] -> [B,N,dv]."""
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        _chk(t, "attention " + n)
    B, N, d = q.shape
    dv = v.shape[-1]
    out = torch.empty_like(v)
    zq = AttnQuant(1.0, 0.0, 0)
    call("attndm_attention", ptr(q), ptr(k), ptr(v), ptr(out), B, N, d, dv, float(scale), int(heads),
         float(softmax_scale), AttnQuant(*qk_q) if qk_q else zq, AttnQuant(*p_q) if p_q else zq, stream())
    return out


def silu(x: torch.Tensor) -> torch.Tensor:
    """x / (1 + exp(-x)) through the act_quant kernel's SiLU producer with the quantizer off."""
    _chk(x, "silu input")
    x4 = x if x.dim() == 4 else x.reshape(1, 1, -1, x.shape[-1])
    B, H, W, Cc = x4.shape
    y = torch.empty_like(x4)
    call("attndm_act_quant", ptr(x4), B, H, W, Cc, None, None, 0, PRE_SILU, None, None, None, 0.0, None, None,
         ROWS_PLAIN, ptr(y), stream())
    return y.view(x.shape)


def scale_add(a, x, gamma):
    out = torch.empty_like(x)
    call("attndm_scale_add", ptr(a), ptr(x), ptr(gamma), ptr(out), x.numel(), stream())
    return out


def maxpool2(x):
    _chk(x, "maxpool2 input")
    B, H, W, Cc = x.shape
    y = torch.empty(B, H // 2, W // 2, Cc, dtype=torch.float32, device=x.device)
    call("attndm_maxpool2", ptr(x), B, H, W, Cc, ptr(y), stream())
    return y


def upsample_concat(x, skip):
    _chk(x, "upsample_concat x")
    _chk(skip, "upsample_concat skip")
    B, H, W, Cx = x.shape
    _, Hs, Ws, Cs = skip.shape
    out = torch.empty(B, Hs, Ws, Cx + Cs, dtype=torch.float32, device=x.device)
    call("attndm_upsample_concat", ptr(x), B, H, W, Cx, ptr(skip), Hs, Ws, Cs, ptr(out), stream())
    return out


def timestep_embedding(t: torch.Tensor, dim: int):
    t = t.float().contiguous()
    emb = torch.empty(t.numel(), dim, dtype=torch.float32, device=t.device)
    call("attndm_timestep_embedding", ptr(t), t.numel(), dim, ptr(emb), stream())
    return emb


def ddim_step(xt, eps, coef, noise=None, x_next=None, x0_out=None, want_x0=False, hist=None):
    """hist = (hist_x [T, ...], hist_x0 [T, ...], step_after int32[1]): also record this step's x_next / x0 in
    slot (step_after - 1) mod T of the device-side history rings (attndm_ddim_step_hist)."""
    if x_next is None:
        x_next = torch.empty_like(xt)
    x0 = x0_out if x0_out is not None else (torch.empty_like(xt) if want_x0 else None)
    if hist is not None:
        hx, hx0, step_after = hist
        call("attndm_ddim_step_hist", ptr(xt), ptr(eps), ptr(coef), ptr(noise), ptr(x_next), ptr(x0), xt.numel(),
             ptr(hx), ptr(hx0), ptr(step_after), hx.shape[0], stream())
    else:
        call("attndm_ddim_step", ptr(xt), ptr(eps), ptr(coef), ptr(noise), ptr(x_next), ptr(x0), xt.numel(), stream())
    return x_next, x0


def stage_tables(table: torch.Tensor, step: torch.Tensor, dst: torch.Tensor, advance: bool = True):
    T, n = table.shape
    call("attndm_stage_tables", ptr(table), n, T, ptr(step), 1 if advance else 0, ptr(dst), stream())


def ddpm_step(xt, eps, coef, noise, want_x0=True):
    """One ddpm_steps update (functions/denoising.py:137-149); coef: device float[6] (denoising.ddpm_coefficients)."""
    x_next = torch.empty_like(xt)
    x0 = torch.empty_like(xt) if want_x0 else None
    call("attndm_ddpm_step", ptr(xt), ptr(eps), ptr(coef), ptr(noise), ptr(x_next), ptr(x0), xt.numel(), stream())
    return x_next, x0


def noise_mix(x0, e, coef):
    """x0 * a.sqrt() + e * (1 - a).sqrt(); coef: device float[2]."""
    x = torch.empty_like(x0)
    call("attndm_noise_mix", ptr(x0), ptr(e), ptr(coef), ptr(x), x0.numel(), stream())
    return x


def sq_err(a, b):
    """Per-sample sum of squared differences, double [B]."""
    B = a.shape[0]
    out = torch.empty(B, dtype=torch.float64, device=a.device)
    call("attndm_sq_err", ptr(a), ptr(b), B, a.numel() // B, ptr(out), stream())
    return out


def alpha_entropy_grad(alpha_t, weight, grad_out, value=None):
    """grad_out[G,C] = weight * d(cal_entropy(softmax(alpha_t, 0)) / (G*C)) / d alpha_t; value (double[1]) += weight * term."""
    G, Cc = alpha_t.shape
    call("attndm_alpha_entropy_grad", ptr(alpha_t), int(G), int(Cc), float(weight), ptr(grad_out), ptr(value), stream())
    return grad_out
