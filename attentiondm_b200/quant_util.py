"""Drop-in for the reference's utils/quant_util.py operator surface
(QModule / QConv2d / Quant / GroupWise_Quantizaion / lp_loss / percentile
helpers), executing through the sm_100a kernels behind include/attndm_b200.h.

Same constructor signatures, attributes, state_dict keys ({weight, bias,
groups_range, alpha_activ}) and stateful behaviour (index_seq advance/wrap,
set_calibrate / first_calibrate / set_quantize) as the reference; citations are
to the reference file utils/quant_util.py unless stated otherwise.
There is no CPU path: a non-CUDA input raises.
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.modules.utils import _pair

from . import ops
from .quantization_utils import (AsymmetricQuantFunction, asymmetric_linear_quantization_params)

# Optional hook: (min_c, max_c) -> (min_c, max_c) all-reduced over ranks
# (attentiondm_b200.dist.install()).  SURVEY.md section 8(e).
calib_allreduce = None
# Optional hook: double tensor -> the same tensor summed over ranks (first-calibrate scores, :237-254).
calib_allreduce_sum = None


def lp_loss(pred, tgt, p=2.0, reduction='none'):
    """:37-44."""
    if reduction == 'none':
        return (pred - tgt).abs().pow(p).sum(1).mean()
    return (pred - tgt).abs().pow(p).mean()


class Quant(nn.Module):
    """Scalar-range fake-quant branch (:47-66)."""

    def __init__(self, range_left=-6, range_right=6, dim=128, device=None):
        super().__init__()
        self.range_left = torch.as_tensor([float(range_left)], dtype=torch.float32, device=device)
        self.range_right = torch.as_tensor([float(range_right)], dtype=torch.float32, device=device)

    def forward(self, inputs, a_bit):
        x = ops.to_nhwc(inputs) if inputs.dim() == 4 else inputs.contiguous()
        dev = x.device
        gr = torch.stack([self.range_left.to(dev), self.range_right.to(dev)], dim=1)      # [1,2]
        Cc = x.shape[-1]
        sw = torch.ones(1, Cc, device=dev)
        y = ops.calib_mix(x, gr, sw, a_bit)
        return ops.to_nchw(y) if inputs.dim() == 4 else y


def GroupWise_Quantizaion(x, dim=128, group_n=8, maxmin='max'):
    """:403-437, on device (attndm_group_ranges with the init floor disabled)."""
    x = x.detach().float().contiguous()
    gr = torch.empty(group_n, 2, device=x.device)
    inf = float("inf")
    xq_min, xq_max = ops.group_ranges(x, x, group_n, inf, -inf, gr)
    if maxmin in 'max':
        return xq_max, gr[:, 1].clone()
    return xq_min, gr[:, 0].clone()


def find_scale_by_percentile_min(x, percentile=0.9999):
    """:440-444, without the host round trip of the whole tensor."""
    n = x.numel()
    return ops.kth_value(x.detach().float().contiguous(), int(n * (1 - percentile))).item()


def find_scale_by_percentile_max(x, percentile=0.9999):
    """:446-450."""
    n = x.numel()
    return ops.kth_value(x.detach().float().contiguous(), int(n * percentile)).item()


class QModule(nn.Module):
    """:70-337."""

    def __init__(self, in_channels=128, out_channels=128, w_bit=8, a_bit=8, half_wave=False, sequence=None,
                 args=None):
        super().__init__()
        self._a_bit = a_bit
        self._w_bit = w_bit
        self._b_bit = 32
        self._half_wave = half_wave
        self.sequence = list(reversed(sequence))
        self.len_seq = len(self.sequence)
        self.index_seq = 0
        self.args = args
        self.init_range_min = -4.0 * torch.ones(self.len_seq)
        self.init_range_max = 6.0 * torch.ones(self.len_seq)
        self.group_num = 8
        self.groups_range = nn.Parameter(torch.zeros([self.len_seq, self.group_num, 2]), requires_grad=False)
        self._quantized = True
        self._tanh_weight = False
        self._fix_weight = False
        self._trainable_activation_range = True
        self._calibrate = False
        self._first_calibrate = False
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.weight_function = AsymmetricQuantFunction.apply
        self.act_function = AsymmetricQuantFunction.apply
        self.activation_range_min1 = torch.zeros(self.len_seq, self.in_channels)
        self.activation_range_max1 = torch.zeros(self.len_seq, self.in_channels)
        self.weight_range_min = torch.zeros(self.out_channels)
        self.weight_range_max = torch.zeros(self.out_channels)
        self.alpha_activ = nn.Parameter(torch.Tensor(self.len_seq, self.group_num, in_channels), requires_grad=True)
        self.alpha_activ.data.fill_(0.01)
        # -- B200 side state (not part of the reference surface) --
        self._tab = None            # per-step tables, see _tables()
        self._tab_key = None
        self._pack = None           # packed weights, see _packed()
        self._pack_center = None    # centre-tap pack of a 3x3 conv (1x1 feature maps)
        self._pack_key = None
        self._state_version = 0     # bumped when a kernel writes groups_range in place
        self._staged = None         # engine-provided "current step" table row (CUDA-graph mode)
        self.force_f32 = False      # force the fp32 conv path (tests / debugging)

    # ---- reference property / setter surface (:123-184) ----
    @property
    def w_bit(self):
        return self._w_bit

    @w_bit.setter
    def w_bit(self, w_bit):
        self._w_bit = w_bit
        self.invalidate_cache()

    @property
    def a_bit(self):
        return self._a_bit

    @a_bit.setter
    def a_bit(self, a_bit):
        self._a_bit = a_bit
        self.invalidate_cache()

    @property
    def b_bit(self):
        return self._b_bit

    @property
    def half_wave(self):
        return self._half_wave

    @property
    def quantized(self):
        return self._quantized

    @property
    def tanh_weight(self):
        return self._tanh_weight

    def set_quantize(self, quantized):
        self._quantized = quantized

    def set_fix_weight(self, fix_weight):
        self._fix_weight = fix_weight

    def set_calibrate(self, calibrate=True):
        self._calibrate = calibrate
        self.invalidate_cache(weights=False)

    def first_calibrate(self, calibrate=True):
        self._first_calibrate = calibrate

    def set_tanh(self, tanh=True):
        self._tanh_weight = tanh

    # ---- cache control ----
    def invalidate_cache(self, weights=True):
        """Call after editing groups_range / alpha_activ / weight / weight_range_* through `.data`."""
        self._tab = None
        self._tab_key = None
        if weights:
            self._pack = None
            self._pack_key = None

    def _load_from_state_dict(self, *a, **k):
        super()._load_from_state_dict(*a, **k)
        self.invalidate_cache()

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self.invalidate_cache()
        return r

    # ---- weights (D3 / H1 helpers) ----
    def init_weight_range(self):
        """weight_range_min/max := per-out-channel min/max, which makes the reference clamp
        (:284-303) the identity.  The reference leaves them at zero (so every weight clamps to 0)."""
        flat = self.weight.detach().reshape(self.out_channels, -1)
        self.weight_range_min = flat.min(1)[0].clone()
        self.weight_range_max = flat.max(1)[0].clone()
        self.invalidate_cache()

    def snap_weights_(self):
        """Snap weight onto the w_bit per-out-channel grid with the reference's own
        AsymmetricQuantFunction (utils/quantization_utils/quant_utils.py:136-162), then
        init_weight_range().  This is what makes the int8 x int8 tensor-core path exact."""
        w = self.weight.data
        flat = w.reshape(w.shape[0], -1)
        self.weight.data = AsymmetricQuantFunction.apply(w, self._w_bit, flat.min(1)[0], flat.max(1)[0]).detach()
        self.init_weight_range()

    def _packed(self):
        w = self.weight
        lo = self.weight_range_min
        hi = self.weight_range_max
        if lo.device != w.device:
            lo = self.weight_range_min = lo.to(w.device)
        if hi.device != w.device:
            hi = self.weight_range_max = hi.to(w.device)
        key = (w.data_ptr(), w._version, lo.data_ptr(), lo._version, hi.data_ptr(), hi._version, self._w_bit)
        if self._pack is None or self._pack_key != key:
            w_eff = ops.weight_clamp_pack(w.detach(), lo, hi)
            i8 = ops.weight_to_i8(w_eff, self._w_bit)
            grid = (i8.w_scale, i8.w_zp.float())     # the recovered grid (zero points as finally used)
            self._pack = (w_eff, i8)
            self._pack_center = None
            if w_eff.shape[1] == 9:
                # On a 1x1 feature map a 3x3/pad-1 conv only ever sees its centre tap (the other eight
                # multiply zero padding), so it is exactly the 1x1 conv with w[:, :, 1, 1] on the same grid.
                wc = w_eff[:, 4:5, :].contiguous()
                self._pack_center = (wc, ops.weight_to_i8(wc, self._w_bit, grid))
            self._pack_key = key
            self._tab = None
        return self._pack

    # ---- per-step tables: scale / zero-point of :260-271, for every index_seq at once ----
    def table_layout(self):
        Cq = (self.in_channels + 3) // 4 * 4
        Oq = (self.out_channels + 3) // 4 * 4
        return dict(scale=0, zp=Cq, mult=2 * Cq, act_zp=2 * Cq + Oq, width=2 * Cq + Oq + 4)

    def _tables(self):
        gr, al = self.groups_range, self.alpha_activ
        w_eff, i8 = self._packed()           # first: an in-place weight update must invalidate the tables too
        key = (gr.data_ptr(), gr._version, al.data_ptr(), al._version, self._a_bit, self._state_version,
               self._pack_key)
        if self._tab is not None and self._tab_key == key:
            return self._tab
        T, G, Cc, O = self.len_seq, self.group_num, self.in_channels, self.out_channels
        with torch.no_grad():
            sw = F.softmax(al.detach().float(), dim=1)                   # [T,G,C]
            grd = gr.detach().float()
            lo = 0
            hi = 0
            for g in range(G):                                           # this accumulation order (:263-267)
                lo = lo + grd[:, g, 0:1] * sw[:, g]
                hi = hi + grd[:, g, 1:2] * sw[:, g]
            scale, zp = asymmetric_linear_quantization_params(self._a_bit, lo, hi)   # [T,C]
            sc, zc = scale.cpu(), zp.cpu()
            n_half = 2 ** (self._a_bit - 1)
            uniform, zero_ok = [], []
            for t in range(T):
                u = bool(torch.isfinite(sc[t]).all() and (sc[t] == sc[t, 0]).all() and (zc[t] == zc[t, 0]).all()
                         and torch.isfinite(zc[t]).all())
                uniform.append(u)
                zero_ok.append(u and (-n_half <= -float(zc[t, 0]) <= n_half - 1))
            lay = self.table_layout()
            tab = torch.zeros(T, lay["width"], dtype=torch.float32, device=gr.device)
            tab[:, lay["scale"]:lay["scale"] + Cc] = scale
            tab[:, lay["zp"]:lay["zp"] + Cc] = zp
            if i8.on_grid:
                mult = 1.0 / (scale[:, 0:1].double() * i8.w_scale.double()[None, :])
                tab[:, lay["mult"]:lay["mult"] + O] = torch.nan_to_num(mult.float(), nan=0.0, posinf=0.0, neginf=0.0)
            azp = torch.nan_to_num(zp[:, 0], nan=0.0, posinf=0.0, neginf=0.0).clamp(-1e6, 1e6).to(torch.int32)
            tab[:, lay["act_zp"]] = azp.view(torch.float32)
        self._tab = dict(tab=tab, lay=lay, uniform=uniform, zero_ok=zero_ok,
                         i8_ok=[i8.on_grid and z for z in zero_ok])
        self._tab_key = key
        return self._tab

    def int8_status(self) -> dict:
        """Why (not) the integer path: weights on the w_bit grid, per-step scale uniform over channels,
        and 0.0 representable (the halo ring needs the code of zero)."""
        tb = self._tables()
        return dict(on_grid=self._pack[1].on_grid, uniform=all(tb["uniform"]), zero_ok=all(tb["zero_ok"]),
                    forced_f32=self.force_f32)

    def int8_ok_all_steps(self) -> bool:
        return (not self.force_f32) and all(self._tables()["i8_ok"])

    # ---- calibration branch (:186-224, :235-258) ----
    def _calib_minmax(self, x):
        """Per-channel (min, max) over (B, H, W) of the GLOBAL batch (:187-191): all-reduced over ranks when
        attentiondm_b200.dist.install() is active."""
        min_c, max_c = ops.minmax_c(x)
        if calib_allreduce is not None:
            min_c, max_c = calib_allreduce(min_c, max_c)
        return min_c, max_c

    def calibrate_quantization(self, x, init_min, init_max, want_lp=False, minmax=None):
        """x: NHWC fp32 (already through any producer op). Returns the G-branch mix."""
        t = self.index_seq
        dev = x.device
        min_c, max_c = self._calib_minmax(x) if minmax is None else minmax
        gr_t = self.groups_range.data[t]
        xq_min, xq_max = ops.group_ranges(min_c, max_c, self.group_num, float(init_min), float(init_max), gr_t)
        self._state_version += 1
        if self.activation_range_min1.device != dev:
            self.activation_range_min1 = self.activation_range_min1.to(dev)
            self.activation_range_max1 = self.activation_range_max1.to(dev)
        self.activation_range_min1[t] = xq_min
        self.activation_range_max1[t] = xq_max
        self.activation_range_min = self.activation_range_min1
        self.activation_range_max = self.activation_range_max1
        with torch.no_grad():
            sw = F.softmax(self.alpha_activ.detach()[t].float(), dim=0)
        self.sw = sw
        if want_lp:
            y, lp = ops.calib_mix(x, gr_t, sw, self._a_bit, lp_p=0.5)
            return y, lp
        return ops.calib_mix(x, gr_t, sw, self._a_bit)

    def _calibrate_step(self, x):
        t = self.index_seq
        minmax = self._calib_minmax(x)        # the statistics do not depend on the candidate range: reduce once
        if self._first_calibrate:
            cand, lps = [], []
            for aa in range(9):
                new_max = self.init_range_max[t] * (1.0 - (aa * 0.1))
                new_min = self.init_range_min[t] * (1.0 - (aa * 0.1))
                _, lp = self.calibrate_quantization(x, new_min, new_max, want_lp=True, minmax=minmax)
                cand.append((new_min, new_max))
                lps.append(lp)
            # lp_loss(p=0.5, reduction='all') is a mean over the GLOBAL batch: SUM the nine lp sums and the
            # element count over ranks, so every replica scores -- and therefore chooses -- identically
            # (SURVEY.md section 8e).  One host read for all nine candidates.
            tot = torch.cat(lps + [torch.full((1,), float(x.numel()), dtype=torch.float64, device=x.device)])
            if calib_allreduce_sum is not None:
                tot = calib_allreduce_sum(tot)
            tot = tot.cpu()
            best_score = 1e+10
            best_max = self.init_range_max[t]
            best_min = self.init_range_min[t]
            for aa in range(9):
                score = float(np.float32(float(tot[aa]) / float(tot[9])))
                if score < best_score:
                    (best_min, best_max), best_score = cand[aa], score
            if best_score < 0.2:
                self.init_range_max[t] = best_max
                self.init_range_min[t] = best_min
        return self.calibrate_quantization(x, self.init_range_min[t], self.init_range_max[t], minmax=minmax)

    def forward(self, *inputs):
        raise NotImplementedError

    def extra_repr(self):
        return 'w_bit={}, a_bit={}, half_wave={}, tanh_weight={}'.format(
            self.w_bit if self.w_bit > 0 else -1, self.a_bit if self.a_bit > 0 else -1, self.half_wave,
            self._tanh_weight)


class QConv2d(QModule):
    """:351-401.  forward(x) takes / returns logical NCHW like the reference;
    forward_fused is the NHWC entry the blocks use to fuse the producer
    (GroupNorm+SiLU / SiLU) and the consumer adds (residual, time embedding)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 w_bit=8, a_bit=8, half_wave=False, sequence=None, args=None):
        super().__init__(in_channels=in_channels, out_channels=out_channels, w_bit=w_bit, a_bit=a_bit,
                         half_wave=half_wave, sequence=sequence, args=args)
        if in_channels % groups != 0:
            raise ValueError('in_channels must be divisible by groups')
        if out_channels % groups != 0:
            raise ValueError('out_channels must be divisible by groups')
        self.kernel_size = _pair(kernel_size)
        self.stride = _pair(stride)
        self.padding = _pair(padding)
        self.dilation = _pair(dilation)
        self.groups = groups
        self.weight = nn.Parameter(torch.zeros(out_channels, in_channels // groups, *self.kernel_size))
        if bias:
            self.bias = nn.Parameter(torch.zeros(out_channels))
        else:
            self.register_parameter('bias', None)
        self.reset_parameters()
        k = self.kernel_size
        ok = (k == (3, 3) and self.padding == (1, 1)) or (k == (1, 1) and self.padding == (0, 0))
        if not ok or self.stride != (1, 1) or self.dilation != (1, 1) or groups != 1:
            raise NotImplementedError(
                "attentiondm_b200.QConv2d: the hot path only has 3x3/s1/p1 and 1x1/s1/p0, groups=1 "
                "(every QConv2d models/diffusion.py and models/self_attention.py construct)")

    def reset_parameters(self):
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in, _ = nn.init._calculate_fan_in_and_fan_out(self.weight)
            nn.init.uniform_(self.bias, -1 / math.sqrt(fan_in), 1 / math.sqrt(fan_in))

    @property
    def taps(self):
        return self.kernel_size[0] * self.kernel_size[1]

    def use_staged_row(self, row: Optional[torch.Tensor]):
        """Engine hook: `row` is this layer's slice of the staged current-step table."""
        self._staged = row

    def quant_request(self, H: int, W: int):
        """(scale, zp, a_bit, halo) this layer's next forward_fused will quantize its input with, or None when that
        call does not take the integer path (calibration, fp32 fallback).  Does not advance index_seq."""
        if self._calibrate:
            return None
        t = 0 if self.index_seq >= self.args.timesteps else self.index_seq
        tb = self._tables()
        lay = tb["lay"]
        if self._staged is not None:
            row, use_i8 = self._staged, (not self.force_f32) and all(tb["i8_ok"])
        else:
            row, use_i8 = tb["tab"][t], (not self.force_f32) and tb["i8_ok"][t]
        if not use_i8:
            return None
        taps = 1 if (self.taps == 9 and H == 1 and W == 1) else self.taps
        return row[lay["scale"]:], row[lay["zp"]:], self._a_bit, taps == 9

    def forward_fused(self, x, pre=ops.PRE_NONE, gn: Optional[ops.GnArgs] = None, residual=None, temb=None,
                      want_stats=False):
        """x: NHWC fp32 CUDA.  residual: NHWC like the output.  temb: [B, O].  want_stats: the output feeds a
        GroupNorm -- on the integer path its statistics come out of the conv's epilogue, attached to the result as
        `out._gn_stats` (double [B, 32, 2]) for diffusion._gn_args."""
        if self.index_seq >= self.args.timesteps:          # :228-229
            self.index_seq = 0
        t = self.index_seq
        B, H, W, Cc = x.shape
        if Cc != self.in_channels:
            raise RuntimeError(f"QConv2d: expected {self.in_channels} input channels, got {Cc}")
        w_eff, i8 = self._packed()
        taps = self.taps
        if taps == 9 and H == 1 and W == 1:
            w_eff, i8 = self._pack_center
            taps = 1
        bias = self.bias.detach() if self.bias is not None else None
        if self._calibrate:
            if isinstance(x, ops.CatView):
                x = x.materialize()
            if pre == ops.PRE_GN_SILU:
                xa = ops.gn_silu(x, gn)
            elif pre == ops.PRE_SILU:
                xa = ops.silu(x)
            else:
                xa = x
            y = self._calibrate_step(xa)
            out = ops.conv_f32(y, w_eff, bias, residual, temb)
            self.index_seq += 1
            return out
        tb = self._tables()
        lay = tb["lay"]
        if self._staged is not None:
            row = self._staged
            use_i8 = (not self.force_f32) and all(tb["i8_ok"])
        else:
            row = tb["tab"][t]
            use_i8 = (not self.force_f32) and tb["i8_ok"][t]
        scale = row[lay["scale"]:]
        zp = row[lay["zp"]:]
        if use_i8:
            codes, rowsum, _ = ops.act_quant(x, scale, zp, self._a_bit, pre, gn, want_codes=True, halo=(taps == 9))
            stats = None
            if (want_stats and ops.conv_stats_enabled() and self.out_channels % ops.GN_GROUPS == 0
                    and not ops.gn_fits_fused(H, W, self.out_channels)):
                stats = ops.gn_stats_buffer(B, x.device)
            out = ops.qconv_i8(codes, rowsum, B, H, W, Cc, i8, taps, row[lay["mult"]:], row[lay["act_zp"]:],
                               bias, residual, temb, gn_stats_out=stats)
            if stats is not None:
                out._gn_stats = stats
        else:
            _, _, y = ops.act_quant(x, scale, zp, self._a_bit, pre, gn, want_codes=False, want_f32=True)
            out = ops.conv_f32(y, w_eff, bias, residual, temb)
        self.index_seq += 1                                 # :281
        return out

    def forward(self, inputs):
        """:383-385."""
        return ops.to_nchw(self.forward_fused(ops.to_nhwc(inputs)))

    def quantize_activation_codes(self, inputs):
        """Test/debug surface: integer codes and fake-quant output of :260-282 for the
        current index_seq WITHOUT advancing it.  inputs: logical NCHW."""
        x = ops.to_nhwc(inputs)
        t = 0 if self.index_seq >= self.args.timesteps else self.index_seq
        tb = self._tables()
        row, lay = tb["tab"][t], tb["lay"]
        codes, rowsum, y = ops.act_quant(x, row[lay["scale"]:], row[lay["zp"]:], self._a_bit, want_codes=True,
                                         halo=False, want_f32=True)
        B, H, W, Cc = x.shape
        codes = codes[:, :Cc].reshape(B, H, W, Cc).permute(0, 3, 1, 2)
        return codes, ops.to_nchw(y), rowsum.reshape(B, H, W)

    def extra_repr(self):
        return (f'{self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}, stride={self.stride}, '
                f'padding={self.padding}, w_bit={self.w_bit}, a_bit={self.a_bit}')


class FConv2d(nn.Conv2d):
    """The un-quantized conv of the reference's FP model (`quantization=False`: every QConv2d site is a plain
    nn.Conv2d, models/diffusion.py:106-116,166-169,281-285,341-345; models/self_attention.py:56-59), which
    `Diffusion.generate_calibrate_set` samples from (runners/diffusion.py:198-216).  Same parameters / state_dict keys
    as nn.Conv2d; the arithmetic runs through the fp32 conv kernel behind the C-ABI (batch-invariant summation
    order), with the same fused producer (GroupNorm+SiLU / SiLU) and consumer adds as QConv2d.forward_fused."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, **kw):
        super().__init__(in_channels, out_channels, kernel_size, stride=stride, padding=padding, **kw)
        k = self.kernel_size
        ok = (k == (3, 3) and self.padding == (1, 1)) or (k == (1, 1) and self.padding == (0, 0))
        if not ok or self.stride != (1, 1) or self.dilation != (1, 1) or self.groups != 1:
            raise NotImplementedError("attentiondm_b200.FConv2d: 3x3/s1/p1 and 1x1/s1/p0, groups=1 only")
        self._wp = None
        self._wp_key = None

    def _w_pack(self):
        """weight [O,C,kh,kw] -> [O, taps, C] (the fp32 kernel's layout) and the centre tap for 1x1 maps."""
        w = self.weight
        key = (w.data_ptr(), w._version)
        if self._wp is None or self._wp_key != key:
            O, Cc, KH, KW = w.shape
            full = w.detach().float().permute(0, 2, 3, 1).reshape(O, KH * KW, Cc).contiguous()
            centre = full[:, 4:5, :].contiguous() if KH * KW == 9 else None
            self._wp, self._wp_key = (full, centre), key
        return self._wp

    def forward_fused(self, x, pre=ops.PRE_NONE, gn: Optional[ops.GnArgs] = None, residual=None, temb=None,
                      want_stats=False):
        if isinstance(x, ops.CatView):
            x = x.materialize()
        B, H, W, Cc = x.shape
        if Cc != self.in_channels:
            raise RuntimeError(f"FConv2d: expected {self.in_channels} input channels, got {Cc}")
        full, centre = self._w_pack()
        w = centre if (centre is not None and H == 1 and W == 1) else full
        if pre == ops.PRE_GN_SILU:
            xa = ops.gn_silu(x, gn)
        elif pre == ops.PRE_SILU:
            xa = ops.silu(x)
        else:
            xa = x
        return ops.conv_f32(xa, w, self.bias.detach() if self.bias is not None else None, residual, temb)

    def forward(self, inputs):
        return ops.to_nchw(self.forward_fused(ops.to_nhwc(inputs)))
