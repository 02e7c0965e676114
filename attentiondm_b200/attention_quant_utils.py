"""Drop-in for utils/attention_quant_utils.py: MixedPrecisionAttention and
AttentionCalibrator, with the attention core (QK^T -> [fake-quant] -> softmax ->
[fake-quant] -> .V) running in attndm_attention.

The shipped reference forward is shape-broken (K is permuted to [B,h,HW,d],
attention_quant_utils.py:70, so matmul(q,k) only works when HW == d) and
update_quantization_params assigns Python floats to registered buffers
(:113-118, a TypeError).  This module implements the evident intent: K as
[B,h,d,HW] and the parameters kept as 1-element buffers.  Parity is pinned at
quantize_tensor (:30-38) only; see DESIGN.md.
"""
import torch
import torch.nn as nn

from . import ops


class MixedPrecisionAttention(nn.Module):
    def __init__(self, head_dim, num_heads, bit_width, scaling_factor=None):
        super().__init__()
        self.head_dim = head_dim
        self.num_heads = num_heads
        self.base_bit_width = bit_width
        self.scaling_factor = scaling_factor or (head_dim ** -0.5)
        self.register_buffer('quant_scale_qk', torch.ones(1))
        self.register_buffer('quant_zero_qk', torch.zeros(1))
        self.register_buffer('quant_scale_attn', torch.ones(1))
        self.register_buffer('quant_zero_attn', torch.zeros(1))
        self.timestep_importance = nn.Parameter(torch.zeros(1000))
        self.timestep_importance.data.fill_(0.5)
        self.softmax_scale = nn.Parameter(torch.ones(1))
        self._host = None          # cached host copies of the scalar parameters

    def quantize_tensor(self, x, scale, zero_point, bits):
        """:30-38.  Elementwise fake-quant with scalar parameters (tiny tensors: torch ops)."""
        qmin, qmax = 0, (1 << bits) - 1
        scale = scale.to(x.device)
        zero_point = zero_point.to(x.device)
        x_q = torch.clamp(torch.round(x / scale) + zero_point, qmin, qmax)
        return (x_q - zero_point) * scale

    def get_effective_bits(self, timestep=None):
        """:40-49."""
        if timestep is None:
            return self.base_bit_width
        importance = self.timestep_importance[timestep]
        return self.base_bit_width + 2.0 * torch.sigmoid(importance)

    def refresh_host_params(self):
        self._host = dict(
            scale_qk=float(self.quant_scale_qk.item()), zero_qk=float(self.quant_zero_qk.item()),
            scale_attn=float(self.quant_scale_attn.item()), zero_attn=float(self.quant_zero_attn.item()),
            softmax_scale=float(self.softmax_scale.item()))
        return self._host

    def forward(self, query, key, value, timestep=None):
        """query [B,HW,C'], key [B,C',HW], value [B,HW,C] -> [B,HW,C]   (:51-107)."""
        k_nd = key.permute(0, 2, 1).contiguous()          # the kernel wants [B,HW,C'] rows
        return self.forward_nhwc(query.contiguous(), k_nd, value.contiguous(), timestep)

    def forward_nhwc(self, q, k, v, timestep=None):
        """q,k [B,N,C'], v [B,N,C] (the NHWC projections as they come out of the 1x1 convs)."""
        h = self._host or self.refresh_host_params()
        eff = self.get_effective_bits(timestep)
        eff = float(eff) if not isinstance(eff, (int, float)) else eff
        qk_q = (h["scale_qk"], h["zero_qk"], max(4, int(eff))) if eff <= 6 else None
        p_q = (h["scale_attn"], h["zero_attn"], max(3, int(eff - 1))) if eff <= 4 else None
        return ops.attention(q, k, v, self.scaling_factor, heads=self.num_heads, softmax_scale=h["softmax_scale"],
                             qk_q=qk_q, p_q=p_q)

    def update_quantization_params(self, qk_min, qk_max, attn_min, attn_max):
        """:109-118, keeping the buffers as tensors."""
        qk_range = qk_max - qk_min
        s = qk_range / (2 ** self.base_bit_width - 1)
        self.quant_scale_qk.fill_(float(s))
        self.quant_zero_qk.fill_(float(-qk_min / s))
        self.quant_scale_attn.fill_(1.0 / (2 ** self.base_bit_width - 1))
        self.quant_zero_attn.fill_(0.0)
        self._host = None


class AttentionCalibrator:
    """:121-182.  Collects the range of the attention modules' outputs over a few
    timesteps with forward hooks and updates the attention-internal quantizers."""

    def __init__(self, model, device='cuda'):
        self.model = model
        self.device = device
        self.attention_modules = []
        for module in model.modules():
            if hasattr(module, 'attention_processor') and hasattr(module, 'mixed_precision'):
                if module.mixed_precision and module.quantization:
                    self.attention_modules.append(module)

    def calibrate(self, sample_batch, timesteps=None):
        if not self.attention_modules:
            print("No mixed-precision attention modules found to calibrate")
            return
        if timesteps is None:
            timesteps = [0, 250, 500, 750, 999]
        for t in timesteps:
            t_tensor = torch.tensor([t], device=self.device).repeat(sample_batch.size(0))
            with torch.no_grad():
                qk_mins, qk_maxs = [], []

                def qk_hook(module, input, output):
                    qk_mins.append(output.min().item())
                    qk_maxs.append(output.max().item())

                hooks = [m.register_forward_hook(qk_hook) for m in self.attention_modules]
                _ = self.model(sample_batch, t=t_tensor)
                for hook in hooks:
                    hook.remove()
                for module in self.attention_modules:
                    module.attention_processor.update_quantization_params(min(qk_mins), max(qk_maxs), 0.0, 1.0)
        print(f"Calibrated {len(self.attention_modules)} attention modules")
