"""Drop-in for models/self_attention.py: EnhancedQSelfAttention.

Four quantized 1x1 projections (QConv2d, per-projection bit widths and group
counts as in :24-30,74-97), the attention core in attndm_attention, and the
gamma-weighted residual.  In NHWC the q/k/v projections ARE the [B,HW,d] matrices
the reference builds with view/permute (:132-134), so no data movement remains.
"""
import torch
import torch.nn as nn

from . import ops
from .attention_quant_utils import MixedPrecisionAttention
from .quant_util import FConv2d, QConv2d


class EnhancedQSelfAttention(nn.Module):
    def __init__(self, in_channels, quantization=False, sequence=None, args=None, mixed_precision=False,
                 bit_config=None):
        super().__init__()
        self.quantization = quantization
        self.in_channels = in_channels
        self.key_channels = in_channels // 8
        self.value_channels = in_channels
        self.heads = 8
        self.temperature = nn.Parameter(torch.ones(1))
        self.mixed_precision = mixed_precision
        if bit_config is None:
            self.bit_config = {
                "query": args.bitwidth if args else 8,
                "key": max(4, args.bitwidth - 2) if args else 6,
                "value": args.bitwidth if args else 8,
                "output": args.bitwidth if args else 8,
            }
        else:
            self.bit_config = bit_config
        if quantization and sequence is not None:
            mk = lambda cin, cout, b: QConv2d(cin, cout, kernel_size=1, w_bit=b, a_bit=b, sequence=sequence, args=args)
        else:                                                    # the FP model (:56-59)
            mk = lambda cin, cout, b: FConv2d(cin, cout, kernel_size=1)
        self.query_conv = mk(in_channels, self.key_channels, self.bit_config["query"])
        self.key_conv = mk(in_channels, self.key_channels, self.bit_config["key"])
        self.value_conv = mk(in_channels, self.value_channels, self.bit_config["value"])
        self.output_conv = mk(self.value_channels, in_channels, self.bit_config["output"])
        if quantization and sequence is not None:
            self.configure_group_quantization()
        self.gamma = nn.Parameter(torch.zeros(1))
        if mixed_precision and quantization:
            self.enable_mixed_precision()
        else:
            self.softmax = nn.Softmax(dim=-1)

    def enable_mixed_precision(self):
        """What `mixed_precision=True` constructs (:64-70)."""
        self.mixed_precision = True
        self.attention_processor = MixedPrecisionAttention(
            head_dim=self.key_channels // self.heads, num_heads=self.heads,
            bit_width=min(self.bit_config.values()), scaling_factor=self.key_channels ** -0.5)
        self.attention_processor.to(self.gamma.device)

    def configure_group_quantization(self):
        """:74-116.  q/k: 8 groups, v: max(2, heads//2) = 4 groups, out: 8; alpha_activ resized to
        match.  Unlike the reference (defect D4) groups_range is resized too."""
        self.query_conv.group_num = self.heads
        self.key_conv.group_num = self.heads
        self.value_conv.group_num = max(2, self.heads // 2)
        self.output_conv.group_num = 8
        for conv in (self.query_conv, self.key_conv, self.value_conv, self.output_conv):
            old = conv.alpha_activ.data
            new = torch.zeros(old.size(0), conv.group_num, conv.in_channels)
            for g in range(conv.group_num):
                new[:, g] = old[:, min(g * old.size(1) // conv.group_num, old.size(1) - 1)]
            conv.alpha_activ = nn.Parameter(new)
            if conv.groups_range.shape[1] != conv.group_num:
                conv.groups_range = nn.Parameter(torch.zeros(conv.len_seq, conv.group_num, 2), requires_grad=False)
            conv.invalidate_cache()

    def forward_fused(self, x, timestep=None):
        """x: NHWC.  :127-151."""
        B, H, W, Cc = x.shape
        N = H * W
        q = self.query_conv.forward_fused(x).view(B, N, self.key_channels)
        k = self.key_conv.forward_fused(x).view(B, N, self.key_channels)
        v = self.value_conv.forward_fused(x).view(B, N, self.value_channels)
        if self.mixed_precision and self.quantization:
            out = self.attention_processor.forward_nhwc(q, k, v, timestep)
        else:
            out = ops.attention(q, k, v, self.key_channels ** -0.5)
        out = self.output_conv.forward_fused(out.view(B, H, W, self.value_channels))
        return ops.scale_add(out, x, self.gamma.detach())

    def forward(self, x, t=None, timestep=None):
        if t is not None:
            raise NotImplementedError("the `t` concat branch (:123-125) is never taken by Model.forward")
        return ops.to_nchw(self.forward_fused(ops.to_nhwc(x), timestep))


def create_enhanced_attention(in_channels, sequence, args):
    """:155-168."""
    return EnhancedQSelfAttention(
        in_channels, quantization=True, sequence=sequence, args=args, mixed_precision=True,
        bit_config={"query": args.bitwidth, "key": max(4, args.bitwidth - 2), "value": args.bitwidth,
                    "output": args.bitwidth})
