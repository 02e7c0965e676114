"""Drop-in for models/diffusion.py: the fake-quantized UNet (`Model`) and its
blocks, same constructor arguments, module names and state_dict keys.

Internally activations are NHWC fp32 and each block drives the fused kernels:
GroupNorm statistics -> (GroupNorm+SiLU+quantize) -> int8 conv with the residual
and time-embedding adds in its epilogue.  `forward(x, t)` keeps the reference
signature (x logical NCHW, t float [B]).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .quant_util import FConv2d, QConv2d
from .self_attention import EnhancedQSelfAttention


def get_timestep_embedding(timesteps, embedding_dim):
    """models/diffusion.py:11-29."""
    assert len(timesteps.shape) == 1
    return ops.timestep_embedding(timesteps, embedding_dim)


def _conv(quantization, sequence, args, cin, cout, k):
    """A QConv2d when `quantization and sequence is not None`, else the FP model's plain conv
    (models/diffusion.py:90-116 and every other such switch in that file)."""
    pad = 1 if k == 3 else 0
    if quantization and sequence is not None:
        return QConv2d(cin, cout, kernel_size=k, stride=1, padding=pad, w_bit=args.bitwidth, a_bit=args.bitwidth,
                       sequence=sequence, args=args)
    return FConv2d(cin, cout, kernel_size=k, stride=1, padding=pad)


def _gn_args(norm: nn.GroupNorm, x):
    """GroupNorm arguments for the consumer conv.  Small feature maps use the fused per-sample kernel
    (statistics computed in-kernel, stats=None); the rest get a statistics pass."""
    _, H, W, C = x.shape
    if ops.gn_fits_fused(H, W, C):
        stats = None
    else:
        stats = getattr(x, "_gn_stats", None)                # left by the conv that produced x (QConv2d.forward_fused)
        if stats is None:
            stats = ops.gn_stats(x)
    return ops.GnArgs(stats=stats, gamma=norm.weight.detach(), beta=norm.bias.detach(), eps=norm.eps)


class ResidualBlock(nn.Module):
    """models/diffusion.py:82-136."""

    def __init__(self, in_channels, out_channels=None, conv_shortcut=False, dropout=0.1, quantization=False,
                 sequence=None, args=None):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = in_channels if out_channels is None else out_channels
        out_channels = self.out_channels
        self.use_conv_shortcut = conv_shortcut
        mk = lambda cin, cout, k: _conv(quantization, sequence, args, cin, cout, k)
        self.norm1 = nn.GroupNorm(num_groups=32, num_channels=in_channels, eps=1e-6)
        self.conv1 = mk(in_channels, out_channels, 3)
        self.norm2 = nn.GroupNorm(num_groups=32, num_channels=out_channels, eps=1e-6)
        self.dropout = nn.Dropout(dropout)
        self.conv2 = mk(out_channels, out_channels, 3)
        if self.in_channels != self.out_channels:
            if self.use_conv_shortcut:
                self.conv_shortcut = mk(in_channels, out_channels, 3)
            else:
                self.nin_shortcut = mk(in_channels, out_channels, 1)

    def forward_fused(self, x, temb=None, out_stats=False):
        """x NHWC.  temb [B, out_channels] (optional) is the block's `x + time_mlp(t_emb)`
        (models/diffusion.py:175-177), applied after the residual add, fused in conv2's epilogue.
        out_stats: the block's output goes straight into a GroupNorm (conv2 then also emits its statistics)."""
        if self.training and self.dropout.p > 0:
            raise NotImplementedError("attentiondm_b200 runs the eval-mode path (model.eval()); dropout is identity")
        gn1 = _gn_args(self.norm1, x)
        if isinstance(x, ops.CatView) and self.in_channels != self.out_channels:
            # UpBlock.res1: conv1 and the shortcut conv quantize the same concat -- one pass over it for both
            sc_conv = self.conv_shortcut if self.use_conv_shortcut else self.nin_shortcut
            if isinstance(self.conv1, QConv2d) and isinstance(sc_conv, QConv2d):
                ra, rb = self.conv1.quant_request(x.shape[1], x.shape[2]), sc_conv.quant_request(x.shape[1], x.shape[2])
                if ra is not None and rb is not None:
                    x.prepare_pair(ra, rb, gn1)
        h = self.conv1.forward_fused(x, ops.PRE_GN_SILU, gn1, want_stats=True)
        gn2 = _gn_args(self.norm2, h)
        if self.in_channels != self.out_channels:
            sc = (self.conv_shortcut if self.use_conv_shortcut else self.nin_shortcut).forward_fused(x)
        else:
            sc = x.materialize() if isinstance(x, ops.CatView) else x
        return self.conv2.forward_fused(h, ops.PRE_GN_SILU, gn2, residual=sc, temb=temb, want_stats=out_stats)

    def forward(self, x):
        return ops.to_nchw(self.forward_fused(ops.to_nhwc(x)))


def _time_mlp(time_emb_dim, out_channels, quantization, sequence, args):
    return nn.Sequential(nn.SiLU(), _conv(quantization, sequence, args, time_emb_dim, out_channels, 1))


class _TimeBlock(nn.Module):
    def _temb(self, time_emb):
        """SiLU -> QConv2d 1x1 on [B,1,1,temb] -> [B, out] (models/diffusion.py:157-161)."""
        pre = getattr(self, "_temb_fused", None)        # written by the fused time_mlp launch (rowprog.py)
        if pre is not None:
            return pre
        if self.time_mlp is None or time_emb is None:
            return None
        y = self.time_mlp[1].forward_fused(time_emb, ops.PRE_SILU)
        return y.view(y.shape[0], -1)

    def _tail(self, x, time_emb):
        x = self.res1.forward_fused(x, self._temb(time_emb), out_stats=True)     # res2.norm1 follows
        x = self.res2.forward_fused(x, out_stats=getattr(self, "_out_feeds_gn", False))
        if isinstance(self.attn, EnhancedQSelfAttention):
            x = self.attn.forward_fused(x)
        return x


class DownBlock(_TimeBlock):
    """models/diffusion.py:139-190."""

    def __init__(self, in_channels, out_channels, time_emb_dim=None, dropout=0.1, quantization=False, sequence=None,
                 args=None, use_attention=True):
        super().__init__()
        self.maxpool = nn.MaxPool2d(2)
        kw = dict(dropout=dropout, quantization=quantization, sequence=sequence, args=args)
        self.res1 = ResidualBlock(in_channels, out_channels, **kw)
        self.res2 = ResidualBlock(out_channels, out_channels, **kw)
        self.attn = (EnhancedQSelfAttention(out_channels, quantization=quantization, sequence=sequence, args=args)
                     if use_attention else nn.Identity())
        self.time_mlp = _time_mlp(time_emb_dim, out_channels, quantization, sequence, args) if time_emb_dim is not None else None

    def forward_fused(self, x, time_emb=None):
        if not (x.shape[1] <= 1 or x.shape[2] <= 1):       # :172 (NHWC: dims 1,2 are H,W)
            x = ops.maxpool2(x)
        return self._tail(x, time_emb)

    def forward(self, x, time_emb=None):
        te = ops.to_nhwc(time_emb) if time_emb is not None else None
        return ops.to_nchw(self.forward_fused(ops.to_nhwc(x), te))


class UpBlock(_TimeBlock):
    """models/diffusion.py:193-252."""

    def __init__(self, in_channels, out_channels, time_emb_dim=None, dropout=0.1, quantization=False, sequence=None,
                 args=None, use_attention=True):
        super().__init__()
        self.upsample = nn.Upsample(scale_factor=2, mode='nearest')
        kw = dict(dropout=dropout, quantization=quantization, sequence=sequence, args=args)
        self.res1 = ResidualBlock(in_channels + out_channels, out_channels, **kw)
        self.res2 = ResidualBlock(out_channels, out_channels, **kw)
        self.attn = (EnhancedQSelfAttention(out_channels, quantization=quantization, sequence=sequence, args=args)
                     if use_attention else nn.Identity())
        self.time_mlp = _time_mlp(time_emb_dim, out_channels, quantization, sequence, args) if time_emb_dim is not None else None

    def forward_fused(self, x, skip_x, time_emb=None):
        expected = self.res1.in_channels
        if x.shape[-1] + skip_x.shape[-1] == expected and ops.CatView.fits(x, skip_x):
            # the concat feeds only res1's GroupNorm statistics and its two quantizers: they read x and skip_x in place
            return self._tail(ops.CatView(x, skip_x), time_emb)
        combined = ops.upsample_concat(x, skip_x)            # :225-229 + cat
        actual = combined.shape[-1]
        if actual != expected:
            if not hasattr(self, 'channel_proj'):             # lazily created, fresh RNG (:238-241)
                self.channel_proj = nn.Conv2d(actual, expected, kernel_size=1, stride=1, padding=0).to(x.device)
            # un-quantized fp32 1x1 conv.  Our own fp32 kernel rather than a cuBLAS GEMM: its per-output
            # summation order does not depend on the batch size, which keeps every sample's result
            # independent of how the batch is sharded over GPUs.
            # On spatial maps it runs on the tensor cores with fp32-level accuracy (3xTF32 operand splitting, also
            # batch-independent); 1x1 maps keep the SIMT kernel, whose order the fused programs reproduce.
            Bc, Hc, Wc, _ = combined.shape
            bias = self.channel_proj.bias.detach()
            if Hc * Wc > 1 and ops.conv_f32_tc_fits(Bc * Hc * Wc, actual, expected):
                wt = self.channel_proj.weight
                key = (wt.data_ptr(), wt._version)
                if getattr(self, "_proj_split_key", None) != key:
                    self._proj_split = ops.split_tf32(wt.detach().view(expected, actual).contiguous())
                    self._proj_split_key = key
                combined = ops.conv1x1_f32_tc(combined, self._proj_split, bias)
            else:
                w = self.channel_proj.weight.detach().view(expected, 1, actual)
                combined = ops.conv_f32(combined, w.contiguous(), bias)
        return self._tail(combined, time_emb)

    def forward(self, x, skip_x, time_emb=None):
        te = ops.to_nhwc(time_emb) if time_emb is not None else None
        return ops.to_nchw(self.forward_fused(ops.to_nhwc(x), ops.to_nhwc(skip_x), te))


class Model(nn.Module):
    """models/diffusion.py:255-382."""

    def __init__(self, config, quantization=False, sequence=None, args=None):
        super().__init__()
        self.config = config
        self.quantization = quantization
        self.sequence = sequence
        self.args = args
        if not hasattr(config.model, 'time_embed_dim'):
            config.model.time_embed_dim = 256
        if not hasattr(config.model, 'attention_resolutions'):
            config.model.attention_resolutions = 1
        ted = config.model.time_embed_dim
        self.time_embed = nn.Sequential(nn.Linear(ted, ted * 4), nn.SiLU(), nn.Linear(ted * 4, ted * 4))
        ch = config.model.ch
        self.init_conv = _conv(quantization, sequence, args, config.data.channels, ch, 3)
        kw = dict(time_emb_dim=ted * 4, dropout=config.model.dropout, quantization=quantization, sequence=sequence,
                  args=args)
        self.down_blocks = nn.ModuleList()
        ch_mult = config.model.ch_mult
        now_ch = ch
        for i, mult in enumerate(ch_mult):
            out_ch = ch * mult
            for _ in range(config.model.num_res_blocks):
                self.down_blocks.append(DownBlock(now_ch, out_ch, use_attention=(i >= config.model.attention_resolutions), **kw))
                now_ch = out_ch
            if i < len(ch_mult) - 1:
                self.down_blocks.append(DownBlock(now_ch, now_ch, use_attention=False, **kw))
        rk = dict(dropout=config.model.dropout, quantization=quantization, sequence=sequence, args=args)
        self.middle_block1 = ResidualBlock(now_ch, now_ch, **rk)
        self.middle_attn = EnhancedQSelfAttention(now_ch, quantization=quantization, sequence=sequence, args=args)
        self.middle_block2 = ResidualBlock(now_ch, now_ch, **rk)
        self.up_blocks = nn.ModuleList()
        for i, mult in reversed(list(enumerate(ch_mult))):
            out_ch = ch * mult
            for j in range(config.model.num_res_blocks + 1):
                blk_in = now_ch + ch * mult if j == 0 else now_ch
                self.up_blocks.append(UpBlock(blk_in, out_ch, use_attention=(i >= config.model.attention_resolutions), **kw))
                now_ch = out_ch
        self.norm_out = nn.GroupNorm(num_groups=32, num_channels=now_ch, eps=1e-6)
        self.conv_out = _conv(quantization, sequence, args, now_ch, config.data.channels, 3)
        if not isinstance(self.up_blocks[-1].attn, EnhancedQSelfAttention):
            self.up_blocks[-1]._out_feeds_gn = True            # its output goes straight into norm_out

    # ---- helpers over all quantized layers ----
    def qconvs(self):
        return [(n, m) for n, m in self.named_modules() if isinstance(m, QConv2d)]

    def set_calibrate(self, flag=True, first=False):
        for _, m in self.qconvs():
            m.set_calibrate(flag)
            m.first_calibrate(first)

    def reset_index_seq(self, value=0):
        for _, m in self.qconvs():
            m.index_seq = value

    def init_weight_ranges(self):
        for _, m in self.qconvs():
            m.init_weight_range()

    def snap_weights_(self):
        for _, m in self.qconvs():
            m.snap_weights_()

    def materialize_lazy_layers(self):
        """Create the `channel_proj` 1x1 convs that UpBlock.forward would create lazily on its first
        call (models/diffusion.py:235-242), so that a full state_dict can be loaded before any forward.
        Channel bookkeeping follows Model.forward: skips = [init_conv] + every down block."""
        skips = [self.init_conv.out_channels] + [b.res2.out_channels for b in self.down_blocks]
        now = self.middle_block2.out_channels
        dev = self.init_conv.weight.device
        for blk in self.up_blocks:
            sk = skips.pop() if skips else now
            actual, expected = now + sk, blk.res1.in_channels
            if actual != expected and not hasattr(blk, 'channel_proj'):
                blk.channel_proj = nn.Conv2d(actual, expected, kernel_size=1, stride=1, padding=0).to(dev)
            now = blk.res2.out_channels
        return self

    def forward_nhwc(self, x, t):
        """x NHWC [B,H,W,C]; t float [B].  models/diffusion.py:347-382."""
        B = x.shape[0]
        if not hasattr(self, "_n_gn"):
            self._n_gn = sum(1 for mm in self.modules() if isinstance(mm, nn.GroupNorm))
        ops.gn_pool_begin(self._n_gn, B, x.device)
        try:
            return self._forward_nhwc(x, t)
        finally:
            ops.gn_pool_end()

    def _linear_f32(self, x, lin):
        """An un-quantized fp32 Linear on [B, 1, 1, C]: the 3xTF32 tensor-core GEMM when the shape fits (the
        1024-wide time_embed layers are 64 CTAs of pure latency on the FP32 pipe), else the fp32 SIMT kernel."""
        B = x.shape[0]
        O, Cc = lin.weight.shape
        if ops.conv_f32_tc_fits(B, Cc, O):
            key = (lin.weight.data_ptr(), lin.weight._version)
            cache = self.__dict__.setdefault("_lin_split", {})
            if cache.get(id(lin), (None,))[0] != key:
                cache[id(lin)] = (key, ops.split_tf32(lin.weight.detach().contiguous()))
            return ops.conv1x1_f32_tc(x, cache[id(lin)][1], lin.bias.detach())
        return ops.conv_f32(x, lin.weight.detach().unsqueeze(1).contiguous(), lin.bias.detach())

    def time_embedding(self, t):
        """t [B] -> [B, 1, 1, 4 * time_embed_dim] (models/diffusion.py:347-351, 273-277)."""
        B = t.shape[0]
        t_emb = ops.timestep_embedding(t, self.config.model.time_embed_dim).view(B, 1, 1, -1)
        # the two un-quantized Linears as batch-invariant fp32 1x1 convs
        l0, l2 = self.time_embed[0], self.time_embed[2]
        t_emb = self._linear_f32(t_emb, l0)
        t_emb = ops.silu(t_emb)
        return self._linear_f32(t_emb, l2)

    def _forward_nhwc(self, x, t):
        fp = getattr(self, "_fused", None)                 # rowprog.FusedPlans, set by the CUDA-graph engine
        if fp is not None and fp.hoisted:
            # the time path of this step was evaluated at the start of the pass (engine.py): fan its rows out
            t_emb = None
            fp.bcast_temb(self._fused_cur)
        else:
            t_emb = self.time_embedding(t)
            if fp is not None:
                fp.run_time_mlps(t_emb, self._fused_cur)
        h = self.init_conv.forward_fused(x)
        skips = [h]
        if fp is not None and fp.trunk_plan is not None:
            # every block that works on a 1x1 map runs inside ONE kernel (attndm_rowprog)
            for layer in self.down_blocks[:fp.first_down]:
                h = layer.forward_fused(h, t_emb)
                skips.append(h)
            h = fp.run_trunk(h, self._fused_cur)
            rest = self.up_blocks[fp.n_up:]
        else:
            for layer in self.down_blocks:
                h = layer.forward_fused(h, t_emb)
                skips.append(h)
            h = self.middle_block1.forward_fused(h)
            h = self.middle_attn.forward_fused(h)
            h = self.middle_block2.forward_fused(h)
            rest = self.up_blocks
        for layer in rest:
            skip = skips.pop() if len(skips) else torch.zeros_like(h)
            h = layer.forward_fused(h, skip, t_emb)
        gn = _gn_args(self.norm_out, h)
        return self.conv_out.forward_fused(h, ops.PRE_GN_SILU, gn)

    def forward(self, x, t):
        with torch.no_grad():
            return ops.to_nchw(self.forward_nhwc(ops.to_nhwc(x), t))
