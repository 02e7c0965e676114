"""Host-side planner for attndm_rowprog (include/attndm_b200.h): turns the run of UNet blocks that
work on 1x1 feature maps, and the time_mlp projections of every block, into op lists that ONE kernel
launch each interprets.

What is fused (reference citations; CIFAR numbers):
  * `time_mlp` of every Down/UpBlock (models/diffusion.py:157-161,175-177): SiLU -> quantize ->
    1x1 QConv2d of the same [B, 1024] embedding -- 23 independent programs, one launch;
  * the "trunk": DownBlocks whose pooled map is 1x1, the middle blocks and the UpBlocks whose skip is
    1x1 (models/diffusion.py:119-136, 170-190, 224-252, 362-376; models/self_attention.py:127-151)
    -- ~150 QConv2d, 97 GroupNorms -> one launch.
The planner mirrors `forward_fused` of the modules op for op; the kernel reproduces each stand-alone
kernel's arithmetic bit for bit, so the fused step equals the unfused one exactly (tested).

Only the CUDA-graph engine uses these plans (SamplerEngine); the eager module path stays layer by layer,
which is what the per-layer parity tests observe.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import List, Optional

import torch
import torch.nn as nn

from . import _ffi
from . import ops
from ._ffi import PRE_GN_SILU, PRE_NONE, PRE_SILU

OP_END, OP_LOAD, OP_LOAD_POOL, OP_STORE, OP_COPY, OP_CONV, OP_FCONV, OP_ATTN1, OP_SCALE_ADD = range(9)
MAX_SMEM = 220 * 1024
MAX_O = 4096          # FCONV
MAX_CONV_O = 512      # CONV (two MMA tiles per warp, two k16 steps per 16 KB ring slot)


class RowOp(C.Structure):
    """Mirror of `attndm_rowop`."""
    _fields_ = [("type", C.c_int32), ("C", C.c_int32), ("O", C.c_int32),
                ("src_off", C.c_int32), ("src_ld", C.c_int32), ("dst_off", C.c_int32), ("dst_ld", C.c_int32),
                ("add0_off", C.c_int32), ("add0_ld", C.c_int32), ("aux_off", C.c_int32), ("aux_ld", C.c_int32),
                ("pre", C.c_int32), ("a_bit", C.c_int32), ("tab_off", C.c_int32),
                ("nx_C", C.c_int32), ("nx_O", C.c_int32), ("nx_tab_off", C.c_int32),
                ("g_ld", C.c_int32), ("fparam", C.c_float), ("g0_ext", C.c_int32),
                ("g0", C.c_void_p), ("g1", C.c_void_p), ("qw", C.c_void_p), ("stat", C.c_void_p),
                ("nx_qw", C.c_void_p), ("nx_stat", C.c_void_p), ("rsv0", C.c_void_p), ("rsv1", C.c_void_p)]


assert C.sizeof(RowOp) == 144


class Unfusable(Exception):
    pass


last_unfusable = None


@dataclass
class Buf:
    """A [ns][width] fp32 buffer of the per-CTA arena: units are floats PER SAMPLE."""
    u: int          # per-sample offset of the parent allocation
    width: int      # leading dimension (channels of the parent allocation)
    c0: int = 0     # first channel of this view
    ch: int = 0     # channels of this view

    def view(self, c0, ch):
        return Buf(self.u, self.width, self.c0 + c0, ch)


class _Arena:
    """First-fit allocator over per-sample float units."""

    def __init__(self):
        self.free: List[List[int]] = []     # [start, size]
        self.top = 0
        self.peak = 0

    def alloc(self, ch: int) -> Buf:
        size = (ch + 3) // 4 * 4
        for blk in self.free:
            if blk[1] >= size:
                start = blk[0]
                blk[0] += size
                blk[1] -= size
                if blk[1] == 0:
                    self.free.remove(blk)
                return Buf(start, ch, 0, ch)
        start = self.top
        self.top += size
        self.peak = max(self.peak, self.top)
        return Buf(start, ch, 0, ch)

    def release(self, b: Buf):
        size = (b.width + 3) // 4 * 4
        self.free.append([b.u, size])
        self.free.sort()
        merged = []
        for blk in self.free:
            if merged and merged[-1][0] + merged[-1][1] == blk[0]:
                merged[-1][1] += blk[1]
            else:
                merged.append(list(blk))
        if merged and merged[-1][0] + merged[-1][1] == self.top:
            self.top = merged[-1][0]
            merged.pop()
        self.free = merged


class Program:
    """One op list (ends with OP_END) over one arena."""

    def __init__(self, plan: "Plan"):
        self.plan = plan
        self.arena = _Arena()
        self.ops: List[dict] = []
        self.cp_max = 16
        self.pbuf = 16

    # buffer references are resolved to arena offsets once ns is known
    def _emit(self, **kw):
        self.ops.append(kw)

    def load(self, dst: Buf, ext: int = -1, tensor: Optional[torch.Tensor] = None, g_ld: int = 0, pool=False):
        self._emit(type=OP_LOAD_POOL if pool else OP_LOAD, C=dst.ch, dst=dst, g0_ext=ext,
                   g0=self.plan.keep(tensor), g_ld=g_ld or dst.ch)

    def store(self, src: Buf, ext: int = -1, tensor: Optional[torch.Tensor] = None, g_ld: int = 0):
        self._emit(type=OP_STORE, C=src.ch, src=src, g0_ext=ext, g0=self.plan.keep(tensor), g_ld=g_ld or src.ch)

    def copy(self, src: Buf, dst: Buf):
        self._emit(type=OP_COPY, C=src.ch, src=src, dst=dst)

    def conv(self, q, src: Buf, dst: Buf, pre=PRE_NONE, norm: Optional[nn.GroupNorm] = None, add0: Optional[Buf] = None,
             temb: Optional[torch.Tensor] = None):
        """One QConv2d on a 1x1 map (QConv2d.forward_fused's integer branch)."""
        if q._calibrate:
            raise Unfusable("layer is calibrating")
        f32 = not q.int8_ok_all_steps()              # the engine's rule: the fp32 path unless EVERY step may use int8
        if f32 and (add0 is not None or temb is not None or q.in_channels % 4):
            raise Unfusable("fp32-path layer with a fused residual / time embedding")
        if q.in_channels % 16 or q.out_channels % 4 or q.out_channels > MAX_CONV_O or q.in_channels != src.ch or q.out_channels != dst.ch:
            raise Unfusable("channel counts outside the fused kernel's range")
        if norm is not None and (norm.num_groups != 32 or q.in_channels % 32):
            raise Unfusable("GroupNorm is not 32 groups")
        w_eff, i8 = q._packed()
        if q.taps == 9:
            w_eff, i8 = q._pack_center
        off, width = self.plan.layer_slice[id(q)]
        Cc, O = q.in_channels, q.out_channels
        lay = q.table_layout()
        if (lay["zp"], lay["mult"], lay["act_zp"]) != ((Cc + 3) // 4 * 4, 2 * ((Cc + 3) // 4 * 4),
                                                        2 * ((Cc + 3) // 4 * 4) + (O + 3) // 4 * 4):
            raise Unfusable("unexpected table layout")
        self.cp_max = max(self.cp_max, Cc)
        self.pbuf = max(self.pbuf, (2 * ((Cc + 3) // 4 * 4) + (O + 3) // 4 * 4 + 4 + 2 * Cc + 3 * O + 3) // 4 * 4)
        # static block: gamma[C] beta[C] bias[O] | wsum[O] w_zp[O] (int32 bit patterns)
        dev = w_eff.device
        z = lambda n: torch.zeros(n, dtype=torch.float32, device=dev)
        stat = torch.cat([norm.weight.detach().float() if norm is not None else z(Cc),
                          norm.bias.detach().float() if norm is not None else z(Cc),
                          q.bias.detach().float() if q.bias is not None else z(O),
                          i8.wsum.view(torch.float32) if i8 is not None else z(O),
                          i8.w_zp.view(torch.float32) if i8 is not None else z(O)]).contiguous()
        if f32:
            # QConv2d.forward_fused's fp32 branch on a 1x1 map: quantize -> de-quantize, then conv_f32 with the clamped fp32
            # weights (centre tap); the kernel runs the same quantizer phases and the FCONV loop (rowprog.cu)
            if O % 4 or O > MAX_O:
                raise Unfusable("fp32-path layer width outside the fused kernel's range")
            wt = w_eff.detach().reshape(O, Cc).t().contiguous()              # [C][O]
            if wt.numel() * 4 < ((O + 15) // 16) * ((Cc + 31) // 32) * 512:
                raise Unfusable("fp32 weights smaller than the fragment prefetch")
            y = self.arena.alloc(Cc)                                         # the fake-quantized row
            self._emit(type=OP_CONV, C=Cc, O=O, src=src, dst=dst, aux=y, pre=pre,
                       a_bit=q._a_bit, tab_off=off, fparam=float(norm.eps) if norm is not None else 0.0,
                       qw=self.plan.keep(wt), stat=self.plan.keep(stat), rsv0=1)
            self.arena.release(y)
            return
        self._emit(type=OP_CONV, C=Cc, O=O, src=src, dst=dst, add0=add0, pre=pre,
                   a_bit=q._a_bit, tab_off=off, fparam=float(norm.eps) if norm is not None else 0.0,
                   g1=self.plan.keep(temb), qw=self.plan.packed_weights(i8), stat=self.plan.keep(stat))

    def fconv(self, conv: nn.Conv2d, src: Buf, dst: Buf):
        """Un-quantized fp32 1x1 conv (the lazily created channel_proj, models/diffusion.py:235-242)."""
        O, Cc = conv.weight.shape[0], conv.weight.shape[1]
        if O % 4 or O > MAX_O:
            raise Unfusable("channel_proj width outside the fused kernel's range")
        wt = conv.weight.detach().view(O, Cc).t().contiguous()           # [C][O]: lanes read consecutive o
        self._emit(type=OP_FCONV, C=Cc, O=O, src=src, dst=dst, g0=self.plan.keep(wt),
                   g1=self.plan.keep(conv.bias.detach() if conv.bias is not None else None))

    def attn1(self, q: Buf, k: Buf, v: Buf, dst: Buf, scale: float, params: Optional[torch.Tensor] = None):
        """params: the 8 floats of a MixedPrecisionAttention (rowprog.cu, ATTN1), or None for the plain softmax branch."""
        self._emit(type=OP_ATTN1, C=q.ch, O=v.ch, src=q, add0=k, aux=v, dst=dst, fparam=float(scale),
                   g0=self.plan.keep(params))

    def scale_add(self, a: Buf, x: Buf, dst: Buf, gamma: torch.Tensor):
        self._emit(type=OP_SCALE_ADD, C=a.ch, src=a, add0=x, dst=dst, g0=self.plan.keep(gamma))

    def encode(self, ns: int) -> List[RowOp]:
        if not self.ops or self.ops[0]["type"] == OP_CONV:
            raise Unfusable("a program must start with a non-CONV op (it prefetches the first conv)")
        # link every CONV (and the first op) to the next CONV of the program
        nxt = None
        for i in range(len(self.ops) - 1, -1, -1):
            o = self.ops[i]
            if o["type"] == OP_CONV or i == 0:
                o["nx"] = nxt
            if o["type"] == OP_CONV:
                nxt = o
        out = []
        for o in self.ops + [dict(type=OP_END)]:
            r = RowOp()
            r.type = o["type"]
            r.C = o.get("C", 0)
            r.O = o.get("O", 0)
            for name in ("src", "dst", "add0", "aux"):
                b = o.get(name)
                if b is None:
                    setattr(r, name + "_off", -1)
                    setattr(r, name + "_ld", 0)
                else:
                    if b.c0 % 4 or b.width % 4:
                        raise Unfusable("unaligned buffer view")
                    setattr(r, name + "_off", b.u * ns + b.c0)
                    setattr(r, name + "_ld", b.width)
            for name in ("pre", "a_bit", "tab_off", "g_ld"):
                setattr(r, name, int(o.get(name, 0)))
            r.fparam = o.get("fparam", 0.0)
            r.g0_ext = int(o.get("g0_ext", -1))
            for name in ("g0", "g1", "qw", "stat", "rsv0"):
                setattr(r, name, o.get(name))
            nx = o.get("nx")
            if nx is not None:
                r.nx_C, r.nx_O, r.nx_tab_off = nx["C"], nx["O"], nx["tab_off"]
                r.nx_qw, r.nx_stat = nx["qw"], nx["stat"]
            out.append(r)
        return out


class Plan:
    """A set of programs launched together (grid.y = program)."""

    def __init__(self, layer_slice, device):
        self.layer_slice = layer_slice      # id(QConv2d) -> (offset, width) in the staged table
        self.device = device
        self.programs: List[Program] = []
        self._keep = []                      # tensors whose storage the op list points into
        self.attn_params = []                # (MixedPrecisionAttention, the 8 values baked into its ATTN1 op)
        self._packed = {}
        self.ns = 0
        self.dev_ops = None
        self.dev_start = None
        self.arena_floats = 0
        self.cp_max = 16
        self.pbuf = 16

    def keep(self, t: Optional[torch.Tensor]):
        if t is None:
            return None
        if not t.is_cuda:
            raise Unfusable("parameter is not on the GPU")
        t = t.contiguous()
        self._keep.append(t)
        return t.data_ptr()

    def packed_weights(self, i8: ops.I8Pack):
        key = i8.qw.data_ptr()
        if key not in self._packed:
            O, Cp = i8.qw.shape
            nbytes = _ffi.lib().attndm_rowprog_packed_weight_bytes(O, Cp)
            out = torch.empty(nbytes, dtype=torch.int8, device=i8.qw.device)
            _ffi.call("attndm_rowprog_pack_weights", _ffi.ptr(i8.qw), O, Cp, _ffi.ptr(out), _ffi.stream())
            self._packed[key] = out
            self._keep.append(i8.qw)
        return self._packed[key].data_ptr()

    def new_program(self) -> Program:
        p = Program(self)
        self.programs.append(p)
        return p

    def finalize(self, B: int, force_ns: Optional[int] = None):
        peak = max(p.arena.peak for p in self.programs)
        self.cp_max = max(p.cp_max for p in self.programs)
        self.pbuf = max(p.pbuf for p in self.programs)
        L = _ffi.lib()
        for ns in (8, 4, 2):
            if L.attndm_rowprog_smem_bytes(ns, peak * ns, self.cp_max, self.pbuf) <= MAX_SMEM:
                break
        else:
            raise Unfusable("program does not fit in shared memory")
        # the programs are latency-bound per CTA and every CTA re-reads the weights from L2: spread the
        # samples over enough CTAs to cover the chip, but no thinner
        target = int(os.environ.get("ATTNDM_ROWPROG_CTAS", "128"))
        while ns > 2 and (B + ns - 1) // ns * len(self.programs) < target:
            ns //= 2
        if force_ns is not None:
            if force_ns > ns:
                raise Unfusable("program does not fit in shared memory at the requested samples per CTA")
            ns = force_ns
        self.ns = ns
        self.arena_floats = peak * ns
        recs, starts = [], []
        for p in self.programs:
            starts.append(len(recs))
            recs.extend(p.encode(ns))
        arr = (RowOp * len(recs))(*recs)
        raw = torch.frombuffer(bytearray(C.string_at(C.addressof(arr), C.sizeof(arr))), dtype=torch.uint8)
        self.dev_ops = raw.to(self.device)
        self.dev_start = torch.tensor(starts, dtype=torch.int32, device=self.device)
        self.n_ops = len(recs)
        return self

    def run(self, B: int, cur: torch.Tensor, ext: List[torch.Tensor], cur_cta_stride: int = 0):
        """cur_cta_stride (floats): 0 = every sample reads the staged row `cur`; else CTA x reads row x of the table."""
        arr = (C.c_void_p * 4)(*([t.data_ptr() for t in ext] + [None] * (4 - len(ext))))
        _ffi.call("attndm_rowprog", _ffi.ptr(self.dev_ops), _ffi.ptr(self.dev_start), len(self.programs), B, self.ns,
                  self.arena_floats, self.cp_max, self.pbuf, _ffi.ptr(cur), int(cur_cta_stride), arr, len(ext),
                  _ffi.stream())


# ---------------------------------------------------------------------------------------------
# planners: mirror forward_fused of the modules
# ---------------------------------------------------------------------------------------------
def _res_block(p: Program, rb, x: Buf, temb: Optional[torch.Tensor]) -> Buf:
    """ResidualBlock.forward_fused: conv1(GN+SiLU) -> [shortcut conv] -> conv2(GN+SiLU) + sc (+ temb)."""
    h = p.arena.alloc(rb.out_channels)
    p.conv(rb.conv1, x, h, PRE_GN_SILU, rb.norm1)
    sc, own = x, False
    if rb.in_channels != rb.out_channels:
        sc, own = p.arena.alloc(rb.out_channels), True
        p.conv(rb.conv_shortcut if rb.use_conv_shortcut else rb.nin_shortcut, x, sc)
    p.conv(rb.conv2, h, h, PRE_GN_SILU, rb.norm2, add0=sc, temb=temb)
    if own:
        p.arena.release(sc)
    return h


def _attention(p: Program, at, x: Buf) -> Buf:
    """EnhancedQSelfAttention.forward_fused at one position (plain softmax branch)."""
    params = None
    scale = at.key_channels ** -0.5
    if at.mixed_precision and at.quantization:
        # MixedPrecisionAttention.forward_nhwc as Model.forward reaches it (timestep=None: the base bit width)
        mpa = at.attention_processor
        if mpa.num_heads > 8 or at.key_channels % mpa.num_heads or at.value_channels % mpa.num_heads:
            raise Unfusable("attention heads outside what the fused program takes")
        h = mpa.refresh_host_params()
        eff = mpa.get_effective_bits(None)
        qk_bits = max(4, int(eff)) if eff <= 6 else 0
        p_bits = max(3, int(eff - 1)) if eff <= 4 else 0
        vals = [float(mpa.num_heads), h["softmax_scale"], h["scale_qk"], h["zero_qk"], float(qk_bits),
                h["scale_attn"], h["zero_attn"], float(p_bits)]
        params = torch.tensor(vals, dtype=torch.float32, device=p.plan.device)
        p.plan.attn_params.append((mpa, tuple(vals)))
        scale = mpa.scaling_factor
    q = p.arena.alloc(at.key_channels)
    k = p.arena.alloc(at.key_channels)
    v = p.arena.alloc(at.value_channels)
    p.conv(at.query_conv, x, q)
    p.conv(at.key_conv, x, k)
    p.conv(at.value_conv, x, v)
    p.attn1(q, k, v, v, scale, params)
    o = p.arena.alloc(at.in_channels)
    p.conv(at.output_conv, v, o)
    p.scale_add(o, x, o, at.gamma.detach())
    for b in (q, k, v):
        p.arena.release(b)
    return o


def _tail(p: Program, blk, x: Buf, temb, release_x: bool) -> Buf:
    from .self_attention import EnhancedQSelfAttention
    h = _res_block(p, blk.res1, x, temb)
    if release_x:
        p.arena.release(x)
    h2 = _res_block(p, blk.res2, h, None)
    p.arena.release(h)
    if isinstance(blk.attn, EnhancedQSelfAttention):
        h3 = _attention(p, blk.attn, h2)
        p.arena.release(h2)
        h2 = h3
    return h2


@dataclass
class FusedPlans:
    time_plan: Plan
    temb: dict                      # id(block) -> persistent [B, O] tensor written by time_plan
    trunk_plan: Optional[Plan]
    first_down: int                 # down_blocks[first_down:] are inside the trunk
    n_up: int                       # up_blocks[:n_up] are inside the trunk
    trunk_out_ch: int
    B: int
    # The time path hoisted out of the step (engine.py): every time_mlp for ALL T sampler steps in one launch
    # (time_all: two identical samples per step, CTA x = step x, reading row x of the whole table), its results kept
    # as table columns and fanned out per step to the [B, O] tensors of `temb` (views of temb_flat).
    time_all: Optional[Plan] = None
    temb_all: Optional[list] = None      # [2T, O] output of each time_mlp, in the order of temb
    temb_flat: Optional[torch.Tensor] = None
    temb_cols: int = 0                   # sum of the O (each rounded up to 4)
    bcast_desc: Optional[torch.Tensor] = None
    bcast_n: int = 0
    bcast_wmax: int = 0
    hoisted: bool = False

    def run_time_mlps(self, t_emb: torch.Tensor, cur: torch.Tensor):
        self.time_plan.run(self.B, cur, [t_emb])

    def set_hoisted(self, col0: int):
        """The engine keeps the hoisted results in table columns [col0, col0 + temb_cols)."""
        desc, off_dst, off_src = [], 0, col0
        for o in self.temb_all:
            w = o.shape[1]
            desc.append((off_dst, off_src, w))
            off_dst += self.B * w
            off_src += (w + 3) // 4 * 4
        self.bcast_desc = torch.tensor(desc, dtype=torch.int32, device=self.temb_flat.device)
        self.bcast_n, self.bcast_wmax = len(desc), max(d[2] for d in desc)
        self.hoisted = True

    def run_time_all(self, t_emb2: torch.Tensor, table: torch.Tensor) -> torch.Tensor:
        """t_emb2: [2T, ted4] (row 2k and 2k+1 = step k).  Returns the [T, temb_cols] block of table columns."""
        T2 = t_emb2.shape[0]
        self.time_all.run(T2, table, [t_emb2], cur_cta_stride=table.stride(0))
        cols = []
        for o in self.temb_all:
            w = o.shape[1]
            c = o[::2]
            if w % 4:
                c = torch.cat([c, c.new_zeros(c.shape[0], 4 - w % 4)], dim=1)
            cols.append(c)
        return torch.cat(cols, dim=1)

    def bcast_temb(self, cur: torch.Tensor):
        _ffi.call("attndm_bcast_rows", _ffi.ptr(cur), _ffi.ptr(self.bcast_desc), self.bcast_n, self.B, self.bcast_wmax,
                  _ffi.ptr(self.temb_flat), _ffi.stream())

    def run_trunk(self, h: torch.Tensor, cur: torch.Tensor) -> torch.Tensor:
        out = torch.empty(self.B, 1, 1, self.trunk_out_ch, dtype=torch.float32, device=h.device)
        self.trunk_plan.run(self.B, cur, [h, out])
        return out


def build(model, B: int, layer_slice: dict, device, T: int = 0) -> Optional[FusedPlans]:
    """Plans for `model` at batch B, or None when the model/state is outside what the kernel fuses.
    T > 0: also the all-steps time plan (FusedPlans.time_all)."""
    if os.environ.get("ATTNDM_FUSED", "1") == "0":
        return None
    global last_unfusable
    last_unfusable = None
    try:
        fp = _build(model, B, layer_slice, device)
        if T > 0:
            try:
                _add_time_all(fp, model, layer_slice, device, T)
            except Unfusable:
                fp.time_all = None
        return fp
    except Unfusable as e:
        last_unfusable = str(e)          # why the model / state is outside what the kernel fuses
        return None


def _add_time_all(fp: FusedPlans, model, layer_slice, device, T: int):
    """The time_mlp programs once more, for 2T samples with two samples per CTA: CTA x is sampler step x."""
    ted4 = model.config.model.time_embed_dim * 4
    tp = Plan(layer_slice, device)
    outs = []
    for blk in list(model.down_blocks) + list(model.up_blocks):
        if blk.time_mlp is None:
            continue
        q = blk.time_mlp[1]
        out = torch.empty(2 * T, q.out_channels, dtype=torch.float32, device=device)
        p = tp.new_program()
        x = p.arena.alloc(ted4)
        y = p.arena.alloc(q.out_channels)
        p.load(x, ext=0, g_ld=ted4)
        p.conv(q, x, y, PRE_SILU)
        p.store(y, tensor=out)
        outs.append(out)
    tp.finalize(2 * T, force_ns=2)
    fp.time_all, fp.temb_all = tp, outs
    fp.temb_cols = sum((o.shape[1] + 3) // 4 * 4 for o in outs)


def _build(model, B, layer_slice, device) -> FusedPlans:
    ted4 = model.config.model.time_embed_dim * 4
    blocks = list(model.down_blocks) + list(model.up_blocks)
    # ---- every time_mlp as its own program of one launch ----
    tp = Plan(layer_slice, device)
    temb = {}
    n_flat = sum(B * blk.time_mlp[1].out_channels for blk in blocks if blk.time_mlp is not None)
    flat = torch.empty(max(1, n_flat), dtype=torch.float32, device=device)     # every [B, O] output, back to back
    off_flat = 0
    for blk in blocks:
        if blk.time_mlp is None:
            continue
        q = blk.time_mlp[1]
        out = flat[off_flat:off_flat + B * q.out_channels].view(B, q.out_channels)
        off_flat += B * q.out_channels
        p = tp.new_program()
        x = p.arena.alloc(ted4)
        y = p.arena.alloc(q.out_channels)
        p.load(x, ext=0, g_ld=ted4)
        p.conv(q, x, y, PRE_SILU)
        p.store(y, tensor=out)
        temb[id(blk)] = out
    if not tp.programs:
        raise Unfusable("no time_mlp")
    tp.finalize(B)
    # ---- the trunk ----
    S = int(model.config.data.image_size)
    sizes = []                                    # (input, output) spatial size of each down block
    s = S
    for _ in model.down_blocks:
        o = s // 2 if s > 1 else s
        sizes.append((s, o))
        s = o
    first = len(sizes)
    while first > 0 and sizes[first - 1][1] == 1:
        first -= 1
    n_down = len(model.down_blocks)
    if first == n_down or s != 1:
        return FusedPlans(tp, temb, None, n_down, 0, 0, B, temb_flat=flat)
    n_up = min(n_down - first, len(model.up_blocks))
    plan = Plan(layer_slice, device)
    p = plan.new_program()
    blk0 = model.down_blocks[first]
    x = p.arena.alloc(blk0.res1.in_channels)
    p.load(x, ext=0, pool=(sizes[first][0] == 2))
    if sizes[first][0] not in (1, 2):
        raise Unfusable("trunk input is neither 1x1 nor 2x2")
    skips = []
    h = x
    for i in range(first, n_down):
        blk = model.down_blocks[i]
        h = _tail(p, blk, h, temb.get(id(blk)), release_x=(i == first))     # later inputs are skips: kept
        skips.append(h)
    m1 = _res_block(p, model.middle_block1, h, None)                        # h is also the last skip: kept
    m2 = _attention(p, model.middle_attn, m1)
    p.arena.release(m1)
    h = _res_block(p, model.middle_block2, m2, None)
    p.arena.release(m2)
    for j in range(n_up):
        blk = model.up_blocks[j]
        skip = skips.pop()
        cat = p.arena.alloc(h.ch + skip.ch)                                  # upsample_concat at 1x1 = channel concat
        p.copy(h, cat.view(0, h.ch))
        p.copy(skip, cat.view(h.ch, skip.ch))
        p.arena.release(h)
        p.arena.release(skip)
        expected = blk.res1.in_channels
        if cat.ch != expected:
            if not hasattr(blk, "channel_proj"):
                raise Unfusable("channel_proj not materialised yet")
            proj = p.arena.alloc(expected)
            p.fconv(blk.channel_proj, cat, proj)
            p.arena.release(cat)
            cat = proj
        h = _tail(p, blk, cat, temb.get(id(blk)), release_x=True)
    p.store(h, ext=1)
    plan.finalize(B)
    return FusedPlans(tp, temb, plan, first, n_up, h.ch, B, temb_flat=flat)
