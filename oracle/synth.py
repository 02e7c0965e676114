"""Deterministic synthetic weights for parity fixtures (TEST INFRASTRUCTURE).

There is no network for checkpoints and the reference publishes no golden
vectors, so fixtures are built on weights that every side can
regenerate bit-identically from (spec, seed): the real reference in the build
container (oracle/make_golden.py), the CPU restatement (oracle/restate.py) and
the CUDA path on the GPU box.  Keys follow the reference's state_dict names
(models/diffusion.py:255-345) plus the lazily created `channel_proj` convs
(models/diffusion.py:235-242).
"""
from __future__ import annotations

import hashlib
import math
import zlib
from typing import Dict

import torch

from . import restate as R


def _gen(name: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def _uniform(shape, bound, name, seed):
    return (torch.rand(shape, generator=_gen(name, seed), dtype=torch.float32) * 2 - 1) * bound


def channel_proj_table(spec: R.UNetSpec):
    """Which up blocks get a lazy fp32 channel_proj and with what shape.
    Derived by tracing channel counts exactly as UpBlock.forward does
    (models/diffusion.py:232-242)."""
    lay = R.unet_layout(spec)
    skips = [spec.ch] + [b["cout"] for b in lay["down"]]
    now = lay["mid"]
    out = {}
    for b in lay["up"]:
        sk = skips.pop() if skips else now
        actual = now + sk
        if actual != b["cin"]:
            out[b["name"] + ".channel_proj"] = (actual, b["cin"])
        now = b["cout"]
    return out


def synth_state_dict(spec: R.UNetSpec, seed: int = 0, weight_gain: float = 1.0, gamma: float = 0.5,
                     alpha_mode: str = "uniform", snap: bool = True) -> Dict[str, torch.Tensor]:
    """alpha_mode: 'uniform' (reference init 0.01, utils/quant_util.py:119-120) or
    'attn_random' (N(0,1) on the attention projections, emulating a post-
    calibrate_attention state; BASELINE.json config 3)."""
    sd: Dict[str, torch.Tensor] = {}
    T = spec.len_seq
    tab = R.qconv_table(spec)
    lay = R.unet_layout(spec)

    def lin(name, cin, cout):
        b = 1.0 / math.sqrt(cin)
        sd[name + ".weight"] = _uniform((cout, cin), b, name + ".weight", seed)
        sd[name + ".bias"] = _uniform((cout,), b, name + ".bias", seed)

    def gn(name, c):
        sd[name + ".weight"] = 1.0 + _uniform((c,), 0.2, name + ".weight", seed)
        sd[name + ".bias"] = _uniform((c,), 0.2, name + ".bias", seed)

    lin("time_embed.0", spec.time_embed_dim, spec.time_embed_dim * 4)
    lin("time_embed.2", spec.time_embed_dim * 4, spec.time_embed_dim * 4)
    for name, d in tab.items():
        fan_in = d["cin"] * d["k"] * d["k"]
        b = weight_gain / math.sqrt(fan_in)
        w = _uniform((d["cout"], d["cin"], d["k"], d["k"]), b, name + ".weight", seed)
        if snap:
            w = R.snap_weight(w, d["w_bit"])[0]
        sd[name + ".weight"] = w
        sd[name + ".bias"] = _uniform((d["cout"],), 1.0 / math.sqrt(fan_in), name + ".bias", seed)
        sd[name + ".groups_range"] = torch.zeros(T, d["group_num"], 2)
        is_attn = any(s in name for s in ("query_conv", "key_conv", "value_conv", "output_conv"))
        if alpha_mode == "attn_random" and is_attn:
            sd[name + ".alpha_activ"] = torch.randn(T, d["group_num"], d["cin"],
                                                    generator=_gen(name + ".alpha", seed))
        else:
            sd[name + ".alpha_activ"] = torch.full((T, d["group_num"], d["cin"]), 0.01)

    def res_norms(p, cin, cout):
        gn(p + ".norm1", cin)
        gn(p + ".norm2", cout)

    for b in lay["down"] + lay["up"]:
        res_norms(b["name"] + ".res1", b["cin"], b["cout"])
        res_norms(b["name"] + ".res2", b["cout"], b["cout"])
        if b["attn"]:
            sd[b["name"] + ".attn.gamma"] = torch.full((1,), gamma)
            sd[b["name"] + ".attn.temperature"] = torch.ones(1)
    res_norms("middle_block1", lay["mid"], lay["mid"])
    res_norms("middle_block2", lay["mid"], lay["mid"])
    sd["middle_attn.gamma"] = torch.full((1,), gamma)
    sd["middle_attn.temperature"] = torch.ones(1)
    gn("norm_out", lay["final"])
    for name, (cin, cout) in channel_proj_table(spec).items():
        b = 1.0 / math.sqrt(cin)
        sd[name + ".weight"] = _uniform((cout, cin, 1, 1), b, name + ".weight", seed)
        sd[name + ".bias"] = _uniform((cout,), b, name + ".bias", seed)
    return sd


def state_digest(sd: Dict[str, torch.Tensor]) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def tiny_spec(T=4, bitwidth=8, ch=32, ch_mult=(1, 2), num_res_blocks=1, image_size=16, channels=3):
    seq = tuple(range(0, 1000, 1000 // T))
    return R.UNetSpec(ch=ch, ch_mult=tuple(ch_mult), num_res_blocks=num_res_blocks, channels=channels,
                      image_size=image_size, bitwidth=bitwidth, timesteps=T, seq=seq)


def cifar_spec(T=100, bitwidth=8):
    """configs/cifar10.yml: ch 128, ch_mult [1,2,2,2], 2 res blocks, 32x32x3."""
    seq = tuple(range(0, 1000, 1000 // T))
    return R.UNetSpec(ch=128, ch_mult=(1, 2, 2, 2), num_res_blocks=2, channels=3, image_size=32,
                      bitwidth=bitwidth, timesteps=T, seq=seq)


def celeba_spec(T=100, bitwidth=8):
    """configs/celeba.yml: ch 128, ch_mult [1,2,2,2,4], 2 res blocks, 64x64x3."""
    seq = tuple(range(0, 1000, 1000 // T))
    return R.UNetSpec(ch=128, ch_mult=(1, 2, 2, 2, 4), num_res_blocks=2, channels=3, image_size=64,
                      bitwidth=bitwidth, timesteps=T, seq=seq)


def church_spec(T=100, bitwidth=8):
    """configs/church.yml: ch 128, ch_mult [1,1,2,2,4,4], 2 res blocks, 256x256x3."""
    seq = tuple(range(0, 1000, 1000 // T))
    return R.UNetSpec(ch=128, ch_mult=(1, 1, 2, 2, 4, 4), num_res_blocks=2, channels=3, image_size=256,
                      bitwidth=bitwidth, timesteps=T, seq=seq)
