"""Harness that drives the UNMODIFIED reference (/root/reference) on CPU.

TEST INFRASTRUCTURE ONLY.  Only tests/, oracle/make_golden.py and bench.py's
cpu_baseline / --impl reference legs may import this.  It exists only in the
build container (the GPU box has no /root/reference); everything that must run
on the GPU box uses oracle/restate.py plus the committed fixtures in
tests/golden/ that this harness generated.

The reference's quantized path does not run as shipped.
The harness applies, without editing any reference source:
  D2  stub modules `progress`, `progress.bar`, `lmdb`            (missing deps)
  D1  alias utils.attention_quant_util -> utils.attention_quant_utils
  D3  weight_range_min/max := per-out-channel min/max of the weight
  D4  groups_range re-created where shape[1] != group_num
  H1  (optional) weights snapped to the w_bit grid with the reference's own
      AsymmetricQuantFunction, per out-channel (utils/quantization_utils/
      quant_utils.py:136-167)
"""
import argparse
import os
import sys
import types

REF_ROOT = os.environ.get("ATTNDM_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "utils"))


_loaded = None


def load():
    """Import the reference packages; returns a namespace of the symbols used."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference not present at {REF_ROOT}")
    sys.dont_write_bytecode = True
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    p, pb = types.ModuleType("progress"), types.ModuleType("progress.bar")
    pb.Bar = object
    p.bar = pb
    sys.modules.setdefault("progress", p)
    sys.modules.setdefault("progress.bar", pb)
    sys.modules.setdefault("lmdb", types.ModuleType("lmdb"))
    import utils.attention_quant_utils as aq  # noqa: E402  (reference module)
    sys.modules["utils.attention_quant_util"] = aq
    import utils.quant_util as qu
    import utils.quantization_utils.quant_utils as qz
    import models.diffusion as md
    import models.self_attention as sa
    import functions.denoising as dn
    import runners.diffusion as rd
    ns = types.SimpleNamespace(qu=qu, qz=qz, md=md, sa=sa, dn=dn, rd=rd, aq=aq)
    _loaded = ns
    return ns


def dict2namespace(d):
    ns = argparse.Namespace()
    for k, v in d.items():
        setattr(ns, k, dict2namespace(v) if isinstance(v, dict) else v)
    return ns


def tiny_config(ch=32, ch_mult=(1, 2), num_res_blocks=1, image_size=8, channels=3):
    """A small config with the same structure as configs/cifar10.yml."""
    return dict2namespace(dict(
        data=dict(dataset="CIFAR10", image_size=image_size, channels=channels),
        model=dict(type="simple", in_channels=channels, out_ch=channels, ch=ch,
                   ch_mult=list(ch_mult), num_res_blocks=num_res_blocks,
                   attn_resolutions=[16], dropout=0.1, var_type="fixedlarge",
                   ema_rate=0.9999, ema=True, resamp_with_conv=True),
        diffusion=dict(beta_schedule="linear", beta_start=0.0001, beta_end=0.02,
                       num_diffusion_timesteps=1000),
    ))


def named_config(name):
    import yaml
    with open(os.path.join(REF_ROOT, "configs", name)) as f:
        return dict2namespace(yaml.safe_load(f))


def build_model(config, timesteps, bitwidth, seed=0, snap_weights=True):
    """Reference Model(quantization=True) with fixes D3/D4 (+H1)."""
    import torch
    import torch.nn as nn
    ref = load()
    args = argparse.Namespace(bitwidth=bitwidth, timesteps=timesteps)
    seq = range(0, 1000, 1000 // timesteps)
    torch.manual_seed(seed)
    m = ref.md.Model(config, quantization=True, sequence=seq, args=args).eval()
    fix_model(m, snap_weights=snap_weights)
    return m, seq, args


def fix_model(m, snap_weights=True):
    import torch
    import torch.nn as nn
    ref = load()
    for q in m.modules():
        if isinstance(q, ref.qu.QConv2d):
            if snap_weights:
                w = q.weight.data
                flat = w.reshape(w.shape[0], -1)
                q.weight.data = ref.qz.AsymmetricQuantFunction.apply(
                    w, q.w_bit, flat.min(1)[0], flat.max(1)[0]).detach().clone()
            flat = q.weight.data.reshape(q.weight.shape[0], -1)
            q.weight_range_min = flat.min(1)[0].clone()
            q.weight_range_max = flat.max(1)[0].clone()
            if q.groups_range.shape[1] != q.group_num:
                q.groups_range = nn.Parameter(
                    torch.zeros(q.len_seq, q.group_num, 2), requires_grad=False)


def qconvs(m):
    ref = load()
    return [(n, q) for n, q in m.named_modules() if isinstance(q, ref.qu.QConv2d)]


def set_calibrate(m, flag, first=False):
    for _, q in qconvs(m):
        q.set_calibrate(flag)
        q.first_calibrate(first)


def reset_index(m):
    for _, q in qconvs(m):
        q.index_seq = 0


def betas(config):
    import torch
    ref = load()
    d = config.diffusion
    b = ref.rd.get_beta_schedule(d.beta_schedule, beta_start=d.beta_start, beta_end=d.beta_end,
                                 num_diffusion_timesteps=d.num_diffusion_timesteps)
    return torch.from_numpy(b).float()
