"""Shared helpers for the test-suite."""
import argparse

import numpy as np
import torch

from oracle import restate as R
from oracle import synth as S


def T(a):
    return torch.from_numpy(np.asarray(a))


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def config_for(spec: R.UNetSpec):
    return argparse.Namespace(
        data=argparse.Namespace(channels=spec.channels, image_size=spec.image_size, dataset="CIFAR10",
                                rescaled=True, logit_transform=False),
        model=argparse.Namespace(ch=spec.ch, ch_mult=list(spec.ch_mult), num_res_blocks=spec.num_res_blocks,
                                 dropout=0.1, var_type="fixedlarge"),
        diffusion=argparse.Namespace(beta_schedule="linear", beta_start=0.0001, beta_end=0.02,
                                     num_diffusion_timesteps=1000))


def args_for(spec: R.UNetSpec):
    return argparse.Namespace(bitwidth=spec.bitwidth, timesteps=spec.timesteps, skip_type="uniform", eta=0.0)


def build_cuda_model(spec, sd, device="cuda"):
    import attentiondm_b200 as A
    m = A.Model(config_for(spec), quantization=True, sequence=spec.seq, args=args_for(spec)).to(device).eval()
    m.materialize_lazy_layers()
    m.load_state_dict(sd, strict=True)
    m.init_weight_ranges()
    return m


def make_qconv(cin, cout, k, a_bit, Tn, G, device="cuda", w_bit=None):
    import attentiondm_b200 as A
    args = argparse.Namespace(bitwidth=a_bit, timesteps=Tn)
    q = A.QConv2d(cin, cout, k, padding=k // 2, w_bit=w_bit or a_bit, a_bit=a_bit,
                  sequence=range(0, 1000, 1000 // Tn), args=args)
    if G != q.group_num:
        q.group_num = G
        q.alpha_activ = torch.nn.Parameter(torch.zeros(Tn, G, cin))
        q.groups_range = torch.nn.Parameter(torch.zeros(Tn, G, 2), requires_grad=False)
    return q.to(device)
