"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference):
    python -m oracle.make_golden
The reference ships no golden vectors for this path, so these
fixtures ARE the pin: oracle/restate.py is checked against them on CPU and the
CUDA path is checked against them on the GPU box.
"""
import os
import sys
import warnings

import numpy as np
import torch

from . import ref_harness as H
from . import restate as R
from . import synth as S

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
warnings.filterwarnings("ignore")


def _np(t):
    return t.detach().cpu().numpy()


def make_qconv(ref, cin, cout, k, a_bit, T, G):
    import argparse
    args = argparse.Namespace(bitwidth=a_bit, timesteps=T)
    q = ref.qu.QConv2d(cin, cout, k, padding=k // 2, w_bit=a_bit, a_bit=a_bit,
                       sequence=range(0, 1000, 1000 // T), args=args)
    if G != q.group_num:
        q.group_num = G
        q.alpha_activ = torch.nn.Parameter(torch.zeros(T, G, cin))
        q.groups_range = torch.nn.Parameter(torch.zeros(T, G, 2), requires_grad=False)
    return q


def quant_unit(ref):
    out = {}
    g = torch.Generator().manual_seed(11)
    T = 4
    cases = [(3, 8, 8, (2, 3, 5, 5)), (64, 8, 8, (2, 64, 4, 4)), (128, 6, 8, (1, 128, 3, 3)),
             (256, 4, 4, (2, 256, 2, 2)), (96, 8, 8, (3, 96, 1, 1))]
    for ci, (C, a_bit, G, shape) in enumerate(cases):
        for mode in ("uniform", "random"):
            q = make_qconv(ref, C, 8, 1, a_bit, T, G)
            gr = torch.zeros(T, G, 2)
            gr[..., 0] = -4.0 - 3.0 * torch.rand(T, G, generator=g)
            gr[..., 1] = 6.0 + 5.0 * torch.rand(T, G, generator=g)
            q.groups_range.data.copy_(gr)
            if mode == "random":
                q.alpha_activ.data.copy_(torch.randn(T, G, C, generator=g))
            x = torch.randn(*shape, generator=g) * 4.0
            ys = []
            with torch.no_grad():
                for t in range(T + 1):          # T+1 calls exercises the index_seq wrap
                    ys.append(q._quantize_activation(x))
            key = f"inf{ci}_{mode}"
            out[key + "_x"] = _np(x)
            out[key + "_gr"] = _np(gr)
            out[key + "_alpha"] = _np(q.alpha_activ.data)
            out[key + "_y"] = np.stack([_np(y) for y in ys])
            out[key + "_meta"] = np.array([C, a_bit, G, T])
    # calibration (plain and first_calibrate)
    ccases = [(32, 8, 8, (2, 32, 4, 4), 5.0), (64, 4, 4, (2, 64, 2, 2), 8.0), (16, 8, 8, (4, 16, 3, 3), 0.05),
              (128, 6, 8, (2, 128, 1, 1), 3.0)]
    for ci, (C, a_bit, G, shape, amp) in enumerate(ccases):
        for first in (False, True):
            q = make_qconv(ref, C, 8, 1, a_bit, T, G)
            q.alpha_activ.data.copy_(torch.randn(T, G, C, generator=g) * 0.5)
            q.set_calibrate(True)
            q.first_calibrate(first)
            x = torch.randn(*shape, generator=g) * amp * (0.2 + torch.rand(1, C, 1, 1, generator=g))
            with torch.no_grad():
                y0 = q._quantize_activation(x)
                y1 = q._quantize_activation(x * 0.7)
            key = f"cal{ci}_{int(first)}"
            out[key + "_x"] = _np(x)
            out[key + "_alpha"] = _np(q.alpha_activ.data)
            out[key + "_y0"] = _np(y0)
            out[key + "_y1"] = _np(y1)
            out[key + "_gr"] = _np(q.groups_range.data)
            out[key + "_init"] = np.stack([_np(q.init_range_min), _np(q.init_range_max)])
            out[key + "_meta"] = np.array([C, a_bit, G, T])
    # group-wise
    vecs = []
    for n in (3, 16, 128, 257):
        for _ in range(6):
            vecs.append(torch.randn(n, generator=g) * 3)
    vecs.append(torch.full((32,), -4.0))                     # degenerate max == min
    v = torch.full((64,), 6.0); v[::7] = 9.25; vecs.append(v)  # ties at the bottom edge
    v = torch.randn(128, generator=g).clamp(min=-4.0); vecs.append(torch.minimum(v, torch.tensor(-4.0)))
    v = torch.randn(40, generator=g) * 10; v[5] = v.max(); vecs.append(v)
    for vi, v in enumerate(vecs):
        for G in (4, 8):
            for mm in ("max", "min"):
                xq, gm = ref.qu.GroupWise_Quantizaion(v.clone(), dim=v.numel(), group_n=G, maxmin=mm)
                out[f"gw{vi}_{G}_{mm}_xq"] = _np(xq)
                out[f"gw{vi}_{G}_{mm}_gm"] = _np(gm)
        out[f"gw{vi}_x"] = _np(v)
    out["gw_count"] = np.array([len(vecs)])
    # weight clamp + snap
    for wi, (shape, bits) in enumerate([((8, 5, 3, 3), 8), ((16, 32, 1, 1), 4), ((3, 64, 3, 3), 6)]):
        w = torch.randn(*shape, generator=g) * 0.1
        q = make_qconv(ref, shape[1], shape[0], shape[2], 8, T, 8)
        lo = w.reshape(shape[0], -1).min(1)[0] * 0.6
        hi = w.reshape(shape[0], -1).max(1)[0] * 0.7
        q.weight_range_min, q.weight_range_max = lo, hi
        out[f"wc{wi}_w"] = _np(w)
        out[f"wc{wi}_lo"], out[f"wc{wi}_hi"] = _np(lo), _np(hi)
        out[f"wc{wi}_out"] = _np(q._quantize_weight(w))
        flat = w.reshape(shape[0], -1)
        out[f"wc{wi}_snap"] = _np(ref.qz.AsymmetricQuantFunction.apply(w, bits, flat.min(1)[0], flat.max(1)[0]))
        out[f"wc{wi}_bits"] = np.array([bits])
    # attention-internal quantizer
    mpa = ref.aq.MixedPrecisionAttention(head_dim=4, num_heads=8, bit_width=4)
    for ai, (bits, sc, zp) in enumerate([(4, 0.37, 7.3), (6, 0.05, 31.0), (3, 1.0 / 15, 0.0), (8, 0.011, 128.5)]):
        x = torch.randn(2, 8, 16, 16, generator=g) * 2
        y = mpa.quantize_tensor(x, torch.tensor([sc]), torch.tensor([zp]), bits)
        out[f"aq{ai}_x"], out[f"aq{ai}_y"] = _np(x), _np(y)
        out[f"aq{ai}_p"] = np.array([bits, sc, zp], dtype=np.float64)
    # percentile helpers
    x = torch.randn(50000, generator=g)
    out["pct_x"] = _np(x)
    out["pct_min"] = np.array([ref.qu.find_scale_by_percentile_min(x)])
    out["pct_max"] = np.array([ref.qu.find_scale_by_percentile_max(x)])
    np.savez_compressed(os.path.join(OUT, "quant_unit.npz"), **out)
    print("quant_unit.npz", len(out), "arrays")


def ddim_unit(ref):
    out = {}
    betas = H.betas(H.tiny_config())
    for ci, (T, eta) in enumerate([(5, 0.0), (10, 0.0), (4, 0.5), (20, 1.0)]):
        seq = range(0, 1000, 1000 // T)
        x = torch.randn(3, 3, 4, 4, generator=torch.Generator().manual_seed(5 + ci))

        def model(xt, t):
            return 0.3 * xt + torch.sin(t / 100.0).view(-1, 1, 1, 1) * 0.1

        torch.manual_seed(77)
        xs, x0s = ref.dn.generalized_steps(x, seq, model, betas, eta=eta)
        out[f"d{ci}_x"] = _np(x)
        out[f"d{ci}_xs"] = np.stack([_np(t) for t in xs])
        out[f"d{ci}_x0"] = np.stack([_np(t) for t in x0s])
        out[f"d{ci}_meta"] = np.array([T, eta], dtype=np.float64)
    out["abar"] = _np(ref.dn.compute_alpha(betas, torch.arange(-1, 1000)).view(-1))
    np.savez_compressed(os.path.join(OUT, "ddim_unit.npz"), **out)
    print("ddim_unit.npz")


def tiny_unet(ref, name, bitwidth, alpha_mode, weight_gain, T=4, batch=2, first=False):
    spec = S.tiny_spec(T=T, bitwidth=bitwidth)
    cfg = H.tiny_config(ch=spec.ch, ch_mult=spec.ch_mult, num_res_blocks=spec.num_res_blocks,
                        image_size=spec.image_size)
    m, seq, args = H.build_model(cfg, T, bitwidth, seed=0, snap_weights=False)
    x = torch.randn(batch, 3, spec.image_size, spec.image_size, generator=torch.Generator().manual_seed(123))
    with torch.no_grad():
        m(x, torch.zeros(batch))                    # creates the lazy channel_proj convs
    sd = S.synth_state_dict(spec, seed=3, weight_gain=weight_gain, alpha_mode=alpha_mode)
    missing, unexpected = m.load_state_dict(sd, strict=True)
    H.fix_model(m, snap_weights=False)               # weights are already on-grid (synth snap=True)
    H.reset_index(m)
    betas = H.betas(cfg)
    out = {"x": _np(x), "digest": np.frombuffer(bytes.fromhex(S.state_digest(sd)), dtype=np.uint8),
           "meta": np.array([T, bitwidth, batch, 3, weight_gain], dtype=np.float64)}
    eps = []
    def _eps_hook(mod, i, o):
        eps.append(_np(o))
    hook = m.register_forward_hook(_eps_hook)
    # FP-free: calibration pass on the same x0
    H.set_calibrate(m, True, first=first)
    with torch.no_grad():
        xs_c, _ = ref.dn.generalized_steps(x, seq, m, betas, eta=0.0)
    out["calib_eps"] = np.stack(eps); eps.clear()
    out["calib_final"] = _np(xs_c[-1])
    H.set_calibrate(m, False)
    for n, q in H.qconvs(m):
        out["gr/" + n] = _np(q.groups_range.data)
        if first:
            out["init/" + n] = np.stack([_np(q.init_range_min), _np(q.init_range_max)])
    # one traced quantized forward at the first step for per-layer pins
    trace = {}
    hooks = []
    picks = ["init_conv", "down_blocks.0.res1.conv2", "down_blocks.1.res1.nin_shortcut",
             "down_blocks.1.attn.key_conv", "down_blocks.1.attn.output_conv", "down_blocks.1.time_mlp.1",
             "middle_block1.conv1", "up_blocks.3.res1.conv1", "conv_out"]
    for n, q in H.qconvs(m):
        if n in picks:
            def _h(mod, i, o, n=n):
                trace.setdefault(n, (_np(i[0]), _np(o)))
            hooks.append(q.register_forward_hook(_h))
    with torch.no_grad():
        xs, x0s = ref.dn.generalized_steps(x, seq, m, betas, eta=0.0)
    for h in hooks:
        h.remove()
    hook.remove()
    out["eps"] = np.stack(eps)
    out["xs"] = np.stack([_np(t) for t in xs])
    for n, (i, o) in trace.items():
        out["trace_in/" + n], out["trace_out/" + n] = i, o
    np.savez_compressed(os.path.join(OUT, name), **out)
    print(name, "eps rms per step:", [float(np.sqrt((e ** 2).mean())) for e in out["eps"]])


def main():
    torch.set_num_threads(1)
    os.makedirs(OUT, exist_ok=True)
    ref = H.load()
    quant_unit(ref)
    ddim_unit(ref)
    tiny_unet(ref, "tiny_unet_w8.npz", 8, "uniform", 1.0)
    tiny_unet(ref, "tiny_unet_w8_scaled.npz", 8, "uniform", 0.5, first=True)
    tiny_unet(ref, "tiny_unet_w4_attn.npz", 4, "attn_random", 1.0)


if __name__ == "__main__":
    main()
