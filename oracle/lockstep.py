"""Lock-step parity analysis (TEST INFRASTRUCTURE): one teacher-forced denoising step run through the CUDA
path and through the oracle, compared layer by layer -- integer codes, not just outputs.

Why: a fake-quantized network is chaotic at the LSB scale (one activation that lands on the other side of a
rounding boundary changes 9*C_out conv outputs, which flips more codes downstream).  A whole-step eps error
therefore says little by itself; what can be held to a hard bar is
  * seed flips  -- codes our quantizer kernels produce vs the oracle arithmetic on the SAME layer input
                   (includes the GroupNorm+SiLU producer): the kernel-level "bit-exact codes" claim, in situ;
  * trajectory flips -- codes of the CUDA run vs codes of the oracle's own whole-network run; a step in which
                   no code differs anywhere must agree to fp32 summation noise (<= 1e-3 rel-L2);
  * the same two numbers for the reference arithmetic itself in CUDA eager vs CPU (the comparator): how far the
    reference diverges from itself across devices.
Nothing here is imported by the product path."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import restate as R

PRE_NONE, PRE_SILU, PRE_GN_SILU = 0, 1, 2


def record_layers_with_codes(model):
    """Wrap every QConv2d.forward_fused of the CUDA model: record the raw layer input, the producer, the output
    and the integer codes / de-quantized values the CUDA quantizer kernels produce for that call."""
    from attentiondm_b200 import ops
    rec = []
    for n, q in model.qconvs():
        orig = q.forward_fused

        def wrap(x, pre=PRE_NONE, gn=None, residual=None, temb=None, want_stats=False, _o=orig, _n=n, _q=q):
            t = 0 if _q.index_seq >= _q.args.timesteps else _q.index_seq
            tb = _q._tables()
            row, lay = tb["tab"][t], tb["lay"]
            B, H, W, C = x.shape
            i8 = bool(tb["i8_ok"][t]) and not _q.force_f32
            codes, _, yq = ops.act_quant(x, row[lay["scale"]:], row[lay["zp"]:], _q._a_bit, pre, gn,
                                         want_codes=i8, halo=False, want_f32=not i8)
            if i8:
                codes = codes[:, :C].reshape(B, H, W, C).cpu()
            else:
                yq = yq.cpu()
            y = _o(x, pre, gn, residual, temb, want_stats)
            rec.append(dict(name=_n, t=t, x=x.cpu(), pre=pre,
                            gn=(gn.gamma.cpu(), gn.beta.cpu(), gn.eps) if gn is not None else None,
                            residual=None if residual is None else residual.cpu(),
                            temb=None if temb is None else temb.cpu(), y=y.cpu(),
                            codes=codes if i8 else None, yq=None if i8 else yq))
            return y
        q.forward_fused = wrap
    return rec


def unwrap(model):
    for _, q in model.qconvs():
        if "forward_fused" in q.__dict__:
            del q.__dict__["forward_fused"]


def _producer(x_nchw, r):
    if r["pre"] == PRE_GN_SILU:
        gam, bet, eps = r["gn"]
        return F.silu(F.group_norm(x_nchw, 32, gam.to(x_nchw.device), bet.to(x_nchw.device), eps=eps))
    if r["pre"] == PRE_SILU:
        return F.silu(x_nchw)
    return x_nchw


def analyze_step(model, orc: R.Oracle, x_t, t_value, step, cuda_comparator=True):
    """Teacher-forced step `step` (index_seq == step on both sides) on the same x_t [B,C,H,W] (CPU tensor).
    Returns dict(eps_rel, eps_rel_torch_cuda, layers=[...], totals)."""
    dev = next(model.parameters()).device
    B = x_t.shape[0]
    tt = torch.full((B,), float(t_value))
    # ---- oracle whole-network run on the CPU, traced ----
    orc.set_index(step)
    orc.trace = {}
    with torch.no_grad():
        eps_o = orc.forward(x_t, tt)
    trace = orc.trace
    orc.trace = None
    # ---- CUDA run, recorded ----
    model.reset_index_seq(step)
    rec = record_layers_with_codes(model)
    try:
        with torch.no_grad():
            eps_c = model(x_t.to(dev), tt.to(dev)).float().cpu()
    finally:
        unwrap(model)
    mods = dict(model.qconvs())
    layers = []
    tot = dict(elements=0, seed_flips=0, traj_flips=0, seed_flips_torch_cuda=0)
    for r in rec:
        q = mods[r["name"]]
        t = r["t"]
        gr, al = q.groups_range.data.cpu()[t], q.alpha_activ.data.cpu()[t]
        xc = r["x"].permute(0, 3, 1, 2)                           # the CUDA path's own layer input, NCHW
        xp = _producer(xc, r)
        yq_ref, codes_ref, scale, zp = R.act_fake_quant(xp, gr, al, q.a_bit, return_codes=True)
        if r["codes"] is not None:
            ours = r["codes"].permute(0, 3, 1, 2).float()
            seed = int((ours != codes_ref).sum())
        else:                                                    # fp32 path (non-uniform alpha / off-grid weights)
            ours_y = r["yq"].permute(0, 3, 1, 2)
            lsb = (1.0 / scale).reshape(1, -1, 1, 1) if scale.dim() else 1.0 / scale
            seed = int(((ours_y - yq_ref).abs() > 0.25 * lsb).sum())
            ours = None
        # trajectory: the oracle's own run reached this layer with input x_o
        x_o, y_o = trace[r["name"]][0]
        _, codes_o, _, _ = R.act_fake_quant(x_o, gr, al, q.a_bit, return_codes=True)
        if ours is not None:
            traj = int((ours != codes_o).sum())
        else:
            traj = int((codes_ref != codes_o).sum()) + seed
        # in-situ operator error: oracle arithmetic on the CUDA path's input vs the CUDA output
        w = R.weight_clamp(q.weight.data.cpu(), q.weight_range_min.cpu(), q.weight_range_max.cpu())
        want = F.conv2d(yq_ref, w, q.bias.data.cpu(), padding=q.kernel_size[0] // 2)
        if want.shape[-2:] != r["y"].shape[1:3]:                 # 3x3 on a 1x1 map etc. never changes the size
            raise AssertionError((r["name"], want.shape, r["y"].shape))
        if r["residual"] is not None:
            want = want + r["residual"].permute(0, 3, 1, 2)
        if r["temb"] is not None:
            want = want + r["temb"][:, :, None, None]
        got = r["y"].permute(0, 3, 1, 2)
        insitu = float((got.double() - want.double()).norm() / want.double().norm().clamp_min(1e-30))
        out_rel = float((got.double() - y_o.double()).norm() / y_o.double().norm().clamp_min(1e-30)) \
            if r["residual"] is None and r["temb"] is None else None
        seed_tc = None
        if cuda_comparator:
            # the reference arithmetic in CUDA eager on the same input: producer + quantizer through torch's CUDA kernels
            xg = r["x"].to(dev).permute(0, 3, 1, 2)
            _, codes_g, _, _ = R.act_fake_quant(_producer(xg, r), gr.to(dev), al.to(dev), q.a_bit, return_codes=True)
            seed_tc = int((codes_g.cpu() != codes_ref).sum())
            tot["seed_flips_torch_cuda"] += seed_tc
        n = xc.numel()
        tot["elements"] += n
        tot["seed_flips"] += seed
        tot["traj_flips"] += traj
        layers.append(dict(name=r["name"], elements=n, pre=r["pre"], seed_flips=seed, traj_flips=traj,
                           seed_flips_torch_cuda=seed_tc, insitu_rel=insitu, out_rel_vs_oracle_run=out_rel))
    eps_rel = float((eps_c.double() - eps_o.double()).norm() / eps_o.double().norm())
    out = dict(step=step, t=float(t_value), eps_rel=eps_rel, layers=layers, totals=tot,
               first_traj_flip=next((l["name"] for l in layers if l["traj_flips"]), None),
               first_seed_flip=next((l["name"] for l in layers if l["seed_flips"]), None),
               worst_insitu=max(l["insitu_rel"] for l in layers))
    if cuda_comparator:
        # the reference arithmetic, whole network, CUDA eager (fp32, TF32 off) vs its own CPU run
        saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        try:
            og = R.Oracle(orc.spec, orc.sd, qs=None)
            og.mixed_precision, og.mp_state = orc.mixed_precision, orc.mp_state
            og.to(dev)
            og.set_index(step)
            with torch.no_grad():
                eps_g = og.forward(x_t.to(dev), tt.to(dev)).cpu()
            del og
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved
        out["eps_rel_torch_cuda"] = float((eps_g.double() - eps_o.double()).norm() / eps_o.double().norm())
    return out, eps_o


def summarize(steps):
    """Compact per-config summary of analyze_step results."""
    el = sum(s["totals"]["elements"] for s in steps)
    sf = sum(s["totals"]["seed_flips"] for s in steps)
    tc = sum(s["totals"]["seed_flips_torch_cuda"] for s in steps)
    return dict(
        steps=len(steps), activations_compared=el,
        seed_flips=sf, seed_flip_rate=sf / max(1, el),
        seed_flips_torch_cuda=tc, seed_flip_rate_torch_cuda=tc / max(1, el),
        eps_rel_per_step=[s["eps_rel"] for s in steps],
        eps_rel_torch_cuda_per_step=[s.get("eps_rel_torch_cuda") for s in steps],
        traj_flips_per_step=[s["totals"]["traj_flips"] for s in steps],
        steps_under_1e_3=sum(1 for s in steps if s["eps_rel"] <= 1e-3),
        flip_free_steps=sum(1 for s in steps if s["totals"]["traj_flips"] == 0),
        worst_insitu_rel=max(s["worst_insitu"] for s in steps),
        first_traj_flip_layer=[s["first_traj_flip"] for s in steps],
    )
