"""In-situ operator parity (TEST INFRASTRUCTURE): record every QConv2d call of the CUDA path and
replay the reference arithmetic (oracle/restate.py) on the SAME layer input on the CPU.

Whole-network comparisons of a fake-quantized UNet are chaotic at the LSB scale (one flipped code
avalanches to ~1e-2 within three layers), so operator-level parity along the CUDA path's own
trajectory is the check that can hold a 1e-3 bar at every layer of every step."""
import torch
import torch.nn.functional as F

from . import restate as R

PRE_NONE, PRE_SILU, PRE_GN_SILU = 0, 1, 2


def record_layers(model):
    rec = []
    for n, q in model.qconvs():
        orig = q.forward_fused

        def wrap(x, pre=PRE_NONE, gn=None, residual=None, temb=None, want_stats=False, _o=orig, _n=n, _q=q):
            t = 0 if _q.index_seq >= _q.args.timesteps else _q.index_seq
            y = _o(x, pre, gn, residual, temb, want_stats)
            rec.append(dict(name=_n, t=t, x=x.cpu(), pre=pre,
                            gn=(gn.gamma.cpu(), gn.beta.cpu(), gn.eps) if gn is not None else None,
                            residual=None if residual is None else residual.cpu(),
                            temb=None if temb is None else temb.cpu(), y=y.cpu()))
            return y
        q.forward_fused = wrap
    return rec


def oracle_layer(q, r, calibrate):
    """Reference arithmetic of one recorded call: producer (GroupNorm+SiLU / SiLU), activation
    fake-quant or calibration mix, weight clamp, conv2d, and the fused adds."""
    x = r["x"].permute(0, 3, 1, 2)
    if r["pre"] == PRE_GN_SILU:
        gam, bet, eps = r["gn"]
        x = F.silu(F.group_norm(x, 32, gam, bet, eps=eps))
    elif r["pre"] == PRE_SILU:
        x = F.silu(x)
    t = r["t"]
    gr, al = q.groups_range.data.cpu(), q.alpha_activ.data.cpu()
    if calibrate:
        xq, gr_t = R.calibrate_activation(x, al[t], q.group_num, q.a_bit, q.init_range_min[t], q.init_range_max[t])
        if not torch.allclose(gr_t, gr[t], rtol=1e-5, atol=1e-6):
            raise AssertionError((r["name"], "group table differs", gr_t, gr[t]))
    else:
        xq = R.act_fake_quant(x, gr[t], al[t], q.a_bit)
    w = R.weight_clamp(q.weight.data.cpu(), q.weight_range_min.cpu(), q.weight_range_max.cpu())
    y = F.conv2d(xq, w, q.bias.data.cpu(), padding=q.kernel_size[0] // 2)
    if r["residual"] is not None:
        y = y + r["residual"].permute(0, 3, 1, 2)
    if r["temb"] is not None:
        y = y + r["temb"][:, :, None, None]
    return y


def layer_errors(model, rec, calibrate):
    """[(rel-L2 error, (layer name, step))] of every recorded call against the reference arithmetic."""
    mods = dict(model.qconvs())
    out = []
    for r in rec:
        want = oracle_layer(mods[r["name"]], r, calibrate)
        got = r["y"].permute(0, 3, 1, 2)
        e = float((got.double() - want.double()).norm() / want.double().norm().clamp_min(1e-30))
        out.append((e, (r["name"], r["t"])))
    return out


def worst_layer_error(model, rec, calibrate):
    worst, where = 0.0, None
    for e, w in layer_errors(model, rec, calibrate):
        if e > worst:
            worst, where = e, w
    return worst, where
