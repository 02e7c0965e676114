"""Generate tests/golden/tiny_calib.npz by running the UNMODIFIED reference on CPU: the callers either side of the
hot path.

    python -m oracle.make_golden_calib          (build container only: needs /root/reference)

Contents:
  fp_*     the FP model (Model(quantization=False), models/diffusion.py:106-116,281-345): eps of two forwards and a
           DDIM trajectory through it;
  ddpm_*   functions/denoising.py:119-151 (ddpm_steps) on a closed-form model with recorded noise;
  ca_*     runners/diffusion.py:266-306 (calibrate_attention -> generalized_steps_loss, functions/denoising.py:62-116)
           on the tiny UNet: recorded noise draws, the returned trajectories, alpha_activ / groups_range of the
           attention convs after the AdamW steps, the per-step losses -- and the measured fact the product relies
           on: the gradient of the noise-estimation loss w.r.t. every attention alpha_activ is EXACTLY zero
           (torch.round in every downstream quantizer, utils/quant_util.py:271, has a zero derivative and conv_out
           is a QConv2d), so the entropy regulariser is the only non-zero gradient source;
  gcs_*    runners/diffusion.py:198-264 (generate_calibrate_set) for t_mode real / range / random / diff.
TEST INFRASTRUCTURE ONLY (see oracle/ref_harness.py).
"""
import argparse
import os
import warnings

import numpy as np
import torch
import torch.nn.functional as F

from . import ref_harness as H
from . import synth as S

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
warnings.filterwarnings("ignore")


def _np(t):
    return t.detach().cpu().numpy().copy()      # a copy: later in-place updates must not reach the saved array


class NoiseTape:
    """Records every torch.randn_like / torch.randn draw of the code under test (in order)."""

    def __init__(self):
        self.draws = []

    def __enter__(self):
        self._rl, self._rn = torch.randn_like, torch.randn
        tape = self

        def randn_like(x, *a, **k):
            v = tape._rl(x, *a, **k)
            tape.draws.append(v.detach().clone())
            return v

        def randn(*a, **k):
            v = tape._rn(*a, **k)
            tape.draws.append(v.detach().clone())
            return v

        torch.randn_like, torch.randn = randn_like, randn
        return self

    def __exit__(self, *exc):
        torch.randn_like, torch.randn = self._rl, self._rn


def fp_subset(sd):
    """The FP model's state_dict: the same tensors without the quantizer state."""
    return {k: v for k, v in sd.items() if not (k.endswith(".groups_range") or k.endswith(".alpha_activ"))}


def build_pair(ref, T=4, bitwidth=8):
    spec = S.tiny_spec(T=T, bitwidth=bitwidth)
    cfg = H.tiny_config(ch=spec.ch, ch_mult=spec.ch_mult, num_res_blocks=spec.num_res_blocks, image_size=spec.image_size)
    m, seq, args = H.build_model(cfg, T, bitwidth, seed=0, snap_weights=False)
    x = torch.randn(2, 3, spec.image_size, spec.image_size, generator=torch.Generator().manual_seed(123))
    with torch.no_grad():
        m(x, torch.zeros(2))                        # creates the lazy channel_proj convs
    sd = S.synth_state_dict(spec, seed=3, weight_gain=0.5)
    m.load_state_dict(sd, strict=True)
    H.fix_model(m, snap_weights=False)
    H.reset_index(m)
    fp = ref.md.Model(cfg, quantization=False, sequence=seq, args=args).eval()
    with torch.no_grad():
        fp(x, torch.zeros(2))
    fp.load_state_dict(fp_subset(sd), strict=True)
    return spec, cfg, m, fp, seq, args, x, sd


def bare_runner(ref, cfg, seq, betas, **args):
    r = object.__new__(ref.rd.Diffusion)
    r.args = argparse.Namespace(**args)
    r.config = cfg
    r.device = torch.device("cpu")
    r.seq = seq
    r.betas = betas
    return r


def main():
    torch.set_num_threads(1)
    ref = H.load()
    out = {}
    spec, cfg, m, fp, seq, args, x, sd = build_pair(ref)
    betas = H.betas(cfg)
    out["x"] = _np(x)
    out["digest"] = np.frombuffer(bytes.fromhex(S.state_digest(sd)), dtype=np.uint8)

    # ---- FP model ----
    with torch.no_grad():
        out["fp_eps_t0"] = _np(fp(x, torch.zeros(2)))
        out["fp_eps_t750"] = _np(fp(x, torch.full((2,), 750.0)))
        xs, x0s = ref.dn.generalized_steps(x, seq, fp, betas, eta=0.0)
    out["fp_xs"] = np.stack([_np(t) for t in xs])

    # ---- ddpm_steps (its `.to('cuda')` at :133 is patched to a no-op for this CPU run; nothing else touched) ----
    def toy(xt, t):
        return 0.3 * xt + torch.sin(t / 100.0).view(-1, 1, 1, 1) * 0.1
    xd = torch.randn(3, 3, 4, 4, generator=torch.Generator().manual_seed(9))
    seq10 = range(0, 1000, 100)
    orig_to = torch.Tensor.to
    def to_nocuda(self, *a, **k):
        if a and isinstance(a[0], str) and a[0] == "cuda":
            return self
        return orig_to(self, *a, **k)
    torch.Tensor.to = to_nocuda
    try:
        torch.manual_seed(31)
        with NoiseTape() as tape:
            xs, x0s = ref.dn.ddpm_steps(xd, seq10, toy, betas)
    finally:
        torch.Tensor.to = orig_to
    out["ddpm_x"] = _np(xd)
    out["ddpm_noise"] = np.stack([_np(t) for t in tape.draws])
    out["ddpm_xs"] = np.stack([_np(t) for t in xs])
    out["ddpm_x0"] = np.stack([_np(t) for t in x0s])

    # ---- general calibration pass, then calibrate_attention ----
    H.set_calibrate(m, True)
    with torch.no_grad():
        ref.dn.generalized_steps(x, seq, m, betas, eta=0.0)
    H.set_calibrate(m, False)
    H.reset_index(m)
    attn_names = [n for n, q in H.qconvs(m) if any(s in n for s in ("query_conv", "key_conv", "value_conv", "output_conv"))]
    g = torch.Generator().manual_seed(41)
    for n, q in H.qconvs(m):
        if n in attn_names:                         # a non-trivial starting point for the optimiser
            q.alpha_activ.data.copy_(0.3 * torch.randn(q.alpha_activ.shape, generator=g))
    out["ca_attn_names"] = np.array(attn_names)
    for n, q in H.qconvs(m):
        out["ca_gr0/" + n] = _np(q.groups_range.data)
        if n in attn_names:
            out["ca_alpha0/" + n] = _np(q.alpha_activ.data)

    # the gradient of the noise-estimation loss alone w.r.t. the attention alphas (one step, no optimiser)
    for n, q in H.qconvs(m):
        q.set_calibrate(n in attn_names)
        q.alpha_activ.grad = None
    torch.manual_seed(5)
    t = torch.ones(2) * list(seq)[-1]
    loss, _ = ref.dn.noise_estimation_loss(m, x, t, torch.randn_like(x), betas)
    loss.backward()
    gmax = max(float(q.alpha_activ.grad.abs().max()) for n, q in H.qconvs(m) if n in attn_names and q.alpha_activ.grad is not None)
    nn_ = sum(1 for n, q in H.qconvs(m) if n in attn_names and q.alpha_activ.grad is not None)
    g_out = float(dict(H.qconvs(m))["conv_out"].alpha_activ.grad.abs().max())
    print(f"main-loss gradient on attention alphas: max |g| = {gmax} over {nn_} tensors (conv_out alpha: {g_out:.3e})")
    assert gmax == 0.0
    out["ca_mainloss_grad_max"] = np.array([gmax, g_out])
    for n, q in H.qconvs(m):
        q.alpha_activ.grad = None
        q.set_calibrate(False)
    # restore the state the probe touched (it ran one calibrating forward on the attention convs at index 0)
    for n, q in H.qconvs(m):
        q.groups_range.data.copy_(torch.from_numpy(out["ca_gr0/" + n]))
    H.reset_index(m)

    runner = bare_runner(ref, cfg, seq, betas, eta=0.0, diff_loss_weight=0.5, timesteps=len(list(seq)))
    runner.t_mode, runner.timestep_select, runner.first_flag = "real", None, False
    losses = []
    orig_nel = ref.dn.noise_estimation_loss
    def nel(*a, **k):
        r = orig_nel(*a, **k)
        losses.append(float(r[0]))
        return r
    ref.dn.noise_estimation_loss = nel
    torch.manual_seed(2024)
    with NoiseTape() as tape:
        runner.calibrate_attention(m, x, "cpu", 2)
    ref.dn.noise_estimation_loss = orig_nel
    out["ca_noise"] = np.stack([_np(t) for t in tape.draws])            # per step: e, then the DDIM noise
    out["ca_loss"] = np.array(losses)
    for n, q in H.qconvs(m):
        if n in attn_names:
            out["ca_alpha1/" + n] = _np(q.alpha_activ.data)
            out["ca_gr1/" + n] = _np(q.groups_range.data)
    out["ca_meta"] = np.array([0.05, 0.05, 0.5, 0.0])                     # AdamW lr, weight_decay; diff_loss_weight; eta
    print("calibrate_attention: losses", losses, "alpha moved by",
          max(float(np.abs(out["ca_alpha1/" + n] - out["ca_alpha0/" + n]).max()) for n in attn_names))

    # ---- generate_calibrate_set (T = 40: the `diff` branch drops the first 30 timesteps, :242-243) ----
    spec40, cfg40, m40, fp40, seq40, args40, x40, sd40 = build_pair(ref, T=40)
    cfg40.data.logit_transform, cfg40.data.rescaled = False, True          # configs/cifar10.yml
    g = torch.Generator().manual_seed(43)
    for n, q in H.qconvs(m40):                                              # non-uniform alphas: a non-trivial entropy profile
        q.alpha_activ.data.copy_(torch.randn(q.alpha_activ.shape, generator=g))
    out["gcs_alpha_seed"] = np.array([43])
    Tn = len(list(seq40))
    for mode in ("real", "range", "random", "diff"):
        r = bare_runner(ref, cfg40, seq40, betas, eta=0.0, timesteps=Tn, sample_weight=0.3)
        r.sample_count = torch.zeros(Tn)
        r.sample_count[35] = 2.0
        torch.manual_seed(77)
        with NoiseTape() as tape:
            cs = r.generate_calibrate_set(fp40, m40, mode, 4)
        out[f"gcs_{mode}"] = _np(cs)
        out[f"gcs_{mode}_noise0"] = _np(tape.draws[0])
        if mode == "diff":
            out["gcs_diff_t"] = np.array([int(r.timestep_select)])
            out["gcs_diff_count"] = _np(r.sample_count)
        print("generate_calibrate_set", mode, "ok", tuple(cs.shape), "draws", len(tape.draws))
    np.savez_compressed(os.path.join(OUT, "tiny_calib.npz"), **out)
    print("tiny_calib.npz", len(out), "arrays")


if __name__ == "__main__":
    main()
