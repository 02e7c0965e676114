"""GPU parity of the callers either side of the sampler against tests/golden/tiny_calib.npz
(generated from the UNMODIFIED reference by oracle/make_golden_calib.py):
  f2  the FP model and generate_calibrate_set (t_mode real / range / random / diff),
  f1  calibrate_attention -> generalized_steps_loss (entropy-regularised AdamW on the attention alphas),
  f4  ddpm_steps (bit-exact) and the ablation driver's bit-width variants,
  f3  quantizer-state checkpoints: a reloaded model samples bit-identically.
Tolerances are stated per test; the fp32 conv kernels differ from MKLDNN/cuDNN by summation order only."""
import argparse
import os

import numpy as np
import pytest
import torch

from oracle import restate as R
from oracle import synth as S
from tests.util import T, args_for, build_cuda_model, config_for, rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _fp_subset(sd):
    return {k: v for k, v in sd.items() if not (k.endswith(".groups_range") or k.endswith(".alpha_activ"))}


def _pair(Tn=4):
    spec = S.tiny_spec(T=Tn, bitwidth=8)
    return spec, S.synth_state_dict(spec, seed=3, weight_gain=0.5)


def _fp_model(spec, sd):
    import attentiondm_b200 as A
    m = A.Model(config_for(spec), quantization=False, sequence=spec.seq, args=args_for(spec)).to(DEV).eval()
    m.materialize_lazy_layers()
    m.load_state_dict(_fp_subset(sd), strict=True)
    return m


def _runner(spec, **extra):
    import attentiondm_b200 as A
    cfg = config_for(spec)
    args = argparse.Namespace(bitwidth=spec.bitwidth, timesteps=spec.timesteps, skip_type="uniform", eta=0.0, **extra)
    r = A.Diffusion(args, cfg, torch.device(DEV))
    r.seq = list(spec.seq)
    return r


def test_fp_model_matches_reference(golden):
    """f2 prerequisite: Model(quantization=False) through the fp32 kernels vs the reference FP model (CPU): 1e-5."""
    import attentiondm_b200 as A
    g = golden("tiny_calib.npz")
    spec, sd = _pair()
    m = _fp_model(spec, sd)
    x = T(g["x"]).to(DEV)
    e0 = rel_l2(m(x, torch.zeros(2, device=DEV)), T(g["fp_eps_t0"]))
    e1 = rel_l2(m(x, torch.full((2,), 750.0, device=DEV)), T(g["fp_eps_t750"]))
    xs, _ = A.generalized_steps(x, spec.seq, m, R.beta_schedule_linear().to(DEV), eta=0.0)
    et = max(rel_l2(a, T(b)) for a, b in zip(xs, g["fp_xs"]))
    print(f"\n[fp model] eps rel-L2 {e0:.2e} / {e1:.2e}; trajectory worst {et:.2e}")
    assert e0 < 1e-5 and e1 < 1e-5 and et < 1e-5


def test_ddpm_steps_bit_exact(golden):
    """f4: the ancestral sampler with the reference's recorded noise -- torch.equal on every x_t and x0."""
    import attentiondm_b200 as A
    g = golden("tiny_calib.npz")
    noise = T(g["ddpm_noise"]).to(DEV)

    def toy(xt, t):                                 # evaluated on the CPU: the update kernel, not torch's sin, is under test
        return (0.3 * xt.cpu() + torch.sin(t.cpu() / 100.0).view(-1, 1, 1, 1) * 0.1).to(DEV)
    xs, x0s = A.ddpm_steps(T(g["ddpm_x"]).to(DEV), range(0, 1000, 100), toy, R.beta_schedule_linear().to(DEV),
                           noise_fn=lambda k, like: noise[k])
    assert len(xs) == 11 and len(x0s) == 10 and not xs[1].is_cuda
    assert all(torch.equal(a.cpu(), T(b)) for a, b in zip(xs, g["ddpm_xs"]))
    assert all(torch.equal(a.cpu(), T(b)) for a, b in zip(x0s, g["ddpm_x0"]))


def _load_ca_state(m, g):
    for n, q in m.qconvs():
        q.groups_range.data.copy_(T(g["ca_gr0/" + n]).to(DEV))
        if "ca_alpha0/" + n in g.files:
            q.alpha_activ.data.copy_(T(g["ca_alpha0/" + n]).to(DEV))
        q.invalidate_cache(weights=False)
    m.reset_index_seq()


def test_calibrate_attention_matches_reference(golden):
    """f1: runners/diffusion.py:266-306 on the tiny UNet, same start state and noise draws as the reference run.
    alpha_activ after the four AdamW steps: 2e-5 absolute (AdamW's normalised update is insensitive to the 1e-7
    differences of the entropy gradient; measured 1.8e-5); the attention convs' calibrated group tables: 2e-3 relative
    (they are the min/max of activations along a fake-quantized trajectory, which inherits the code-flip sensitivity
    documented in tests/test_gpu_lockstep.py; measured 5e-4); per-step losses 2e-3 relative."""
    import attentiondm_b200 as A
    from attentiondm_b200 import ops
    g = golden("tiny_calib.npz")
    spec, sd = _pair()
    m = build_cuda_model(spec, sd)
    _load_ca_state(m, g)
    lr, wd, w, eta = [float(v) for v in g["ca_meta"]]
    r = _runner(spec, diff_loss_weight=w)
    r.betas = R.beta_schedule_linear().to(DEV)
    noise = T(g["ca_noise"]).to(DEV)
    losses = []
    r.calibrate_attention(m, T(g["x"]).to(DEV), noise_fn=lambda k, which, like: noise[2 * k + (0 if which == "e" else 1)],
                          losses=losses)
    names = [str(n) for n in g["ca_attn_names"]]
    mods = dict(m.qconvs())
    worst_a = max(float((mods[n].alpha_activ.data.cpu() - T(g["ca_alpha1/" + n])).abs().max()) for n in names)
    moved = max(float((T(g["ca_alpha1/" + n]) - T(g["ca_alpha0/" + n])).abs().max()) for n in names)
    worst_g = max(rel_l2(mods[n].groups_range.data, T(g["ca_gr1/" + n])) for n in names)
    ls = [float(v) for v in losses]
    print(f"\n[calibrate_attention] alpha moved by {moved:.3f} in the reference; ours differs by {worst_a:.2e}; "
          f"group tables rel {worst_g:.2e}; losses {ls} vs {list(g['ca_loss'])}")
    per = {n: (float((mods[n].alpha_activ.data.cpu() - T(g["ca_alpha1/" + n])).abs().max()),
               rel_l2(mods[n].groups_range.data, T(g["ca_gr1/" + n]))) for n in names}
    print("  per layer (alpha abs, table rel):", {n.split("blocks.")[-1]: (f"{a:.1e}", f"{b:.1e}") for n, (a, b) in per.items()})
    assert worst_a < 5e-5 and worst_g < 2e-3
    assert np.allclose(ls, g["ca_loss"], rtol=2e-3)
    assert all(not q._calibrate for q in mods.values())
    assert len(r.last_calibration[0]) == spec.timesteps + 1 and not r.last_calibration[0][1].is_cuda
    # the entropy gradient kernel against autograd through the reference expression
    a = torch.randn(4, 8, 24, generator=torch.Generator().manual_seed(3))
    for t in range(4):
        term, grad = R.entropy_term_and_grad(a, t)
        out = torch.zeros(8, 24, device=DEV)
        val = torch.zeros(1, dtype=torch.float64, device=DEV)
        ops.alpha_entropy_grad(a[t].to(DEV).contiguous(), 0.5, out, val)
        assert float((out.cpu() - 0.5 * grad[t]).abs().max()) < 1e-9
        assert abs(float(val) - 0.5 * float(term)) < 1e-8


def test_generate_calibrate_set_matches_reference(golden):
    """f2: runners/diffusion.py:198-264, T = 40, all four t_modes with the reference's initial latent (and, for
    `random`, its timestep draw replayed from the same CPU RNG state).  1e-4: forty FP-model steps in fp32."""
    g = golden("tiny_calib.npz")
    spec, sd = _pair(40)
    fp = _fp_model(spec, sd)
    m = build_cuda_model(spec, sd)
    gen = torch.Generator().manual_seed(int(g["gcs_alpha_seed"][0]))
    for n, q in m.qconvs():                                   # same alphas as the reference run (same iteration order)
        q.alpha_activ.data.copy_(torch.randn(q.alpha_activ.shape, generator=gen).to(DEV))
    for mode in ("real", "range", "random", "diff"):
        r = _runner(spec, sample_weight=0.3)
        r.betas = R.beta_schedule_linear().to(DEV)
        r.sample_count = torch.zeros(40)
        r.sample_count[35] = 2.0
        x0 = T(g[f"gcs_{mode}_noise0"]).to(DEV)
        t_random = None
        if mode == "random":                                  # replay the reference's RNG stream up to the normal_() draw
            torch.manual_seed(77)
            torch.randn(4, 3, spec.image_size, spec.image_size)
            for _ in range(40):
                torch.randn(4, 3, spec.image_size, spec.image_size)
            nv = torch.nn.init.normal_(torch.Tensor(4), mean=0.4, std=0.4) * 40
            t_random = nv.clone().type(torch.int).clamp(0, 39)
        cs = r.generate_calibrate_set(fp, m, mode, 4, x=x0, t_random=t_random)
        err = rel_l2(cs, T(g[f"gcs_{mode}"]))
        print(f"\n[generate_calibrate_set {mode}] rel-L2 {err:.2e}" + (f" t={r.timestep_select}" if mode == "diff" else ""))
        assert cs.is_cuda and tuple(cs.shape) == (4, 3, spec.image_size, spec.image_size)
        assert err < 1e-4, mode
        if mode == "diff":
            assert r.timestep_select == int(g["gcs_diff_t"][0])
            assert torch.equal(r.sample_count, T(g["gcs_diff_count"]))


def test_quant_state_reload_samples_identically(tmp_path):
    """f3: save_quant_state -> a fresh model -> load_quant_state reproduces the calibrated sampler bit for bit
    (group tables, first-calibrate init ranges, per-layer bit widths, lazily created channel_proj)."""
    import attentiondm_b200 as A
    from attentiondm_b200 import runner
    spec, sd = _pair()
    m = build_cuda_model(spec, sd)
    betas = R.beta_schedule_linear().to(DEV)
    x = torch.randn(2, 3, spec.image_size, spec.image_size, generator=torch.Generator().manual_seed(5)).to(DEV)
    m.set_calibrate(True, first=True)
    A.generalized_steps(x, spec.seq, m, betas, eta=0.0, keep="last")
    m.set_calibrate(False)
    m.reset_index_seq()
    dict(m.qconvs())["down_blocks.2.attn.key_conv"].a_bit = 6            # a per-layer bit width (the ablation mutates them)
    ref_xs, _ = A.generalized_steps(x, spec.seq, m, betas, eta=0.0)
    m.reset_index_seq()
    path = os.path.join(tmp_path, "quant_state.pt")
    runner.save_quant_state(m, path)
    m2 = A.Model(config_for(spec), quantization=True, sequence=spec.seq, args=args_for(spec)).to(DEV).eval()
    runner.load_quant_state(m2, path)
    xs2, _ = A.generalized_steps(x, spec.seq, m2, betas, eta=0.0)
    assert all(torch.equal(a, b) for a, b in zip(ref_xs, xs2))
    q1, q2 = dict(m.qconvs()), dict(m2.qconvs())
    assert all(torch.equal(q1[n].init_range_min, q2[n].init_range_min) and q1[n]._a_bit == q2[n]._a_bit for n in q1)


def test_ablation_variants():
    """f4: ablation_study_attention_quantization.py:111-207 -- the four bit-width variants get their per-layer widths
    through the w_bit / a_bit setters, stay on the integer path, and order as expected against the FP model
    (uniform 8-bit closest, uniform 4-bit farthest) on metrics that need no external network."""
    from attentiondm_b200.ablation import AttentionQuantizationAblation, VARIANTS
    spec, sd = _pair()
    cfg = config_for(spec)
    ab = AttentionQuantizationAblation(cfg, DEV, timesteps=4)
    assert ab.sequence == list(spec.seq)
    variants = ab.prepare_model_variants({k: v for k, v in sd.items() if "groups_range" not in k and "alpha_activ" not in k})
    for name, (cb, abits) in VARIANTS.items():
        bw = ab.bit_widths(variants[name])
        attn = {n: v for n, v in bw.items() if any(s_ in n for s_ in ("query_conv", "key_conv", "value_conv", "output_conv"))}
        rest = {n: v for n, v in bw.items() if n not in attn}
        assert attn and all(v == (abits, abits) for v in attn.values()), name
        assert all(v == (cb, cb) for v in rest.values()), name
    fp = ab.build_fp_model(sd)
    x = torch.randn(4, 3, spec.image_size, spec.image_size, generator=torch.Generator().manual_seed(8)).to(DEV)
    ab.calibrate_models(variants, x)
    res = ab.evaluate(variants, fp, x)
    print("\n[ablation] " + "; ".join(f"{k}: eps {['%.3f' % e for e in v['eps_rel_l2']]} image {v['image_rel_l2']:.3f} "
                                       f"int8 {v['int8_layers']}/{v['layers']}" for k, v in res.items()))
    assert all(v["int8_layers"] == v["layers"] for v in res.values())
    assert max(res["D"]["eps_rel_l2"]) < min(res["A"]["eps_rel_l2"])          # 8-bit beats 4-bit
    imgs = ab.generate_samples({"D": variants["D"]}, num_samples=4, batch_size=4, sampler="ddpm")
    assert tuple(imgs["D"].shape) == (4, 3, spec.image_size, spec.image_size) and bool(torch.isfinite(imgs["D"]).all())
    assert float(imgs["D"].min()) >= 0.0 and float(imgs["D"].max()) <= 1.0
    assert all(v != v for v in ab.compute_fid({"A": None}).values())           # NaN, like the reference without pytorch-fid
