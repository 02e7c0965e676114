"""CPU: the oracle restatements of the callers either side of the sampler against
tests/golden/tiny_calib.npz, which oracle/make_golden_calib.py generated from the UNMODIFIED reference:
the FP model, ddpm_steps, calibrate_attention -> generalized_steps_loss, and the host logic of the product's
mirrors (coefficients, state_dict layout of the FP model)."""
import numpy as np
import torch

import attentiondm_b200 as A
from attentiondm_b200 import denoising
from oracle import restate as R
from oracle import synth as S
from tests.util import T, args_for, config_for, rel_l2


def _fp_subset(sd):
    return {k: v for k, v in sd.items() if not (k.endswith(".groups_range") or k.endswith(".alpha_activ"))}


def _pair(Tn=4):
    spec = S.tiny_spec(T=Tn, bitwidth=8)
    return spec, S.synth_state_dict(spec, seed=3, weight_gain=0.5)


def test_fixture_is_for_this_state_dict(golden):
    g = golden("tiny_calib.npz")
    spec, sd = _pair()
    assert bytes(g["digest"]).hex() == S.state_digest(sd)
    # the fact the product's gradient path relies on, measured on the reference itself
    assert float(g["ca_mainloss_grad_max"][0]) == 0.0


def test_fp_model_oracle_matches_reference(golden):
    g = golden("tiny_calib.npz")
    spec, sd = _pair()
    orc = R.Oracle(spec, sd)
    orc.quantized = False
    x = T(g["x"])
    assert rel_l2(orc.forward(x, torch.zeros(2)), T(g["fp_eps_t0"])) < 2e-6
    assert rel_l2(orc.forward(x, torch.full((2,), 750.0)), T(g["fp_eps_t750"])) < 2e-6
    xs = R.ddim_sample(orc.forward, x, spec.seq, R.beta_schedule_linear())[0]
    assert rel_l2(xs[-1], T(g["fp_xs"][-1])) < 1e-5


def test_fp_model_state_dict_layout():
    """The product's FP model takes the reference FP model's state_dict (strict), i.e. the quantized model's minus
    the quantizer tables; construction needs no CUDA."""
    spec, sd = _pair()
    m = A.Model(config_for(spec), quantization=False, sequence=spec.seq, args=args_for(spec))
    m.materialize_lazy_layers()
    m.load_state_dict(_fp_subset(sd), strict=True)
    assert not m.qconvs()
    assert all(isinstance(c, A.FConv2d) for c in (m.init_conv, m.conv_out, m.down_blocks[0].res1.conv1,
                                                  m.down_blocks[0].time_mlp[1]))


def test_ddpm_oracle_and_coefficients_bit_exact(golden):
    g = golden("tiny_calib.npz")
    betas = R.beta_schedule_linear()
    seq = list(range(0, 1000, 100))
    noise = T(g["ddpm_noise"])

    def toy(xt, t):
        return 0.3 * xt + torch.sin(t / 100.0).view(-1, 1, 1, 1) * 0.1
    xs, x0s = R.ddpm_sample(toy, T(g["ddpm_x"]), seq, betas, lambda k, like: noise[k])
    assert all(torch.equal(a, T(b)) for a, b in zip(xs, g["ddpm_xs"]))
    assert all(torch.equal(a, T(b)) for a, b in zip(x0s, g["ddpm_x0"]))
    # the product's per-step coefficient table reproduces the same trajectory with the kernel's op sequence
    coef = denoising.ddpm_coefficients(seq, betas)
    x = T(g["ddpm_x"])
    for k in range(len(seq)):
        e = toy(x, torch.full((3,), float(seq[len(seq) - 1 - k])))
        ca, cb, cc, cd, ce, cf = [coef[k, i] for i in range(6)]
        x0 = (ca * x - cb * e).clamp(-1, 1)
        x = (cc * x0 + cd * x) / ce + cf * noise[k]
        assert torch.equal(x, T(g["ddpm_xs"][k + 1])) and torch.equal(x0, T(g["ddpm_x0"][k]))
        assert float(coef[k, 6]) == float(seq[len(seq) - 1 - k])


def test_entropy_gradient_closed_form():
    """attndm_alpha_entropy_grad's formula against autograd through the reference expression."""
    a = torch.randn(4, 8, 16, generator=torch.Generator().manual_seed(2))
    for t in range(4):
        term, grad = R.entropy_term_and_grad(a, t)
        s = torch.softmax(a[t].double(), dim=0)
        ls = torch.log(s)
        m1 = (s * (ls + 1)).sum(0, keepdim=True)
        g = -(1.0 / (8 * 8 * 16)) * s * ((ls + 1) - m1)
        assert float((g.float() - grad[t]).abs().max()) < 1e-9
        assert float(grad[[i for i in range(4) if i != t]].abs().max()) == 0.0
        assert abs(float(-(s * ls).sum() / 8 / (8 * 16)) - float(term)) < 1e-7


def test_calibrate_attention_oracle_matches_reference(golden):
    g = golden("tiny_calib.npz")
    spec, sd = _pair()
    names = [str(n) for n in g["ca_attn_names"]]
    sd = dict(sd)
    for k in g.files:
        if k.startswith("ca_gr0/"):
            sd[k[len("ca_gr0/"):] + ".groups_range"] = T(g[k])
        if k.startswith("ca_alpha0/"):
            sd[k[len("ca_alpha0/"):] + ".alpha_activ"] = T(g[k])
    orc = R.Oracle(spec, sd)
    noise = T(g["ca_noise"])
    lr, wd, w, eta = [float(v) for v in g["ca_meta"]]
    xs, x0s, losses = R.calibrate_attention(orc, T(g["x"]), spec.seq, R.beta_schedule_linear(), w,
                                            lambda k, which, like: noise[2 * k + (0 if which == "e" else 1)], lr, wd, eta)
    for n in names:
        assert float((orc.sd[n + ".alpha_activ"] - T(g["ca_alpha1/" + n])).abs().max()) < 2e-6, n
        assert torch.allclose(orc.sd[n + ".groups_range"], T(g["ca_gr1/" + n]), rtol=1e-5, atol=1e-6), n
    assert np.allclose(losses, g["ca_loss"], rtol=1e-5)
