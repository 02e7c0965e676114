"""The CPU restatement (oracle/restate.py) against fixtures produced by the
unmodified reference (oracle/make_golden.py).  Integer/quantizer arithmetic is
compared bit-exactly; UNet eps within fp32 summation-order noise."""
import numpy as np
import pytest
import torch

from oracle import restate as R
from oracle import synth as S


def T(a):
    return torch.from_numpy(np.asarray(a))


def test_inference_fake_quant_bit_exact(golden):
    g = golden("quant_unit.npz")
    for ci in range(5):
        for mode in ("uniform", "random"):
            k = f"inf{ci}_{mode}"
            C, a_bit, G, Tn = [int(v) for v in g[k + "_meta"]]
            st = R.QState(C, 8, a_bit, a_bit, Tn, Tn, group_num=G)
            gr, alpha, x = T(g[k + "_gr"]), T(g[k + "_alpha"]), T(g[k + "_x"])
            for t in range(Tn + 1):
                y = R.quantize_activation(x, st, gr, alpha)
                assert torch.equal(y, T(g[k + "_y"][t])), (k, t)
            assert st.index_seq == 1          # wrapped after T calls


@pytest.mark.parametrize("first", [0, 1])
def test_calibration_bit_exact(golden, first):
    g = golden("quant_unit.npz")
    for ci in range(4):
        k = f"cal{ci}_{first}"
        C, a_bit, G, Tn = [int(v) for v in g[k + "_meta"]]
        st = R.QState(C, 8, a_bit, a_bit, Tn, Tn, group_num=G, calibrate=True, first_calibrate=bool(first))
        gr = torch.zeros(Tn, G, 2)
        alpha, x = T(g[k + "_alpha"]), T(g[k + "_x"])
        y0 = R.quantize_activation(x, st, gr, alpha)
        y1 = R.quantize_activation(x * 0.7, st, gr, alpha)
        assert torch.equal(y0, T(g[k + "_y0"])), k
        assert torch.equal(y1, T(g[k + "_y1"])), k
        assert torch.equal(gr, T(g[k + "_gr"])), k
        assert torch.equal(torch.stack([st.init_range_min, st.init_range_max]), T(g[k + "_init"])), k


def test_group_wise_bit_exact(golden):
    g = golden("quant_unit.npz")
    for vi in range(int(g["gw_count"][0])):
        x = T(g[f"gw{vi}_x"])
        for G in (4, 8):
            for mm in ("max", "min"):
                xq, gm = R.group_wise(x.clone(), G, mm)
                assert torch.equal(xq, T(g[f"gw{vi}_{G}_{mm}_xq"])), (vi, G, mm)
                assert torch.equal(gm, T(g[f"gw{vi}_{G}_{mm}_gm"])), (vi, G, mm)


def test_weight_clamp_and_snap(golden):
    g = golden("quant_unit.npz")
    for wi in range(3):
        w = T(g[f"wc{wi}_w"])
        out = R.weight_clamp(w, T(g[f"wc{wi}_lo"]), T(g[f"wc{wi}_hi"]))
        assert torch.equal(out, T(g[f"wc{wi}_out"]))
        snapped, q, s, zp = R.snap_weight(w, int(g[f"wc{wi}_bits"][0]))
        assert torch.equal(snapped, T(g[f"wc{wi}_snap"]))
        # snapping is idempotent and codes are integral
        assert torch.equal(q, q.round())
        again = R.snap_weight(snapped, int(g[f"wc{wi}_bits"][0]))[0]
        assert torch.allclose(again, snapped, atol=0, rtol=1e-6)


def test_attention_quantize_tensor(golden):
    g = golden("quant_unit.npz")
    for ai in range(4):
        bits, sc, zp = g[f"aq{ai}_p"]
        y = R.attn_quantize_tensor(T(g[f"aq{ai}_x"]), torch.tensor([sc], dtype=torch.float32),
                                   torch.tensor([zp], dtype=torch.float32), int(bits))
        assert torch.equal(y, T(g[f"aq{ai}_y"]))


def test_percentiles(golden):
    g = golden("quant_unit.npz")
    x = T(g["pct_x"])
    assert R.percentile_min(x) == g["pct_min"][0]
    assert R.percentile_max(x) == g["pct_max"][0]


def test_ddim_sampler(golden):
    g = golden("ddim_unit.npz")
    betas = R.beta_schedule_linear()
    abar = R.compute_alpha(betas, torch.arange(-1, 1000)).view(-1)
    assert torch.equal(abar, T(g["abar"]))
    for ci in range(4):
        Tn, eta = g[f"d{ci}_meta"]
        seq = range(0, 1000, 1000 // int(Tn))

        def model(xt, t):
            return 0.3 * xt + torch.sin(t / 100.0).view(-1, 1, 1, 1) * 0.1

        torch.manual_seed(77)
        xs, x0s = R.ddim_sample(model, T(g[f"d{ci}_x"]), seq, betas, eta=float(eta))
        assert torch.equal(torch.stack(xs), T(g[f"d{ci}_xs"]))
        assert torch.equal(torch.stack(x0s), T(g[f"d{ci}_x0"]))


def rel_l2(a, b):
    return float((a - b).norm() / b.norm())


@pytest.mark.parametrize("name,bw,alpha,gain,first", [
    ("tiny_unet_w8.npz", 8, "uniform", 1.0, False),
    ("tiny_unet_w8_scaled.npz", 8, "uniform", 0.5, True),
    ("tiny_unet_w4_attn.npz", 4, "attn_random", 1.0, False),
])
def test_tiny_unet_matches_reference(golden, name, bw, alpha, gain, first):
    g = golden(name)
    Tn = int(g["meta"][0])
    spec = S.tiny_spec(T=Tn, bitwidth=bw)
    sd = S.synth_state_dict(spec, seed=3, weight_gain=gain, alpha_mode=alpha)
    assert bytes.fromhex(S.state_digest(sd)) == g["digest"].tobytes(), "synthetic weights not reproducible"
    orc = R.Oracle(spec, sd)
    betas = R.beta_schedule_linear()
    x = T(g["x"])
    orc.set_calibrate(True, first=first)
    xs, _, eps = R.ddim_sample(orc.forward, x, spec.seq, betas, eta=0.0, return_eps=True)
    for t in range(Tn):
        assert rel_l2(eps[t], T(g["calib_eps"][t])) < 2e-6
    for n in orc.qs:
        assert torch.equal(orc.sd[n + ".groups_range"], T(g["gr/" + n])), n
        if first:
            init = torch.stack([orc.qs[n].init_range_min, orc.qs[n].init_range_max])
            assert torch.equal(init, T(g["init/" + n])), n
    orc.set_calibrate(False)
    orc.trace = {}
    xs, _, eps = R.ddim_sample(orc.forward, x, spec.seq, betas, eta=0.0, return_eps=True)
    for t in range(Tn):
        assert rel_l2(eps[t], T(g["eps"][t])) < 2e-6, t
    assert rel_l2(torch.stack(xs), T(g["xs"])) < 2e-6
    for key in g.files:
        if key.startswith("trace_in/"):
            n = key[len("trace_in/"):]
            xi, yo = orc.trace[n][0]
            assert rel_l2(xi, T(g[key])) < 2e-6, n
            assert rel_l2(yo, T(g["trace_out/" + n])) < 2e-6, n
