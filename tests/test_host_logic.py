"""CPU-only checks: the C-ABI library loads and exports every declared symbol,
the module surface mirrors the reference (state_dict keys, counters, errors),
and the host-side helpers match the oracle."""
import os
import re

import pytest
import torch

import attentiondm_b200 as A
from attentiondm_b200 import _ffi, denoising, dist as adist, runner
from oracle import restate as R
from oracle import synth as S
from tests.util import args_for, config_for

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from attentiondm_b200 import build
    build.build()
    hdr = open(os.path.join(ROOT, "include", "attndm_b200.h")).read()
    declared = set(re.findall(r"\b(attndm_[a-z0-9_]+)\s*\(", hdr))
    declared.discard("attndm_attn_quant")
    L = _ffi.lib()
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/attndm_b200.h but not exported"
    assert set(_ffi.SIGNATURES) | {"attndm_last_error"} == declared
    assert L.attndm_version() >= 100
    assert L.attndm_minmax_workspace_blocks() > 0


def test_state_dict_keys_match_reference_layout():
    spec = S.tiny_spec()
    m = A.Model(config_for(spec), quantization=True, sequence=spec.seq, args=args_for(spec))
    m.materialize_lazy_layers()
    sd = S.synth_state_dict(spec, seed=3)          # loaded strict=True into the real reference by make_golden
    assert set(m.state_dict().keys()) == set(sd.keys())
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == tuple(sd[k].shape), k
    m.load_state_dict(sd, strict=True)


@pytest.mark.parametrize("maker", [S.cifar_spec, S.celeba_spec])
def test_full_config_layer_tables(maker):
    spec = maker(T=10)
    tab = R.qconv_table(spec)
    m = A.Model(config_for(spec), quantization=True, sequence=spec.seq, args=args_for(spec))
    got = {n: (q.in_channels, q.out_channels, q.kernel_size[0], q.a_bit, q.group_num) for n, q in m.qconvs()}
    want = {n: (d["cin"], d["cout"], d["k"], d["a_bit"], d["group_num"]) for n, d in tab.items()}
    assert got == want
    if maker is S.cifar_spec:
        assert len(got) == 198


def test_no_cpu_fallback():
    spec = S.tiny_spec()
    m = A.Model(config_for(spec), quantization=True, sequence=spec.seq, args=args_for(spec))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(1, 3, 8, 8), torch.zeros(1))
    with pytest.raises(RuntimeError, match="CUDA"):
        A.generalized_steps(torch.randn(1, 3, 8, 8), spec.seq, m, R.beta_schedule_linear())
    fp = A.Model(config_for(spec), quantization=False, sequence=spec.seq, args=args_for(spec))     # the FP model builds ...
    with pytest.raises(RuntimeError, match="no CPU fallback"):                                     # ... and has no CPU path either
        fp(torch.randn(1, 3, 8, 8), torch.zeros(1))


def test_ddim_coefficients_match_oracle():
    betas = R.beta_schedule_linear()
    for Tn, eta in [(5, 0.0), (10, 0.3), (100, 1.0)]:
        seq = list(range(0, 1000, 1000 // Tn))
        tab = denoising.ddim_coefficients(seq, betas, eta)
        seq_next = [-1] + seq[:-1]
        for k, (i, j) in enumerate(zip(reversed(seq), reversed(seq_next))):
            at = R.compute_alpha(betas, torch.tensor([i])).view(())
            an = R.compute_alpha(betas, torch.tensor([j])).view(())
            c1 = eta * ((1 - at / an) * (1 - an) / (1 - at)).sqrt()
            c2 = ((1 - an) - c1 ** 2).sqrt()
            want = torch.stack([(1 - at).sqrt(), at.sqrt(), an.sqrt(), torch.as_tensor(c1).float(), c2])
            assert torch.equal(tab[k, :5], want)
            assert tab[k, 5] == i


def test_beta_schedule_and_seq():
    b = torch.from_numpy(runner.get_beta_schedule("linear", beta_start=1e-4, beta_end=0.02,
                                                  num_diffusion_timesteps=1000)).float()
    assert torch.equal(b, R.beta_schedule_linear())
    import argparse
    assert list(runner.make_seq(argparse.Namespace(skip_type="uniform", timesteps=100), 1000)) == list(range(0, 1000, 10))
    q = runner.make_seq(argparse.Namespace(skip_type="quad", timesteps=10), 1000)
    assert len(q) == 10 and q[0] == 0 and q[-1] in (799, 800)


def test_shard_bounds_cover_batch():
    for n in (1, 7, 256, 1000):
        for world in (1, 2, 3, 8):
            spans = [adist.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_quantization_params_match_oracle():
    lo = torch.tensor([-4.0, -5.5, -0.3])
    hi = torch.tensor([6.0, 7.25, 9.0])
    for bits in (4, 6, 8):
        s, z = A.asymmetric_linear_quantization_params(bits, lo, hi)
        s2, z2 = R.asym_params(bits, lo, hi)
        assert torch.equal(s, s2) and torch.equal(z, z2)
    w = torch.randn(8, 4, 3, 3)
    flat = w.reshape(8, -1)
    snapped = A.AsymmetricQuantFunction.apply(w, 8, flat.min(1)[0], flat.max(1)[0])
    assert torch.equal(snapped, R.snap_weight(w, 8)[0])


def test_quant_state_round_trip(tmp_path):
    """runner.quant_state / load_quant_state keep what state_dict() alone loses (weight ranges, init ranges,
    bit widths, counters, lazily created channel_proj)."""
    spec = S.tiny_spec()
    mk = lambda: A.Model(config_for(spec), quantization=True, sequence=spec.seq, args=args_for(spec))
    m = mk()
    m.materialize_lazy_layers()
    m.load_state_dict(S.synth_state_dict(spec, seed=3), strict=True)
    m.init_weight_ranges()
    names = [n for n, _ in m.qconvs()]
    q0 = dict(m.qconvs())[names[0]]
    q0.init_range_min[1] = -2.5
    q0.init_range_max[2] = 3.25
    q0.a_bit = 6
    q0.index_seq = 3
    path = tmp_path / "quant_state.pt"
    runner.save_quant_state(m, path)
    m2 = runner.load_quant_state(mk(), str(path))
    for (n1, a), (n2, b) in zip(m.qconvs(), m2.qconvs()):
        assert n1 == n2
        assert torch.equal(a.weight_range_min, b.weight_range_min) and torch.equal(a.weight_range_max, b.weight_range_max)
        assert torch.equal(a.init_range_min, b.init_range_min) and torch.equal(a.init_range_max, b.init_range_max)
        assert (a.a_bit, a.w_bit, a.index_seq, a.group_num) == (b.a_bit, b.w_bit, b.index_seq, b.group_num)
    sd1, sd2 = m.state_dict(), m2.state_dict()
    assert set(sd1) == set(sd2) and all(torch.equal(sd1[k], sd2[k]) for k in sd1)
    bad = runner.quant_state(m)
    bad["layers"].pop(names[0])
    with pytest.raises(RuntimeError):
        runner.load_quant_state(mk(), bad)
