"""GPU parity tests: the CUDA path (through the C-ABI) against the golden
fixtures generated from the unmodified reference and against the CPU oracle
(oracle/restate.py) on the same seeded inputs.

Bars: integer codes / clip indices / group tables bit-exact given identical fp32
inputs and identical (scale, zero_point) tables; floating-point results within
1e-3 relative L2 (the north_star tolerance), most far tighter (stated per test).
Weights are on the w_bit grid (H1): both sides consume the same
snapped tensor.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import restate as R
from oracle import synth as S
from tests.util import T, build_cuda_model, make_qconv, rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module", autouse=True)
def _lib():
    from attentiondm_b200 import _ffi
    assert _ffi.lib().attndm_device_supported() == 1, "these tests need an sm_100 (B200) device"
    yield


# ---------------------------------------------------------------------------
# a1: activation fake-quant (inference branch)
# ---------------------------------------------------------------------------
def test_act_quant_codes_bit_exact_vs_reference(golden):
    """Given the reference's own (scale, zp) tables the kernel's codes / outputs are bit-exact."""
    from attentiondm_b200 import ops
    g = golden("quant_unit.npz")
    for ci in range(5):
        for mode in ("uniform", "random"):
            k = f"inf{ci}_{mode}"
            C, a_bit, G, Tn = [int(v) for v in g[k + "_meta"]]
            gr, alpha, x = T(g[k + "_gr"]), T(g[k + "_alpha"]), T(g[k + "_x"])
            xn = ops.to_nhwc(x.to(DEV))
            for t in range(Tn):
                lo, hi = R.mixed_range(gr[t], alpha[t])
                s, z = R.asym_params(a_bit, lo, hi)
                codes, rowsum, y = ops.act_quant(xn, s.to(DEV).contiguous(), z.to(DEV).contiguous(), a_bit,
                                                 want_codes=True, want_f32=True)
                want = T(g[k + "_y"][t])
                assert torch.equal(ops.to_nchw(y).cpu(), want), (k, t)
                _, wc, _, _ = R.act_fake_quant(x, gr[t], alpha[t], a_bit, return_codes=True)
                B, _, H, W = x.shape
                got = codes[:, :C].reshape(B, H, W, C).permute(0, 3, 1, 2).cpu().float()
                assert torch.equal(got, wc), (k, t)
                assert torch.equal(rowsum.reshape(B, H, W).cpu().float(), wc.sum(1)), (k, t)
                n = 2 ** (a_bit - 1)
                assert got.min() >= -n and got.max() <= n - 1


def test_qconv_module_quantizer_and_index_wrap(golden):
    """QConv2d's own table build + index_seq wrap; uniform alpha gives bit-exact tables on any device."""
    g = golden("quant_unit.npz")
    for ci in range(5):
        k = f"inf{ci}_uniform"
        C, a_bit, G, Tn = [int(v) for v in g[k + "_meta"]]
        q = make_qconv(C, 8, 1, a_bit, Tn, G)
        q.groups_range.data.copy_(T(g[k + "_gr"]))
        q.invalidate_cache()
        q.init_weight_range()
        x = T(g[k + "_x"]).to(DEV)
        for t in range(Tn + 1):
            _, y, _ = q.quantize_activation_codes(x)
            assert torch.equal(y.cpu(), T(g[k + "_y"][t])), (k, t)
            q(x)                                   # advances index_seq like the reference
        assert q.index_seq == 1
        # random alpha: tables come from the device softmax (ulp-level differences allowed)
        k = f"inf{ci}_random"
        q = make_qconv(C, 8, 1, a_bit, Tn, G)
        q.groups_range.data.copy_(T(g[k + "_gr"]))
        q.alpha_activ.data.copy_(T(g[k + "_alpha"]))
        q.invalidate_cache()
        _, y, _ = q.quantize_activation_codes(T(g[k + "_x"]).to(DEV))
        want = T(g[k + "_y"][0])
        assert rel_l2(y, want) < 1e-3
        # a 1-ulp scale difference moves every element of that channel by an ulp; real code flips are rare
        assert (~torch.isclose(y.cpu(), want, rtol=1e-5, atol=1e-6)).float().mean() < 0.02


# ---------------------------------------------------------------------------
# a4-a6: calibration collectors
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("first", [0, 1])
def test_calibration_vs_reference(golden, first):
    g = golden("quant_unit.npz")
    for ci in range(4):
        k = f"cal{ci}_{first}"
        C, a_bit, G, Tn = [int(v) for v in g[k + "_meta"]]
        q = make_qconv(C, 8, 1, a_bit, Tn, G)
        q.alpha_activ.data.copy_(T(g[k + "_alpha"]))
        q.invalidate_cache()
        q.init_weight_range()
        q.set_calibrate(True)
        q.first_calibrate(bool(first))
        from attentiondm_b200 import ops
        x = T(g[k + "_x"])
        y0 = q._calibrate_step(ops.to_nhwc(x.to(DEV)))
        q.index_seq += 1
        y1 = q._calibrate_step(ops.to_nhwc((x * 0.7).to(DEV)))
        q.index_seq += 1
        # group tables: bit-exact (min/max and the fp32 bin-edge arithmetic are order independent)
        assert torch.equal(q.groups_range.data.cpu(), T(g[k + "_gr"])), k
        init = torch.stack([q.init_range_min, q.init_range_max])
        assert torch.equal(init, T(g[k + "_init"])), k
        # mix output: the G-term fp32 sum may associate differently -> 1e-6
        assert rel_l2(ops.to_nchw(y0), T(g[k + "_y0"])) < 1e-6, k
        assert rel_l2(ops.to_nchw(y1), T(g[k + "_y1"])) < 1e-6, k


def test_minmax_ragged_shapes():
    from attentiondm_b200 import ops
    g = torch.Generator().manual_seed(0)
    for shape in [(1, 1, 1, 3), (2, 5, 7, 3), (3, 4, 4, 32), (2, 1, 1, 1024), (5, 9, 3, 96), (2, 16, 16, 256),
                  (1, 3, 3, 1536)]:
        x = torch.randn(*shape, generator=g).to(DEV) * 5
        mn, mx = ops.minmax_c(x)
        flat = x.reshape(-1, shape[-1])
        assert torch.equal(mn, flat.min(0)[0]) and torch.equal(mx, flat.max(0)[0]), shape


def test_group_wise_bit_exact(golden):
    import attentiondm_b200 as A
    g = golden("quant_unit.npz")
    for vi in range(int(g["gw_count"][0])):
        x = T(g[f"gw{vi}_x"]).to(DEV)
        for G in (4, 8):
            for mm in ("max", "min"):
                xq, gm = A.GroupWise_Quantizaion(x.clone(), dim=x.numel(), group_n=G, maxmin=mm)
                assert torch.equal(xq.cpu(), T(g[f"gw{vi}_{G}_{mm}_xq"])), (vi, G, mm)
                assert torch.equal(gm.cpu(), T(g[f"gw{vi}_{G}_{mm}_gm"])), (vi, G, mm)


def test_percentile_collectors(golden):
    import attentiondm_b200 as A
    g = golden("quant_unit.npz")
    x = T(g["pct_x"]).to(DEV)
    assert A.find_scale_by_percentile_min(x) == float(g["pct_min"][0])
    assert A.find_scale_by_percentile_max(x) == float(g["pct_max"][0])
    # ragged / tied / signed-zero inputs against numpy
    gen = torch.Generator().manual_seed(3)
    for n in (1, 2, 257, 100003):
        v = torch.randn(n, generator=gen)
        v[::3] = 0.0
        v[1::7] = -0.0
        for p in (0.9999, 0.5, 0.01):
            kk = int(n * p)
            if kk >= n:
                continue
            from attentiondm_b200 import ops
            got = ops.kth_value(v.to(DEV), kk).item()
            want = float(np.partition(v.numpy(), kk)[kk])
            assert got == want, (n, p)


# ---------------------------------------------------------------------------
# a2: weights
# ---------------------------------------------------------------------------
def test_weight_clamp_and_grid(golden):
    from attentiondm_b200 import ops
    g = golden("quant_unit.npz")
    for wi in range(3):
        w = T(g[f"wc{wi}_w"]).to(DEV)
        lo, hi = T(g[f"wc{wi}_lo"]).to(DEV), T(g[f"wc{wi}_hi"]).to(DEV)
        w_eff = ops.weight_clamp_pack(w, lo, hi)                     # [O, taps, C]
        O, C, KH, KW = w.shape
        got = w_eff.reshape(O, KH, KW, C).permute(0, 3, 1, 2).cpu()
        assert torch.equal(got, T(g[f"wc{wi}_out"])), wi
        bits = int(g[f"wc{wi}_bits"][0])
        snapped = T(g[f"wc{wi}_snap"]).to(DEV)
        flat = snapped.reshape(O, -1)
        w_eff2 = ops.weight_clamp_pack(snapped, flat.min(1)[0], flat.max(1)[0])
        pack = ops.weight_to_i8(w_eff2, bits)
        assert pack.on_grid
        _, qref, _, zref = R.snap_weight(T(g[f"wc{wi}_w"]), bits)
        Cp = ops.cp_of(C)
        qw = pack.qw.reshape(O, KH * KW, Cp)[:, :, :C].reshape(O, KH, KW, C).permute(0, 3, 1, 2).cpu().float()
        # codes are defined up to the per-channel zero-point shift: compare de-biased integers
        assert torch.equal(qw + pack.w_zp.cpu().float().view(-1, 1, 1, 1), qref + zref.view(-1, 1, 1, 1))
        assert torch.equal(pack.wsum.cpu().float(), qw.reshape(O, -1).sum(1))
        # a channel with exactly symmetric extremes ties at a rounding boundary when snapped and never
        # attains its top code; the grid must still be recovered (span 2^b - 2 steps)
        wt = T(g[f"wc{wi}_w"]).clone()
        b0 = wt[1].abs().max()
        wt[1].clamp_(-b0, b0)
        wt[1].view(-1)[0] = b0
        wt[1].view(-1)[1] = -b0
        st, _, _, _ = R.snap_weight(wt, bits)
        fs = st.reshape(O, -1)
        pk = ops.weight_to_i8(ops.weight_clamp_pack(st.to(DEV), fs.min(1)[0].to(DEV), fs.max(1)[0].to(DEV)), bits)
        assert pk.on_grid
        deq = (pk.qw.reshape(O, KH * KW, Cp)[:, :, :C].float() + pk.w_zp.float().view(-1, 1, 1)) / pk.w_scale.view(-1, 1, 1)
        assert rel_l2(deq.reshape(O, KH, KW, C).permute(0, 3, 1, 2), st) < 1e-6
        # off-grid weights are detected
        assert not ops.weight_to_i8(ops.weight_clamp_pack(w, flat.min(1)[0] * 0 - 10, flat.max(1)[0] * 0 + 10), bits).on_grid


def test_attention_quantize_tensor(golden):
    import attentiondm_b200 as A
    g = golden("quant_unit.npz")
    mpa = A.MixedPrecisionAttention(head_dim=4, num_heads=8, bit_width=4).to(DEV)
    for ai in range(4):
        bits, sc, zp = g[f"aq{ai}_p"]
        y = mpa.quantize_tensor(T(g[f"aq{ai}_x"]).to(DEV), torch.tensor([sc], dtype=torch.float32),
                                torch.tensor([zp], dtype=torch.float32), int(bits))
        assert torch.equal(y.cpu(), T(g[f"aq{ai}_y"]))


# ---------------------------------------------------------------------------
# a3: convolutions
# ---------------------------------------------------------------------------
CONV_SHAPES = [
    # B, H, W, C, O, k
    (2, 8, 8, 32, 32, 3), (1, 5, 7, 64, 48, 3), (3, 4, 4, 128, 128, 3), (2, 16, 16, 256, 128, 3),
    (2, 1, 1, 256, 256, 3), (4, 2, 2, 96, 64, 3), (2, 8, 8, 3, 32, 3), (2, 8, 8, 32, 3, 3),
    (5, 1, 1, 1024, 64, 1), (2, 4, 4, 64, 8, 1), (2, 6, 6, 160, 512, 1), (1, 32, 32, 128, 128, 3),
    (130, 1, 1, 32, 16, 1), (1, 9, 9, 384, 272, 3),
]


def _conv_case(B, H, W, C, O, k, a_bit=8, w_bit=8, seed=0):
    from attentiondm_b200 import ops
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, C, H, W, generator=g) * 3
    w = (torch.rand(O, C, k, k, generator=g) * 2 - 1) / (C * k * k) ** 0.5
    w = R.snap_weight(w, w_bit)[0]
    bias = torch.randn(O, generator=g) * 0.1
    lo, hi = torch.tensor(-4.3), torch.tensor(6.1)
    s, z = R.asym_params(a_bit, lo, hi)
    return x, w, bias, s, z


@pytest.mark.parametrize("shape", CONV_SHAPES)
def test_qconv_i8_simt_and_tcgen05_vs_oracle(shape):
    """Both int8 kernels: bit-identical to each other, and equal to F.conv2d on the
    de-quantized operands up to fp32 summation order (1e-5)."""
    from attentiondm_b200 import ops
    B, H, W, C, O, k = shape
    a_bit = 8
    x, w, bias, s, z = _conv_case(B, H, W, C, O, k)
    taps = k * k
    xn = ops.to_nhwc(x.to(DEV))
    sv = torch.full((C,), float(s), device=DEV)
    zv = torch.full((C,), float(z), device=DEV)
    codes, rowsum, y = ops.act_quant(xn, sv, zv, a_bit, want_codes=True, halo=(k == 3), want_f32=True)
    flat = w.reshape(O, -1)
    w_eff = ops.weight_clamp_pack(w.to(DEV), flat.min(1)[0].to(DEV), flat.max(1)[0].to(DEV))
    pack = ops.weight_to_i8(w_eff, 8)
    assert pack.on_grid
    mult = (1.0 / (float(s) * pack.w_scale.double())).float().contiguous()
    azp = torch.tensor([int(z)], dtype=torch.int32, device=DEV)
    g = torch.Generator().manual_seed(1)
    res = torch.randn(B, H, W, O, generator=g).to(DEV)
    temb = torch.randn(B, O, generator=g).to(DEV)
    outs = {}
    for impl in (ops.CONV_SIMT, ops.CONV_TCGEN05):
        outs[impl] = ops.qconv_i8(codes, rowsum, B, H, W, C, pack, taps, mult, azp, bias.to(DEV), res, temb, impl=impl)
    torch.cuda.synchronize()
    assert torch.equal(outs[ops.CONV_SIMT], outs[ops.CONV_TCGEN05]), "tcgen05 and dp4a kernels disagree"
    xq = R.act_fake_quant(x, torch.tensor([[-4.3, 6.1]]), torch.zeros(1, C), a_bit)
    want = F.conv2d(xq.double(), w.double(), bias.double(), padding=k // 2)
    want = want + res.cpu().double().permute(0, 3, 1, 2) + temb.cpu().double()[:, :, None, None]
    assert rel_l2(ops.to_nchw(outs[ops.CONV_TCGEN05]), want) < 1e-5
    # fp32 kernel on the de-quantized activations
    of = ops.conv_f32(y, w_eff, bias.to(DEV), res, temb)
    assert rel_l2(ops.to_nchw(of), want) < 1e-5


@pytest.mark.parametrize("shape,adds", [((3, 16, 16, 128, 128, 3), False), ((3, 16, 16, 128, 128, 3), True),
                                        ((2, 32, 32, 256, 128, 3), True), ((5, 8, 8, 128, 256, 3), False),
                                        ((2, 64, 64, 128, 128, 3), True), ((3, 8, 12, 256, 384, 1), False),
                                        ((7, 4, 4, 256, 256, 3), True), ((2, 16, 16, 128, 64, 3), False)])
def test_qconv_epilogue_groupnorm_statistics(shape, adds):
    """GroupNorm {sum, sumsq} of the conv output from the tcgen05 epilogue (models/diffusion.py:119-127: every conv of a
    ResidualBlock feeds a GroupNorm): equal to the statistics pass over the stored output up to fp32 summation order
    (the fp32 mean / rstd the consumers form agree to ~1 ulp), and IDENTICAL -- up to the order of the double-precision
    additions -- to the dp4a path's twin kernel (quad order, csrc/conv_common.cuh).  The fp32 part of the order is per
    pixel, so a sample's statistics do not depend on where it sits in the batch (what batch sharding relies on)."""
    from attentiondm_b200 import ops
    B, H, W, C, O, k = shape
    x, w, bias, s, z = _conv_case(B, H, W, C, O, k, seed=3)
    taps = k * k
    xn = ops.to_nhwc(x.to(DEV))
    sv, zv = torch.full((C,), float(s), device=DEV), torch.full((C,), float(z), device=DEV)
    codes, rowsum, _ = ops.act_quant(xn, sv, zv, 8, want_codes=True, halo=(k == 3))
    flat = w.reshape(O, -1)
    w_eff = ops.weight_clamp_pack(w.to(DEV), flat.min(1)[0].to(DEV), flat.max(1)[0].to(DEV))
    pack = ops.weight_to_i8(w_eff, 8)
    mult = (1.0 / (float(s) * pack.w_scale.double())).float().contiguous()
    azp = torch.tensor([int(z)], dtype=torch.int32, device=DEV)
    g = torch.Generator().manual_seed(1)
    res = torch.randn(B, H, W, O, generator=g).to(DEV) if adds else None
    temb = torch.randn(B, O, generator=g).to(DEV) if adds else None
    st, outs = {}, {}
    for impl in (ops.CONV_SIMT, ops.CONV_TCGEN05):
        st[impl] = torch.zeros(B, 32, 2, dtype=torch.float64, device=DEV)
        outs[impl] = ops.qconv_i8(codes, rowsum, B, H, W, C, pack, taps, mult, azp, bias.to(DEV), res, temb, impl=impl,
                                  gn_stats_out=st[impl])
    plain = ops.qconv_i8(codes, rowsum, B, H, W, C, pack, taps, mult, azp, bias.to(DEV), res, temb)
    assert torch.equal(outs[ops.CONV_TCGEN05], plain) and torch.equal(outs[ops.CONV_SIMT], plain)
    a, b = st[ops.CONV_TCGEN05], st[ops.CONV_SIMT]
    assert torch.allclose(a, b, rtol=1e-13, atol=1e-10), (a - b).abs().max()
    ref = plain.double().reshape(B, H * W, 32, O // 32)
    want = torch.stack([ref.sum((1, 3)), (ref * ref).sum((1, 3))], -1)
    n = H * W * (O // 32)
    mean_a, mean_w = a[..., 0] / n, want[..., 0] / n
    var_a, var_w = a[..., 1] / n - mean_a ** 2, want[..., 1] / n - mean_w ** 2
    assert (mean_a - mean_w).abs().max() < 2e-7 * ref.abs().max()
    assert ((var_a - var_w).abs() / var_w).max() < 1e-6
    assert torch.allclose(ops.gn_stats(plain), want, rtol=1e-12, atol=1e-9)
    # the last sample alone (another alignment of its rows to the 128-row tiles): same statistics
    if B > 1 and not adds:
        cl, rl, _ = ops.act_quant(xn[B - 1:].contiguous(), sv, zv, 8, want_codes=True, halo=(k == 3))
        sl = torch.zeros(1, 32, 2, dtype=torch.float64, device=DEV)
        ops.qconv_i8(cl, rl, 1, H, W, C, pack, taps, mult, azp, bias.to(DEV), None, None, gn_stats_out=sl)
        assert torch.allclose(sl, a[B - 1:], rtol=1e-13, atol=1e-10), (sl - a[B - 1:]).abs().max()


@pytest.mark.parametrize("bits", [(4, 4), (6, 8), (8, 4)])
def test_qconv_low_bit(bits):
    from attentiondm_b200 import ops
    a_bit, w_bit = bits
    q = make_qconv(64, 32, 3, a_bit, 4, 8, w_bit=w_bit)
    g = torch.Generator().manual_seed(2)
    q.weight.data.copy_(torch.randn(32, 64, 3, 3, generator=g) * 0.05)
    q.snap_weights_()
    q.groups_range.data[..., 0] = -4.0
    q.groups_range.data[..., 1] = 6.0
    q.invalidate_cache()
    assert q.int8_ok_all_steps()
    x = torch.randn(2, 64, 6, 6, generator=g) * 3
    y = q(x.to(DEV))
    xq = R.act_fake_quant(x, torch.tensor([[-4.0, 6.0]] * 8), torch.full((8, 64), 0.01), a_bit)
    want = F.conv2d(xq.double(), q.weight.data.cpu().double(), q.bias.data.cpu().double(), padding=1)
    assert rel_l2(y, want) < 1e-5


def test_qconv_3x3_on_1x1_map_uses_centre_tap():
    """Most of the CIFAR trunk runs 3x3 convs on 1x1 feature maps: dispatched as the exact 1x1 conv."""
    g = torch.Generator().manual_seed(12)
    for cin, cout in [(256, 256), (768, 256), (64, 32)]:
        q = make_qconv(cin, cout, 3, 8, 4, 8)
        q.weight.data.copy_(torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5)
        q.snap_weights_()
        q.groups_range.data[..., 0] = -4.0
        q.groups_range.data[..., 1] = 6.0
        q.invalidate_cache()
        assert q.int8_ok_all_steps()
        x = torch.randn(5, cin, 1, 1, generator=g) * 3
        y = q(x.to(DEV))
        xq = R.act_fake_quant(x, torch.tensor([[-4.0, 6.0]] * 8), torch.full((8, cin), 0.01), 8)
        want = F.conv2d(xq.double(), q.weight.data.cpu().double(), q.bias.data.cpu().double(), padding=1)
        assert rel_l2(y, want) < 1e-5, (cin, cout)
        q.set_calibrate(True)                      # calibration branch (fp32 kernel) takes the same shortcut
        q.index_seq = 0
        yc = q(x.to(DEV))
        xc, _ = R.calibrate_activation(x, torch.full((8, cin), 0.01), 8, 8, -4.0, 6.0)
        wantc = F.conv2d(xc.double(), q.weight.data.cpu().double(), q.bias.data.cpu().double(), padding=1)
        assert rel_l2(yc, wantc) < 1e-5, (cin, cout)


def test_qconv_falls_back_to_f32_when_not_integer_exact():
    """Non-uniform alpha (per-channel scales, H2) and off-grid weights take the fp32 kernel and
    still match the reference arithmetic."""
    g = torch.Generator().manual_seed(4)
    for case in ("alpha", "weights", "one_sided_channel"):
        Cin, k = (1024, 3) if case == "one_sided_channel" else (32, 1)
        q = make_qconv(Cin, 16, k, 8, 4, 8)
        q.groups_range.data[..., 0] = -4.0 - torch.rand(4, 8, generator=g).to(DEV)
        q.groups_range.data[..., 1] = 6.0 + torch.rand(4, 8, generator=g).to(DEV)
        if case == "alpha":
            q.alpha_activ.data.copy_(torch.randn(4, 8, 32, generator=g))
            q.snap_weights_()
        elif case == "one_sided_channel":
            # a same-sign out-channel: exactly on the 8-bit grid, but with K = 9*1024 its zero point
            # 128 + round(255 * lo / (hi - lo)) = 978 would overflow the int32 epilogue bracket
            # acc + zp*wsum + w_zp*cs: weight_to_i8 must declare it off-grid -> fp32 kernel
            q.weight.data[3] = 1.0 + 0.3 * torch.rand(Cin, 3, 3, generator=g).to(DEV)
            q.snap_weights_()
            assert not q._packed()[1].on_grid
            q.weight.data[3] = -0.02 + 0.04 * torch.rand(Cin, 3, 3, generator=g).to(DEV)     # control: straddles 0
            q.snap_weights_()
            assert q._packed()[1].on_grid
            q.weight.data[3] = 1.0 + 0.3 * torch.rand(Cin, 3, 3, generator=g).to(DEV)
            q.snap_weights_()
        else:
            q.init_weight_range()
        q.invalidate_cache()
        assert not q.int8_ok_all_steps()
        x = torch.randn(3, Cin, 5, 5, generator=g) * 3
        y = q(x.to(DEV))
        xq = R.act_fake_quant(x, q.groups_range.data[0].cpu(), q.alpha_activ.data[0].cpu(), 8)
        want = F.conv2d(xq.double(), q.weight.data.cpu().double(), q.bias.data.cpu().double(), padding=k // 2)
        assert rel_l2(y, want) < 1e-3, case      # device softmax -> ulp-different tables -> a few code flips


# ---------------------------------------------------------------------------
# a8-a10, a13: blocks
# ---------------------------------------------------------------------------
def test_groupnorm_silu_quant_fused():
    from attentiondm_b200 import ops
    g = torch.Generator().manual_seed(5)
    for (B, H, W, C) in [(2, 8, 8, 32), (3, 4, 4, 64), (2, 16, 16, 128), (2, 1, 1, 256), (1, 3, 5, 96), (2, 2, 2, 768),
                         (2, 1, 1, 192), (2, 3, 3, 160)]:   # 192/160: 6 / 5 channels per group, quads straddle groups
        x = torch.randn(B, C, H, W, generator=g) * 2 + 0.5
        gamma = 1 + 0.2 * torch.randn(C, generator=g)
        beta = 0.2 * torch.randn(C, generator=g)
        want = F.silu(F.group_norm(x, 32, gamma, beta, eps=1e-6))
        xn = ops.to_nhwc(x.to(DEV))
        gn = ops.GnArgs(ops.gn_stats(xn), gamma.to(DEV), beta.to(DEV), 1e-6)
        y = ops.gn_silu(xn, gn)
        assert rel_l2(ops.to_nchw(y), want) < 2e-6, (B, H, W, C)
        s, z = R.asym_params(8, torch.tensor(-4.0), torch.tensor(6.0))
        sv, zv = torch.full((C,), float(s), device=DEV), torch.full((C,), float(z), device=DEV)
        codes, rowsum, yq = ops.act_quant(xn, sv, zv, 8, ops.PRE_GN_SILU, gn, want_codes=True, halo=True, want_f32=True)
        # fused == unfused on the kernel's own GN+SiLU output (bit-exact), and within a code of torch's
        c2, r2, y2 = ops.act_quant(y, sv, zv, 8, want_codes=True, halo=True, want_f32=True)
        assert torch.equal(codes, c2) and torch.equal(rowsum, r2) and torch.equal(yq, y2)
        wq = R.act_fake_quant(want, torch.tensor([[-4.0, 6.0]]), torch.zeros(1, C), 8)
        diff = (ops.to_nchw(yq).cpu() - wq).abs()
        assert diff.max() <= 10.0 / 255 * 1.001 and (diff > 0).float().mean() < 2e-3
        # halo ring holds the code of 0.0
        ring = codes.reshape(B, H + 2, W + 2, -1)[:, 0, :, :C]
        assert (ring.float() == -float(z)).all()
        # the one-kernel variant (statistics in-kernel, sample tile in shared memory) agrees with the
        # two-kernel path up to the last bit of the fp64 statistics (a rare +-1 code)
        if ops.gn_fits_fused(H, W, C):
            gl = ops.GnArgs(None, gamma.to(DEV), beta.to(DEV), 1e-6)
            c3, r3, y3 = ops.act_quant(xn, sv, zv, 8, ops.PRE_GN_SILU, gl, want_codes=True, halo=True, want_f32=True)
            dc = (c3.int() - codes.int()).abs()
            assert dc.max() <= 1 and (dc > 0).float().mean() < 1e-4
            assert torch.equal(r3.long() - rowsum.long(), (c3.long() - codes.long()).sum(1))
            assert rel_l2(ops.gn_silu(xn, gl), ops.gn_silu(xn, gn)) < 1e-6

@pytest.mark.parametrize("B,H,W,C1,C2,a_bit", [(3, 16, 16, 128, 128, 8), (2, 8, 12, 384, 128, 8), (2, 16, 8, 128, 128, 4),
                                               (2, 32, 32, 256, 256, 8), (5, 4, 4, 128, 128, 6)])
def test_upblock_concat_read_in_place(B, H, W, C1, C2, a_bit):
    """UpBlock.res1's input cat([upsample_x2(x), skip]) (models/diffusion.py:225-229,244) is never written: GroupNorm
    statistics and both quantizers read x and skip in place.  Codes and row sums are bit-identical to the same kernels
    run on the materialised concat (which test_unet_glue_ops pins to torch), plain and halo layouts, with and without
    the GroupNorm+SiLU producer; the statistics agree to the last bits of a double."""
    from attentiondm_b200 import ops
    g = torch.Generator().manual_seed(11 + C1 + H)
    lo = (torch.randn(B, H // 2, W // 2, C1, generator=g) * 1.5 + 0.3).to(DEV)
    skip = (torch.randn(B, H, W, C2, generator=g) * 0.8 - 0.2).to(DEV)
    C = C1 + C2
    assert ops.CatView.fits(lo, skip) == (not ops.gn_fits_fused(H, W, C))
    if not ops.CatView.fits(lo, skip):
        return
    view = ops.CatView(lo, skip)
    full = ops.upsample_concat(lo, skip)
    want = torch.cat([lo.repeat_interleave(2, 1).repeat_interleave(2, 2), skip], -1)
    assert torch.equal(full, want) and torch.equal(view.materialize(), want)
    st_v, st_f = ops.gn_stats(view), ops.gn_stats(full)
    assert torch.allclose(st_v, st_f, rtol=1e-13, atol=1e-9)
    gamma = (1 + 0.2 * torch.randn(C, generator=g)).to(DEV)
    beta = (0.2 * torch.randn(C, generator=g)).to(DEV)
    s, z = R.asym_params(a_bit, torch.tensor(-4.0), torch.tensor(6.0))
    sv, zv = torch.full((C,), float(s), device=DEV), torch.full((C,), float(z), device=DEV)
    for halo in (False, True):
        for pre in (ops.PRE_NONE, ops.PRE_GN_SILU, ops.PRE_SILU):
            gn = ops.GnArgs(st_f, gamma, beta, 1e-6) if pre == ops.PRE_GN_SILU else None
            c_v, r_v, _ = ops.act_quant(ops.CatView(lo, skip), sv, zv, a_bit, pre, gn, want_codes=True, halo=halo)
            c_f, r_f, _ = ops.act_quant(full, sv, zv, a_bit, pre, gn, want_codes=True, halo=halo)
            assert torch.equal(c_v, c_f) and torch.equal(r_v, r_f), (halo, pre)
    # with its own (in-place) statistics: at most a last-bit difference of the fp32 mean / rstd
    gn_v = ops.GnArgs(st_v, gamma, beta, 1e-6)
    c_v, _, _ = ops.act_quant(view, sv, zv, a_bit, ops.PRE_GN_SILU, gn_v, want_codes=True, halo=True)
    c_f, _, _ = ops.act_quant(full, sv, zv, a_bit, ops.PRE_GN_SILU, ops.GnArgs(st_f, gamma, beta, 1e-6), want_codes=True, halo=True)
    assert ((c_v.int() - c_f.int()).abs() > 0).float().mean() < 1e-6
    # both quantizers of UpBlock.res1 in one pass (attndm_act_quant_cat2): conv1 behind GroupNorm+SiLU and the shortcut
    # conv on the raw concat, each with its own tables and layout -- bit-identical to the two separate passes
    s2, z2 = R.asym_params(a_bit, torch.tensor(-2.5), torch.tensor(3.5))
    sv2 = torch.full((C,), float(s2), device=DEV) * (1 + 0.01 * torch.arange(C, device=DEV) / C)
    zv2 = torch.full((C,), float(z2), device=DEV)
    gn = ops.GnArgs(st_f, gamma, beta, 1e-6)
    for h1, h2 in ((True, False), (True, True), (False, False)):
        v2 = ops.CatView(lo, skip)
        assert v2.prepare_pair((sv, zv, a_bit, h1), (sv2, zv2, a_bit, h2), gn)
        ca, ra, _ = ops.act_quant(v2, sv, zv, a_bit, ops.PRE_GN_SILU, gn, want_codes=True, halo=h1)
        cb, rb, _ = ops.act_quant(v2, sv2, zv2, a_bit, ops.PRE_NONE, None, want_codes=True, halo=h2)
        assert not v2._prepared                                  # both came from the prepared pair
        wa = ops.act_quant(full, sv, zv, a_bit, ops.PRE_GN_SILU, gn, want_codes=True, halo=h1)
        wb = ops.act_quant(full, sv2, zv2, a_bit, ops.PRE_NONE, None, want_codes=True, halo=h2)
        assert torch.equal(ca, wa[0]) and torch.equal(ra, wa[1]), (h1, h2)
        assert torch.equal(cb, wb[0]) and torch.equal(rb, wb[1]), (h1, h2)


@pytest.mark.gpu
@pytest.mark.parametrize("B,H,W,C,O", [(4, 8, 8, 768, 512), (3, 5, 7, 320, 128), (2, 4, 4, 100, 256), (300, 2, 2, 64, 128)])
def test_conv1x1_f32_tensor_core_3xtf32(B, H, W, C, O):
    """The fp32 `channel_proj` GEMM on the tensor cores (tcgen05 kind::tf32, operands split exactly into
    big + small): within 2e-5 of an fp64 reference (fp32 SIMT kernel: ~1e-6), and batch-independent."""
    from attentiondm_b200 import ops
    assert ops.conv_f32_tc_fits(B * H * W, C, O)
    g = torch.Generator().manual_seed(7 + C)
    x = (torch.randn(B, H, W, C, generator=g) * 1.7 + 0.2).to(DEV)
    w = (torch.randn(O, C, generator=g) / C ** 0.5).to(DEV)
    bias = torch.randn(O, generator=g).to(DEV)
    big, small = ops.split_tf32(x)
    assert torch.equal(big + small, x) and (big.view(torch.int32) & 0x1FFF).eq(0).all()
    y_tc = ops.conv1x1_f32_tc(x, ops.split_tf32(w), bias)
    y_simt = ops.conv_f32(x, w.view(O, 1, C).contiguous(), bias)
    want = (x.double().reshape(-1, C) @ w.double().t() + bias.double()).reshape(B, H, W, O)
    scale = want.abs().max().item()
    err_tc = (y_tc.double() - want).abs().max().item() / scale
    err_simt = (y_simt.double() - want).abs().max().item() / scale
    # measured ~5e-6 of the output range (the tensor core's fp32 accumulation is not round-to-nearest) against ~1e-6
    # for sequential fmaf; the bar for this operator is 1e-3 (the reference's own GPU conv runs in plain TF32)
    assert err_tc < 2e-5 and err_simt < 4e-6, (err_tc, err_simt)
    # every sample's rows do not depend on the rest of the batch
    y1 = ops.conv1x1_f32_tc(x[1:2].contiguous(), ops.split_tf32(w), bias)
    assert torch.equal(y1, y_tc[1:2])


def test_silu_quantizer_forms():
    """SiLU in front of the quantizer (csrc/common.cuh).  The default build evaluates it on the special-function
    unit: on random pre-activations fewer than 5e-6 of the codes may differ from quantizing the library's accurate
    fp32 SiLU.  The A/B builds (ATTNDM_LIB=..._silu_guard.so: SFU first, accurate next to a rounding boundary;
    ..._silu_accurate.so) must give EXACTLY the accurate codes, also on inputs constructed to sit right at the
    boundaries.  In every build the accurate SiLU agrees with torch's to <= 2 ulp."""
    import os
    from attentiondm_b200 import ops, _ffi
    exact = any(v in os.path.basename(_ffi.LIB_PATH) for v in ("silu_guard", "silu_accurate"))
    g = torch.Generator().manual_seed(12)
    C = 128
    for a_bit, lo, hi in ((8, -4.0, 6.0), (8, -0.5, 3.0), (6, -4.0, 6.0), (4, -4.0, 6.0), (8, -30.0, 50.0)):
        s, z = R.asym_params(a_bit, torch.tensor(lo), torch.tensor(hi))
        sv = torch.full((C,), float(s), device=DEV)
        zv = torch.full((C,), float(z), device=DEV)
        # random pre-activations + adversarial ones: u with s * silu(u) - zp = k + 0.5 +- delta
        u_rand = torch.randn(64, 32, 32, C, generator=g) * 2.5
        k = torch.randint(-2 ** (a_bit - 1), 2 ** (a_bit - 1), (8, 32, 32, C), generator=g).double()
        delta = (torch.rand(8, 32, 32, C, generator=g).double() - 0.5) * 4e-4
        y_t = (k + 0.5 + delta + float(z)) / float(s)                      # target silu value
        y_t = y_t.clamp_min(-0.27)                                         # silu >= -0.2785
        u = y_t.clone().clamp_min(0.1)
        for _ in range(60):                                                # Newton on silu(u) = y_t, u > -1.2785 branch
            sg = torch.sigmoid(u)
            f = u * sg - y_t
            u = u - f / (sg * (1 + u * (1 - sg))).clamp_min(1e-3)
        x = torch.cat([u_rand, u.float()]).to(DEV)
        y_acc = ops.silu(x)                                                # accurate fp32 SiLU of the library
        want, _, _ = ops.act_quant(y_acc, sv, zv, a_bit, ops.PRE_NONE, want_codes=True, halo=False)
        got, rs, _ = ops.act_quant(x, sv, zv, a_bit, ops.PRE_SILU, want_codes=True, halo=False)
        got_h, _, _ = ops.act_quant(x, sv, zv, a_bit, ops.PRE_SILU, want_codes=True, halo=True)   # rows kernel
        B, H, W, _ = x.shape
        inner = got_h.view(B, H + 2, W + 2, -1)[:, 1:-1, 1:-1].reshape(-1, got_h.shape[-1])
        assert torch.equal(inner, got)                                     # every quantizer kernel shares the form
        if exact:
            assert torch.equal(got, want), (a_bit, lo, hi, int((got != want).sum()))
        else:
            nr = u_rand.shape[0] * H * W
            d = (got[:nr].int() - want[:nr].int()).abs()
            assert int(d.max()) <= 1 and float((d > 0).float().mean()) < 5e-6, (a_bit, lo, hi, float((d > 0).float().mean()))
        # how many of the adversarial elements really are boundary cases (sanity of the construction)
        t = float(s) * F.silu(u.double()) - float(z)
        near = ((t - t.floor() - 0.5).abs() < 3e-4).float().mean()
        assert near > 0.5 or a_bit < 8 or lo < -10, float(near)
    ref = F.silu(x.cpu())
    ulp = (y_acc.cpu().view(torch.int32).long() - ref.view(torch.int32).long()).abs()
    assert int(ulp.max()) <= 2 and float((ulp > 0).float().mean()) < 0.02, (int(ulp.max()), float((ulp > 0).float().mean()))


def test_attention_core():
    from attentiondm_b200 import ops
    g = torch.Generator().manual_seed(6)
    for (B, N, d, dv) in [(2, 16, 16, 128), (3, 64, 32, 256), (1, 1, 4, 32), (2, 1024, 16, 128), (2, 4, 8, 64)]:
        q = torch.randn(B, N, d, generator=g)
        k = torch.randn(B, N, d, generator=g)
        v = torch.randn(B, N, dv, generator=g)
        att = F.softmax(torch.bmm(q.double(), k.double().transpose(1, 2)) * d ** -0.5, dim=-1)
        want = torch.bmm(att, v.double())
        got = ops.attention(q.to(DEV), k.to(DEV), v.to(DEV), d ** -0.5)
        assert rel_l2(got, want) < 2e-6, (B, N, d, dv)


def test_mixed_precision_attention_vs_restatement():
    """PARITY PINNED AT `quantize_tensor` ONLY: MixedPrecisionAttention.forward is shape-broken in the
    reference (K permuted to [B, h, HW, d], utils/attention_quant_utils.py:70), so no reference output exists for it.
    This test compares the kernel with the oracle's restatement of the evident intent (K as [B, h, d, HW]); the
    quantizer inside it is pinned bit-exactly to reference fixtures by test_attention_quantize_tensor."""
    import attentiondm_b200 as A
    g = torch.Generator().manual_seed(7)
    for bits in (4, 6, 8):
        B, N, C = 2, 16, 256
        kc = C // 8
        mpa = A.MixedPrecisionAttention(head_dim=kc // 8, num_heads=8, bit_width=bits, scaling_factor=kc ** -0.5).to(DEV)
        mpa.update_quantization_params(-3.0, 4.0, 0.0, 1.0)
        q = torch.randn(B, N, kc, generator=g)
        k = torch.randn(B, kc, N, generator=g)
        v = torch.randn(B, N, C, generator=g)
        st = dict(num_heads=8, base_bits=bits, scaling_factor=kc ** -0.5, scale_qk=mpa.quant_scale_qk.cpu(),
                  zero_qk=mpa.quant_zero_qk.cpu(), scale_attn=mpa.quant_scale_attn.cpu(),
                  zero_attn=mpa.quant_zero_attn.cpu(), softmax_scale=torch.ones(1))
        want = R.mixed_precision_attention(q, k, v, st)
        got = mpa(q.to(DEV), k.to(DEV), v.to(DEV))
        # quantized logits/probabilities: a rounding flip moves one probability by one step
        assert rel_l2(got, want) < (2e-2 if bits <= 6 else 2e-6), bits


def test_unet_glue_ops():
    from attentiondm_b200 import ops
    g = torch.Generator().manual_seed(8)
    x = torch.randn(2, 32, 8, 6, generator=g)
    assert torch.equal(ops.to_nchw(ops.maxpool2(ops.to_nhwc(x.to(DEV)))).cpu(), F.max_pool2d(x, 2))
    for (h, w, hs, ws) in [(4, 4, 8, 8), (1, 1, 1, 1), (2, 2, 3, 3), (1, 1, 2, 2), (3, 5, 4, 7)]:
        a = torch.randn(2, 16, h, w, generator=g)
        sk = torch.randn(2, 8, hs, ws, generator=g)
        up = F.interpolate(a, scale_factor=2, mode="nearest")
        if up.shape[2:] != sk.shape[2:]:
            up = F.interpolate(up, size=sk.shape[2:], mode="nearest")
        want = torch.cat([up, sk], dim=1)
        got = ops.upsample_concat(ops.to_nhwc(a.to(DEV)), ops.to_nhwc(sk.to(DEV)))
        assert torch.equal(ops.to_nchw(got).cpu(), want), (h, w, hs, ws)
    t = torch.tensor([0.0, 1.0, 10.0, 500.0, 990.0])
    emb = ops.timestep_embedding(t.to(DEV), 256)
    d = (emb.cpu() - R.timestep_embedding(t, 256)).abs()
    # correctly rounded here vs the CPU's 1-ulp exp: a 1-ulp frequency moves sin(t*f) by up to |t*f| * 6e-8
    assert d.max() < 1e-4 and (d > 1e-6).float().mean() < 0.05
    a, b = torch.randn(1000, generator=g), torch.randn(1000, generator=g)
    gam = torch.tensor([0.37])
    assert torch.equal(ops.scale_add(a.to(DEV), b.to(DEV), gam.to(DEV)).cpu(), gam * a + b)


# ---------------------------------------------------------------------------
# a12: sampler
# ---------------------------------------------------------------------------
def test_ddim_loop_bit_exact_vs_reference(golden):
    import attentiondm_b200 as A
    g = golden("ddim_unit.npz")
    betas = R.beta_schedule_linear().to(DEV)
    for ci in range(4):
        Tn, eta = g[f"d{ci}_meta"]
        seq = range(0, 1000, 1000 // int(Tn))

        def model(xt, t):
            return (0.3 * xt.cpu() + torch.sin(t.cpu() / 100.0).view(-1, 1, 1, 1) * 0.1).to(DEV)

        torch.manual_seed(77)                     # the reference draws randn_like(x) on the CPU every step

        def noise_fn(k, like):
            return torch.randn(like.shape).to(DEV)

        xs, x0s = A.generalized_steps(T(g[f"d{ci}_x"]).to(DEV), seq, model, betas, eta=float(eta), noise_fn=noise_fn)
        assert len(xs) == int(Tn) + 1 and len(x0s) == int(Tn)
        assert all(not t.is_cuda for t in xs[1:]) and all(not t.is_cuda for t in x0s)
        assert torch.equal(torch.stack([t.cpu() for t in xs]), T(g[f"d{ci}_xs"]))
        assert torch.equal(torch.stack(x0s), T(g[f"d{ci}_x0"]))


def test_engine_noise_source_can_alternate():
    """eta > 0 on ONE CUDA-graph engine, alternating a caller-supplied noise_fn with engine-generated noise:
    the captured graph must neither overwrite the caller's noise nor replay stale noise."""
    import attentiondm_b200 as A
    spec = S.tiny_spec(T=4, bitwidth=8)
    sd = S.synth_state_dict(spec, seed=5)
    m = build_cuda_model(spec, sd)
    betas = R.beta_schedule_linear().to(DEV)
    x = torch.randn(2, 3, 16, 16, generator=torch.Generator().manual_seed(2)).to(DEV)
    m.set_calibrate(True)
    A.generalized_steps(x, spec.seq, m, betas, eta=0.0, keep="last")
    m.set_calibrate(False)
    fixed = [torch.randn(2, 3, 16, 16, generator=torch.Generator().manual_seed(50 + k)).to(DEV) for k in range(4)]

    def run(noise_fn=None, seed=None, graph=True):
        m.reset_index_seq()
        if seed is not None:
            torch.manual_seed(seed)
        xs, _ = A.generalized_steps(x, spec.seq, m, betas, eta=0.5, noise_fn=noise_fn, use_graph=graph)
        return torch.stack(xs[1:])

    ext_eager = run(lambda k, like: fixed[k], graph=False)
    ext_a = run(lambda k, like: fixed[k])                  # graph captured while the caller supplies the noise
    own_b = run(seed=9)                                    # same engine, engine-generated noise
    ext_c = run(lambda k, like: fixed[k])                  # and back
    own_d = run(seed=9)
    assert torch.equal(ext_a, ext_eager) and torch.equal(ext_c, ext_a)
    assert torch.equal(own_b, own_d) and not torch.equal(own_b, ext_a)
    # engine-generated noise is one normal_() draw per step on the NHWC buffer: reproduce it through noise_fn
    buf = torch.empty(2, 16, 16, 3, device=DEV)
    own_e = run(lambda k, like: buf.normal_().permute(0, 3, 1, 2), seed=9)
    assert torch.equal(own_e, own_b)
    # opposite capture order on a fresh engine
    from attentiondm_b200.engine import SamplerEngine
    SamplerEngine._cache.clear()
    own_f = run(seed=9)
    ext_g = run(lambda k, like: fixed[k])
    assert torch.equal(own_f, own_b) and torch.equal(ext_g, ext_a)


# ---------------------------------------------------------------------------
# end to end: tiny UNet against the reference fixtures
#
# A fake-quantized network is chaotic at the LSB scale: GroupNorm/SiLU results that differ from
# torch's by <= 2 ulp flip an activation code with probability ~1e-6 per element, one flip moves 9*C_out
# conv outputs by |w|*LSB, each of which flips a few per cent of the next quantizer's codes (branching
# factor > 1), and within three layers the deviation saturates near the quantization noise (~1e-2
# relative).  The same happens between the reference's own CPU and CUDA runs.  So parity is asserted
#   (1) operator by operator IN SITU (every QConv2d / attention call of every step, the oracle applied
#       to the CUDA path's own layer input): <= 1e-3, no avalanche possible;
#   (2) per layer on the reference's recorded inputs (test_per_layer_trace_vs_reference_fixture);
#   (3) on whole steps: <= 1e-3 when no code flipped upstream, and bounded by the avalanche level
#       (5e-2) otherwise -- the measured values are printed.
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("bw,alpha,mode", [(8, "uniform", "calibrate"), (8, "uniform", "sample"),
                                           (4, "attn_random", "sample"), (6, "uniform", "sample")])
def test_unet_operator_parity_in_situ(bw, alpha, mode):
    import attentiondm_b200 as A
    spec = S.tiny_spec(T=4, bitwidth=bw)          # 1000 // 4 = 250 -> exactly 4 steps (len_seq == timesteps)
    sd = S.synth_state_dict(spec, seed=5, alpha_mode=alpha)
    m = build_cuda_model(spec, sd)
    mods = dict(m.qconvs())
    betas = R.beta_schedule_linear().to(DEV)
    x = torch.randn(2, 3, 16, 16, generator=torch.Generator().manual_seed(21)).to(DEV)
    m.set_calibrate(True)
    A.generalized_steps(x, spec.seq, m, betas, eta=0.0, keep="last")       # calibrate the tables
    m.set_calibrate(mode == "calibrate")
    m.reset_index_seq()
    from oracle import insitu
    rec = insitu.record_layers(m)
    A.generalized_steps(x, spec.seq, m, betas, eta=0.0, keep="last", use_graph=False)
    assert len(rec) == len(mods) * spec.len_seq
    errs = sorted(insitu.layer_errors(m, rec, mode == "calibrate"), reverse=True)
    worst, where = errs[0]
    over = [e for e in errs if e[0] >= 1e-3]
    print(f"in-situ operator parity [{bw}-bit {alpha} {mode}]: worst layer rel-L2 {worst:.2e} at {where}, "
          f"median {errs[len(errs) // 2][0]:.1e}, {len(over)} of {len(rec)} calls over 1e-3")
    # Bar: 1e-3 relative L2 per operator call.  The GroupNorm+SiLU producer is not bit-identical to torch's
    # (special-function-unit exp and reciprocal, a few ulp), so once in a few hundred calls ONE activation lands
    # on the other side of a rounding boundary; on this tiny model's 4x4x32 maps (1024 activations) a single
    # flipped 8-bit code is a ~1e-3 change of the conv output by itself.  Hence: at most 1 % of the calls may
    # exceed 1e-3 and none may exceed 5e-3 (measured: median 1e-7, 99th percentile 6e-7, tools/debug_insitu.py).
    assert len(over) <= max(1, len(rec) // 100) and worst < 5e-3, (where, worst, over[:5])


@pytest.mark.parametrize("name,bw,alpha,gain,first", [
    ("tiny_unet_w8.npz", 8, "uniform", 1.0, False),
    ("tiny_unet_w8_scaled.npz", 8, "uniform", 0.5, True),
    ("tiny_unet_w4_attn.npz", 4, "attn_random", 1.0, False),
])
@pytest.mark.parametrize("impl", ["tcgen05", "simt"])
def test_tiny_unet_vs_reference_fixture(golden, name, bw, alpha, gain, first, impl):
    import attentiondm_b200 as A
    from attentiondm_b200 import ops
    ops.DEFAULT_CONV_IMPL = ops.CONV_TCGEN05 if impl == "tcgen05" else ops.CONV_SIMT
    AVALANCHE = 5e-2            # see the section comment: deviation level once any code has flipped
    try:
        g = golden(name)
        Tn = int(g["meta"][0])
        spec = S.tiny_spec(T=Tn, bitwidth=bw)
        sd = S.synth_state_dict(spec, seed=3, weight_gain=gain, alpha_mode=alpha)
        m = build_cuda_model(spec, sd)
        betas = R.beta_schedule_linear().to(DEV)
        bcpu = R.beta_schedule_linear()
        x = T(g["x"]).to(DEV)
        rseq = list(reversed(spec.seq))
        rnext = list(reversed([-1] + list(spec.seq)[:-1]))
        # --- calibration pass, teacher-forced with the reference's trajectory ---
        m.set_calibrate(True, first=first)
        xt = T(g["x"])
        errs = []
        for t in range(Tn):
            tt = torch.full((x.shape[0],), float(rseq[t]))
            eps = m(xt.to(DEV), tt.to(DEV))
            ge = T(g["calib_eps"][t])
            errs.append(rel_l2(eps, ge))
            at = R.compute_alpha(bcpu, tt.long())
            an = R.compute_alpha(bcpu, torch.full_like(tt, rnext[t]).long())
            xt, _ = R.ddim_update(xt, ge, at, an, 0.0, torch.zeros_like(xt))
        m.set_calibrate(False)
        print(f"{name} [{impl}] calibration eps rel-L2 per step: {['%.1e' % e for e in errs]}")
        assert max(errs) < AVALANCHE, errs
        if not first:
            for n, q in m.qconvs():          # activations stay inside the [-4, 6] floor -> identical tables
                assert rel_l2(q.groups_range.data, T(g["gr/" + n])) < 2e-2, n
        # --- quantized sampling with the REFERENCE's calibrated tables (isolates the sampler) ---
        for n, q in m.qconvs():
            q.groups_range.data.copy_(T(g["gr/" + n]))
            q.invalidate_cache(weights=False)
        xs_gold = T(g["xs"])
        errs = []
        for t in range(Tn):                  # step level, teacher-forced x_t
            m.reset_index_seq(t)
            tt = torch.full((x.shape[0],), float(rseq[t]), device=DEV)
            errs.append(rel_l2(m(xs_gold[t].to(DEV), tt), T(g["eps"][t])))
        print(f"{name} [{impl}] sampling eps rel-L2 per step (teacher-forced): {['%.1e' % e for e in errs]}")
        assert max(errs) < AVALANCHE, errs
        m.reset_index_seq()
        xs, _ = A.generalized_steps(x, spec.seq, m, betas, eta=0.0, use_graph=False)
        e_run = rel_l2(torch.stack([t.cpu() for t in xs]), xs_gold)
        print(f"{name} [{impl}] free-running trajectory rel-L2: {e_run:.1e}")
        assert e_run < AVALANCHE
        # --- same thing through the CUDA-graph engine: identical to the eager kernels ---
        m.reset_index_seq()
        xs_g, _ = A.generalized_steps(x, spec.seq, m, betas, eta=0.0, use_graph=True)
        assert torch.equal(torch.stack(xs_g[1:]), torch.stack(xs[1:]))
        assert all(q.index_seq == Tn for _, q in m.qconvs())
    finally:
        ops.DEFAULT_CONV_IMPL = ops.CONV_TCGEN05


def test_per_layer_trace_vs_reference_fixture(golden):
    """Teacher-forced per-layer check: feed the reference's recorded layer input, compare the output."""
    g = golden("tiny_unet_w8.npz")
    spec = S.tiny_spec(T=int(g["meta"][0]), bitwidth=8)
    sd = S.synth_state_dict(spec, seed=3)
    m = build_cuda_model(spec, sd)
    mods = dict(m.qconvs())
    for key in g.files:
        if not key.startswith("trace_in/"):
            continue
        n = key[len("trace_in/"):]
        q = mods[n]
        q.groups_range.data.copy_(T(g["gr/" + n]))
        q.invalidate_cache(weights=False)
        q.index_seq = 0
        assert q.int8_ok_all_steps(), n
        y = q(T(g[key]).to(DEV))
        assert rel_l2(y, T(g["trace_out/" + n])) < 1e-5, n


# ---------------------------------------------------------------------------
# fused per-sample programs (attndm_rowprog): the CUDA-graph engine runs every time_mlp and every block
# that works on a 1x1 map inside one kernel each; the result must equal the layer-by-layer kernels
# (which the in-situ tests above pin to the oracle) bit for bit.
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("ch,ch_mult,size,bw,alpha,B,mixed", [
    (32, (1, 2), 4, 8, "uniform", 3, False),   # 4 -> 2 -> 1 -> 1: pooled trunk input, 32/64 channels (1-2 per group)
    (64, (1, 2, 2), 8, 8, "uniform", 9, False),    # 8 -> 4 -> 2 -> 1 ..., 128 channels, ragged sample tile
    (32, (1, 2), 4, 6, "uniform", 2, False),   # 6-bit activations (4-bit key projection)
    (32, (1, 2), 4, 4, "attn_random", 2, False),   # trained attention alphas -> those layers leave the integer path
    (64, (1, 2, 2), 8, 4, "attn_random", 5, True),     # the same with mixed-precision attention (BASELINE.json config 3)
    # MixedPrecisionAttention (8 heads, softmax scale, fake-quantized scores at <= 6 bits and probabilities at <= 4) at one
    # position inside the fused trunk (utils/attention_quant_utils.py:51-107)
    (64, (1, 2), 4, 8, "uniform", 3, True),
    (64, (1, 2), 4, 6, "uniform", 2, True),
    (64, (1, 2), 4, 4, "uniform", 5, True),
])
def test_fused_rowprog_equals_layerwise(ch, ch_mult, size, bw, alpha, B, mixed):
    import attentiondm_b200 as A
    from attentiondm_b200.engine import SamplerEngine
    spec = S.tiny_spec(T=4, bitwidth=bw, ch=ch, ch_mult=ch_mult, image_size=size)
    sd = S.synth_state_dict(spec, seed=11, alpha_mode=alpha)
    m = build_cuda_model(spec, sd)
    if mixed:
        n_mixed = 0
        for mod in m.modules():
            if isinstance(mod, A.EnhancedQSelfAttention):
                mod.gamma.data.fill_(0.5)
                mod.enable_mixed_precision()
                mod.attention_processor.update_quantization_params(-6.0, 7.0, 0.0, 1.0)
                mod.attention_processor.softmax_scale.data.fill_(1.25)
                n_mixed += 1
        assert n_mixed > 0
    betas = R.beta_schedule_linear().to(DEV)
    x = torch.randn(B, 3, size, size, generator=torch.Generator().manual_seed(4)).to(DEV)
    m.set_calibrate(True)
    A.generalized_steps(x, spec.seq, m, betas, eta=0.0, keep="last")
    m.set_calibrate(False)
    m.reset_index_seq()
    xs_e, x0_e = A.generalized_steps(x, spec.seq, m, betas, eta=0.0, use_graph=False)
    m.reset_index_seq()
    xs_g, x0_g = A.generalized_steps(x, spec.seq, m, betas, eta=0.0, use_graph=True)
    eng = SamplerEngine.for_model(m, spec.seq, betas, 0.0, tuple(x.shape))
    # (attn_random: the attention convs take the fp32 path -- quantize, de-quantize, fp32 conv -- inside the fused trunk)
    assert eng.fused is not None and eng.fused.trunk_plan is not None, "the fused plan was not built"
    if alpha == "uniform":
        assert eng.fused.n_up >= 1
        # the time path (timestep embedding -> time_embed -> every time_mlp) is evaluated once per pass for all steps
        # instead of B times per step (engine.py); the eager run above evaluates it per step, per sample
        assert eng.hoist and eng.fused.hoisted
    assert torch.isfinite(xs_g[-1]).all()
    assert torch.equal(torch.stack(xs_g[1:]), torch.stack(xs_e[1:]))
    assert torch.equal(torch.stack(x0_g), torch.stack(x0_e))
    # the fused kernel prefetches the next conv's parameters asynchronously: replays must not depend on timing
    for _ in range(20):
        m.reset_index_seq()
        xs_r, _ = A.generalized_steps(x, spec.seq, m, betas, eta=0.0, keep="last", use_graph=True)
        assert torch.equal(xs_r[-1], xs_e[-1])



def test_bcast_rows_and_time_path_all_steps():
    """The time path of the sampler is evaluated once per pass for all steps and fanned out per step (engine.py):
    attndm_bcast_rows copies the staged row's column ranges to [B, width] tensors, and the all-steps evaluation
    (time_embedding on [T] rows + the time_mlp programs with one table row per CTA) gives, for every step, exactly what
    the per-step per-sample evaluation gives."""
    import attentiondm_b200 as A
    from attentiondm_b200 import _ffi, ops
    from attentiondm_b200.engine import SamplerEngine
    cur = torch.arange(200, dtype=torch.float32, device=DEV) * 0.5
    desc = torch.tensor([[0, 17, 5], [35, 100, 12], [119, 4, 3]], dtype=torch.int32, device=DEV)
    B = 7
    dst = torch.full((B * (5 + 12 + 3),), -1.0, device=DEV)
    _ffi.call("attndm_bcast_rows", _ffi.ptr(cur), _ffi.ptr(desc), 3, B, 12, _ffi.ptr(dst), _ffi.stream())
    assert torch.equal(dst[:35].view(B, 5), cur[17:22].expand(B, 5))
    assert torch.equal(dst[35:119].view(B, 12), cur[100:112].expand(B, 12))
    assert torch.equal(dst[119:].view(B, 3), cur[4:7].expand(B, 3))
    # the hoisted time path against the per-step one, on the tiny model
    spec = S.tiny_spec(T=4, bitwidth=8)
    m = build_cuda_model(spec, S.synth_state_dict(spec, seed=5))
    betas = R.beta_schedule_linear().to(DEV)
    x = torch.randn(3, 3, 16, 16, generator=torch.Generator().manual_seed(2)).to(DEV)
    m.set_calibrate(True)
    A.generalized_steps(x, spec.seq, m, betas, eta=0.0, keep="last")
    m.set_calibrate(False)
    m.reset_index_seq()
    eng = SamplerEngine(m, spec.seq, betas, 0.0, tuple(x.shape))
    assert eng.hoist
    eng.load_input(x)                                        # runs the all-steps time path into the table columns
    fp = eng.fused
    blocks = [b for b in list(m.down_blocks) + list(m.up_blocks) if b.time_mlp is not None]
    for k in range(eng.T):
        t = eng.table[k, eng.t_off:eng.t_off + 3].contiguous()
        te = m.time_embedding(t)
        col = eng.temb_off
        for blk in blocks:
            q = blk.time_mlp[1]
            q.index_seq = k
            want = q.forward_fused(te, ops.PRE_SILU).view(3, -1)          # per step, per sample (the eager path)
            w = q.out_channels
            got = eng.table[k, col:col + w]
            assert torch.equal(want, got.expand(3, w)), (k, w)
            col += (w + 3) // 4 * 4
    m.reset_index_seq()

# ---------------------------------------------------------------------------
# full-size properties (oracle too slow): CIFAR config, batch 8
# ---------------------------------------------------------------------------
def test_cifar_full_size_properties():
    import attentiondm_b200 as A
    from attentiondm_b200 import ops
    spec = S.cifar_spec(T=4)
    sd = S.synth_state_dict(spec, seed=1)
    m = build_cuda_model(spec, sd)
    assert len(m.qconvs()) == 198
    betas = R.beta_schedule_linear().to(DEV)
    x = torch.randn(8, 3, 32, 32, generator=torch.Generator().manual_seed(9)).to(DEV)
    m.set_calibrate(True)
    A.generalized_steps(x, spec.seq, m, betas, eta=0.0, keep="last")
    m.set_calibrate(False)
    m.reset_index_seq()
    assert all(q.int8_ok_all_steps() for _, q in m.qconvs())
    xs, _ = A.generalized_steps(x, spec.seq, m, betas, eta=0.0, keep="last")
    assert torch.isfinite(xs[-1]).all()
    # samples are independent: a batch of 8 equals two batches of 4 (what batch sharding relies on)
    m.reset_index_seq()
    xa, _ = A.generalized_steps(x[:4].contiguous(), spec.seq, m, betas, eta=0.0, keep="last")
    m.reset_index_seq()
    xb, _ = A.generalized_steps(x[4:].contiguous(), spec.seq, m, betas, eta=0.0, keep="last")
    assert torch.equal(torch.cat([xa[-1], xb[-1]]), xs[-1])
    # the dp4a twin gives the same images bit-for-bit
    ops.DEFAULT_CONV_IMPL = ops.CONV_SIMT
    try:
        m.reset_index_seq()
        xs2, _ = A.generalized_steps(x, spec.seq, m, betas, eta=0.0, keep="last", use_graph=False)
    finally:
        ops.DEFAULT_CONV_IMPL = ops.CONV_TCGEN05
    assert torch.equal(xs2[-1], xs[-1])


# ---------------------------------------------------------------------------
# the other named configs (BASELINE.json configs 4 and 5): CelebA 64x64 and LSUN church 256x256 UNets.
# The oracle is far too slow at these sizes, so the checks are the size-independent properties the path
# offers: every layer stays on the integer path, the CUDA-graph engine (with the fused 1x1 programs) equals the
# layer-by-layer kernels bit for bit, the tcgen05 conv equals its dp4a twin bit for bit, samples are independent
# of how the batch is sharded, outputs are finite.
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("name,B", [("celeba", 2), ("church", 1)])
def test_large_configs_properties(name, B):
    import attentiondm_b200 as A
    from attentiondm_b200 import ops
    from attentiondm_b200.engine import SamplerEngine
    spec = (S.celeba_spec if name == "celeba" else S.church_spec)(T=2)
    sd = S.synth_state_dict(spec, seed=2)
    m = build_cuda_model(spec, sd)
    del sd
    n_layers = {"celeba": 252, "church": 305}[name]
    assert len(m.qconvs()) == n_layers
    betas = R.beta_schedule_linear().to(DEV)
    size = spec.image_size
    x = torch.randn(B, 3, size, size, generator=torch.Generator().manual_seed(17)).to(DEV)
    m.set_calibrate(True)
    A.generalized_steps(x, spec.seq, m, betas, eta=0.0, keep="last")
    m.set_calibrate(False)
    m.reset_index_seq()
    assert all(q.int8_ok_all_steps() for _, q in m.qconvs())
    xs_g, _ = A.generalized_steps(x, spec.seq, m, betas, eta=0.0, keep="last")                  # CUDA-graph engine
    eng = SamplerEngine.for_model(m, spec.seq, betas, 0.0, tuple(x.shape))
    from attentiondm_b200 import rowprog
    assert eng.fused is not None and eng.fused.trunk_plan is not None, rowprog.last_unfusable
    assert torch.isfinite(xs_g[-1]).all()
    m.reset_index_seq()
    xs_e, _ = A.generalized_steps(x, spec.seq, m, betas, eta=0.0, keep="last", use_graph=False)
    assert torch.equal(xs_e[-1], xs_g[-1])
    ops.DEFAULT_CONV_IMPL = ops.CONV_SIMT
    try:
        m.reset_index_seq()
        xs_s, _ = A.generalized_steps(x, spec.seq, m, betas, eta=0.0, keep="last", use_graph=False)
    finally:
        ops.DEFAULT_CONV_IMPL = ops.CONV_TCGEN05
    assert torch.equal(xs_s[-1], xs_g[-1])
    if B > 1:
        m.reset_index_seq()
        xa, _ = A.generalized_steps(x[:1].contiguous(), spec.seq, m, betas, eta=0.0, keep="last")
        assert torch.equal(xa[-1], xs_g[-1][:1])
