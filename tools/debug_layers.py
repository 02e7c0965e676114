"""Debug aid (GPU box): run one UNet forward through the CUDA modules and through the CPU
oracle, and print the per-QConv2d input/output rel-L2 in execution order."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import restate as R, synth as S
from tests.util import build_cuda_model, rel_l2
from attentiondm_b200 import ops

torch.set_num_threads(1)
mode = sys.argv[1] if len(sys.argv) > 1 else "calib"
spec = S.tiny_spec(T=4, bitwidth=8)
sd = S.synth_state_dict(spec, seed=3)
m = build_cuda_model(spec, sd)
orc = R.Oracle(spec, sd)
x = torch.randn(2, 3, 16, 16, generator=torch.Generator().manual_seed(123))
t = torch.full((2,), 750.0)
if mode == "calib":
    m.set_calibrate(True); orc.set_calibrate(True)
else:
    for k in sd:
        if k.endswith("groups_range"):
            pass
    for n, q in m.qconvs():
        q.groups_range.data[..., 0] = -4.0; q.groups_range.data[..., 1] = 6.0; q.invalidate_cache(weights=False)
        orc.sd[n + ".groups_range"][..., 0] = -4.0; orc.sd[n + ".groups_range"][..., 1] = 6.0
rec = {}
for n, q in m.qconvs():
    orig = q.forward_fused
    def wrap(xx, pre=ops.PRE_NONE, gn=None, residual=None, temb=None, _o=orig, _n=n):
        y = _o(xx, pre, gn, residual, temb)
        rec[_n] = (xx.detach().clone(), y.detach().clone(), pre, residual is not None, temb is not None)
        return y
    q.forward_fused = wrap
orc.trace = {}
with torch.no_grad():
    eo = orc.forward(x, t)
    eg = m(x.cuda(), t.cuda())
print("eps rel", rel_l2(eg, eo))
for n in orc.trace:
    xi, yo = orc.trace[n][0]
    gx, gy, pre, hr, ht = rec[n]
    # the CUDA layer input is BEFORE its fused producer; only comparable when pre == NONE
    rin = rel_l2(ops.to_nchw(gx), xi) if pre == ops.PRE_NONE else float("nan")
    # the CUDA output includes the fused residual/temb adds; the oracle's does not
    note = ("+res" if hr else "") + ("+temb" if ht else "")
    rout = rel_l2(ops.to_nchw(gy), yo) if not (hr or ht) else float("nan")
    print(f"{n:40s} in={rin:9.2e} out={rout:9.2e} pre={pre} {note} shape={tuple(gy.shape)}")
