"""Debug aid (GPU box): teacher-forced calibration/inference step k, per-layer output diffs CUDA vs oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import restate as R, synth as S
from tests.util import build_cuda_model, rel_l2, T
from attentiondm_b200 import ops
torch.set_num_threads(1)
name, bw, alpha, gain, first, kstep = sys.argv[1], int(sys.argv[2]), sys.argv[3], float(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
g = np.load(os.path.join("tests/golden", name))
Tn = int(g["meta"][0])
spec = S.tiny_spec(T=Tn, bitwidth=bw)
sd = S.synth_state_dict(spec, seed=3, weight_gain=gain, alpha_mode=alpha)
m = build_cuda_model(spec, sd)
orc = R.Oracle(spec, sd)
betas = R.beta_schedule_linear()
x = T(g["x"])
seq = list(spec.seq); seq_next = [-1] + seq[:-1]
m.set_calibrate(True, first=bool(first)); orc.set_calibrate(True, first=bool(first))
rec = {}
for n, q in m.qconvs():
    orig = q.forward_fused
    def wrap(xx, pre=ops.PRE_NONE, gn=None, residual=None, temb=None, _o=orig, _n=n):
        y = _o(xx, pre, gn, residual, temb)
        rec[_n] = (xx.detach().clone(), y.detach().clone(), pre, residual is not None, temb is not None)
        return y
    q.forward_fused = wrap
xt = x
for k, (i, j) in enumerate(zip(reversed(seq), reversed(seq_next))):
    tt = torch.full((x.shape[0],), float(i))
    orc.trace = {}
    with torch.no_grad():
        eo = orc.forward(xt, tt)
        eg = m(xt.cuda(), tt.cuda()).float().cpu()
    ge = T(g["calib_eps"][k])
    print(f"step {k}: cuda-vs-golden {rel_l2(eg, ge):.3e}  oracle-vs-golden {rel_l2(eo, ge):.3e}")
    if k == kstep:
        for n in orc.trace:
            xi, yo = orc.trace[n][0]
            gx, gy, pre, hr, ht = rec[n]
            rin = rel_l2(ops.to_nchw(gx), xi) if pre == ops.PRE_NONE else float("nan")
            rout = rel_l2(ops.to_nchw(gy), yo) if not (hr or ht) else float("nan")
            flag = " <<<" if (rout == rout and rout > 1e-5) or (rin == rin and rin > 1e-5) else ""
            print(f"   {n:36s} in={rin:9.2e} out={rout:9.2e}{flag}")
        break
    at = R.compute_alpha(betas, tt.long()); an = R.compute_alpha(betas, torch.full_like(tt, j).long())
    xt, _ = R.ddim_update(xt, ge, at, an, 0.0, torch.zeros_like(xt))
