"""Debug aid (GPU box): teacher-forced calibration pass vs the reference fixture; prints per-step
eps rel-L2 and which layers' group tables / init ranges differ."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import restate as R, synth as S
from tests.util import build_cuda_model, rel_l2, T
name, bw, alpha, gain, first = sys.argv[1], int(sys.argv[2]), sys.argv[3], float(sys.argv[4]), int(sys.argv[5])
g = np.load(os.path.join("tests/golden", name))
Tn = int(g["meta"][0])
spec = S.tiny_spec(T=Tn, bitwidth=bw)
sd = S.synth_state_dict(spec, seed=3, weight_gain=gain, alpha_mode=alpha)
m = build_cuda_model(spec, sd)
betas = R.beta_schedule_linear()
x = T(g["x"])
seq = list(spec.seq); seq_next = [-1] + seq[:-1]
m.set_calibrate(True, first=bool(first))
xt = x
for k, (i, j) in enumerate(zip(reversed(seq), reversed(seq_next))):
    tt = torch.full((x.shape[0],), float(i))
    eps = m(xt.cuda(), tt.cuda()).float().cpu()
    ge = T(g["calib_eps"][k])
    print(f"step {k}: eps rel {rel_l2(eps, ge):.3e}")
    at = R.compute_alpha(betas, tt.long()); an = R.compute_alpha(betas, torch.full_like(tt, j).long())
    xt, _ = R.ddim_update(xt, ge, at, an, 0.0, torch.zeros_like(xt))      # teacher forcing with the reference eps
    bad = []
    for n, q in m.qconvs():
        d = (q.groups_range.data[k].cpu() - T(g["gr/" + n])[k]).abs().max().item()
        if d > 1e-5:
            bad.append((n, d))
        if first:
            init = torch.stack([q.init_range_min, q.init_range_max])[:, k]
            gi = T(g["init/" + n])[:, k]
            if not torch.allclose(init, gi):
                bad.append((n + " INIT", (init - gi).abs().max().item(), init.tolist(), gi.tolist()))
    print("   layers with different tables:", len(bad))
    for b in bad[:12]:
        print("     ", b)
