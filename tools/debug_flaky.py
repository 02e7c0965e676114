"""Debug: find which fused launch makes step 1 flaky (engine step bodies run eagerly, per-launch checks)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import attentiondm_b200 as A
from attentiondm_b200 import rowprog, ops
from attentiondm_b200.engine import SamplerEngine
from oracle import restate as R, synth as S
from tests.util import build_cuda_model
name, B = sys.argv[1], int(sys.argv[2])
spec = {"celeba": S.celeba_spec, "church": S.church_spec, "cifar": S.cifar_spec}[name](T=2)
sd = S.synth_state_dict(spec, seed=2)
m = build_cuda_model(spec, sd)
dev = torch.device("cuda")
betas = R.beta_schedule_linear().to(dev)
size = spec.image_size
x = torch.randn(B, 3, size, size, generator=torch.Generator().manual_seed(17)).to(dev)
m.set_calibrate(True)
A.generalized_steps(x, spec.seq, m, betas, eta=0.0, keep="last")
m.set_calibrate(False)
m.reset_index_seq()
eng = SamplerEngine(m, spec.seq, betas, 0.0, tuple(x.shape))
fp = eng.fused
# record what the two fused launches produce, per step, over many trials
orig_time, orig_trunk = fp.run_time_mlps, fp.run_trunk
rec = {}
def run_time(t_emb, cur):
    orig_time(t_emb, cur)
    rec.setdefault(("time", state["step"]), []).append(torch.cat([v.flatten() for v in fp.temb.values()]).clone())
def run_trunk(h, cur):
    out = orig_trunk(h, cur)
    rec.setdefault(("trunk_in", state["step"]), []).append(h.clone())
    rec.setdefault(("cur", state["step"]), []).append(cur.clone())
    rec.setdefault(("temb_at_trunk", state["step"]), []).append(torch.cat([v.flatten() for v in fp.temb.values()]).clone())
    rec.setdefault(("trunk", state["step"]), []).append(out.clone())
    return out
fp.run_time_mlps, fp.run_trunk = run_time, run_trunk
state = {"step": 0}
N = int(os.environ.get("TRIALS", "30"))
with torch.no_grad():
    for i in range(N):
        eng.load_input(x)
        for k in range(2):
            state["step"] = k
            eng._with_staged(eng._step_body)
torch.cuda.synchronize()
for key, vals in sorted(rec.items()):
    ref = vals[0]
    bad = [i for i, v in enumerate(vals) if not torch.equal(v, ref)]
    print(key, "trials differing from trial 0:", bad[:10], "count", len(bad))
    if bad and key[0] == "trunk":
        d = (vals[bad[0]] - ref).abs().view(B, -1)
        print("   per-sample max diff", d.max(1)[0].tolist(), "differing channels per sample", (d > 0).sum(1).tolist())
# ---- isolate the trunk kernel: same input, same staged table, many launches ----
fp.run_time_mlps, fp.run_trunk = orig_time, orig_trunk
with torch.no_grad():
    eng.load_input(x)
    for k in range(2):
        eng._with_staged(eng._step_body)      # leaves `cur` = step-1 tables, temb = step-1 values
    h1 = rec[("trunk_in", 1)][0]
    outs = [orig_trunk(h1, eng.cur).clone() for _ in range(300)]
torch.cuda.synchronize()
bad = [i for i, o in enumerate(outs) if not torch.equal(o, outs[0])]
print("isolated trunk launches differing from the first:", len(bad), bad[:10])
if bad:
    d = (outs[bad[0]] - outs[0]).abs().view(B, -1)
    print(" per-sample max diff", d.max(1)[0].tolist(), " differing channels of sample 0:", int((d[0] > 0).sum()))
