"""Debug: bisect a fused-vs-layerwise mismatch (time_mlp plan / trunk plan) on a named config."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import attentiondm_b200 as A
from attentiondm_b200 import rowprog
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
print("fused", eng.fused is not None, rowprog.last_unfusable)
fp = eng.fused
if fp is not None and "--bisect" in sys.argv:
    print("trunk ns", fp.trunk_plan.ns, "first_down", fp.first_down, "n_up", fp.n_up, "time ns", fp.time_plan.ns)
    def run(mode):
        eng.load_input(x)
        saved = eng.fused
        if mode == "none": eng.fused = None
        elif mode == "time": eng.fused = rowprog.FusedPlans(fp.time_plan, fp.temb, None, len(m.down_blocks), 0, 0, fp.B)
        else: eng.fused = fp
        with torch.no_grad():
            eps = eng._with_staged(eng._step_body)
        eng.fused = saved
        return eps.clone()
    e0 = run("none"); e1 = run("time"); e2 = run("all")
    print("time-only == none:", torch.equal(e0, e1), float((e0 - e1).abs().max()))
    print("all == none:", torch.equal(e0, e2), float((e0 - e2).abs().max()))
# ---- whole trajectories: eager twice, graph twice, per step ----
def traj(graph):
    m.reset_index_seq()
    xs, _ = A.generalized_steps(x, spec.seq, m, betas, eta=0.0, use_graph=graph)
    return torch.stack(xs[1:])
ea, eb = traj(False), traj(False)
ga, gb = traj(True), traj(True)
print("eager run-to-run equal:", torch.equal(ea, eb), " graph run-to-run equal:", torch.equal(ga, gb))
for k in range(ea.shape[0]):
    print(f" step {k}: eager==graph {torch.equal(ea[k], ga[k])} max|d| {float((ea[k]-ga[k]).abs().max()):.3e}")
