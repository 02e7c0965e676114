"""Debug: graph-first vs eager ordering on a named config (reproduces test_large_configs_properties)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import attentiondm_b200 as A
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
def traj(graph):
    m.reset_index_seq()
    xs, _ = A.generalized_steps(x, spec.seq, m, betas, eta=0.0, use_graph=graph)
    return torch.stack(xs[1:])
e1 = traj(False)
bad = 0
N = int(os.environ.get("TRIALS", "30"))
for i in range(N):
    g = traj(True)
    if not torch.equal(g, e1):
        bad += 1
        d = [float((g[k] - e1[k]).abs().max()) for k in range(g.shape[0])]
        print(f" trial {i}: mismatch per step {d}")
print(f"mismatching graph runs: {bad}/{N}")
