"""Per-call time breakdown of one ENGINE step (CIFAR config, batch 256): the step body the CUDA graph
captures (staged tables, fused rowprog launches, layer kernels, DDIM update), run eagerly with CUDA
events around every C-ABI call.  Also the launch list source for ncu (--events 0)."""
import argparse, os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import attentiondm_b200 as A
from attentiondm_b200 import _ffi, ops, rowprog
from attentiondm_b200.engine import SamplerEngine

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=0)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--events", type=int, default=1)
a = ap.parse_args()
dev = torch.device("cuda")
bench.T_STEPS = 100
CFG = bench.CONFIGS[os.environ.get("ATTNDM_CONFIG", "cifar10_w8a8")]
SIZE = CFG["size"]
if a.batch <= 0:
    a.batch = CFG["batch"]
m, seq = bench.build_model(dev, CFG)
for n, q in m.qconvs():          # skip calibration: every activation range = the reference floor [-4, 6]
    q.groups_range.data[..., 0] = -4.0
    q.groups_range.data[..., 1] = 6.0
    q.invalidate_cache(weights=False)
betas = torch.linspace(1e-4, 0.02, 1000, dtype=torch.float64).float().to(dev)
eng = SamplerEngine(m, seq, betas, 0.0, (a.batch, 3, SIZE, SIZE))
print("fused:", eng.fused is not None, "trunk:", eng.fused is not None and eng.fused.trunk_plan is not None)
if eng.fused is not None and eng.fused.trunk_plan is not None:
    tp = eng.fused.trunk_plan
    print(f"trunk: ns={tp.ns} ops={tp.n_ops} arena={tp.arena_floats*4} B cp_max={tp.cp_max} first_down={eng.fused.first_down} n_up={eng.fused.n_up}")
    print(f"time_mlp: ns={eng.fused.time_plan.ns} programs={len(eng.fused.time_plan.programs)}")
x = torch.randn(a.batch, 3, SIZE, SIZE, device=dev)
eng.load_input(x)
with torch.no_grad():
    for _ in range(2):
        eng._with_staged(eng._step_body)
    torch.cuda.synchronize()
    rec = []
    orig = _ffi.call
    def timed(name, *args):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); orig(name, *args); e1.record()
        ints = tuple(v for v in args if isinstance(v, int) and not isinstance(v, bool) and 0 <= v < 100000)[:8]
        rec.append((name, ints, e0, e1))
    if a.events:
        _ffi.call = timed; ops.call = timed; rowprog._ffi.call = timed
    w0 = torch.cuda.Event(enable_timing=True); w1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    w0.record()
    for _ in range(a.steps):
        eng._with_staged(eng._step_body)
    w1.record()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print(f"eager wall per step: {w0.elapsed_time(w1)/a.steps:.3f} ms, C-ABI calls per step: {len(rec)//max(1,a.steps)}")
agg = collections.defaultdict(lambda: [0, 0.0])
for name, ints, e0, e1 in rec:
    k = (name, ints)
    agg[k][0] += 1
    agg[k][1] += e0.elapsed_time(e1)
tot = sum(v[1] for v in agg.values())
byname = collections.defaultdict(lambda: [0, 0.0])
for (name, ints), (c, ms) in agg.items():
    byname[name][0] += c; byname[name][1] += ms
print(f"sum of kernel times per step: {tot/a.steps:.3f} ms")
for name, (c, ms) in sorted(byname.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:28s} calls/step={c//a.steps:4d} ms/step={ms/a.steps:8.3f} share={ms/tot*100:5.1f}%")
print("--- top 45 (name, int args) ---")
for (name, ints), (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{name:24s} {str(ints):50s} n/step={c//a.steps:3d} ms/step={ms/a.steps:7.3f} avg_us={ms/c*1e3:8.1f}")
