"""Per-kernel time breakdown of one denoising step (CIFAR config, batch 256) measured with CUDA
events around every C-ABI call in eager mode; writes a table to stdout.  Also usable under
`ncu --metrics gpu__time_duration.sum` for the launch list (use --steps 1)."""
import argparse, os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import attentiondm_b200 as A
from attentiondm_b200 import _ffi, ops

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--events", type=int, default=1)
a = ap.parse_args()
dev = torch.device("cuda")
bench.T_STEPS = 100
m, seq = bench.build_model(dev, bench.CONFIGS[os.environ.get("ATTNDM_CONFIG", "cifar10_w8a8")])
for n, q in m.qconvs():          # skip calibration: every activation range = the reference floor [-4, 6]
    q.groups_range.data[..., 0] = -4.0
    q.groups_range.data[..., 1] = 6.0
    q.invalidate_cache(weights=False)
x = torch.randn(a.batch, 3, 32, 32, device=dev)
t = torch.full((a.batch,), 990.0, device=dev)
with torch.no_grad():
    for _ in range(2):
        m(x, t)
    torch.cuda.synchronize()
    rec = []
    orig = _ffi.call
    def timed(name, *args):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); orig(name, *args); e1.record()
        ints = tuple(v for v in args if isinstance(v, int) and not isinstance(v, bool) and 0 <= v < 100000)[:8]
        rec.append((name, ints, e0, e1))
    if a.events:
        _ffi.call = timed
        ops.call = timed
    w0 = torch.cuda.Event(enable_timing=True); w1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    w0.record()
    for _ in range(a.steps):
        m(x, t)
    w1.record()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print(f"eager wall per forward: {w0.elapsed_time(w1)/a.steps:.3f} ms, C-ABI calls per forward: {len(rec)//max(1,a.steps)}")
agg = collections.defaultdict(lambda: [0, 0.0])
for name, ints, e0, e1 in rec:
    k = (name, ints)
    agg[k][0] += 1
    agg[k][1] += e0.elapsed_time(e1)
tot = sum(v[1] for v in agg.values())
byname = collections.defaultdict(lambda: [0, 0.0])
for (name, ints), (c, ms) in agg.items():
    byname[name][0] += c; byname[name][1] += ms
print(f"sum of kernel times per forward: {tot/a.steps:.3f} ms")
for name, (c, ms) in sorted(byname.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:28s} calls/fwd={c//a.steps:4d} ms/fwd={ms/a.steps:8.3f} share={ms/tot*100:5.1f}%")
print("--- top 40 (name, int args) ---")
for (name, ints), (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{name:24s} {str(ints):50s} n/fwd={c//a.steps:3d} ms/fwd={ms/a.steps:7.3f} avg_us={ms/c*1e3:8.1f}")
