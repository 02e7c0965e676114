"""Where the end-to-end pass (keep='all') spends its wall time: pinned allocation, replay loop, history copies."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import attentiondm_b200 as A
from attentiondm_b200.engine import SamplerEngine
dev = torch.device("cuda")
c = bench.CONFIGS["cifar10_w8a8"]
m, seq = bench.build_model(dev, c)
for n, q in m.qconvs():
    q.groups_range.data[..., 0] = -4.0
    q.groups_range.data[..., 1] = 6.0
    q.invalidate_cache(weights=False)
betas = torch.linspace(1e-4, 0.02, 1000, dtype=torch.float64).float().to(dev)
x_host = torch.randn(256, 3, 32, 32).pin_memory()
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    h = torch.empty((2, 100, 256, 32, 32, 3), dtype=torch.float32, device="cpu", pin_memory=True)
    t1 = time.perf_counter()
    del h
    xs, x0s = A.generalized_steps(x_host.to(dev, non_blocking=True), seq, m, betas, eta=0.0, keep="all")
    torch.cuda.synchronize(); t2 = time.perf_counter()
    xs2, _ = A.generalized_steps(x_host.to(dev, non_blocking=True), seq, m, betas, eta=0.0, keep="last")
    torch.cuda.synchronize(); t3 = time.perf_counter()
    print(f"iter {it}: pinned alloc 630 MB {1e3*(t1-t0):.1f} ms; keep=all {1e3*(t2-t1):.1f} ms; keep=last {1e3*(t3-t2):.1f} ms")
    m.reset_index_seq()
# raw D2H rate
src = torch.empty(100, 256, 32, 32, 3, device=dev)
dst = torch.empty(100, 256, 32, 32, 3, pin_memory=True)
torch.cuda.synchronize(); t0 = time.perf_counter(); dst.copy_(src, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"D2H 315 MB: {1e3*(t1-t0):.1f} ms = {0.3146/(t1-t0):.1f} GB/s")
