"""Launch the HBM-bound kernels of the path on the bench shape (batch 256, 32x32, 128 channels: 134 MB fp32 > L2),
for `ncu --set full` captures (tools/gpu_round2_b.sh) -- GroupNorm statistics, GroupNorm+SiLU+quantize, plain
quantize, calibration min/max, calibration mix, DDIM update.  Prints CUDA-event times when run plainly."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from attentiondm_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B, H, W, C = 256, 32, 32, 128
g = torch.Generator().manual_seed(1)
x = torch.randn(B, H, W, C, generator=g).to(dev)
sv = torch.full((C,), 25.5, device=dev)
zv = torch.full((C,), 26.0, device=dev)
gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
stats = ops.gn_stats(x)
gn = ops.GnArgs(stats, gamma, beta, 1e-6)
gr = torch.tensor([[-4.0, 6.0]] * 8, device=dev)
sw = torch.full((8, C), 0.125, device=dev)
eps = torch.randn_like(x)
coef = torch.tensor([0.6, 0.8, 0.81, 0.0, 0.59, 500.0, 0, 0], device=dev)
xo = torch.empty_like(x)
cases = [
    ("gn_stats", 4, lambda: ops.gn_stats(x, out=stats)),
    ("act_quant_rows<GN+SiLU>", 5, lambda: ops.act_quant(x, sv, zv, 8, ops.PRE_GN_SILU, gn, want_codes=True, halo=True)),
    ("act_quant_rows<none>", 5, lambda: ops.act_quant(x, sv, zv, 8, want_codes=True, halo=True)),
    ("minmax_c", 4, lambda: ops.minmax_c(x)),
    ("calib_mix", 8, lambda: ops.calib_mix(x, gr, sw, 8)),
    ("ddim_step", 12, lambda: ops.ddim_step(x, eps, coef, None, x_next=xo)),
]
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 5
for name, bpe, fn in cases:
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"{name:28s} {ms * 1e3:8.1f} us  {bpe * x.numel() / ms / 1e6:8.0f} GB/s (algorithmic {bpe} B/elem)")
