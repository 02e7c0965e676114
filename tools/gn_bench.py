"""Timing of GroupNorm+SiLU+quantize on the large feature maps: two-pass (gn_stats + act_quant_rows) vs the
one-pass cluster kernel.  CUDA events, batch 256; the tensors are larger than L2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from attentiondm_b200 import ops
dev = torch.device("cuda")
for (B, H, W, C) in [(256, 32, 32, 128), (256, 32, 32, 256)]:
    if not ops.gn_fits_cluster(H, W, C):
        print(H, W, C, "cluster kernel not taken"); continue
    g = torch.Generator().manual_seed(0)
    xs = [torch.randn(B, H, W, C, generator=g).to(dev) for _ in range(2)]
    gamma = torch.ones(C, device=dev); beta = torch.zeros(C, device=dev)
    sv = torch.full((C,), 25.5, device=dev); zv = torch.full((C,), 26.0, device=dev)
    def two(x):
        gn = ops.GnArgs(ops.gn_stats(x), gamma, beta, 1e-6)
        return ops.act_quant(x, sv, zv, 8, ops.PRE_GN_SILU, gn, want_codes=True, halo=True)
    def one(x):
        gn = ops.GnArgs(None, gamma, beta, 1e-6)
        return ops.act_quant(x, sv, zv, 8, ops.PRE_GN_SILU, gn, want_codes=True, halo=True)
    for name, f in (("two-pass", two), ("one-pass cluster", one)):
        for i in range(4): f(xs[i & 1])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        e0.record()
        for i in range(n): f(xs[i & 1])
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / n * 1e3
        by = xs[0].numel() * 4 + B * (H + 2) * (W + 2) * C
        print(f"{H}x{W}x{C} {name:18s} {us:7.1f} us   {by / us / 1e3:6.0f} GB/s (one read + code write)")
