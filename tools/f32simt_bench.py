"""Small-M fp32 convs (the time_embed linears: 256 x 1024 -> 1024) on the SIMT kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from attentiondm_b200 import ops
dev = torch.device("cuda")
for (B, C, O) in [(256, 256, 1024), (256, 1024, 1024), (256, 128, 512)]:
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, 1, 1, C, generator=g).to(dev)
    w = (torch.randn(O, 1, C, generator=g) / C ** 0.5).to(dev)
    bias = torch.randn(O, generator=g).to(dev)
    gph = torch.cuda.CUDAGraph()
    for _ in range(3): ops.conv_f32(x, w, bias)
    torch.cuda.synchronize()
    with torch.cuda.graph(gph):
        for _ in range(20): y = ops.conv_f32(x, w, bias)
    gph.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gph.replay(); e1.record(); torch.cuda.synchronize()
    want = (x.double().reshape(B, C) @ w.double().reshape(O, C).t() + bias.double())
    err = (y.double().reshape(B, O) - want).abs().max().item() / want.abs().max().item()
    print(f"{B} x {C} -> {O}: {e0.elapsed_time(e1) / 20 * 1e3:6.1f} us per launch (graph of 20)   err {err:.1e}   checksum {y.double().sum().item():.10f}")
