"""fp32 channel_proj GEMM (CIFAR: 16384 x 768 -> 512): SIMT fp32 kernel vs the 3xTF32 tensor-core path (split + GEMM);
error of both against an fp64 reference."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from attentiondm_b200 import ops
dev = torch.device("cuda")
B, H, W, C, O = 256, 8, 8, 768, 512
g = torch.Generator().manual_seed(0)
x = (torch.randn(B, H, W, C, generator=g) * 1.7 + 0.2).to(dev)
w = (torch.randn(O, C, generator=g) / C ** 0.5).to(dev)
bias = torch.randn(O, generator=g).to(dev)
ws = ops.split_tf32(w)
w3 = w.view(O, 1, C).contiguous()
def t(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
print("simt fp32      %7.1f us" % t(lambda: ops.conv_f32(x, w3, bias)))
print("3xTF32 (total) %7.1f us" % t(lambda: ops.conv1x1_f32_tc(x, ws, bias)))
want = (x[:16].double().reshape(-1, C) @ w.double().t() + bias.double())
for name, y in (("simt", ops.conv_f32(x, w3, bias)), ("3xTF32", ops.conv1x1_f32_tc(x, ws, bias))):
    e = (y[:16].double().reshape(-1, O) - want).abs()
    print(f"{name:7s} max err / max|out| = {e.max().item() / want.abs().max().item():.2e}   rms err / rms out = {(e.pow(2).mean().sqrt() / want.pow(2).mean().sqrt()).item():.2e}")
