"""Standalone timing of the int8 conv kernels on the shapes that dominate the CIFAR step (batch 256).
CUDA events on the launching stream, 3 warm-ups, operands larger than L2 for the big shapes."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import attentiondm_b200 as A
from attentiondm_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--shapes", default="all")
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--impl", default="tc")
ap.add_argument("--graph", type=int, default=1)
ap.add_argument("--stats", type=int, default=0, help="1: the conv also emits the GroupNorm statistics of its output")
a = ap.parse_args()
dev = torch.device("cuda")
SHAPES = {
    "c128_32": (256, 32, 32, 128, 128, 3), "c256_32": (256, 32, 32, 256, 128, 3), "c128_16": (256, 16, 16, 128, 128, 3),
    "c128_8": (256, 8, 8, 128, 128, 3), "n256_32": (256, 32, 32, 256, 128, 1), "t256_1": (256, 1, 1, 256, 256, 1),
    "t1024_1": (256, 1, 1, 1024, 256, 1), "out_32": (256, 32, 32, 128, 3, 3), "in_32": (256, 32, 32, 3, 128, 3),
    # LSUN church (batch 8): the 256x256 and 128x128 layers; CelebA (batch 64): 64x64
    "c128_256": (8, 256, 256, 128, 128, 3), "c128_128": (8, 128, 128, 128, 128, 3), "c128_64": (64, 64, 64, 128, 128, 3),
}
names = [n for n in SHAPES if not n.startswith("c128_") or n in ("c128_32", "c128_16", "c128_8")] if a.shapes == "all" else a.shapes.split(",")
impl = ops.CONV_TCGEN05 if a.impl == "tc" else ops.CONV_SIMT
for name in names:
    B, H, W, C, O, k = SHAPES[name]
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, H, W, C, generator=g).to(dev)
    w = ((torch.rand(O, C, k, k, generator=g) * 2 - 1) / (C * k * k) ** 0.5).to(dev)
    flat = w.reshape(O, -1)
    ws = A.AsymmetricQuantFunction.apply(w, 8, flat.min(1)[0], flat.max(1)[0])
    fl = ws.reshape(O, -1)
    pack = ops.weight_to_i8(ops.weight_clamp_pack(ws, fl.min(1)[0], fl.max(1)[0]), 8)
    assert pack.on_grid
    sv = torch.full((C,), 25.5, device=dev); zv = torch.full((C,), 26.0, device=dev)
    codes, rowsum, _ = ops.act_quant(x, sv, zv, 8, want_codes=True, halo=(k == 3))
    mult = (1.0 / (25.5 * pack.w_scale.double())).float().contiguous()
    azp = torch.tensor([26], dtype=torch.int32, device=dev)
    bias = torch.zeros(O, device=dev)
    out = torch.empty(B, H, W, O, device=dev)
    res = torch.randn(B, H, W, O, device=dev)
    gst = torch.zeros(B, 32, 2, dtype=torch.float64, device=dev) if a.stats and O % 32 == 0 else None
    for use_res in (False, True):
        for _ in range(3):
            ops.qconv_i8(codes, rowsum, B, H, W, C, pack, k * k, mult, azp, bias, res if use_res else None, impl=impl, out=out, gn_stats_out=gst)
        if a.graph:
            # the launches are replayed from a CUDA graph: for kernels shorter than ~35 us the host-side launch path
            # (tensor-map encodes + ctypes) is slower than the kernel and event timing of eager launches measures IT
            st = torch.cuda.Stream()
            st.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(st):
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr, stream=st):
                    for i in range(a.iters):
                        ops.qconv_i8(codes, rowsum, B, H, W, C, pack, k * k, mult, azp, bias, res if use_res else None, impl=impl, out=out, gn_stats_out=gst)
                gr.replay()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                gr.replay()
                e1.record(st)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.iters
        else:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.iters + 1)]
            ev[0].record()
            for i in range(a.iters):
                ops.qconv_i8(codes, rowsum, B, H, W, C, pack, k * k, mult, azp, bias, res if use_res else None, impl=impl, out=out, gn_stats_out=gst)
                ev[i + 1].record()
            torch.cuda.synchronize()
            ms = sum(ev[i].elapsed_time(ev[i + 1]) for i in range(a.iters)) / a.iters
        flops = 2.0 * B * H * W * O * C * k * k
        by = codes.numel() + out.numel() * 4 * (2 if use_res else 1) + pack.qw.numel()
        print(f"{name:8s} res={int(use_res)} {ms*1e3:8.1f} us  {flops/ms/1e9:8.1f} TOP/s  {by/ms/1e6:7.0f} GB/s (algorithmic){' +stats' if gst is not None else ''}")
