"""Debug: per-role timeline of the halo conv kernel (CTA 0, first tiles), via the trace hook."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import attentiondm_b200 as A
from attentiondm_b200 import ops, _ffi
dev = torch.device("cuda")
shape = {"c128_32": (256, 32, 32, 128, 128, 3), "out_32": (256, 32, 32, 128, 3, 3), "c256_32": (256, 32, 32, 256, 128, 3)}[sys.argv[1] if len(sys.argv) > 1 else "c128_32"]
B, H, W, C, O, k = shape
g = torch.Generator().manual_seed(0)
x = torch.randn(B, H, W, C, generator=g).to(dev)
w = ((torch.rand(O, C, k, k, generator=g) * 2 - 1) / (C * k * k) ** 0.5).to(dev)
flat = w.reshape(O, -1)
ws = A.AsymmetricQuantFunction.apply(w, 8, flat.min(1)[0], flat.max(1)[0])
fl = ws.reshape(O, -1)
pack = ops.weight_to_i8(ops.weight_clamp_pack(ws, fl.min(1)[0], fl.max(1)[0]), 8)
sv = torch.full((C,), 25.5, device=dev); zv = torch.full((C,), 26.0, device=dev)
codes, rowsum, _ = ops.act_quant(x, sv, zv, 8, want_codes=True, halo=True)
mult = (1.0 / (25.5 * pack.w_scale.double())).float().contiguous()
azp = torch.tensor([26], dtype=torch.int32, device=dev)
bias = torch.zeros(O, device=dev); out = torch.empty(B, H, W, O, device=dev)
for _ in range(3):
    ops.qconv_i8(codes, rowsum, B, H, W, C, pack, 9, mult, azp, bias, out=out)
tr = torch.zeros(4 * 32 * 4, dtype=torch.int64, device=dev)
L = _ffi.lib(); L.attndm_debug_set_tc_trace.argtypes = [ctypes.c_void_p]
L.attndm_debug_set_tc_trace(ctypes.c_void_p(tr.data_ptr()))
ops.qconv_i8(codes, rowsum, B, H, W, C, pack, 9, mult, azp, bias, out=out)
torch.cuda.synchronize()
L.attndm_debug_set_tc_trace(None)
t = tr.cpu().view(4, 32, 4)
t0 = int(t[t > 0].min())
names = ["Aprod", "MMA", "epi0", "epi1"]
evn = [["wait_empty", "got_empty", "-", "-"], ["start", "tmem_free", "a_full", "issued"], ["ready", "tmem_full", "tmem_read", "done"], ["ready", "tmem_full", "tmem_read", "done"]]
for it in range(8):
    for r in range(4):
        row = [(int(v) - t0) / 1000.0 if v > 0 else float('nan') for v in t[r, it]]
        print(f"it={it} {names[r]:6s} " + "  ".join(f"{evn[r][e]}={row[e]:8.2f}us" for e in range(4)))
