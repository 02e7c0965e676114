"""Debug: per-role timeline of the halo conv kernel, via the trace hooks.

The hooks are compiled in only with -DATTNDM_TC_TRACE (each costs a global load on the kernel's critical path):
    ATTNDM_NVCC_EXTRA=-DATTNDM_TC_TRACE python -m attentiondm_b200.build      # trace build of the library
    TRACE_CTA=5 TRACE_ITS=9 python tools/conv_trace.py c128_32                # on the GPU box
    python -m attentiondm_b200.build                                          # back to the normal build
Prints, for CTA TRACE_CTA, the globaltimer stamps of the MMA warp (start / accumulator free / halo full / issued),
the true MMA completion (an idle warp waits on the same barrier) and the eight epilogue warps (accumulator seen /
first block loaded / first block done / second block done); the SM clock per tile; and for ALL CTAs the start,
weights-resident and end times and the per-tile issue intervals."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import attentiondm_b200 as A
from attentiondm_b200 import ops, _ffi
dev = torch.device("cuda")
shape = {"c128_32": (256, 32, 32, 128, 128, 3), "out_32": (256, 32, 32, 128, 3, 3), "c256_32": (256, 32, 32, 256, 128, 3), "n256_32": (256, 32, 32, 256, 128, 1)}[sys.argv[1] if len(sys.argv) > 1 else "c128_32"]
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
    ops.qconv_i8(codes, rowsum, B, H, W, C, pack, k * k, mult, azp, bias, out=out)
tr = torch.zeros(4096 + 148 * 32, dtype=torch.int64, device=dev)
os.environ['ATTNDM_TRACE_CTA'] = os.environ.get('TRACE_CTA', '0')
L = _ffi.lib(); L.attndm_debug_set_tc_trace.argtypes = [ctypes.c_void_p]
L.attndm_debug_set_tc_trace(ctypes.c_void_p(tr.data_ptr()))
ops.qconv_i8(codes, rowsum, B, H, W, C, pack, k * k, mult, azp, bias, out=out)
torch.cuda.synchronize()
L.attndm_debug_set_tc_trace(None)
full = tr.cpu()
t = full[:2048].view(16, 32, 4)
t0 = int(t[t > 0].min())
names = ["Aprod", "MMA", "geo1", "geo"] + [f"epi{i}" for i in range(12)]
evn = [["wait_empty", "got_empty", "-", "-"], ["start", "tmem_free", "a_full", "issued"], ["finish_top", "got_empty", "arrived", "-"], ["finish_top", "got_empty", "arrived", "-"]] + [["tmem_full", "c0_loaded", "c0_done", "last_done"]] * 12
for it in range(int(os.environ.get('TRACE_ITS', '8'))):
    for r in (1, 2, 3, 4, 8, 12):
        row = [(int(v) - t0) / 1000.0 if v > 0 else float('nan') for v in t[r, it]]
        print(f"it={it} {names[r]:6s} " + "  ".join(f"{evn[r][e]}={row[e]:8.2f}us" for e in range(4)))

v = t[t > 0]
print(f"CTA {os.environ['ATTNDM_TRACE_CTA']}: span first..last event {(int(v.max()) - int(v.min())) / 1000.0:.2f} us")

ck = full[2048:2048 + 16]
st = t[1, :16, 0]
for i in range(1, 14):
    if ck[i] > 0 and ck[i - 1] > 0:
        print(f"tile {i}: {(int(ck[i]) - int(ck[i-1])) / max(1, int(st[i]) - int(st[i-1])) * 1000:.0f} MHz")

sp = full[2100:2100 + 4 * 148].view(148, 4)
ok = sp[:, 0] > 0
k0 = int(sp[ok, 0].min())
import statistics
st_ = [(int(v) - k0) / 1000 for v in sp[ok, 0]]; wr = [(int(v) - k0) / 1000 for v in sp[ok, 1]]; en = [(int(v) - k0) / 1000 for v in sp[ok, 2]]
print(f"CTAs {int(ok.sum())}: start min/med/max {min(st_):.2f}/{statistics.median(st_):.2f}/{max(st_):.2f}  weights-resident med/max {statistics.median(wr):.2f}/{max(wr):.2f}  end min/med/max {min(en):.2f}/{statistics.median(en):.2f}/{max(en):.2f} us")

pt = full[4096:4096 + 148 * 32].view(148, 32)
import numpy as np
a = pt.numpy().astype(np.int64)
d = np.diff(a, axis=1).astype(np.float64) / 1000.0
valid = (a[:, 1:] > 0) & (a[:, :-1] > 0)
if int(os.environ.get('TRACE_CTA', '0')) < 148:
    valid[int(os.environ.get('TRACE_CTA', '0'))] = False
print("per-tile issue-to-issue interval, all untraced CTAs: mean %.2f  p10 %.2f  p50 %.2f  p90 %.2f  max %.2f us" % (
    d[valid].mean(), np.percentile(d[valid], 10), np.percentile(d[valid], 50), np.percentile(d[valid], 90), d[valid].max()))
for i in range(0, 13):
    col = d[:, i][valid[:, i]]
    print(f"  tile {i+1}: mean {col.mean():.2f} p50 {np.percentile(col,50):.2f} max {col.max():.2f}")
first = (a[:, 0] - k0) / 1000.0
print("first tile issued at: mean %.2f max %.2f us" % (first.mean(), first.max()))
