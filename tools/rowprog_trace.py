"""Debug: per-op phase timeline of the trunk program (CTA 0), via attndm_debug_set_rp_trace.
The hooks are compiled in only with -DATTNDM_RP_TRACE:
    python -m attentiondm_b200.build --variant=rp_trace
    ATTNDM_LIB=attentiondm_b200/libattndm_b200_rp_trace.so python tools/rowprog_trace.py      # on the GPU box"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from attentiondm_b200 import _ffi
from attentiondm_b200.engine import SamplerEngine
dev = torch.device("cuda")
bench.T_STEPS = 100
m, seq = bench.build_model(dev, bench.CONFIGS[os.environ.get("ATTNDM_CONFIG", "cifar10_w8a8")])
for n, q in m.qconvs():
    q.groups_range.data[..., 0] = -4.0
    q.groups_range.data[..., 1] = 6.0
    q.invalidate_cache(weights=False)
betas = torch.linspace(1e-4, 0.02, 1000, dtype=torch.float64).float().to(dev)
eng = SamplerEngine(m, seq, betas, 0.0, (256, 3, 32, 32))
eng.load_input(torch.randn(256, 3, 32, 32, device=dev))
with torch.no_grad():
    for _ in range(2):
        eng._with_staged(eng._step_body)
    torch.cuda.synchronize()
    h = torch.randn(256, 2, 2, 256, device=dev)
    tr = torch.zeros(256 * 8, dtype=torch.int64, device=dev)
    L = _ffi.lib(); L.attndm_debug_set_rp_trace.argtypes = [ctypes.c_void_p]
    L.attndm_debug_set_rp_trace(ctypes.c_void_p(tr.data_ptr()))
    eng.fused.run_trunk(h, eng.cur)
    torch.cuda.synchronize()
    L.attndm_debug_set_rp_trace(None)
t = tr.cpu().view(256, 8)
plan = eng.fused.trunk_plan
ops_ = plan.programs[0].ops
names = {1: "LOAD", 2: "POOL", 3: "STORE", 4: "COPY", 5: "CONV", 6: "FCONV", 7: "ATTN1", 8: "SCADD"}
t0 = int(t[0, 0])
tot = {}
for i, o in enumerate(ops_):
    r = [int(v) for v in t[i]]
    dur = (r[6] - r[0]) / 1000.0
    key = names[o["type"]] + (f" {o.get('C')}->{o.get('O')} pre{o.get('pre',0)}" if o["type"] in (5, 6) else "")
    tot.setdefault(key, [0, 0.0]); tot[key][0] += 1; tot[key][1] += dur
    if i < 40:
        if o["type"] == 5:
            print(f"{i:3d} {key:22s} start={(r[0]-t0)/1e3:8.2f} wait_prm={(r[1]-r[0])/1e3:5.2f} A1={(r[2]-r[1])/1e3:5.2f} A2={(r[3]-r[2])/1e3:5.2f} Bwait={(r[7]-r[3])/1e3:5.2f} B={(r[4]-r[3])/1e3:5.2f} C={(r[5]-r[4])/1e3:5.2f} sync={(r[6]-r[5])/1e3:5.2f} total={dur:5.2f}")
        else:
            print(f"{i:3d} {key:22s} start={(r[0]-t0)/1e3:8.2f} total={dur:5.2f}")
print("--- totals")
for k, (n, d) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:24s} n={n:3d} total={d:8.1f} us avg={d/n:6.2f}")
