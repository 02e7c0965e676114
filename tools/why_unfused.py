import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch, bench
from attentiondm_b200 import rowprog
from attentiondm_b200.engine import SamplerEngine
dev = torch.device("cuda")
bench.T_STEPS = 100
for name in sys.argv[1:] or ("church_w8a8", "celeba_w8a8", "cifar10_w4_attn"):
    CFG = bench.CONFIGS[name]
    m, seq = bench.build_model(dev, CFG)
    for n, q in m.qconvs():
        q.groups_range.data[..., 0] = -4.0
        q.groups_range.data[..., 1] = 6.0
        q.invalidate_cache(weights=False)
    betas = torch.linspace(1e-4, 0.02, 1000, dtype=torch.float64).float().to(dev)
    eng = SamplerEngine(m, seq, betas, 0.0, (CFG["batch"], 3, CFG["size"], CFG["size"]))
    print(name, "fused:", eng.fused is not None, "why:", rowprog.last_unfusable)
    del eng, m
    torch.cuda.empty_cache()
