"""Headline benchmark: CIFAR-10 UNet W8A8 (group-wise activation quant), DDIM-100,
batch 256 per GPU -> images/sec/box (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W            # our arm (torchrun for N > 1)
    python bench.py --impl reference ...                     # the reference's CPU path (oracle port)

One "step" = one full DDIM-100 sampling pass (100 UNet forwards + 100 DDIM updates)
over one batch of 256 synthetic Gaussian latents per GPU, random-init weights
snapped to the int8 grid (H1), calibrated once (untimed) on the same kind of
latents.  `value` times the pass with the latents already in HBM; `e2e` times the
public API call (attentiondm_b200.generalized_steps) with the latents in pinned
host memory and the final images copied back, both inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T_STEPS = 100
BATCH = 256
GFLOP_PER_IMG_STEP = 2.769          # SURVEY.md App. B (conv MACs x 2, CIFAR config)


def peaks():
    p = dict(hbm_gbs=6650.0, bf16_tflops=1590.0, source="fallback (B200_PROFILING.md)")
    f = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(f):
        try:
            d = json.load(open(f))
            p = dict(hbm_gbs=float(d["hbm_gbs"]), bf16_tflops=float(d["bf16_tflops"]),
                     bf16_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                     source="MEASURED_PEAKS.json")
        except Exception:
            pass
    return p


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        sm.sort()
        return dict(sm_mhz=(sm[len(sm) // 2] if sm else None), sm_max_mhz=mx, reasons=sorted(reasons),
                    samples=len(sm))


def cifar_config():
    ns = argparse.Namespace
    return ns(data=ns(channels=3, image_size=32, dataset="CIFAR10", rescaled=True, logit_transform=False),
              model=ns(ch=128, ch_mult=[1, 2, 2, 2], num_res_blocks=2, dropout=0.1, var_type="fixedlarge"),
              diffusion=ns(beta_schedule="linear", beta_start=0.0001, beta_end=0.02, num_diffusion_timesteps=1000))


# ---------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's path on the host cores
# ---------------------------------------------------------------------------
def cpu_oracle_rate(forwards=2, batch=4, threads=None):
    """images/sec of the reference's CPU path (oracle/restate.py), DDIM-100: a bounded sample of
    `forwards` quantized UNet forwards at `batch`, extrapolated to 100 steps per image."""
    import torch
    from oracle import restate as R
    from oracle import synth as S
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    spec = S.cifar_spec(T=T_STEPS, bitwidth=8)
    sd = S.synth_state_dict(spec, seed=0)
    for k in sd:
        if k.endswith("groups_range"):
            sd[k][..., 0] = -4.0
            sd[k][..., 1] = 6.0
    orc = R.Oracle(spec, sd)
    x = torch.randn(batch, 3, 32, 32, generator=torch.Generator().manual_seed(1234))
    t = torch.full((batch,), 990.0)
    with torch.no_grad():
        orc.forward(x, t)                      # warm-up
        t0 = time.perf_counter()
        for _ in range(forwards):
            orc.forward(x, t)
        dt = (time.perf_counter() - t0) / forwards
    return batch / (dt * T_STEPS), dt, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rates = []
    for _ in range(max(1, args.warmup)):
        pass
    rate, dt, threads = cpu_oracle_rate(forwards=max(1, args.steps), batch=4)
    sample = f"{max(1, args.steps)} quantized UNet forwards at batch 4 (oracle port of the reference, CPU), x100 steps/image"
    line = {
        "impl": "reference", "metric": "CIFAR-10 W8A8 DDIM-100 images/sec/box", "value": rate, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3 * T_STEPS * (BATCH / 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "CIFAR-10 32x32 UNet (198 QConv2d) W8A8 fake-quant, DDIM 100 steps, CPU host cores"},
        "cpu_baseline": {"value": rate, "unit": "images/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------
def build_model(dev, seed=0):
    import torch
    import attentiondm_b200 as A
    torch.manual_seed(seed)
    cfg = cifar_config()
    args = argparse.Namespace(bitwidth=8, timesteps=T_STEPS, skip_type="uniform", eta=0.0)
    seq = range(0, 1000, 1000 // T_STEPS)
    m = A.Model(cfg, quantization=True, sequence=seq, args=args).to(dev).eval()
    m.materialize_lazy_layers()
    for mod in m.modules():                       # gamma = 0 makes attention a numerical no-op (H7)
        if isinstance(mod, A.EnhancedQSelfAttention):
            mod.gamma.data.fill_(0.5)
    m.snap_weights_()                             # H1: weights on the int8 grid, shared by every arm
    return m, seq


def dominant_kernel_roofline(dev, pk):
    """The 128->128 3x3 conv at 32x32, batch 256 (33 % of the step's FLOPs), timed
    alone with CUDA events on the launching stream; operands (38 MB codes, 134 MB output) exceed L2."""
    import torch
    from attentiondm_b200 import ops
    B, H, W, C, O = BATCH, 32, 32, 128, 128
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, H, W, C, generator=g).to(dev)
    w = (torch.rand(O, C, 3, 3, generator=g) * 2 - 1).to(dev) / (C * 9) ** 0.5
    flat = w.reshape(O, -1)
    w_eff = ops.weight_clamp_pack(w, flat.min(1)[0], flat.max(1)[0])
    import attentiondm_b200 as A
    ws = A.AsymmetricQuantFunction.apply(w, 8, flat.min(1)[0], flat.max(1)[0])
    fl = ws.reshape(O, -1)
    pack = ops.weight_to_i8(ops.weight_clamp_pack(ws, fl.min(1)[0], fl.max(1)[0]), 8)
    sv = torch.full((C,), 25.5, device=dev)
    zv = torch.full((C,), 26.0, device=dev)
    codes, rowsum, _ = ops.act_quant(x, sv, zv, 8, want_codes=True, halo=True)
    mult = (1.0 / (25.5 * pack.w_scale.double())).float().contiguous()
    azp = torch.tensor([26], dtype=torch.int32, device=dev)
    bias = torch.zeros(O, device=dev)
    out = torch.empty(B, H, W, O, device=dev)
    res = {}
    for name, impl in (("tcgen05", ops.CONV_TCGEN05),):
        for _ in range(3):
            ops.qconv_i8(codes, rowsum, B, H, W, C, pack, 9, mult, azp, bias, impl=impl, out=out)
        n = 10
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
        ev[0].record()
        for i in range(n):
            ops.qconv_i8(codes, rowsum, B, H, W, C, pack, 9, mult, azp, bias, impl=impl, out=out)
            ev[i + 1].record()
        torch.cuda.synchronize()
        ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(n))
        res[name] = sum(ms) / n
    flops = 2.0 * B * H * W * O * C * 9
    ach = flops / (res["tcgen05"] * 1e-3) / 1e12
    peak = 2.0 * pk["bf16_tflops"]
    algo_bytes = codes.numel() + rowsum.numel() * 4 + out.numel() * 4 + pack.qw.numel()
    hbm_s = algo_bytes / (pk["hbm_gbs"] * 1e9)               # what HBM alone would need for the algorithmic bytes
    return {"bound": "tensor", "kernel": "qconv_i8_halo_kernel (128->128 3x3 @32x32, batch 256)", "achieved": ach,
            "peak": peak, "unit": "TOP/s", "frac": ach / peak,
            "peak_source": f"2 x bf16_tflops of {pk['source']}: tcgen05 kind::i8 measured at exactly 2x the bf16 "
                           "MMA rate on this part (tools/umma_rate_test.cu: 64 cycles per 128x128x32 MMA); "
                           "MEASURED_PEAKS.json has no int8 entry",
            # dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full (profiles/ncu_conv_r01.txt)
            "traffic": 119.57e6, "traffic_source": "profiles/ncu_conv_r01.txt (39.3 MB read + 80.3 MB written; "
                                                   "the rest of the 134 MB output is still in L2 at kernel end)",
            "ms_per_launch": res["tcgen05"], "algorithmic_bytes": algo_bytes,
            "hbm_gbs_at_this_time": algo_bytes / (res["tcgen05"] * 1e-3) / 1e9,
            "hbm_floor_ms": hbm_s * 1e3,
            "note": "fp32 activations between layers make this layer HBM-co-bound: 173 MB at the measured "
                    f"{pk['hbm_gbs']:.0f} GB/s is {hbm_s * 1e6:.0f} us, i.e. at most "
                    f"{flops / hbm_s / 1e12 / peak:.2f} of the int8 peak is reachable without changing the layout"}


def run_ours(args):
    import torch
    import torch.distributed as dist

    import attentiondm_b200 as A
    from attentiondm_b200 import _ffi
    from attentiondm_b200 import dist as adist
    from attentiondm_b200.engine import SamplerEngine

    rank, world = adist.init_from_env()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    assert _ffi.lib().attndm_device_supported() == 1, "bench needs an sm_100 (B200) device"
    pk = peaks()
    m, seq = build_model(dev, seed=0)             # same weights on every rank (replicated model)
    betas = torch.from_numpy(A.get_beta_schedule("linear", beta_start=1e-4, beta_end=0.02,
                                                 num_diffusion_timesteps=1000)).float().to(dev)
    gen = torch.Generator().manual_seed(1234 + rank)            # reference default seed 1234 (main.py:23)
    x_host = torch.randn(BATCH, 3, 32, 32, generator=gen).pin_memory()
    # ---- calibration (untimed): ranges all-reduced over ranks so every replica holds identical tables ----
    if world > 1:
        adist.install()
    xc = x_host[:32].to(dev)
    with torch.no_grad():
        m.set_calibrate(True)
        A.generalized_steps(xc, seq, m, betas, eta=0.0, keep="last")
        m.set_calibrate(False)
        m.reset_index_seq()
        n_i8 = sum(1 for _, q in m.qconvs() if q.int8_ok_all_steps())
        if rank == 0:
            for nm, q in m.qconvs():
                if not q.int8_ok_all_steps():
                    print(f"[bench] fp32 fallback layer {nm}: {q.int8_status()}", file=sys.stderr)
        eng = SamplerEngine.for_model(m, seq, betas, 0.0, (BATCH, 3, 32, 32))
        x_dev = x_host.to(dev)

        def one_pass_resident():
            eng.load_input(x_dev)
            eng.run_loaded()

        def one_pass_e2e():
            xs, _ = A.generalized_steps(x_host_dev_view(), seq, m, betas, eta=0.0, keep="last")
            return xs[-1]

        def x_host_dev_view():
            return x_host.to(dev, non_blocking=True)        # H2D inside the timed region

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        for _ in range(max(3, args.warmup)):
            one_pass_resident()
        torch.cuda.synchronize()
        launches0 = _ffi.launches
        barrier()
        cs = ClockSampler(local)
        cs.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            one_pass_resident()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        clocks = cs.stop()
        # ---- end-to-end through the public API, host buffers, copies inside the timed region ----
        m.reset_index_seq()
        one_pass_e2e()
        barrier()
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(args.steps):
            img = one_pass_e2e()
        t1.record()
        barrier()
        ms_e2e = t0.elapsed_time(t1)
        if world > 1:
            tt = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms, ms_e2e = float(tt[0]), float(tt[1])
        roof = dominant_kernel_roofline(dev, pk) if rank == 0 else None
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    ms_per_step = ms / args.steps
    value = world * BATCH / (ms_per_step * 1e-3)
    e2e_v = world * BATCH / (ms_e2e / args.steps * 1e-3)
    try:
        cpu_rate, cpu_dt, cpu_threads = cpu_oracle_rate(forwards=2, batch=4)
        cpu = {"value": cpu_rate, "unit": "images/s", "cores": cpu_threads, "kind": "port",
               "sample": "2 quantized UNet forwards at batch 4 (oracle port, CPU), x100 steps/image"}
    except Exception as e:                         # never lose the GPU line to a CPU hiccup
        cpu = {"value": None, "unit": "images/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}
    step_tops = GFLOP_PER_IMG_STEP * 1e9 * BATCH * T_STEPS / (ms_per_step * 1e-3) / 1e12
    line = {
        "metric": "CIFAR-10 W8A8 DDIM-100 images/sec/box", "value": value, "unit": "images/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int8 x int8 -> s32 (fp32 activations between layers)",
        "data": "synthetic",
        "config": {"workload": "CIFAR-10 32x32 UNet (configs/cifar10.yml, 198 QConv2d, random-init weights on the "
                               "int8 grid) W8A8 group-wise fake-quant, DDIM 100 steps, batch 256/GPU",
                   "batch_per_gpu": BATCH, "ddim_steps": T_STEPS, "parallelism": f"dp{world} (batch sharded)",
                   "l2": "per-step working set (>= 134 MB fp32 activations per layer at 32x32) exceeds the 126 MB L2",
                   "int8_layers": n_i8, "cuda_graph": True},
        "clocks": clocks,
        "e2e": {"value": e2e_v, "unit": "images/s", "h2d_bytes_per_step": int(x_host.numel() * 4),
                "d2h_bytes_per_step": int(img.numel() * 4)},
        "gpu_launches": int((eng.launches_per_step or 0) * T_STEPS * args.steps),
        "whole_step_conv_tops": step_tops,
        "roofline": roof, "cpu_baseline": cpu,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
