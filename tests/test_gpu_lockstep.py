"""Full-size lock-step parity against the oracle (GPU): BASELINE.json configs 2 / 4 / 5 shapes (CIFAR-10, CelebA,
LSUN church UNets) plus the tiny fixture model, teacher-forced step by step.

Per step, oracle/lockstep.py compares the CUDA run with the CPU oracle layer by layer:
  * seed flips: our quantizer kernels' integer codes vs the oracle arithmetic on the SAME layer input
    (GroupNorm+SiLU producer included) -- and the same count for the reference arithmetic run in CUDA eager
    (torch's own kernels) vs CPU, the comparator;
  * trajectory flips: codes of the CUDA run vs codes of the oracle's own whole-network run;
  * in-situ operator error of every QConv2d call; eps rel-L2 of the step (ours, and torch-CUDA-eager's).
Bars (north_star: 1e-3 on each step's eps):
  * every operator call in which no code flipped: rel-L2 <= 1e-3 (measured ~1e-7);
  * every step in which no code differs anywhere from the oracle's run: eps rel-L2 <= 1e-3;
  * our seed-flip rate is no worse than the reference arithmetic's own CUDA-vs-CPU rate (x2 + slack for counting
    noise) and <= 1e-5 of the activations;
  * steps in which codes did flip: their summed eps error is no more than twice that of the reference arithmetic's own
    CUDA-eager-vs-CPU runs on the same steps (both sit at the avalanche level, 1e-2 ... 7e-2), each is capped at
    0.15, and every measured value is recorded.
With ATTNDM_PARITY_OUT=<dir> every case writes <dir>/parity_<name>.json (committed as profiles/parity_r02.json).
"""
import json
import os
import time

import pytest
import torch

from oracle import lockstep as L
from oracle import restate as R
from oracle import synth as S
from tests.util import build_cuda_model, rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda"

SPECS = {
    "tiny": lambda T: S.tiny_spec(T=T, bitwidth=8),
    "cifar10": lambda T: S.cifar_spec(T=T),
    "celeba": lambda T: S.celeba_spec(T=T),
    "church": lambda T: S.church_spec(T=T),
}


def run_lockstep(name, B, T, seed=1, calib_insitu=False):
    import attentiondm_b200 as A
    from oracle import insitu
    threads = torch.get_num_threads()
    torch.set_num_threads(os.cpu_count() or 1)           # the CPU oracle at full size; no fixture depends on it here
    try:
        t_start = time.time()
        spec = SPECS[name](T)
        sd = S.synth_state_dict(spec, seed=seed)
        m = build_cuda_model(spec, sd)
        betas = R.beta_schedule_linear()
        size = spec.image_size
        x = torch.randn(B, 3, size, size, generator=torch.Generator().manual_seed(9))
        # ---- calibrate on the GPU (tables then shared with the oracle, which isolates the sampler) ----
        m.set_calibrate(True)
        out = dict(config=name, batch=B, steps=T, layers=len(m.qconvs()))
        if calib_insitu:
            rec = insitu.record_layers(m)
            A.generalized_steps(x.to(DEV), spec.seq, m, betas.to(DEV), eta=0.0, keep="last")
            errs = insitu.layer_errors(m, rec, calibrate=True)     # also checks every group table against the oracle's
            L.unwrap(m)
            out["calibration_insitu"] = dict(calls=len(errs), worst_rel=max(e for e, _ in errs),
                                             over_1e_3=sum(1 for e, _ in errs if e > 1e-3))
            del rec
        else:
            A.generalized_steps(x.to(DEV), spec.seq, m, betas.to(DEV), eta=0.0, keep="last")
        m.set_calibrate(False)
        m.reset_index_seq()
        out["int8_layers"] = sum(1 for _, q in m.qconvs() if q.int8_ok_all_steps())
        orc = R.Oracle(spec, sd)
        for n, q in m.qconvs():
            orc.sd[n + ".groups_range"] = q.groups_range.data.detach().cpu().clone()
        # ---- teacher-forced steps along the ORACLE's trajectory ----
        rseq = list(reversed(spec.seq))
        rnext = list(reversed([-1] + list(spec.seq)[:-1]))
        xt = x
        steps = []
        for k in range(T):
            res, eps_o = L.analyze_step(m, orc, xt, rseq[k], k)
            steps.append(res)
            tt = torch.full((B,), float(rseq[k]))
            at = R.compute_alpha(betas, tt.long())
            an = R.compute_alpha(betas, torch.full_like(tt, rnext[k]).long())
            xt, _ = R.ddim_update(xt, eps_o, at, an, 0.0, torch.zeros_like(xt))
        out["summary"] = L.summarize(steps)
        # ---- free-running sample through the CUDA-graph engine vs the oracle's final image ----
        m.reset_index_seq()
        xs, _ = A.generalized_steps(x.to(DEV), spec.seq, m, betas.to(DEV), eta=0.0, keep="last")
        out["final_image_rel"] = rel_l2(xs[-1], xt)
        out["per_layer"] = [[dict(name=l["name"], elements=l["elements"], seed_flips=l["seed_flips"],
                                  seed_flips_torch_cuda=l["seed_flips_torch_cuda"], traj_flips=l["traj_flips"],
                                  insitu_rel=l["insitu_rel"]) for l in s["layers"]
                             if l["seed_flips"] or l["seed_flips_torch_cuda"] or l["insitu_rel"] > 1e-5][:40]
                            for s in steps]
        out["lib"] = os.path.basename(os.environ.get("ATTNDM_LIB", "libattndm_b200.so"))
        out["seconds"] = time.time() - t_start
        return out, steps
    finally:
        torch.set_num_threads(threads)


def _check_and_dump(name, out, steps):
    sm = out["summary"]
    print(f"\n[lockstep {name} B={out['batch']} T={out['steps']} lib={out['lib']}] "
          f"eps rel-L2 per step {['%.1e' % e for e in sm['eps_rel_per_step']]} "
          f"(reference arithmetic, CUDA eager vs CPU: {['%.1e' % e for e in sm['eps_rel_torch_cuda_per_step']]}); "
          f"seed flips {sm['seed_flips']} of {sm['activations_compared']} ({sm['seed_flip_rate']:.1e}; "
          f"torch CUDA vs CPU {sm['seed_flips_torch_cuda']}, {sm['seed_flip_rate_torch_cuda']:.1e}); "
          f"trajectory flips per step {sm['traj_flips_per_step']}; worst in-situ {sm['worst_insitu_rel']:.1e}; "
          f"final image {out['final_image_rel']:.1e}; {out['seconds']:.0f} s")
    d = os.environ.get("ATTNDM_PARITY_OUT")
    if d:
        os.makedirs(d, exist_ok=True)
        tag = "" if out["lib"] == "libattndm_b200.so" else "_" + out["lib"].replace("libattndm_b200_", "").replace(".so", "")
        with open(os.path.join(d, f"parity_{name}{tag}.json"), "w") as f:
            json.dump(out, f, indent=1)
    for s in steps:
        for l in s["layers"]:
            if l["seed_flips"] == 0:
                assert l["insitu_rel"] <= 1e-3, (name, s["step"], l)          # operator bar, flip-free call
        if s["totals"]["traj_flips"] == 0:
            assert s["eps_rel"] <= 1e-3, (name, s["step"], s["eps_rel"])       # step bar, flip-free step
        assert s["eps_rel"] <= 0.15, (name, s["step"], s["eps_rel"])           # avalanche cap (measured: 1e-2 ... 7e-2)
    # once codes flip, the deviation is that of the reference arithmetic against itself across devices
    assert sum(sm["eps_rel_per_step"]) <= 2 * sum(sm["eps_rel_torch_cuda_per_step"]) + 2e-2, sm
    assert sm["seed_flip_rate"] <= 1e-5, sm["seed_flip_rate"]
    assert sm["seed_flips"] <= 2 * sm["seed_flips_torch_cuda"] + 8, (sm["seed_flips"], sm["seed_flips_torch_cuda"])
    assert out["int8_layers"] == out["layers"]
    if "calibration_insitu" in out:
        assert out["calibration_insitu"]["worst_rel"] <= 1e-3, out["calibration_insitu"]


def test_lockstep_tiny():
    out, steps = run_lockstep("tiny", B=2, T=4, seed=3, calib_insitu=True)
    _check_and_dump("tiny", out, steps)


def test_lockstep_cifar10_full_size():
    out, steps = run_lockstep("cifar10", B=int(os.environ.get("ATTNDM_LOCKSTEP_B", "2")), T=4, calib_insitu=True)
    assert out["layers"] == 198
    _check_and_dump("cifar10", out, steps)


def test_lockstep_celeba_full_size():
    out, steps = run_lockstep("celeba", B=1, T=2)
    assert out["layers"] == 252
    _check_and_dump("celeba", out, steps)


def test_lockstep_church_full_size():
    out, steps = run_lockstep("church", B=1, T=2)
    assert out["layers"] == 305
    _check_and_dump("church", out, steps)
