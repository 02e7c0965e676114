"""Multi-GPU plumbing: one process per GPU, the sample batch sharded across
ranks, and the ONE exchange the path has -- calibration range statistics.

Sampling is embarrassingly parallel (every op is per-sample), so ranks never
talk while sampling.  Calibration's per-channel activation min/max
(utils/quant_util.py:187-191) are batch statistics: to calibrate on the global
batch each QConv2d all-reduces [min_c || -max_c] with MIN before the init-range
floor and the grouping.  MIN/MAX are order independent, so the result is
bit-identical to single-process calibration on the concatenated batch
(SURVEY.md section 8e).  The first-calibrate search (utils/quant_util.py:237-254) scores nine candidate init
ranges with a mean over the batch: the nine sums and the element count are all-reduced with SUM, so every rank
chooses the same range (equal to the single-process choice up to fp64 summation order).  Backend: NCCL over NVLink on the GPUs, gloo in the CPU tests.
"""
import os

import torch
import torch.distributed as dist

from . import quant_util


def init_from_env(backend=None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / MASTER_* (torchrun).  Returns (rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world


def shard_bounds(n: int, rank: int, world: int):
    """Rows [lo, hi) of a global batch of n owned by `rank` (contiguous, sizes differ by <= 1)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_minmax(min_c: torch.Tensor, max_c: torch.Tensor, group=None):
    """Global per-channel (min, max): one all-reduce(MIN) over [min_c || -max_c]."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return min_c, max_c
    c = min_c.numel()
    buf = torch.cat([min_c, -max_c])
    dist.all_reduce(buf, op=dist.ReduceOp.MIN, group=group)
    return buf[:c].contiguous(), (-buf[c:]).contiguous()


def allreduce_sum(v: torch.Tensor, group=None):
    """SUM over ranks of a small vector (the nine first-calibrate lp sums and the element count)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return v
    v = v.clone()
    dist.all_reduce(v, op=dist.ReduceOp.SUM, group=group)
    return v


def install(group=None):
    """Make every QConv2d calibration call all-reduce its range statistics (MIN/MAX) and, in first-calibrate
    mode, the scores of the nine candidate init ranges (SUM) -- every rank then holds identical tables."""
    quant_util.calib_allreduce = lambda mn, mx: allreduce_minmax(mn, mx, group)
    quant_util.calib_allreduce_sum = lambda v: allreduce_sum(v, group)


def uninstall():
    quant_util.calib_allreduce = None
    quant_util.calib_allreduce_sum = None


def gather_images(x_local: torch.Tensor, group=None):
    """Optional: all ranks' final images on every rank (rank order)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return x_local
    out = [torch.empty_like(x_local) for _ in range(dist.get_world_size(group))]
    dist.all_gather(out, x_local.contiguous(), group=group)
    return torch.cat(out, dim=0)
