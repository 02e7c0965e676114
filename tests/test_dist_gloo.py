"""world_size-2 gloo test of the one exchange on the path: the calibration
range all-reduce.  Sharding the batch over two ranks and all-reducing
[min || -max] must reproduce single-process calibration on the whole batch
bit-for-bit.  The CPU oracle plays the local statistic."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import restate as R


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    from attentiondm_b200 import dist as adist
    from attentiondm_b200 import quant_util
    adist.init_from_env(backend="gloo")
    adist.install()
    g = torch.Generator().manual_seed(5)
    C, G, a_bit = 32, 8, 8
    x = torch.randn(6, C, 4, 4, generator=g) * 6 * (0.2 + torch.rand(1, C, 1, 1, generator=g))
    alpha = torch.randn(G, C, generator=g) * 0.3
    lo, hi = adist.shard_bounds(x.shape[0], rank, world)
    xs = x[lo:hi]

    def ext(xin):
        flat = xin.transpose(0, 1).reshape(C, -1)
        return quant_util.calib_allreduce(flat.min(1)[0], flat.max(1)[0])

    y, gr = R.calibrate_activation(xs, alpha, G, a_bit, -4.0, 6.0, ext_minmax=ext(xs))
    y_full, gr_full = R.calibrate_activation(x, alpha, G, a_bit, -4.0, 6.0)
    ok = torch.equal(gr, gr_full) and torch.equal(y, y_full[lo:hi])
    gathered = adist.gather_images(xs)
    ok = ok and torch.equal(gathered, x)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_calibration_allreduce_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


# ---------------------------------------------------------------------------
# first-calibrate under sharding: the nine candidate scores (utils/quant_util.py:237-254) are means over the
# GLOBAL batch, so the nine lp sums and the element count are all-reduced with SUM; both ranks must pick the
# same init range as single-process calibration on the whole batch, and end with the same group table.
# The host logic under test is attentiondm_b200.QModule._calibrate_step itself; the three device ops it calls
# are replaced by the CPU oracle (there is no GPU in this test).
# ---------------------------------------------------------------------------
def _oracle_ops(monkey):
    from attentiondm_b200 import ops

    def minmax_c(x):                                   # x NHWC
        flat = x.reshape(-1, x.shape[-1])
        return flat.min(0)[0], flat.max(0)[0]

    def group_ranges(min_c, max_c, G, init_min, init_max, out_gr_t):
        C = min_c.numel()
        mn = torch.where(min_c > init_min, torch.full((C,), init_min), min_c)
        mx = torch.where(max_c < init_max, torch.full((C,), init_max), max_c)
        xq_min, gmin = R.group_wise(mn.clone(), G, "min")
        xq_max, gmax = R.group_wise(mx.clone(), G, "max")
        out_gr_t.copy_(torch.stack([gmin, gmax], dim=1))
        return xq_min, xq_max

    def calib_mix(x, gr_t, sw, a_bit, lp_p=None):
        y = None
        for g in range(gr_t.shape[0]):
            b = R.scalar_branch_fake_quant(x, gr_t[g][0], gr_t[g][1], a_bit) * sw[g]
            y = b if y is None else y + b
        if lp_p is None:
            return y
        return y, (y - x).abs().double().pow(lp_p).sum().reshape(1)

    monkey(ops, "minmax_c", minmax_c)
    monkey(ops, "group_ranges", group_ranges)
    monkey(ops, "calib_mix", calib_mix)


def _first_calib_case():
    g = torch.Generator().manual_seed(11)
    C = 16
    # in-range activations whose best shrink factor is decided by a close race between candidates
    x = torch.randn(8, 4, 4, C, generator=g) * 0.9 + 0.3
    x[:4] *= 0.35                                       # the two shards see different spreads
    return x, C


def _run_first_calibrate(x, C):
    import argparse
    import attentiondm_b200 as A
    args = argparse.Namespace(bitwidth=8, timesteps=2)
    q = A.QConv2d(C, 8, 1, w_bit=8, a_bit=8, sequence=range(0, 1000, 500), args=args)
    q.set_calibrate(True)
    q.first_calibrate(True)
    q._calibrate_step(x)
    return float(q.init_range_min[0]), float(q.init_range_max[0]), q.groups_range.data[0].clone()


def _worker_first(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    from attentiondm_b200 import dist as adist
    adist.init_from_env(backend="gloo")
    _oracle_ops(setattr)
    x, C = _first_calib_case()
    want = _run_first_calibrate(x, C)                   # single process, whole batch (hooks not installed yet)
    lo, hi = adist.shard_bounds(x.shape[0], rank, world)
    local_only = _run_first_calibrate(x[lo:hi].contiguous(), C)
    adist.install()
    got = _run_first_calibrate(x[lo:hi].contiguous(), C)
    adist.uninstall()
    ok = got[0] == want[0] and got[1] == want[1] and torch.equal(got[2], want[2])
    q.put((rank, bool(ok), local_only[:2], got[:2]))
    dist.destroy_process_group()


def test_first_calibrate_allreduce_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_first, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert [r[:2] for r in res] == [(0, True), (1, True)], res
    assert res[0][3] == res[1][3]                       # both ranks chose the same init range
    # the case is only meaningful if the un-reduced, shard-local search would have disagreed between ranks
    assert res[0][2] != res[1][2], "test case does not discriminate: local searches agree"
