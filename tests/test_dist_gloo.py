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
