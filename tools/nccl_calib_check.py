"""2-GPU check of the path's only exchange, over NCCL:
   sharded calibration with the [min || -max] all-reduce == single-process calibration on the whole batch
   (bit-identical tables), and sharded sampling == the matching rows of whole-batch sampling.
   torchrun --nproc-per-node 2 tools/nccl_calib_check.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import attentiondm_b200 as A
from attentiondm_b200 import dist as adist
from oracle import restate as R, synth as S
from tests.util import build_cuda_model

rank, world = adist.init_from_env()
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
torch.cuda.set_device(dev)
spec = S.tiny_spec(T=4, bitwidth=8, ch=32, ch_mult=(1, 2), image_size=8)       # reaches 1x1: fused programs included
sd = S.synth_state_dict(spec, seed=7)
betas = R.beta_schedule_linear().to(dev)
xg = torch.randn(8, 3, 8, 8, generator=torch.Generator().manual_seed(99)).to(dev)
lo, hi = adist.shard_bounds(xg.shape[0], rank, world)

def calibrated(x, allreduce):
    m = build_cuda_model(spec, sd, device=dev)
    if allreduce:
        adist.install()
    else:
        adist.uninstall()
    m.set_calibrate(True)
    A.generalized_steps(x, spec.seq, m, betas, eta=0.0, keep="last")
    m.set_calibrate(False)
    m.reset_index_seq()
    adist.uninstall()
    return m

m_shard = calibrated(xg[lo:hi].contiguous(), True)       # this rank's shard, ranges all-reduced over NCCL
m_full = calibrated(xg, False)                           # every rank: the whole batch, no communication
tables_equal = all(torch.equal(a.groups_range.data, b.groups_range.data)
                   for (_, a), (_, b) in zip(m_shard.qconvs(), m_full.qconvs()))
xs_shard, _ = A.generalized_steps(xg[lo:hi].contiguous(), spec.seq, m_shard, betas, eta=0.0, keep="last")
xs_full, _ = A.generalized_steps(xg, spec.seq, m_full, betas, eta=0.0, keep="last")
rows_equal = torch.equal(xs_shard[-1], xs_full[-1][lo:hi])
gathered = adist.gather_images(xs_shard[-1].to(dev))
gather_equal = torch.equal(gathered.cpu(), xs_full[-1])
ok = torch.tensor([int(tables_equal), int(rows_equal), int(gather_equal)], device=dev)
dist.all_reduce(ok, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"world": world, "backend": dist.get_backend(), "calibration_tables_bit_identical": bool(ok[0]),
                      "sharded_sampling_equals_whole_batch_rows": bool(ok[1]), "all_gather_equals_whole_batch": bool(ok[2])}))
dist.barrier()
dist.destroy_process_group()
