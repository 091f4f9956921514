"""1 rank vs N ranks: identical per-instance samples and log-probs, with real ranks.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/shard_equiv_check.py

Every rank builds the same model (seed 0), takes its block of instances from bcnf_b200.sharding, samples with injected
z / evaluates log_prob, and all_gathers the blocks over NCCL; rank 0 compares the gathered result bit for bit with the
unsharded evaluation on its own GPU and prints one line.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from bcnf_b200 import sharding  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
cfg = bench.load_run_config("trajectory_FC_large")
model = bench.build_model(cfg, dev)
g = torch.Generator().manual_seed(5)
n_inst, m = 301, 7
cond = torch.randn(n_inst, 30, 3, generator=g)
z = torch.randn(m, n_inst, 19, generator=g)
y = torch.randn(n_inst, 19, generator=g)
with torch.no_grad():
    (mine,), lo, hi = sharding.shard_conditions([cond])
    block = model._sample(m, mine, outer=True, z=z[:, lo:hi].reshape(-1, 19))
    lp_block = model.log_prob(y[lo:hi], mine)
    got = sharding.gather_instance_blocks(block, n_inst, dim=1)
    lp_got = sharding.gather_instance_blocks(lp_block, n_inst, dim=0)
    # the public sharded entry points (device fast path, own z): shapes and finiteness
    s = sharding.sample_sharded(model.sample, 3, cond, gather=True)
    if rank == 0:
        full = model._sample(m, cond, outer=True, z=z.reshape(-1, 19))
        lp_full = model.log_prob(y, cond)
        ok = torch.equal(got, full) and torch.equal(lp_got, lp_full)
        print(f"shard equivalence over {world} rank(s): samples identical {torch.equal(got, full)}, "
              f"log_prob identical {torch.equal(lp_got, lp_full)}, sample_sharded(gather) {tuple(s.shape)} "
              f"finite {bool(torch.isfinite(s).all())} on {s.device}", flush=True)
        assert ok and tuple(s.shape) == (3, n_inst, 19)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
