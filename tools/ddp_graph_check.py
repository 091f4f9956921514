"""Data-parallel training with the whole step (incl. the NCCL all-reduce) in one CUDA graph: replicas must stay
bit-identical and the loss must fall.  Run: torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/ddp_graph_check.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bcnf_b200

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(rank)          # different initial weights per rank: the trainer must broadcast rank 0's
n_blocks = int(os.environ.get("BCNF_CHECK_BLOCKS", "3"))
# BCNF_CHECK_TRANSFORMER=1: a Transformer encoder in front of the stack -- its forward / backward kernels, the side-stream
# parameter gradients and the all-reduce of the non-stack parameters are then part of the captured step
use_trf = bool(os.environ.get("BCNF_CHECK_TRANSFORMER"))
fnets = [bcnf_b200.ConcatenateCondition(None, 3), bcnf_b200.Transformer(input_size=3, trf_size=32, n_heads=4, ff_size=48, n_blocks=2,
                                                                       output_size=32, dropout=0.2, trf_dropout=0.1)] \
    if use_trf else [bcnf_b200.ConcatenateCondition(None, 32)]
model = bcnf_b200.CondRealNVP_v2(size=19, nested_sizes=[64, 64], n_blocks=n_blocks, n_conditions=32,
                                 feature_networks=fnets, dropout=0.1, act_norm=True).to(dev).train()
if os.environ.get("BCNF_CHECK_FLAT_ADAM"):     # the optimizer's gradient blob is the sink the buckets are cut from
    opt = bcnf_b200.FlatAdam(model, lr=1e-3)
else:
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, capturable=True, fused=True)
tr = bcnf_b200.Trainer(model, opt, cuda_graph=True, process_group=dist.group.WORLD)
g = torch.Generator().manual_seed(100 + rank)
y, c = torch.randn(128, 19, generator=g), (torch.randn(128, 30, 3, generator=g) if use_trf else torch.randn(128, 32, generator=g))
losses = [tr.train_batch(y, c)[0] for _ in range(30)]
flat = torch.cat([p.detach().flatten() for p in model.parameters()])
ref = flat.clone()
dist.broadcast(ref, src=0)
same = bool(torch.equal(flat, ref))
gathered = [None] * world
dist.all_gather_object(gathered, (rank, same, losses[0], losses[-1]))
if rank == 0:
    print(gathered)
    assert all(s for _, s, _, _ in gathered), "replicas diverged"
    assert all(l1 < l0 for _, _, l0, l1 in gathered), "loss did not fall"
    print("ddp_graph_check OK")
tr.close()
del tr
import gc
gc.collect()
torch.cuda.synchronize()
dist.barrier()
dist.destroy_process_group()
