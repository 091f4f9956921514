"""Kernel-level timeline of the training step (torch.profiler / CUPTI) for the trajectory_TRF_large config.
Usage (GPU box): python tools/train_profile.py [--batch 256] [--no-graph]"""
import argparse
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import bcnf_b200

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--no-graph", action="store_true")
ap.add_argument("--config", default="trajectory_TRF_large")
ap.add_argument("--stack-only", action="store_true", help="coupling stack forward + backward only (no feature network, no optimizer)")
args = ap.parse_args()
dev = torch.device("cuda:0")
cfg = bench.load_run_config(args.config)
torch.manual_seed(0)
model = bcnf_b200.CondRealNVP_v2.from_config(cfg)
bench.perturb_actnorm(model)
model = model.to(dev).train()
use_graph = not args.no_graph
opt = torch.optim.Adam(model.parameters(), lr=2e-4, capturable=use_graph, fused=True)
trainer = bcnf_b200.Trainer(model, opt, cuda_graph=use_graph)
mk = cfg["model"]["kwargs"]
y = torch.randn(args.batch, mk["size"], device=dev)
c = torch.randn(args.batch, 30, 3, device=dev)
if args.stack_only:
    from bcnf_b200 import train as _tr
    h = torch.randn(args.batch, mk["n_conditions"], device=dev, requires_grad=True)

    class _T:
        def train_batch(self, y, c):
            z, ld = _tr.stack_forward_train(model, y, h, seed=1)
            (0.5 * (z ** 2).sum(1) - ld).mean().backward()
    trainer = _T()
for _ in range(4):
    trainer.train_batch(y, c)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    trainer.train_batch(y, c)
e1.record(); torch.cuda.synchronize()
print(f"step: {e0.elapsed_time(e1) / 5:.3f} ms (graph={use_graph}, batch={args.batch})")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        trainer.train_batch(y, c)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
tot = {}
for e in evs:
    k = e.name[:90]
    t = tot.setdefault(k, [0.0, 0])
    t[0] += e.device_time if hasattr(e, "device_time") else e.cuda_time
    t[1] += 1
t0 = min(e.time_range.start for e in evs); t1 = max(e.time_range.end for e in evs)
print(f"CUDA events: {len(evs)} over {(t1 - t0) / 1e3:.3f} ms wall (2 steps); sum of kernel time {sum(v[0] for v in tot.values()) / 1e3:.3f} ms")
for k, (us, n) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:30]:
    print(f"{us / 2e3:9.3f} ms/step  n/step={n / 2:7.1f}  avg {us / n:8.2f} us  {k}")
