"""Kernel-level timeline of one log_prob evaluation (torch.profiler / CUPTI): encoder, projection, fused stack.
Usage (GPU box): python tools/lp_profile.py [--config trajectory_TRF_large] [--rows 16384]"""
import argparse
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import bcnf_b200

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="trajectory_TRF_large")
ap.add_argument("--rows", type=int, default=16384)
args = ap.parse_args()
dev = torch.device("cuda:0")
cfg = bench.load_run_config(args.config)
torch.manual_seed(0)
model = bcnf_b200.CondRealNVP_v2.from_config(cfg)
bench.perturb_actnorm(model)
model = model.to(dev).eval()
mk = cfg["model"]["kwargs"]
y = torch.randn(args.rows, mk["size"], device=dev)
c = torch.randn(args.rows, 30, 3, device=dev)
with torch.no_grad():
    for _ in range(3):
        model.log_prob(y, c)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        model.log_prob(y, c)
    e1.record(); torch.cuda.synchronize()
    print(f"log_prob of {args.rows} rows: {e0.elapsed_time(e1) / 5:.3f} ms ({args.config})")
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(2):
            model.log_prob(y, c)
        torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
tot = {}
for e in evs:
    k = e.name[:100]
    t = tot.setdefault(k, [0.0, 0])
    t[0] += e.device_time if hasattr(e, "device_time") else e.cuda_time
    t[1] += 1
t0 = min(e.time_range.start for e in evs); t1 = max(e.time_range.end for e in evs)
print(f"CUDA events: {len(evs)} over {(t1 - t0) / 1e3:.3f} ms wall (2 evaluations); sum of kernel time {sum(v[0] for v in tot.values()) / 1e3:.3f} ms")
for k, (us, n) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:25]:
    print(f"{us / 2e3:9.3f} ms/eval  n/eval={n / 2:7.1f}  avg {us / n:8.2f} us  {k}")
