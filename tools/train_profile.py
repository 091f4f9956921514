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

# per-stream view of ONE step: which stream carries the dependency chain, how much of it is kernel time and how much gaps
half = (t0 + t1) / 2
step_evs = [e for e in evs if e.time_range.start >= half]
if step_evs:
    s0 = min(e.time_range.start for e in step_evs); s1 = max(e.time_range.end for e in step_evs)
    print(f"second step alone: {(s1 - s0) / 1e3:.3f} ms wall, {len(step_evs)} kernels")
    by_stream = {}
    for e in step_evs:
        sid = getattr(e, "stream", None) if hasattr(e, "stream") else None
        if sid is None:
            sid = getattr(e, "device_resource_id", -1)
        by_stream.setdefault(sid, []).append(e)
    for sid, es in sorted(by_stream.items(), key=lambda kv: -sum(x.time_range.end - x.time_range.start for x in kv[1])):
        busy = sum(x.time_range.end - x.time_range.start for x in es)
        print(f"  stream {sid}: {len(es):5d} kernels, busy {busy / 1e3:7.3f} ms, first {(min(x.time_range.start for x in es) - s0) / 1e3:7.3f} ms, last end {(max(x.time_range.end for x in es) - s0) / 1e3:7.3f} ms")
    main_sid = max(by_stream, key=lambda k: len(by_stream[k]))
    es = sorted(by_stream[main_sid], key=lambda x: x.time_range.start)
    names = {}
    for x in es:
        k = x.name[:70]
        t = names.setdefault(k, [0.0, 0])
        t[0] += x.time_range.end - x.time_range.start; t[1] += 1
    gaps = sum(max(0.0, b.time_range.start - a.time_range.end) for a, b in zip(es, es[1:]))
    print(f"  chain stream {main_sid}: gaps between consecutive kernels {gaps / 1e3:.3f} ms")
    for k, (us, n) in sorted(names.items(), key=lambda kv: -kv[1][0])[:22]:
        print(f"    {us / 1e3:8.3f} ms  n={n:4d}  avg {us / n:7.2f} us  {k}")
    # phases of the chain by the first / last kernel of the coupling stack
    stack = [x for x in es if "bcnf::train" in x.name]
    if stack:
        a0, a1 = min(x.time_range.start for x in stack), max(x.time_range.end for x in stack)
        print(f"  chain phases: before the stack's first kernel {(a0 - s0) / 1e3:.3f} ms (encoder forward), "
              f"stack forward + loss + stack backward {(a1 - a0) / 1e3:.3f} ms, after its last kernel {(s1 - a1) / 1e3:.3f} ms "
              f"(encoder backward + optimizer)")

    # what fills the time before the stack's first kernel (encoder forward, step preamble): per kernel name, its run time
    # and the idle time in front of it on that stream
    if stack:
        pre = sorted([x for x in step_evs if x.time_range.end <= a0], key=lambda x: x.time_range.start)
        agg = {}
        prev_end = s0
        for x in pre:
            k = x.name[:70]
            t = agg.setdefault(k, [0.0, 0.0, 0])
            t[0] += x.time_range.end - x.time_range.start
            t[1] += max(0.0, x.time_range.start - prev_end)
            t[2] += 1
            prev_end = max(prev_end, x.time_range.end)
        print(f"  before the stack: {len(pre)} kernels")
        for k, (us, gap, n) in sorted(agg.items(), key=lambda kv: -(kv[1][0] + kv[1][1]))[:16]:
            print(f"    run {us / 1e3:7.3f} ms + idle before {gap / 1e3:7.3f} ms  n={n:4d}  {k}")
