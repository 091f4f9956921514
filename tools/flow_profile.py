"""Kernel-level timeline (torch.profiler / CUPTI) of one inverse pass of the trajectory_FC_large stack.
Usage (GPU box): [BCNF_FLOW_TC=1] python tools/flow_profile.py [rows]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 18944 * 2
dev = torch.device("cuda:0")
cfg = bench.load_run_config("trajectory_FC_large")
model = bench.build_model(cfg, dev)
mk = cfg["model"]["kwargs"]
n_inst = 1000
h = torch.randn(n_inst, mk["n_conditions"], device=dev)
flow = model._flow()
P = flow.project(h)
z = torch.randn(rows, mk["size"], device=dev)
out = torch.empty_like(z)
call = lambda: flow.run(True, z, P, inst_period=n_inst, out=out)
for _ in range(3):
    call()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); call(); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"inverse of {rows} rows: {ms:.3f} ms = {rows / ms / 1e3:.3f} M rows/s")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    call()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
tot = {}
for e in evs:
    t = tot.setdefault(e.name[:80], [0.0, 0])
    t[0] += e.device_time
    t[1] += 1
t0 = min(e.time_range.start for e in evs); t1 = max(e.time_range.end for e in evs)
print(f"{len(evs)} kernels over {(t1 - t0) / 1e3:.3f} ms; sum of kernel time {sum(v[0] for v in tot.values()) / 1e3:.3f} ms")
for k, (us, n) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:8]:
    print(f"{us / 1e3:9.3f} ms  n={n:5d}  avg {us / n:8.2f} us  {k}")
# per-launch durations of the first conditioner network, in order
names = [(e.time_range.start, e.name[:40], e.device_time) for e in evs]
names.sort()
print("first launches:", [(n.split("(")[0][-28:], round(d, 1)) for _, n, d in names[:9]])
