"""Pipeline trace of the second-generation fused kernel (csrc/flow_tc2.cuh) on the trajectory_FC_large stack.

    python tools/tc2_trace.py [rows] [out_file]

Runs one inverse pass with BCNF_TC2_TRACE set: block 0's issuer, epilogue and producer write clock64 stamps per N chunk /
job, which this script turns into a per-layer table (cycles): MMA time of each chunk, its epilogue time, and how long the
issuer waited for the accumulator slot / the first stage.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
path = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/tc2_trace.txt"
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
for _ in range(2):
    flow.run(True, z, P, inst_period=n_inst, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); flow.run(True, z, P, inst_period=n_inst, out=out); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"inverse of {rows} rows: {ms:.3f} ms = {rows / ms / 1e3:.3f} M rows/s")
os.environ["BCNF_TC2_TRACE"] = path              # read once, when the handle is created
model._packed = None
flow = model._flow()
P = flow.project(h)
flow.run(True, z, P, inst_period=n_inst, out=out)
torch.cuda.synchronize()
del os.environ["BCNF_TC2_TRACE"]

ev = {0: {}, 1: {}, 2: {}, 3: {}}
for line in open(path):
    r, i, a, b, c, d = line.split()
    ev[int(r)][int(i)] = (int(a), int(b), int(c), int(d))
iss, epi, prod = ev[0], ev[1], ev[2]
t0 = min(v[0] for v in iss.values())
L = len(mk["nested_sizes"])
chunks_per_layer = [3] * L + [1]            # FC_large: 528 = 192 + 192 + 144; last Linear one chunk
print("chunk  layer.c | issuer: start  slot_wait  stage_wait  mma_issue(all K) | epilogue: wait_from  acc_ready  done  (cycles from kernel start; durations)")
n = 0
for net in range(3):
    for l, nc in enumerate(chunks_per_layer):
        for c in range(nc):
            if n in iss and n in epi:
                s0, s1, s2, s3 = iss[n]
                p0, p1, p2, _ = epi[n]
                print(f"{n:4d}  n{net} L{l}.{c} | {s0 - t0:9d} {s1 - s0:8d} {s2 - s1:8d} {s3 - s2:8d} | {p0 - t0:9d} {p1 - t0:9d} {p2 - t0:9d}  "
                      f"epi {p2 - p1:6d}  epi_idle_before {p1 - p0:6d}")
            n += 1
per_net = sum(chunks_per_layer)
if 2 * per_net in iss:
    print(f"cycles per network (issuer start of network 1 -> network 2): {iss[2 * per_net][0] - iss[per_net][0]}")
jobs = sorted(prod)
print("producer: job  first-A-available  last-A-available (cycles from start)")
for j in jobs[: 3 * (L + 1)]:
    print(f"   {j:3d} {prod[j][0] - t0:9d} {prod[j][1] - t0:9d}")

cp = ev[3]
print("coupling (cycles): tmem+ts | sync1 | affine | sync2 | ldsum | glue ops")
for net in range(3):
    if 2 * net in cp and 2 * net + 1 in cp and (per_net * net + per_net - 1) in epi:
        a0, a1, a2, a3 = cp[2 * net]
        g0, g1, _, _ = cp[2 * net + 1]
        acc = epi[per_net * net + per_net - 1][1]
        print(f"   n{net}: {a0 - acc:6d} | {a1 - a0:6d} | {a2 - a1:6d} | {a3 - a2:6d} | {g0 - a3:6d} | {g1 - g0:6d}")
