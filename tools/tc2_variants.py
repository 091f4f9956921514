"""Timing experiments on the fused kernel: BCNF_TC2_DEBUG variants (results of variants != 0 are wrong by design)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
variants = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1, 2, 4, 8, 3, 7, 15]
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
for v in variants:
    os.environ["BCNF_TC2_DEBUG"] = str(v)        # read once, when the handle is created
    model._packed = None
    flow = model._flow()
    P = flow.project(h)
    for _ in range(2):
        flow.run(True, z, P, inst_period=n_inst, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        flow.run(True, z, P, inst_period=n_inst, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"debug={v:2d}: {ms:8.3f} ms = {rows / ms / 1e3:.3f} M rows/s", flush=True)
