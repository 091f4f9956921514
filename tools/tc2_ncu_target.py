"""One inverse pass of the trajectory_FC_large stack on the fused kernel, sized for an ncu capture.
    python tools/tc2_ncu_target.py [rows]      (default 75776 = 4 tiles of 256 rows per CTA pair)"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 75776
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
for _ in range(3):
    flow.run(True, z, P, inst_period=n_inst, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); flow.run(True, z, P, inst_period=n_inst, out=out); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"inverse of {rows} rows: {ms:.3f} ms = {rows / ms / 1e3:.3f} M rows/s")
