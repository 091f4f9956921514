#!/usr/bin/env python
"""BASELINE.json config 5: coupling-stack sweep (stack alone, h injected).

    python tools/sweep.py [--out profiles/r02_sweep_1gpu.jsonl]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/sweep.py --out ...

Batch 2^10..2^24 rows PER GPU, hidden width 64..1024, 8..32 blocks, L = 5, D = 19, C = 128, ActNorm on, one-way
(SURVEY.md section 8d).  One JSON line per point: rows/s of forward (z + log-det) and inverse summed over the GPUs
(instances sharded, no collective: max over ranks of the CUDA-event time), kernel family, precision, algorithmic
TFLOP/s.  No point is skipped; the largest ones are timed over one repetition instead of three.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bcnf_b200  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r01_sweep.jsonl"))
    ap.add_argument("--budget-s", type=float, default=1e9, help="skip points estimated to take longer (default: never)")
    ap.add_argument("--widths", default="64,128,256,512,1024")
    ap.add_argument("--blocks", default="8,16,32")
    ap.add_argument("--batches", default="10,14,17,20,24")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    est_tflops = {"tcgen05": 150.0, "tiled": 8.0, "rowthread": 10.0}
    with open(args.out if rank == 0 else os.devnull, "w") as out, torch.no_grad():
        for H in map(int, args.widths.split(",")):
            for K in map(int, args.blocks.split(",")):
                torch.manual_seed(0)
                try:
                    model = bcnf_b200.CondRealNVP_v2(size=19, nested_sizes=[H] * 5, n_blocks=K, n_conditions=128,
                                                     feature_networks=[bcnf_b200.ConcatenateCondition(None, 128)],
                                                     dropout=0.3, act_norm=True, precision="auto").to(dev).eval()
                    flow = model._flow()
                except NotImplementedError as e:
                    out.write(json.dumps({"H": H, "K": K, "skipped": str(e)}) + "\n")
                    continue
                flops_row = 2.0 * int(flow.info.macs_per_row)
                n_inst = 1024
                P = flow.project(torch.randn(n_inst, 128, device=dev))
                for lg in map(int, args.batches.split(",")):
                    B = 1 << lg
                    est = B * flops_row / (est_tflops[flow.kernel] * 1e12)
                    rec = {"H": H, "K": K, "L": 5, "D": 19, "C": 128, "rows_per_gpu": B, "n_gpus": world, "kernel": flow.kernel,
                           "rows_per_cta": int(flow.info.rows_per_cta), "precision": flow.precision, "flops_per_row": flops_row}
                    if est > args.budget_s:
                        rec["skipped"] = f"estimated {est:.1f} s > budget"
                        out.write(json.dumps(rec) + "\n")
                        continue
                    y = torch.randn(B, 19, device=dev)
                    for name, inv in (("forward", False), ("inverse", True)):
                        flow.run(inv, y, P, inst_period=n_inst, want_logdet=not inv)
                        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        reps = 1 if est > 0.5 else 3
                        if world > 1:
                            dist.barrier()
                        torch.cuda.synchronize()
                        ev0.record()
                        for _ in range(reps):
                            flow.run(inv, y, P, inst_period=n_inst, want_logdet=not inv)
                        ev1.record()
                        torch.cuda.synchronize()
                        ms = ev0.elapsed_time(ev1) / reps
                        if world > 1:
                            t = torch.tensor([ms], device=dev)
                            dist.all_reduce(t, op=dist.ReduceOp.MAX)
                            ms = float(t.item())
                        rec[f"{name}_rows_per_s"] = world * B / (ms * 1e-3)
                        rec[f"{name}_tflops"] = world * B * flops_row / (ms * 1e-3) / 1e12
                    out.write(json.dumps(rec) + "\n")
                    out.flush()
                    if rank == 0:
                        print(json.dumps(rec), flush=True)
                    del y
                del model, flow, P
                torch.cuda.empty_cache()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
