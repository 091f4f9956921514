#!/usr/bin/env python
"""Benchmark of the CondRealNVP_v2 coupling-stack hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

Default workload = BASELINE.json configs[1]: trajectory_FC_large posterior sampling, 500 samples
per instance x 10 000 instances (5e6 rows), processed as K steps of `--instances-per-step`
instances each (default 1000 -> K=10 steps is the whole job).  One JSON line on stdout (rank 0).

  value        posterior samples/s (rows/s), whole job over all N GPUs, conditions resident in HBM;
               timed region per step = feature network + condition projection + z draw + fused
               inverse stack, CUDA events on the launching stream, max over ranks.
  e2e          the same metric through the public API  model.sample(500, cond_host, outer=True,
               output_device="cpu")  with host conditions (pinned) and the samples copied back.
  roofline     the fused stack kernel alone: algorithmic FLOPs (2 x MACs/row hoisted, SURVEY 8d,
               x rows per launch) / CUDA-event duration, vs the measured bf16 tensor peak.
  cpu_baseline oracle port (torch CPU back end = the ATen kernels the reference's eager path runs
               on a host) on a bounded sample of the same workload, all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

WORKLOADS = {
    # name: (config key in tests/golden/state_dict_keys.json, samples per instance, instances, kind)
    "fc_large_sample": ("trajectory_FC_large", 500, 10_000, "sample"),
    "fc_small_sample": ("trajectory_FC_small", 500, 10_000, "sample"),
    "fc_small_logprob": ("trajectory_FC_small", 1, 1 << 22, "log_prob"),
    "fc_large_logprob": ("trajectory_FC_large", 1, 1 << 17, "log_prob"),
    "lstm_large_logprob": ("trajectory_LSTM_large", 1, 1 << 17, "log_prob"),
    "trf_large_logprob": ("trajectory_TRF_large", 1, 1 << 17, "log_prob"),
    "lstm_large_sample": ("trajectory_LSTM_large", 500, 10_000, "sample"),
    # calibration ranks (compute_y_hat_ranks, M = 10 000 samples per instance) reduced inside the sampler:
    # the (M, N, D) samples are never written, the output is (N, D) counters
    "fc_large_ranks": ("trajectory_FC_large", 10_000, 1_000, "ranks"),
    # BASELINE config 4: training step (forward NLL + backward + Adam), batch 256 per GPU, DDP over N GPUs
    "trf_large_train": ("trajectory_TRF_large", 1, 256, "train"),
    "fc_small_train": ("trajectory_FC_small", 1, 256, "train"),
}


def load_run_config(key: str) -> dict:
    rec = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_keys.json")))[key]
    return rec["config"]


def measured_peaks() -> tuple[dict, str]:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def perturb_actnorm(model, seed: int = 1) -> None:
    """scale ~ U(0.75, 1.25), bias ~ N(0, 0.1): ActNorm is not the identity, the stack stays well conditioned."""
    from bcnf_b200 import ActNorm
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for layer in model.layers:
            if isinstance(layer, ActNorm):
                layer.scale.copy_((0.75 + 0.5 * torch.rand(layer.scale.shape, generator=g)).to(layer.scale.device))
                layer.bias.copy_((0.1 * torch.randn(layer.bias.shape, generator=g)).to(layer.bias.device))


# (kernel family, config) -> (DRAM bytes per row from one ncu --set full capture, the committed summary it comes from)
NCU_DRAM_BYTES_PER_ROW = {("tcgen05", "trajectory_FC_large"): ((3.611e9 + 8.644e9) / 75776, "profiles/r02_flow_tc2.txt"),
                          ("rowthread", "trajectory_FC_small"): (40.8e6 / 5.0e5, "profiles/r01_rowthread_lds.txt")}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], None, [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])); mx = float(r[1]); pw.append(float(r[2]))
            except ValueError:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        # median over the busiest half of the samples = "under load"
        sm_sorted = sorted(sm)
        return {"sm_mhz": float(np.median(sm_sorted[: max(1, len(sm_sorted) // 2)])) if sm else None,
                "sm_max_mhz": mx, "power_w_max": max(pw) if pw else None, "samples": len(sm),
                "reasons": sorted(reasons)}


def build_model(cfg: dict, device, precision: str = "auto"):
    from bcnf_b200 import CondRealNVP_v2
    torch.manual_seed(0)
    model = CondRealNVP_v2.from_config(cfg, precision=precision)
    perturb_actnorm(model)
    return model.to(device).eval()


def oracle_cpu_rate(cfg: dict, kind: str, samples: int, budget_s: float = 12.0) -> dict:
    """Time the oracle port (torch CPU back end) on a bounded sample of the workload."""
    from bcnf_b200 import CondRealNVP_v2
    from oracle import flow_oracle as fo
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    model = CondRealNVP_v2.from_config(cfg)          # CPU parameters; only used as a weight container
    perturb_actnorm(model)
    model.eval()
    layers = fo.layers_from_state_dict(model.state_dict(), convert=lambda v: v.detach().clone())
    mk = cfg["model"]["kwargs"]
    d, big = mk["size"], mk["nested_sizes"][0] > 64
    if kind == "sample":
        m, n_inst = (256, 8) if big else (256, 64)     # BASELINE.md section 2 protocol
    else:
        m, n_inst = 1, (2048 if big else 16384)
    rows = m * n_inst
    g = torch.Generator().manual_seed(3)
    cond = torch.randn(n_inst, 30, 3, generator=g)

    def once():
        with torch.no_grad():
            # the reference tiles the raw conditions and re-runs the feature network per row (cnf.py:579, :497)
            rep = cond.repeat(m, 1, 1)
            h = model.feature_network_stack(rep)
            z = torch.randn(rows, d, generator=g)
            if kind == "sample":
                return fo.stack_inverse(layers, z, h)
            return fo.stack_forward(layers, z, h)

    once()
    best, t_end, reps = float("inf"), time.perf_counter() + budget_s, 0
    while reps < 3 or (time.perf_counter() < t_end and reps < 50):
        t0 = time.perf_counter(); once(); best = min(best, time.perf_counter() - t0); reps += 1
    return {"value": rows / best, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{kind} {m} x {n_inst} instances = {rows} rows, best of {reps}, oracle/flow_oracle.py torch-CPU "
                      f"back end incl. per-row feature network as the reference does"}


def reference_cpu_rate(cfg: dict, kind: str, samples: int, budget_s: float = 12.0) -> dict:
    """Time the UNMODIFIED reference (oracle/_ref: byte-identical copy of /root/reference/src/bcnf made by
    oracle/build_ref.py, imported through oracle/ref_shim.py) on the host cores: its own CondRealNVP_v2.sample /
    forward in eval mode under no_grad, fp32, all threads, BASELINE.md section 2 protocol.  Falls back to the oracle port
    where the reference cannot run the workload (its LSTM encoder only accepts batch == 30) or is absent."""
    try:
        from oracle.ref_shim import import_reference, reference_available
        fn_types = [f["type"] for f in cfg["feature_networks"]]
        if not reference_available() or any("LSTM" in t for t in fn_types):
            raise RuntimeError("reference not usable for this workload")
        ref = import_reference()
    except Exception as e:  # noqa: BLE001
        out = oracle_cpu_rate(cfg, kind, samples, budget_s)
        out["sample"] += f" [port: {e}]"
        return out
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    model = ref.CondRealNVP_v2.from_config(cfg).eval()
    gen = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for layer in model.layers:
            if isinstance(layer, ref.ActNorm):
                layer.scale.copy_(0.75 + 0.5 * torch.rand(layer.scale.shape, generator=gen))
                layer.bias.copy_(0.1 * torch.randn(layer.bias.shape, generator=gen))
    mk = cfg["model"]["kwargs"]
    d, big = mk["size"], mk["nested_sizes"][0] > 64
    if kind == "sample":
        m, n_inst = (256, 8) if big else (256, 64)
    else:
        m, n_inst = 1, (2048 if big else 16384)
    rows = m * n_inst
    g = torch.Generator().manual_seed(3)
    cond = torch.randn(n_inst, 30, 3, generator=g)
    y = torch.randn(rows, d, generator=g)

    def once():
        with torch.no_grad():
            if kind == "sample":       # cnf.py:510-538: tiles the conditions, re-runs the encoder per row, draws z on the CPU
                return model.sample(m, cond, outer=True, batch_size=n_inst, sample_batch_size=m)
            return model.forward(y, cond, log_det_J=True)

    once()
    best, t_end, reps = float("inf"), time.perf_counter() + budget_s, 0
    while reps < 3 or (time.perf_counter() < t_end and reps < 50):
        t0 = time.perf_counter(); once(); best = min(best, time.perf_counter() - t0); reps += 1
    call = f"model.sample({m}, cond[{n_inst}], outer=True)" if kind == "sample" else f"model.forward(y[{rows}], cond, log_det_J=True)"
    return {"value": rows / best, "unit": "samples/s" if kind == "sample" else "evals/s", "cores": torch.get_num_threads(),
            "kind": "reference",
            "sample": f"{kind} {m} x {n_inst} instances = {rows} rows, best of {reps}: the unmodified reference "
                      f"(oracle/_ref, src/bcnf/models/cnf.py) {call}, eval mode, no_grad, fp32"}


def oracle_cpu_train_rate(cfg: dict, batch: int, budget_s: float = 15.0) -> dict:
    """One optimisation step (forward NLL + backward + Adam) of the oracle port on the host cores."""
    from bcnf_b200 import CondRealNVP_v2
    from oracle import flow_oracle as fo
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    model = CondRealNVP_v2.from_config(cfg)          # CPU parameter container + PyTorch feature network
    perturb_actnorm(model)
    model.train()
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.Adam(params, lr=2e-4)
    layers = fo.layers_from_state_dict(dict(model.named_parameters()), convert=lambda v: v)
    g = torch.Generator().manual_seed(3)
    y = torch.randn(batch, cfg["model"]["kwargs"]["size"], generator=g)
    cond = torch.randn(batch, 30, 3, generator=g)

    def once():
        opt.zero_grad()
        h = model.feature_network_stack(cond)
        z, ld = fo.stack_forward(layers, y, h)     # eval-mode conditioner (no dropout): a lower bound on the reference's cost
        fo.inn_nll(z, ld).backward()
        opt.step()

    once()
    best, t_end, reps = float("inf"), time.perf_counter() + budget_s, 0
    while reps < 2 or (time.perf_counter() < t_end and reps < 20):
        t0 = time.perf_counter(); once(); best = min(best, time.perf_counter() - t0); reps += 1
    return {"value": batch / best, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"train step batch {batch}, best of {reps}: oracle/flow_oracle.py torch-CPU back end with autograd + Adam "
                      f"(no dropout kernels, i.e. a lower bound on the reference's step time)"}


def main_train(args, cfg, cfg_key, batch, rank, world, local_rank) -> None:
    """Workload kind 'train': one optimisation step per bench step (Trainer.train_batch)."""
    import torch.distributed as dist
    import bcnf_b200
    mk = cfg["model"]["kwargs"]
    metric, unit = "training samples/sec", "samples/s"
    workload_name = f"{cfg_key} training step (forward NLL + backward + Adam), batch {batch} per GPU, dropout on"
    if args.impl == "reference":
        if rank != 0:
            return
        base = oracle_cpu_train_rate(cfg, batch, budget_s=20.0)
        print(json.dumps({"impl": "reference", "metric": metric, "value": base["value"], "unit": unit, "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * batch / base["value"],
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                          "data": "synthetic", "config": {"workload": workload_name}, "cpu_baseline": base,
                          "e2e": {"value": base["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}))
        return
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    line = measure_train(args, cfg, cfg_key, batch, rank, world, local_rank, args.steps, args.warmup, e2e=True)
    if rank == 0:
        line.update({"metric": metric, "unit": unit, "n_gpus": world, "higher_is_better": True, "scaling": "weak",
                     "vs_baseline": None, "dtype": "f32", "data": "synthetic"})
        line["config"]["workload"] = workload_name
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = oracle_cpu_train_rate(cfg, batch)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def measure_train(args, cfg, cfg_key, batch, rank, world, local_rank, steps, warmup, e2e=True) -> dict:
    """Time Trainer.train_batch (forward NLL + backward + Adam; with world > 1 the NCCL all-reduce of the gradients inside
    the step's CUDA graph) on an initialised process group; returns the fields of the bench line (rank 0: complete)."""
    import torch.distributed as dist
    import bcnf_b200
    mk = cfg["model"]["kwargs"]
    unit = "samples/s"
    device = torch.device("cuda", local_rank)
    torch.manual_seed(0)
    model = bcnf_b200.CondRealNVP_v2.from_config(cfg)
    perturb_actnorm(model)
    model = model.to(device).train()
    # default: the whole step (incl. the NCCL all-reduce of the flat gradient buffer) is one CUDA graph;
    # BCNF_TRAIN_DDP=1: eager steps under torch's DistributedDataParallel; BCNF_NO_TRAIN_GRAPH=1: eager steps
    use_ddp = world > 1 and bool(os.environ.get("BCNF_TRAIN_DDP"))
    if use_ddp:
        model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local_rank])
    use_graph = not use_ddp and not os.environ.get("BCNF_NO_TRAIN_GRAPH")
    flat_adam = os.environ.get("BCNF_BENCH_TORCH_ADAM", "0") != "1"
    opt = (bcnf_b200.FlatAdam(model, lr=2e-4) if flat_adam
           else torch.optim.Adam(model.parameters(), lr=2e-4, capturable=use_graph, fused=True))
    trainer = bcnf_b200.Trainer(model, opt, cuda_graph=use_graph,
                                process_group=dist.group.WORLD if world > 1 and not use_ddp else None)
    g = torch.Generator().manual_seed(100 + rank)
    y_host = torch.randn(batch, mk["size"], generator=g).pin_memory()
    c_host = torch.randn(batch, 30, 3, generator=g).pin_memory()
    y_dev, c_dev = y_host.to(device), c_host.to(device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier(); ev0.record()
        for _ in range(n):
            fn()
        ev1.record(); barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(warmup, 3)):
        trainer.train_batch(y_dev, c_dev)
    # device-timed value: steps enqueued back to back (train_batch_async: the losses stay on the device, one synchronisation
    # when the timed region closes); the e2e figure below goes through train_batch with host batches and reads the losses
    # back every step, as the reference's loop does (trainer.py:271-277)
    ms_total = timed(lambda: trainer.train_batch_async(y_dev, c_dev), steps)
    clocks = sampler.stop() if rank == 0 else {}
    e2e_steps = max(1, min(steps, 5))
    ms_e2e = timed(lambda: trainer.train_batch(y_host, c_host), e2e_steps) if e2e else None
    value = world * batch * steps / (ms_total * 1e-3)
    enc = [fn for fn in (model.module if hasattr(model, "module") else model).feature_network_stack.feature_networks
           if isinstance(fn, bcnf_b200.Transformer)]
    n_enc_launches = sum(2 + 11 * len(fn.layers) for fn in enc)
    n_lin = len(mk["nested_sizes"]) + 1
    n_coupling = mk["n_blocks"] * (2 if mk.get("two_way") else 1)
    from oracle.flow_oracle import macs_per_row
    flops_step = 3 * 2.0 * macs_per_row(mk["size"], mk["nested_sizes"], mk["n_blocks"], mk["n_conditions"], hoisted=False) * batch
    peaks, peak_src = measured_peaks()
    achieved = flops_step / (ms_total / steps * 1e-3) / 1e12
    line = {"value": value, "steps": steps, "warmup": max(warmup, 3), "ms_per_step": ms_total / steps,
            "config": {"size": mk["size"], "nested_sizes": mk["nested_sizes"],
                       "n_blocks": mk["n_blocks"], "n_conditions": mk["n_conditions"],
                       "kernel": "train_tc2_gemm (tcgen05, TMA-fed bf16 hi/lo operand images) + fused pre/post kernels; "
                                 "weight gradients train_tc3_dw (tcgen05, MN-major images)",
                       "precision": "bf16x3 (3-pass split, fp32 accumulate: fp32-class)", "optimizer": ("bcnf_b200.FlatAdam (one launch over the flat parameter / gradient / moment blob)" if flat_adam
                                     else "torch.optim.Adam(fused=True)"),
                       "cuda_graph": use_graph, "l2": "each step touches every parameter, gradient and Adam moment (4 x 195 MB)",
                       "parallelism": f"data parallel over {world} GPU(s), " + ("torch DistributedDataParallel (eager)" if use_ddp else
                                       "gradients written into one flat buffer and all-reduced (NCCL) in buckets underneath the backward, inside the step's CUDA graph")},
            # per conditioner network: pre, hidden GEMMs, post; post_bwd, data-gradient GEMMs, pre_bwd; weight-gradient
            # GEMMs, P and d h GEMMs, two column sums, the operand-image pack (bcnf_b200/train.py)
            # + the Transformer encoder on its own kernels (bcnf_b200/trf_train.py): weight-image pack, embedding, per block
            # four GEMMs + attention + GELU + two add-LayerNorm forward, attention backward + two LayerNorm parameter
            # gradients backward; + the optimizer launches of FlatAdam (one per gradient bucket + the remainder)
            "gpu_launches": steps * (n_coupling * (4 * (n_lin - 2) + 8) + 3 + n_enc_launches + (5 if flat_adam else 0)),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": float(peaks["bf16_tflops"]), "unit": "TFLOP/s",
                         "frac": achieved / float(peaks["bf16_tflops"]), "traffic": None,
                         "peak_source": f"{peak_src} bf16 dense", "kernel": "train_tc2_gemm_kernel / train_tc3_dw_kernel (all launches of the step)",
                         "note": "algorithmic FLOPs (3 x forward) of the step / step time; at batch 256 the step is a chain of ~310 "
                                 "dependent launches of ~10 us each for the stack plus ~150 for the Transformer encoder (own forward "
                                 "kernels, hand-written backward; parameter gradients and the optimizer off the chain) (SURVEY 8d: "
                                 "latency-, not throughput-bound); --instances-per-step 4096 / 32768 shows the tensor-core rate"},
            "clocks": clocks}
    if e2e:
        line["e2e"] = {"value": world * batch * e2e_steps / (ms_e2e * 1e-3), "unit": unit,
                       "h2d_bytes_per_step": (y_host.numel() + c_host.numel()) * 4, "d2h_bytes_per_step": 12,
                       "ms_per_step": ms_e2e / e2e_steps}
    # the step's CUDA graph holds kernels of the NCCL communicator: release it before the group goes away
    trainer.close()
    del trainer
    import gc
    gc.collect()
    torch.cuda.synchronize()
    return line


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="fc_large_sample", choices=sorted(WORKLOADS))
    ap.add_argument("--instances-per-step", type=int, default=0, help="per GPU; 0 = workload default")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", default="auto", choices=["auto", "fp32", "bf16x3", "bf16"],
                    help="conditioner GEMM arithmetic; auto = bf16x3 on tcgen05 for wide conditioners (fp32-class "
                         "accuracy, 1e-5 gate), fp32 FMA for narrow ones")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg_key, m_samples, n_inst_total, kind = WORKLOADS[args.workload]
    cfg = load_run_config(cfg_key)
    mk = cfg["model"]["kwargs"]
    if kind == "train":
        return main_train(args, cfg, cfg_key, args.instances_per_step or n_inst_total, rank, world, local_rank)
    unit = "samples/s" if kind in ("sample", "ranks") else "evals/s"
    metric = "posterior samples/sec" if kind in ("sample", "ranks") else "log_prob evals/sec"
    inst_step = args.instances_per_step or {"sample": 1000, "ranks": 50}.get(kind, n_inst_total // 8)
    what = {"sample": "posterior sampling", "ranks": "calibration ranks (samples reduced in the kernel, never written)"}.get(kind, "log_prob")
    workload_name = (f"{cfg_key} {what} "
                     f"{m_samples} x {n_inst_total} (step = {m_samples} x {inst_step} instances per GPU)")

    if args.impl == "reference":
        # reference arm: the reference's own CPU implementation of the path = the oracle port on
        # the host cores (the Python reference tree does not exist on the GPU box)
        if rank != 0:
            return
        base = reference_cpu_rate(cfg, "sample" if kind == "ranks" else kind, m_samples, budget_s=20.0)
        rows_step = int(base["sample"].split("=")[1].split()[0])
        line = {"impl": "reference", "metric": metric, "value": base["value"], "unit": unit, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * rows_step / base["value"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": {"workload": workload_name}, "cpu_baseline": base,
                "e2e": {"value": base["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the coupling stack has no CPU path")
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)

    from bcnf_b200 import _cabi
    _cabi.lib()                                  # fail loudly if the extension is missing
    model = build_model(cfg, device, args.precision)
    flow = model._flow()
    d = mk["size"]
    g = torch.Generator().manual_seed(100 + rank)
    # shard by conditioning instance, no collective on the data path.  One step of the whole job holds
    # world x inst_step instances; bcnf_b200.sharding hands this rank its contiguous block of them.
    from bcnf_b200 import sharding
    cond_global = torch.randn(world * inst_step, 30, 3, generator=torch.Generator().manual_seed(99))
    (cond_block,), lo, hi = sharding.shard_conditions([cond_global])
    assert hi - lo == inst_step
    cond_host = cond_block.contiguous().pin_memory()
    cond_dev = cond_host.to(device)
    rows_step = m_samples * inst_step
    launches = [0]

    seed = [1234 + rank]
    if kind == "sample":
        out_buf = torch.empty((rows_step, d), device=device)
    elif kind == "ranks":
        y_host = torch.randn(inst_step, d, generator=g).pin_memory()
        y_dev = y_host.to(device)
        ranks_dev = torch.zeros((inst_step, d), dtype=torch.int32, device=device)
    else:
        y_host = torch.randn(rows_step, d, generator=g).pin_memory()
        y_dev = y_host.to(device)

    from bcnf_b200 import feature_tc

    def step_resident():
        with torch.no_grad():
            # features + projection as the model does it: one fused GEMM chain from 2 048 instances up (log-prob workloads),
            # PyTorch feature network + bcnf_cond_project below (the 500-instance sampling steps)
            n0 = feature_tc.N_LAUNCH[0]
            P = model._projection(cond_dev)
            seed[0] += 1
            if kind == "sample":          # the latent is drawn inside the kernel (bcnf_flow_sample), as model.sample does
                out = flow.sample(rows_step, P, seed=seed[0], inst_period=inst_step, out=out_buf)
            elif kind == "ranks":
                out = flow.sample_ranks(rows_step, P, y_dev, ranks_dev, seed=seed[0], inst_period=inst_step)
            else:
                out, _ = flow.run(False, y_dev, P, want_logdet=True)
            # own kernels per step: the fused stack + either the feature-network / projection GEMM chain (counted by
            # feature_tc where it runs) or bcnf_cond_project behind a PyTorch feature network (tensor-core handles:
            # img_pack of h + the CTA-pair GEMM; fp32 handles: one SGEMM)
            own = feature_tc.N_LAUNCH[0] - n0
            launches[0] += 1 + (own if own else (2 if flow.kernel == "tcgen05" else 1))
            return out

    def step_e2e():
        if kind == "sample":
            return model.sample(m_samples, cond_host, outer=True, output_device="cpu")
        if kind == "ranks":
            from bcnf_b200 import compute_y_hat_ranks
            return compute_y_hat_ranks(model, y_host, cond_host, M_samples=m_samples, device=device, output_device="cpu",
                                       verbose=False)
        lp = model.log_prob(y_host.to(device, non_blocking=True), cond_host.to(device, non_blocking=True))
        return lp.to("cpu")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for _ in range(steps):
            fn()
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step_resident()
    launches[0] = 0
    ms_total = timed(step_resident, args.steps)
    n_launch = launches[0]
    clocks = sampler.stop() if rank == 0 else {}
    value = world * rows_step * args.steps / (ms_total * 1e-3)

    # the fused stack kernel alone (roofline numerator/denominator)
    with torch.no_grad():
        P = model._projection(cond_dev)
        zbuf = torch.randn((rows_step, d), device=device)
        k_steps = max(3, min(args.steps, 10))

        def kernel_only():
            if kind == "sample":
                flow.sample(rows_step, P, seed=7, inst_period=inst_step, out=zbuf)
            elif kind == "ranks":
                flow.sample_ranks(rows_step, P, y_dev, ranks_dev, seed=7, inst_period=inst_step)
            else:
                flow.run(False, zbuf, P, want_logdet=True)
        kernel_only()
        ms_kernel = timed(kernel_only, k_steps) / k_steps
    peaks, peak_src = measured_peaks()
    flops_launch = 2.0 * int(flow.info.macs_per_row) * rows_step
    achieved_tf = flops_launch / (ms_kernel * 1e-3) / 1e12
    peak_tf = float(peaks.get("bf16_tflops_sustained" if ms_kernel > 500 else "bf16_tflops"))
    fma_peak_tf = 148 * 128 * 2 * (clocks.get("sm_max_mhz") or 1965.0) * 1e6 / 1e12
    # DRAM bytes per row of the flow kernel from the committed ncu --set full captures (profiles/), scaled to this
    # launch.  flow_tc2 (r02): dirty lines of the L2-resident activation scratch that are written back before they
    # are overwritten, plus the weight images and the projection slices (profiles/r02_flow_tc2.txt)
    ncu = NCU_DRAM_BYTES_PER_ROW.get((flow.kernel, cfg_key))
    traffic = ncu[0] * rows_step if ncu else None
    roofline = {"bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved_tf / peak_tf, "traffic": traffic,
                "traffic_source": f"ncu dram__bytes_read+write per row ({ncu[1]}) x rows" if traffic else None,
                "peak_source": f"{peak_src} bf16 dense",
                "kernel": f"flow_{flow.kernel}", "kernel_ms": ms_kernel,
                "fp32_fma_peak_tflops": fma_peak_tf, "fp32_fma_frac": achieved_tf / fma_peak_tf,
                "algorithmic_bytes_per_row": {"sample": 4 * d, "ranks": 0}.get(kind, 4 * d * 2 + 4), "flops_per_row": 2 * int(flow.info.macs_per_row)}

    # end to end through the public API with host buffers
    e2e_steps = max(1, min(args.steps, 5))
    step_e2e()
    ms_e2e = timed(step_e2e, e2e_steps)
    e2e_value = world * rows_step * e2e_steps / (ms_e2e * 1e-3)
    h2d = cond_host.numel() * 4 + {"sample": 0, "ranks": inst_step * d * 4}.get(kind, rows_step * d * 4)
    d2h = {"sample": rows_step * d * 4, "ranks": inst_step * d * 4}.get(kind, rows_step * 4)

    # auxiliary measurements carried by the default line (not the headline; each names its own workload):
    #   train_*     BASELINE config 4, trajectory_TRF_large training step, batch 256 per GPU -- the one multi-GPU
    #               workload that communicates (NCCL all-reduce of 195 MB of gradients per step), so that the driver's
    #               1/2/4/8-GPU runs record a curve with a collective on it;
    #   fc_small_*  BASELINE config 1, trajectory_FC_small sampling on the fp32 row-per-thread kernel.
    aux = None
    if args.workload == "fc_large_sample" and not os.environ.get("BCNF_BENCH_NO_AUX"):
        aux = {}
        try:
            del zbuf, P
            torch.cuda.empty_cache()
            tcfg = load_run_config("trajectory_TRF_large")
            tl = measure_train(args, tcfg, "trajectory_TRF_large", 256, rank, world, local_rank, steps=20, warmup=5, e2e=False)
            aux.update({"train_workload": "trajectory_TRF_large training step (forward NLL + backward + Adam), batch 256 per GPU, "
                                          "dropout on; gradients all-reduced (NCCL) in buckets underneath the backward, inside the step's CUDA graph",
                        "train_ms_per_step": tl["ms_per_step"], "train_samples_per_s": tl["value"], "train_n_gpus": world})
        except Exception as e:  # noqa: BLE001   (the headline must not depend on the auxiliary run)
            aux["train_error"] = repr(e)[:300]
        try:
            scfg = load_run_config("trajectory_FC_small")
            small = build_model(scfg, device, "auto")
            sflow = small._flow()
            sc = torch.randn(inst_step, 30, 3, device=device)
            with torch.no_grad():
                sP = sflow.project(small.features(sc))
                sz = torch.randn((rows_step, scfg["model"]["kwargs"]["size"]), device=device)
                def small_step():
                    sflow.run(True, sz, sP, inst_period=inst_step)
                for _ in range(3):
                    small_step()
                ms_small = timed(small_step, 10) / 10
            aux.update({"fc_small_workload": f"trajectory_FC_small posterior sampling, {m_samples} x {inst_step} rows per launch, "
                                             f"{sflow.kernel} kernel (fp32 FMA)",
                        "fc_small_samples_per_s": world * rows_step / (ms_small * 1e-3),
                        "fc_small_tflops": 2.0 * int(sflow.info.macs_per_row) * rows_step / (ms_small * 1e-3) / 1e12,
                        "fc_small_fp32_fma_frac": 2.0 * int(sflow.info.macs_per_row) * rows_step / (ms_small * 1e-3) / 1e12 / fma_peak_tf})
        except Exception as e:  # noqa: BLE001
            aux["fc_small_error"] = repr(e)[:300]

    if rank == 0:
        line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None,
                "dtype": {"fp32": "f32", "bf16x3": "bf16x3 split, f32 accumulate (fp32-class, 1e-5 gate)",
                          "bf16": "bf16, f32 accumulate"}[flow.precision],
                "data": "synthetic",
                "config": {"workload": workload_name, "size": d, "nested_sizes": mk["nested_sizes"],
                           "n_blocks": mk["n_blocks"], "n_conditions": mk["n_conditions"],
                           "kernel": flow.kernel, "precision": flow.precision,
                           "l2": "inputs/outputs per step exceed L2 or are regenerated each step; weights "
                                 f"({int(flow.info.packed_bytes) >> 20} MiB packed) stream from L2/HBM",
                           "parallelism": f"instances sharded over {world} GPU(s), no collective"},
                "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / e2e_steps},
                "gpu_launches": n_launch, "roofline": roofline, "clocks": clocks}
        if aux is not None:
            line["aux"] = aux
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = reference_cpu_rate(cfg, "sample" if kind == "ranks" else kind, m_samples)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
