"""Second-generation fused kernel (csrc/flow_tc2.cuh) against the first-generation one and the oracle.

    python tools/tc2_check.py            # every shape, each in its own process (a trap cannot take the rest down)
    python tools/tc2_check.py one IDX    # one shape in this process

Prints one line per shape: max |new - oracle| / max |oracle| for z, log-det, x and the same for the old kernel.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SHAPES = [
    # (size, nested, blocks, C, two_way, rows, precision)
    (19, [64], 1, 8, False, 128, "bf16x3"),
    (19, [64, 64], 1, 8, False, 77, "bf16x3"),
    (19, [128, 128, 128], 2, 16, False, 300, "bf16x3"),
    (19, [256] * 5, 2, 128, False, 257, "bf16x3"),
    (19, [512] * 5, 2, 128, False, 700, "bf16x3"),
    (19, [526] * 5, 3, 1360, False, 300, "bf16x3"),
    (21, [175, 175, 175], 3, 107, True, 129, "bf16x3"),
    (19, [206, 206, 206], 3, 40, False, 513, "bf16x3"),
    (19, [1024] * 3, 2, 64, False, 260, "bf16x3"),
    (19, [526] * 5, 3, 1360, False, 300, "bf16"),
    (19, [526] * 5, 26, 1360, False, 40000, "bf16x3"),
]


def run_one(idx: int) -> None:
    import numpy as np
    import torch

    import bcnf_b200
    from bcnf_b200 import CondRealNVP_v2
    from oracle import flow_oracle as fo

    size, nested, blocks, n_cond, two_way, rows, precision = SHAPES[idx]
    torch.manual_seed(0)

    def make():
        torch.manual_seed(0)
        m = CondRealNVP_v2(size=size, nested_sizes=nested, n_blocks=blocks, n_conditions=n_cond,
                           feature_networks=[bcnf_b200.ConcatenateCondition(None, n_cond)], dropout=0.3,
                           act_norm=True, two_way=two_way, precision=precision)
        g = torch.Generator().manual_seed(1)
        with torch.no_grad():
            for layer in m.layers:
                if isinstance(layer, bcnf_b200.ActNorm):
                    layer.scale.copy_(0.75 + 0.5 * torch.rand(layer.scale.shape, generator=g))
                    layer.bias.copy_(0.1 * torch.randn(layer.bias.shape, generator=g))
        return m.to("cuda:0").eval()

    g = torch.Generator().manual_seed(11)
    y = torch.randn(rows, size, generator=g)
    n_inst = min(rows, 1000)
    h = torch.randn(n_inst, n_cond, generator=g)
    z_in = torch.randn(rows, size, generator=g)
    idx_map = torch.arange(rows) % n_inst
    hh = h[idx_map]

    def run(model):
        with torch.no_grad():
            z = model(y, hh, log_det_J=True)
            ld = model.log_det_J
            x = model.inverse(z_in, hh)
        torch.cuda.synchronize()
        return z.cpu().numpy(), ld.cpu().numpy(), x.cpu().numpy()

    os.environ.pop("BCNF_FLOW_TC", None)
    new = make()
    try:
        zn, ldn, xn = run(new)
    except Exception as e:  # noqa: BLE001
        words = (C.c_uint32 * 4)()
        try:
            new._flow().lib.bcnf_flow_debug_words(new._flow()._handle, words)
        except Exception:  # noqa: BLE001
            pass
        print(f"[{idx}] {SHAPES[idx]} NEW KERNEL FAILED: {e}; watchdog = {[hex(w) for w in words]}", flush=True)
        raise
    os.environ["BCNF_FLOW_TC"] = "1"
    old = make()
    zo, ldo, xo = run(old)
    os.environ.pop("BCNF_FLOW_TC", None)

    def rel(a, b):
        return float(np.abs(a - b).max() / np.abs(b).max())

    if rows <= 1024:
        sd = {k: v.cpu().numpy() for k, v in new.state_dict().items()}
        l32 = fo.layers_from_state_dict(sd)
        z32, ld32 = fo.stack_forward(l32, y.numpy(), hh.numpy())
        x32 = fo.stack_inverse(l32, z_in.numpy(), hh.numpy())
        print(f"[{idx}] {SHAPES[idx]} new vs oracle z {rel(zn, z32):.2e} ld {rel(ldn, ld32):.2e} x {rel(xn, x32):.2e} | "
              f"old vs oracle z {rel(zo, z32):.2e} ld {rel(ldo, ld32):.2e} x {rel(xo, x32):.2e}", flush=True)
    else:
        print(f"[{idx}] {SHAPES[idx]} new vs old z {rel(zn, zo):.2e} ld {rel(ldn, ldo):.2e} x {rel(xn, xo):.2e} "
              f"finite {bool(np.isfinite(zn).all() and np.isfinite(xn).all())}", flush=True)


def main() -> None:
    if len(sys.argv) >= 3 and sys.argv[1] == "one":
        run_one(int(sys.argv[2]))
        return
    rc = 0
    for i in range(len(SHAPES)):
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "one", str(i)], timeout=240,
                               capture_output=True, text=True)
            out = (p.stdout + p.stderr).strip().splitlines()
            keep = [ln for ln in out if ln.startswith("[")] or out[-5:]
            print("\n".join(keep), flush=True)
            if p.returncode != 0:
                rc = 1
                print("\n".join(out[-8:]), flush=True)
                # a trapped kernel leaves this process only; the next shape starts a fresh context
        except subprocess.TimeoutExpired:
            print(f"[{i}] {SHAPES[i]} TIMEOUT", flush=True)
            rc = 1
            break
    sys.exit(rc)


if __name__ == "__main__":
    main()
