"""The benchmarked configuration at the depth it is benchmarked.

``bench.py`` times the K = 26 block, H = 526 x 5, C = 1360 stack of ``trajectory_FC_large`` (reference
configs/runs/old/trajectory_FC_large.yaml:27-44) in the inverse direction (cnf.py:495-508, :572-582).  Rounding
compounds over 26 blocks and the inverse is the ill-conditioned direction, so the gates here are the ones of
``conftest.assert_parity`` evaluated on the whole stack: 1e-5 of max|ref32|, or, where the reference's own fp32
evaluation is further than that from its fp64 evaluation, at most 3x the reference's own fp32 error (SURVEY.md
section 7.2).  Single-pass bf16 has its stated tolerance (z, x 1e-2; log-det 3e-2).
"""
import json
import os

import numpy as np
import pytest
import torch

from bcnf_b200 import ActNorm, CondRealNVP_v2
from conftest import GOLDEN_DIR, assert_parity, rel_err
from oracle import flow_oracle as fo

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROWS = 384            # one and a half 256-row tiles of the fused kernel: a full tile and a ragged one
N_INST = 96


@pytest.fixture(autouse=True)
def _no_grad():
    with torch.no_grad():
        yield


def _fc_large(precision):
    cfg = json.load(open(os.path.join(GOLDEN_DIR, "state_dict_keys.json")))["trajectory_FC_large"]["config"]
    torch.manual_seed(0)
    model = CondRealNVP_v2.from_config(cfg, precision=precision)
    g = torch.Generator().manual_seed(5)
    for layer in model.layers:                      # ActNorm is the identity at init (cnf.py:345-346): perturb it
        if isinstance(layer, ActNorm):
            layer.scale.copy_(0.75 + 0.5 * torch.rand(layer.scale.shape, generator=g))
            layer.bias.copy_(0.1 * torch.randn(layer.bias.shape, generator=g))
    return model.to(DEV).eval()


@pytest.fixture(scope="module")
def case():
    """Inputs, features h of the model's own feature network, and the oracle's fp32 / fp64 results."""
    with torch.no_grad():
        model = _fc_large("bf16x3")
        assert model.n_blocks == 26 and list(model.nested_sizes) == [526] * 5 and model.n_conditions == 1360
        g = torch.Generator().manual_seed(17)
        cond = torch.randn(N_INST, 30, 3, generator=g)
        y = torch.randn(ROWS, 19, generator=g)
        z_in = torch.randn(ROWS, 19, generator=g)
        inst = torch.arange(ROWS) % N_INST
        h = model.features(cond).cpu()
        sd = {k: v.cpu().numpy() for k, v in model.state_dict().items() if k.startswith("layers.")}
        l32 = fo.layers_from_state_dict(sd)
        l64 = fo.layers_from_state_dict(sd, convert=lambda v: np.asarray(v, dtype=np.float64))
        hr = h[inst].numpy()
        z32, ld32 = fo.stack_forward(l32, y.numpy(), hr)
        x32 = fo.stack_inverse(l32, z_in.numpy(), hr)
        z64, ld64 = fo.stack_forward(l64, y.numpy().astype(np.float64), hr.astype(np.float64))
        x64 = fo.stack_inverse(l64, z_in.numpy().astype(np.float64), hr.astype(np.float64))
        return dict(sd=model.state_dict(), cond=cond, y=y, z_in=z_in, inst=inst, h=h, z32=z32, ld32=ld32, x32=x32,
                    z64=z64, ld64=ld64, x64=x64)


def _load(precision, case):
    model = _fc_large(precision)
    model.load_state_dict(case["sd"])
    return model


@pytest.mark.parametrize("precision", ["bf16x3", "fp32", "bf16"])
def test_full_depth_forward_inverse_log_prob(case, precision):
    model = _load(precision, case)
    flow = model._flow()
    assert flow.kernel == ("tiled" if precision == "fp32" else "tcgen05")
    P = flow.project(case["h"].to(DEV))
    z, ld = flow.run(False, case["y"], P, row2inst=case["inst"], want_logdet=True)
    x, _ = flow.run(True, case["z_in"], P, row2inst=case["inst"])
    z, ld, x = z.cpu().numpy(), ld.cpu().numpy(), x.cpu().numpy()
    errs = {"z": rel_err(z, case["z32"]), "logdet": rel_err(ld, case["ld32"]), "x": rel_err(x, case["x32"])}
    print(f"\nFC_large K=26 {precision}: {errs}; reference fp32 vs fp64: z {rel_err(case['z32'], case['z64']):.2e} "
          f"x {rel_err(case['x32'], case['x64']):.2e}")
    if precision == "bf16":
        assert errs["z"] < 1e-2 and errs["x"] < 1e-2 and errs["logdet"] < 3e-2, errs
    else:
        assert_parity(z, case["z32"], case["z64"], what="z")
        assert_parity(ld, case["ld32"], case["ld64"], what="logdet")
        assert_parity(x, case["x32"], case["x64"], what="x")
    # log_prob through the public API (new) on the first N_INST rows, one row per instance
    lp = model.log_prob(case["y"][:N_INST], case["cond"]).cpu().numpy()
    ref = -0.5 * (case["z64"][:N_INST] ** 2).sum(1) + case["ld64"][:N_INST] - 0.5 * 19 * np.log(2 * np.pi)
    # |d lp| <= sum |z| |dz| + |d logdet|: with z and log-det at 1e-5 of scale this is a few 1e-5 of max|lp|
    tol = 3e-2 if precision == "bf16" else 5e-5
    assert rel_err(lp, ref) < tol, rel_err(lp, ref)


def test_sample_outer_with_injected_z_is_the_stack_inverse(case):
    """_sample(outer=True) (cnf.py:572-582) with z injected: row (s, i) must be the oracle's inverse of z[s, i]
    under instance i's features -- the sample-major tiling of `c.repeat(m, 1)` (cnf.py:579)."""
    model = _load("bf16x3", case)
    n_samples = 4
    z = case["z_in"].view(n_samples, N_INST, 19)
    s = model._sample(n_samples, case["cond"], outer=True, z=z.reshape(-1, 19))
    assert s.shape == (n_samples, N_INST, 19)
    assert_parity(s.cpu().numpy().reshape(-1, 19), case["x32"], case["x64"], what="sample(outer=True)")


def test_round_trip_at_full_depth(case):
    model = _load("bf16x3", case)
    flow = model._flow()
    P = flow.project(case["h"].to(DEV))
    z, _ = flow.run(False, case["y"], P, row2inst=case["inst"])
    y2, _ = flow.run(True, z, P, row2inst=case["inst"])
    err = rel_err(y2.cpu().numpy(), case["y"].numpy())
    ref = rel_err(fo.stack_inverse(fo.layers_from_state_dict({k: v.cpu().numpy() for k, v in case["sd"].items()
                                                              if k.startswith("layers.")}),
                                   case["z32"], case["h"][case["inst"]].numpy()), case["y"].numpy())
    print(f"\nround trip K=26: mine {err:.2e}, reference fp32 {ref:.2e}")
    assert err < max(1e-5, 3 * ref)
