"""Parity of the CUDA path (through the Python class -> C ABI) against the golden vectors of the
live reference and against the CPU oracle.  Run on the B200 box:  pytest -m gpu

Tolerance (north_star): 1e-5 of max|ref| in fp32 for z, x and log-det; where the reference's own
fp32 result is further than that from its fp64 evaluation (ill-conditioned inverses), the error
budget of conftest.assert_parity applies.
"""
import json
import os

import numpy as np
import pytest
import torch

import bcnf_b200
from bcnf_b200 import CondRealNVP_v2
from conftest import GOLDEN_DIR, assert_parity, load_golden, rel_err
from oracle import flow_oracle as fo

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _no_grad():
    # inference tests: the feature networks are plain PyTorch modules whose outputs would
    # otherwise carry autograd history
    with torch.no_grad():
        yield


def _model_from_golden(meta, sd, **extra):
    model = CondRealNVP_v2.from_config(meta["config"], **extra)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    return model.to(DEV).eval()


def _t(a):
    return torch.from_numpy(np.asarray(a))


@pytest.fixture(params=["auto", "tiled"])
def kernel_choice(request, monkeypatch):
    if request.param == "tiled":
        monkeypatch.setenv("BCNF_FORCE_KERNEL", "tiled")
    else:
        monkeypatch.delenv("BCNF_FORCE_KERNEL", raising=False)
    return request.param


def test_forward_matches_reference_golden(golden, kernel_choice):
    name, data, sd, meta = golden
    model = _model_from_golden(meta, sd)
    z, h = model(_t(data["y"]), _t(data["cond"]), log_det_J=True, return_features=True)
    if kernel_choice == "tiled":
        assert model._flow().kernel == "tiled"
    assert model.log_det_J.shape == (data["y"].shape[0],) and model.log_det_J.dtype == torch.float32
    assert_parity(h.cpu().numpy(), data["h"], what=name + " h")
    assert_parity(z.cpu().numpy(), data["z"], data["z64"], what=name + " z")
    assert_parity(model.log_det_J.cpu().numpy(), data["logdet"], data["logdet64"], what=name + " logdet")


def test_inverse_matches_reference_golden(golden, kernel_choice):
    name, data, sd, meta = golden
    model = _model_from_golden(meta, sd)
    x = model.inverse(_t(data["z_in"]), _t(data["cond"]))
    assert_parity(x.cpu().numpy(), data["x"], data["x64"], what=name + " x")
    rt = model.inverse(_t(data["z"]), _t(data["cond"]))
    assert_parity(rt.cpu().numpy(), data["roundtrip"], data["roundtrip64"], what=name + " roundtrip")


def test_single_coupling_layer_like_reference_unit_test(golden):
    # tests/test_cnf.py:18-32 of the reference exercises one ConditionalAffineCouplingLayer
    name, data, sd, meta = golden
    model = _model_from_golden(meta, sd)
    layer = model.layers[int(data["layer_index"])]
    h = _t(data["h"]).to(DEV)
    z = layer.forward(_t(data["y"]).to(DEV), h, log_det_J=True)
    assert_parity(z.cpu().numpy(), data["layer_z"], what=name + " layer z")
    assert_parity(layer.log_det_J.cpu().numpy(), data["layer_logdet"], what=name + " layer logdet")
    x = layer.inverse(_t(data["z_in"]).to(DEV), h)
    assert_parity(x.cpu().numpy(), data["layer_x"], what=name + " layer x")
    if not meta["config"]["model"]["kwargs"].get("two_way", False):
        back = layer.inverse(z, h)
        assert torch.allclose(back.cpu(), _t(data["y"]), atol=1e-5)       # the reference's own assertion


def test_actnorm_and_orthonormal_layers_alone():
    data, sd, meta = load_golden("d21_two_way")
    model = _model_from_golden(meta, sd)
    y = _t(data["y"]).to(DEV)
    an = next(l for l in model.layers if isinstance(l, bcnf_b200.ActNorm))
    z = an(y, True)
    ref = an.scale.detach() * y + an.bias.detach()
    assert torch.allclose(z, ref, rtol=1e-6, atol=1e-6)
    assert abs(float(an.log_det_J) - float(torch.log(an.scale.detach().abs()).sum())) < 1e-5
    assert torch.allclose(an.inverse(z), y, rtol=1e-5, atol=1e-6)
    ot = next(l for l in model.layers if isinstance(l, bcnf_b200.OrthonormalTransformation))
    q = ot.orthonormal_matrix.detach()
    assert torch.allclose(ot(y, None), y @ q, rtol=1e-5, atol=1e-6)
    assert torch.allclose(ot.inverse(y, None), y @ q.T, rtol=1e-5, atol=1e-6)


def test_sample_reference_rng_reproduces_reference(golden):
    name, data, sd, meta = golden
    model = _model_from_golden(meta, sd, sample_rng="reference")
    cond = _t(data["cond"])
    torch.manual_seed(1234)
    s = model.sample(5, cond[:7], sigma=0.8, outer=True, batch_size=4, sample_batch_size=2)
    assert s.shape == data["sample_outer"].shape and s.device.type == "cpu"
    assert_parity(s.numpy(), data["sample_outer"], tol=2e-5, what=name + " sample outer")
    torch.manual_seed(1235)
    s = model._sample(7, cond[:7], sigma=1.0, outer=False)
    assert_parity(s.cpu().numpy(), data["sample_inner"], tol=2e-5, what=name + " sample inner")
    if data["sample_1d"].size:
        torch.manual_seed(1236)
        s = model._sample(6, cond[0], sigma=1.0)
        assert_parity(s.cpu().numpy(), data["sample_1d"], tol=2e-5, what=name + " sample 1d")


def test_sample_device_rng_shapes_and_statistics():
    data, sd, meta = load_golden("fc_small")
    model = _model_from_golden(meta, sd)
    cond = _t(data["cond"])[:6]
    s = model.sample(2000, cond, outer=True, output_device="cpu")
    assert s.shape == (2000, 6, 19) and s.device.type == "cpu" and torch.isfinite(s).all()
    # same distribution as the reference-order path: compare per-instance means / stds
    model_r = _model_from_golden(meta, sd, sample_rng="reference")
    torch.manual_seed(0)
    r = model_r.sample(2000, cond, outer=True, batch_size=6, sample_batch_size=2000)
    se = r.std(0) / np.sqrt(2000)
    assert ((s.mean(0) - r.mean(0)).abs() < 6 * se + 1e-3).all()
    assert ((s.std(0) / r.std(0) - 1).abs() < 0.15).all()
    # a fixed generator makes the device path reproducible
    g1 = torch.Generator(device=DEV).manual_seed(5)
    g2 = torch.Generator(device=DEV).manual_seed(5)
    a = model._sample_device(7, cond, sigma=1.0, output_device=DEV, generator=g1)
    b = model._sample_device(7, cond, sigma=1.0, output_device=DEV, generator=g2)
    assert torch.equal(a, b)


def test_log_prob_matches_golden_nll_rows(golden):
    name, data, sd, meta = golden
    model = _model_from_golden(meta, sd)
    lp = model.log_prob(_t(data["y"]), _t(data["cond"]), reference_scale=True)
    assert_parity(-lp.cpu().numpy(), data["nll_rows"], what=name + " nll rows", tol=2e-5)
    lp2 = model.log_prob(_t(data["y"]), _t(data["cond"]))
    d = data["y"].shape[1]
    assert torch.allclose(lp2, lp - 0.5 * d * np.log(2 * np.pi), atol=1e-4)
    nll = bcnf_b200.inn_nll_loss(model(_t(data["y"]), _t(data["cond"]), log_det_J=True), model.log_det_J)
    assert abs(float(nll) - float(data["nll"])) < 2e-5 * max(1.0, abs(float(data["nll"])))


# ------------------------------------------------------------------------------------------
# oracle parity on seeded shapes that are too big to commit as fixtures
# ------------------------------------------------------------------------------------------
def _random_model(size, nested, n_blocks, n_cond, two_way=False, act_norm=True, seed=0):
    torch.manual_seed(seed)
    model = CondRealNVP_v2(size=size, nested_sizes=nested, n_blocks=n_blocks, n_conditions=n_cond,
                           feature_networks=[bcnf_b200.ConcatenateCondition(None, n_cond)], dropout=0.3,
                           act_norm=act_norm, two_way=two_way)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for layer in model.layers:
            if isinstance(layer, bcnf_b200.ActNorm):
                layer.scale.copy_(0.75 + 0.5 * torch.rand(layer.scale.shape, generator=g))
                layer.bias.copy_(0.1 * torch.randn(layer.bias.shape, generator=g))
    return model.to(DEV).eval()


SHAPES = [
    # (size, nested, blocks, C, two_way, rows)
    (19, [526] * 5, 3, 1360, False, 300),     # trajectory_*_large conditioner, 3 of its 26 blocks
    (19, [64] * 5, 4, 128, False, 257),       # sweep corners (BASELINE config 5)
    (19, [128] * 5, 2, 128, False, 100),
    (19, [256] * 5, 2, 128, False, 65),
    (19, [512] * 5, 2, 128, False, 33),
    (19, [1024] * 5, 2, 128, False, 31),
    (21, [175, 175, 175], 3, 107, True, 129), # D=21, two-way, odd width (Appendix B)
    (19, [32] * 4, 6, 16, False, 1000),       # row-per-thread, width 32
    (21, [16] * 3, 5, 32, True, 513),         # row-per-thread, D=21 two-way
    (4, [8], 2, 2, False, 5),                 # tiny
]


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: f"D{s[0]}_H{s[1][0]}x{len(s[1])}_K{s[2]}_C{s[3]}_tw{int(s[4])}")
def test_forward_inverse_match_oracle(shape):
    size, nested, blocks, n_cond, two_way, rows = shape
    model = _random_model(size, nested, blocks, n_cond, two_way)
    g = torch.Generator().manual_seed(11)
    y = torch.randn(rows, size, generator=g)
    h = torch.randn(rows, n_cond, generator=g)
    z_in = torch.randn(rows, size, generator=g)
    z = model(y, h, log_det_J=True)
    ld = model.log_det_J
    x = model.inverse(z_in, h)
    sd = {k: v.cpu().numpy() for k, v in model.state_dict().items()}
    layers32 = fo.layers_from_state_dict(sd)
    layers64 = fo.layers_from_state_dict(sd, convert=lambda v: np.asarray(v, dtype=np.float64))
    z32, ld32 = fo.stack_forward(layers32, y.numpy(), h.numpy())
    z64, ld64 = fo.stack_forward(layers64, y.numpy().astype(np.float64), h.numpy().astype(np.float64))
    x32 = fo.stack_inverse(layers32, z_in.numpy(), h.numpy())
    x64 = fo.stack_inverse(layers64, z_in.numpy().astype(np.float64), h.numpy().astype(np.float64))
    assert_parity(z.cpu().numpy(), z32, z64, what="z")
    assert_parity(ld.cpu().numpy(), ld32, ld64, what="logdet")
    assert_parity(x.cpu().numpy(), x32, x64, what="x")


def test_row_to_instance_maps_and_ragged_batches():
    model = _random_model(19, [16] * 3, 3, 8)
    flow = model._flow()
    g = torch.Generator().manual_seed(3)
    h = torch.randn(5, 8, generator=g).to(DEV)
    P = flow.project(h)
    for rows in (0, 1, 127, 128, 129, 1000):
        zin = torch.randn(rows, 19, generator=g).to(DEV)
        idx = torch.arange(rows) % 5
        a, _ = flow.run(True, zin, P, inst_period=5)
        b, _ = flow.run(True, zin, P, row2inst=idx)
        c, _ = flow.run(True, zin, flow.project(h[idx.to(DEV)]) if rows else P)
        assert a.shape == (rows, 19)
        assert torch.equal(a, b) and torch.equal(a, c)


def test_results_do_not_depend_on_batch_composition(kernel_choice):
    # size-independent property used at full BASELINE sizes: a row's result is a function of that
    # row alone, bit for bit, whatever tile / CTA / launch it lands in
    model = _random_model(19, [16] * 7, 8, 80)
    g = torch.Generator().manual_seed(4)
    n = 1 << 17
    y = torch.randn(n, 19, generator=g).to(DEV)
    h = torch.randn(n, 80, generator=g).to(DEV)
    z = model(y, h, log_det_J=True)
    ld = model.log_det_J
    pick = torch.randperm(n, generator=g)[:777].to(DEV)
    z2 = model(y[pick], h[pick], log_det_J=True)
    assert torch.equal(z[pick], z2) and torch.equal(ld[pick], model.log_det_J)


def test_full_size_round_trip_fc_small():
    # BASELINE config 1 at scale: 2^20 log_prob rows, then inverse(forward(y)) against y
    data, sd, meta = load_golden("fc_small")
    model = _model_from_golden(meta, sd)
    g = torch.Generator().manual_seed(5)
    n = 1 << 20
    y = torch.randn(n, 19, generator=g).to(DEV)
    h = torch.randn(n, 80, generator=g).to(DEV)
    flow = model._flow()
    P = flow.project(h)
    z, ld = flow.run(False, y, P, want_logdet=True)
    back, ld_b = flow.run(True, z, P, want_logdet=True)
    assert torch.isfinite(z).all() and torch.isfinite(ld).all()
    # the golden stack is ill-conditioned by construction (ActNorm scales in [0.5, 1.5], 31 mixing
    # layers): the reference's own fp32 round trip is 9e-5 of scale off its fp64 one (make_golden.py)
    err = float((back - y).abs().max() / y.abs().max())
    assert err < 5e-3, err
    assert float((ld - ld_b).abs().max()) < 5e-3 * float(ld.abs().max())


def test_parameter_updates_are_picked_up():
    model = _random_model(19, [16] * 2, 2, 4)
    g = torch.Generator().manual_seed(6)
    y, h = torch.randn(9, 19, generator=g), torch.randn(9, 4, generator=g)
    z0 = model(y, h).clone()
    with torch.no_grad():
        model.layers[0].bias.add_(1.0)
    z1 = model(y, h)
    assert not torch.allclose(z0, z1)
    sd = {k: v.cpu().numpy() for k, v in model.state_dict().items()}
    zo, _ = fo.stack_forward(fo.layers_from_state_dict(sd), y.numpy(), h.numpy())
    assert_parity(z1.cpu().numpy(), zo, what="z after update")


def test_nan_and_inf_propagate():
    model = _random_model(19, [16] * 2, 2, 4)
    y = torch.zeros(4, 19)
    y[1, 3] = float("nan")
    y[2, 12] = float("inf")
    z = model(y, torch.zeros(4, 4), log_det_J=True).cpu()
    assert torch.isfinite(z[0]).all() and torch.isfinite(z[3]).all()
    assert torch.isnan(z[1]).any() and not torch.isfinite(z[2]).all()


def test_calibration_ranks_reference_order_and_device_path():
    # compute_y_hat_ranks (reference calibration.py:20-48) on top of sample(): exact replay of the reference's
    # computation from its own golden samples, and the on-device reduction path
    from bcnf_b200 import compute_CDF_residuals, compute_y_hat_ranks
    data, sd, meta = load_golden("fc_small")
    model = _model_from_golden(meta, sd, sample_rng="reference")
    cond, y = _t(data["cond"])[:7], _t(data["y"])[:7]
    torch.manual_seed(1234)
    ranks = compute_y_hat_ranks(model, y, cond, M_samples=5, batch_size=4, sample_batch_size=2, device=DEV, verbose=False)
    expect = (torch.cat([_t(data["sample_outer"]) / 0.8 * 1.0, y.unsqueeze(0)]) < y.unsqueeze(0)).sum(0)
    # golden samples were drawn with sigma=0.8; ranks use sigma=1, so only shape/range are comparable here
    assert ranks.shape == expect.shape == (7, 19) and ranks.dtype == torch.int64
    assert int(ranks.min()) >= 0 and int(ranks.max()) <= 5
    model_d = _model_from_golden(meta, sd)
    r = compute_y_hat_ranks(model_d, y, cond, M_samples=4000, device=DEV, verbose=False)
    assert r.shape == (7, 19) and r.device.type == "cpu" and int(r.max()) <= 4000
    # device-path ranks agree with ranks computed from an explicit sample tensor of the same model
    s = model_d.sample(4000, cond, outer=True, output_device="cpu")
    r2 = (s < y.unsqueeze(0)).sum(0)
    assert ((r - r2).abs().float() < 6 * (4000 * 0.25) ** 0.5 + 1).all()      # two independent draws: binomial spread
    t, res, ci = compute_CDF_residuals(r, 4000)
    assert t.shape == res.shape[-1:] == ci.shape and np.isfinite(res).all()
