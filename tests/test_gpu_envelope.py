"""Edges of the supported envelope and the other BASELINE configs, on the GPU."""
import json
import os

import numpy as np
import pytest
import torch

import bcnf_b200
from bcnf_b200 import CondRealNVP_v2
from conftest import GOLDEN_DIR, assert_parity, rel_err
from oracle import flow_oracle as fo

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _no_grad():
    with torch.no_grad():
        yield


def _stack(size, nested, n_blocks, n_cond, precision="auto", two_way=False, act_norm=True, hybrid=False, seed=0):
    torch.manual_seed(seed)
    return CondRealNVP_v2(size=size, nested_sizes=nested, n_blocks=n_blocks, n_conditions=n_cond,
                          feature_networks=[bcnf_b200.ConcatenateCondition(None, n_cond)], dropout=0.1,
                          act_norm=act_norm, two_way=two_way, precision=precision, hybrid=hybrid).to(DEV).eval()


def _check_vs_oracle(model, rows, seed=3):
    g = torch.Generator().manual_seed(seed)
    y = torch.randn(rows, model.size, generator=g)
    h = torch.randn(rows, model.n_conditions, generator=g)
    z = model(y, h, log_det_J=True)
    x = model.inverse(y, h)
    sd = {k: v.cpu().numpy() for k, v in model.state_dict().items() if k.startswith("layers.")}
    l32 = fo.layers_from_state_dict(sd)
    l64 = fo.layers_from_state_dict(sd, convert=lambda v: np.asarray(v, dtype=np.float64))
    z32, ld32 = fo.stack_forward(l32, y.numpy(), h.numpy())
    z64, ld64 = fo.stack_forward(l64, y.numpy().astype(np.float64), h.numpy().astype(np.float64))
    x32 = fo.stack_inverse(l32, y.numpy(), h.numpy())
    x64 = fo.stack_inverse(l64, y.numpy().astype(np.float64), h.numpy().astype(np.float64))
    assert_parity(z.cpu().numpy(), z32, z64, what="z")
    assert_parity(model.log_det_J.cpu().numpy(), ld32, ld64, what="logdet")
    assert_parity(x.cpu().numpy(), x32, x64, what="x")


@pytest.mark.parametrize("case", [
    dict(size=64, nested=[96] * 2, n_blocks=2, n_cond=7, rows=70),                    # widest flow (BCNF_MAX_SIZE)
    dict(size=2, nested=[8], n_blocks=2, n_cond=1, rows=9),                           # narrowest
    dict(size=19, nested=[24] * 8, n_blocks=2, n_cond=5, rows=40),                    # deepest conditioner (8 hidden layers)
    dict(size=19, nested=[1024] * 2, n_blocks=2, n_cond=16, rows=20, precision="fp32"),   # widest conditioner, fp32 tiled kernel
    dict(size=19, nested=[48, 200, 64], n_blocks=2, n_cond=9, rows=130),              # non-uniform widths
    dict(size=21, nested=[64], n_blocks=1, n_cond=3, rows=3, act_norm=False),         # a single coupling layer stack
], ids=["D64", "D2", "L8", "H1024_fp32", "nonuniform", "single_block"])
def test_envelope_edges_match_oracle(case):
    case = dict(case)
    rows = case.pop("rows")
    model = _stack(case.pop("size"), case.pop("nested"), case.pop("n_blocks"), case.pop("n_cond"), **case)
    _check_vs_oracle(model, rows)


def test_too_many_hidden_layers_is_refused():
    with pytest.raises(NotImplementedError):
        _stack(19, [16] * 9, 2, 4)(torch.zeros(1, 19), torch.zeros(1, 4))


def test_explicit_row_to_instance_map_on_the_tensor_core_kernel():
    model = _stack(19, [128, 128], 3, 12, precision="bf16x3")
    flow = model._flow()
    g = torch.Generator().manual_seed(8)
    h = torch.randn(37, 12, generator=g).to(DEV)
    z = torch.randn(1000, 19, generator=g).to(DEV)
    idx = torch.randint(0, 37, (1000,), generator=g)
    P = flow.project(h)
    a, _ = flow.run(True, z, P, row2inst=idx)
    b, _ = flow.run(True, z, flow.project(h[idx.to(DEV)]))
    assert torch.equal(a, b)
    c, _ = flow.run(True, z[:37 * 20], P, inst_period=37)
    d, _ = flow.run(True, z[:37 * 20], P, row2inst=torch.arange(37 * 20) % 37)
    assert torch.equal(c, d)


def test_hybrid_model_has_a_prediction_head_and_trains():
    model = _stack(19, [16] * 2, 2, 8, hybrid=True)
    assert "prediction_head.weight" in model.state_dict()
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    trainer = bcnf_b200.Trainer(model, opt, hybrid_weight=0.5)
    g = torch.Generator().manual_seed(1)
    with torch.enable_grad():
        loss, nll, mse = trainer.train_batch(torch.randn(32, 19, generator=g), torch.randn(32, 8, generator=g))
    assert np.isfinite([loss, nll, mse]).all() and mse > 0
    assert abs(loss - (nll + 0.5 * mse) / 1.5) < 1e-4 * max(1.0, abs(loss))          # trainer.py:269


@pytest.mark.parametrize("name", ["trajectory_LSTM_large", "trajectory_TRF_large"])
def test_other_large_baseline_configs_run_and_match_the_oracle_at_h(name):
    cfg = json.load(open(os.path.join(GOLDEN_DIR, "state_dict_keys.json")))[name]["config"]
    torch.manual_seed(0)
    model = CondRealNVP_v2.from_config(cfg).to(DEV).eval()
    assert model._flow().kernel == "tcgen05"
    g = torch.Generator().manual_seed(2)
    y = torch.randn(64, 19, generator=g)
    cond = torch.randn(64, 30, 3, generator=g)
    z, h = model(y, cond, log_det_J=True, return_features=True)
    assert h.shape == (64, 1360) and torch.isfinite(z).all() and torch.isfinite(model.log_det_J).all()
    # parity of the coupling stack at the h boundary (the feature networks stay PyTorch)
    sd = {k: v.cpu().numpy() for k, v in model.state_dict().items() if k.startswith("layers.")}
    z32, ld32 = fo.stack_forward(fo.layers_from_state_dict(sd), y.numpy(), h.cpu().numpy())
    assert rel_err(z.cpu().numpy(), z32) < 1e-5 and rel_err(model.log_det_J.cpu().numpy(), ld32) < 1e-5
    s = model.sample(7, cond[:5], outer=True)
    assert s.shape == (7, 5, 19) and torch.isfinite(s).all()
    lp = model.log_prob(y, cond)
    assert lp.shape == (64,) and torch.isfinite(lp).all()


def test_projection_and_flow_capture_in_a_cuda_graph_at_a_grown_instance_count():
    """bcnf_cond_project + bcnf_flow_forward make no hidden synchronisation or allocation (include/bcnf_b200.h): after a
    warm-up on the capturing stream with FEW instances, a CUDA graph of project + forward captures at MANY instances
    (the projection's scratch image is sized once, by set_params) and replays to the eager result."""
    model = _stack(19, [128, 128], 3, 12, precision="bf16x3")
    flow = model._flow()
    g = torch.Generator().manual_seed(4)
    h_small, y_small = torch.randn(5, 12, generator=g).to(DEV), torch.randn(5, 19, generator=g).to(DEV)
    h_big, y_big = torch.randn(3000, 12, generator=g).to(DEV), torch.randn(3000, 19, generator=g).to(DEV)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        flow.run(False, y_small, flow.project(h_small), want_logdet=True)     # per-stream activation scratch exists now
    s.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=s):
        z_g, ld_g = flow.run(False, y_big, flow.project(h_big), want_logdet=True)
    z_g.zero_(); ld_g.zero_()
    graph.replay()
    torch.cuda.synchronize()
    z_e, ld_e = flow.run(False, y_big, flow.project(h_big), want_logdet=True)
    assert torch.equal(z_g, z_e) and torch.equal(ld_g, ld_e)


def test_entry_points_leave_the_current_device_alone():
    """Every ABI entry point restores the caller's current CUDA device; only meaningful with >= 2 devices,
    but the single-device path must at least keep device 0 current after create / run / destroy."""
    before = torch.cuda.current_device()
    model = _stack(19, [16] * 2, 2, 4)
    model(torch.zeros(3, 19), torch.zeros(3, 4))
    del model
    assert torch.cuda.current_device() == before
    if torch.cuda.device_count() >= 2:
        with torch.cuda.device(1):
            m0 = _stack(19, [16] * 2, 2, 4)            # lives on cuda:0
            m0(torch.zeros(3, 19), torch.zeros(3, 4))
            assert torch.cuda.current_device() == 1


@pytest.mark.parametrize("world", [2, 3, 8])
def test_instance_shards_reproduce_the_unsharded_samples(world):
    """sharding by conditioning instance must not change any instance's samples.  The N ranks of a
    node are emulated one after the other on this GPU (the blocks of bcnf_b200.sharding.shard_bounds, the z rows of
    each block injected): every (sample, instance) row is bit-identical to the unsharded launch, for the tensor-core
    and the fp32 kernel.  tools/shard_equiv_check.py is the same check with real ranks under torchrun."""
    from bcnf_b200 import sharding
    for precision, nested in (("bf16x3", [128, 128]), ("fp32", [16, 16])):
        model = _stack(19, nested, 3, 12, precision=precision)
        g = torch.Generator().manual_seed(6)
        n_inst, m = 37, 5
        cond = torch.randn(n_inst, 12, generator=g)
        z = torch.randn(m, n_inst, 19, generator=g)
        full = model._sample(m, cond, outer=True, z=z.reshape(-1, 19))
        parts = []
        for r in range(world):
            lo, hi = sharding.shard_bounds(n_inst, r, world)
            if hi > lo:
                parts.append(model._sample(m, cond[lo:hi], outer=True, z=z[:, lo:hi].reshape(-1, 19)))
        assert torch.equal(torch.cat(parts, dim=1), full)
