"""Host-side boundary tests (no GPU): API surface, state_dict layout, config loading, C-ABI exports."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

import bcnf_b200
from bcnf_b200 import CondRealNVP_v2, _cabi
from conftest import GOLDEN_DIR, ROOT, load_golden


def test_library_loads_and_exports_every_declared_symbol():
    lib = _cabi.lib()
    header = open(os.path.join(ROOT, "include", "bcnf_b200.h")).read()
    declared = set(re.findall(r"\b(bcnf_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_cabi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.bcnf_abi_version() == 2


def test_abi_validation_errors_without_gpu():
    lib = _cabi.lib()
    assert lib.bcnf_flow_create(None, None, None) == -1
    assert b"null" in lib.bcnf_last_error()
    desc = _cabi.FlowDesc()
    desc.size, desc.n_conditions, desc.n_hidden, desc.n_ops = 19, 0, 1, 1
    desc.hidden[0] = 16
    h = ctypes.c_void_p()
    arr = (ctypes.c_int32 * 1)(1)
    assert lib.bcnf_flow_create(ctypes.byref(desc), arr, ctypes.byref(h)) == -2     # n_conditions == 0
    desc.n_conditions = 4
    desc.size = 1000
    assert lib.bcnf_flow_create(ctypes.byref(desc), arr, ctypes.byref(h)) == -2
    desc.size = 19
    arr[0] = 7
    assert lib.bcnf_flow_create(ctypes.byref(desc), arr, ctypes.byref(h)) == -1     # unknown layer type
    assert lib.bcnf_flow_forward(None, None, None, None, 0, 0, None, None, None) == -1
    assert lib.bcnf_flow_destroy(None) == 0
    # trainer-step fusion entry points validate before they touch the device
    assert lib.bcnf_adam_flat(None, None, None, None, 0, None, None, 0, None) == -1
    assert lib.bcnf_adam_flat(16, 16, 16, 16, 6, 16, 16, 0, None) == -1          # n not a multiple of 4
    assert lib.bcnf_train_nll(None, None, 4, 19, None, None, None, 0, None) == -1
    assert lib.bcnf_train_nll(16, 16, 0, 19, 16, 16, 16, 0, None) == -1           # empty batch


def test_flat_adam_refuses_what_it_cannot_hold():
    import torch
    import bcnf_b200
    lstm = torch.nn.LSTM(3, 4)
    with pytest.raises(NotImplementedError):
        bcnf_b200.FlatAdam(lstm)                      # cuDNN keeps RNN weights in its own flat buffer
    with pytest.raises(ValueError):
        bcnf_b200.FlatAdam(torch.nn.Linear(3, 4))     # CPU parameters: the step is a CUDA kernel, no fallback


@pytest.mark.parametrize("name", ["trajectory_FC_small", "trajectory_FC_large", "trajectory_LSTM_large",
                                  "trajectory_TRF_large"])
def test_state_dict_layout_matches_reference(name):
    rec = json.load(open(os.path.join(GOLDEN_DIR, "state_dict_keys.json")))[name]
    model = CondRealNVP_v2.from_config(rec["config"])
    mine = {k: list(v.shape) for k, v in model.state_dict().items()}
    assert mine == rec["keys"]
    assert model.n_params == rec["n_params"]
    assert len(model.layers) == rec["n_layers"]


def test_golden_state_dicts_load_strictly(golden):
    name, data, sd, meta = golden
    model = CondRealNVP_v2.from_config(meta["config"])
    res = model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert model.n_params == meta["n_params"]


def test_same_q_in_every_block_when_random_state_is_set():
    kw = dict(size=5, nested_sizes=[8], n_blocks=4, n_conditions=3,
              feature_networks=[bcnf_b200.ConcatenateCondition(None, 3)])
    m = CondRealNVP_v2(random_state=7, **kw)
    qs = [l.orthonormal_matrix for l in m.layers if isinstance(l, bcnf_b200.OrthonormalTransformation)]
    assert all(torch.equal(qs[0], q) for q in qs[1:])        # reference quirk, cnf.py:319-320
    m2 = CondRealNVP_v2(**kw)
    qs2 = [l.orthonormal_matrix for l in m2.layers if isinstance(l, bcnf_b200.OrthonormalTransformation)]
    assert not torch.equal(qs2[0], qs2[1])
    assert not qs[0].requires_grad


def test_unsupported_configurations_are_refused():
    fn = [bcnf_b200.ConcatenateCondition(None, 3)]
    with pytest.raises(NotImplementedError):
        CondRealNVP_v2(5, [8], 2, 3, feature_networks=fn, layer="AnyGLU")
    with pytest.raises(NotImplementedError):
        CondRealNVP_v2(5, [8], 2, 3, feature_networks=fn, activation="ReLU")
    with pytest.raises(NotImplementedError):
        CondRealNVP_v2(5, [8], 2, 0, feature_networks=fn)
    with pytest.raises(ValueError):
        CondRealNVP_v2(5, [8], 2, 3, feature_networks=None)     # feature_network.py:32-33
    with pytest.raises(NotImplementedError):
        bcnf_b200.FeatureNetworkFactory.get_feature_network("Nope", {})


def test_verify_rejects_mismatched_feature_sizes():
    cfg = {"global": {"parameter_selection": list("abcde")},
           "model": {"kwargs": dict(size=5, nested_sizes=[8], n_blocks=2, n_conditions=4)},
           "feature_networks": [{"type": "ConcatenateCondition", "kwargs": {"input_size": None, "output_size": 6}},
                                {"type": "FullyConnected", "kwargs": {"sizes": [6, 3]}}]}
    with pytest.raises(AssertionError):
        CondRealNVP_v2.from_config(cfg)


def test_cpu_device_has_no_fallback():
    cfg = json.load(open(os.path.join(GOLDEN_DIR, "state_dict_keys.json")))["trajectory_FC_small"]["config"]
    model = CondRealNVP_v2.from_config(cfg).eval()
    with pytest.raises(RuntimeError, match="no CPU path"):
        model(torch.randn(2, 19), torch.randn(2, 30, 3))


def test_training_mode_needs_cuda_too_and_sampling_needs_eval():
    cfg = json.load(open(os.path.join(GOLDEN_DIR, "state_dict_keys.json")))["trajectory_FC_small"]["config"]
    model = CondRealNVP_v2.from_config(cfg)          # a fresh module is in training mode
    with pytest.raises(RuntimeError):                  # CPU parameters: no fallback in the training path either
        model(torch.randn(2, 19), torch.randn(2, 30, 3))
    with pytest.raises(NotImplementedError):
        model.sample(3, torch.randn(2, 30, 3), outer=True)


def test_load_config_coerces_yaml_numbers(tmp_path):
    p = tmp_path / "run.yaml"
    p.write_text("global:\n  parameter_selection: [a]\noptimizer:\n  kwargs:\n    lr: 2e-4\n"
                 "training:\n  random_state: 2024_03_25\ndata:\n  path: '{{BCNF_ROOT}}/data'\n  config_file: 'x'\n")
    cfg = bcnf_b200.load_config(str(p), root="/tmp/root")
    assert cfg["optimizer"]["kwargs"]["lr"] == pytest.approx(2e-4)
    assert cfg["training"]["random_state"] == 20240325
    assert cfg["data"]["path"] == "/tmp/root/data"
    assert cfg["global"]["hybrid_weight"] == 0


def test_inn_nll_loss_matches_golden(golden):
    name, data, sd, meta = golden
    v = bcnf_b200.inn_nll_loss(torch.from_numpy(data["z"]), torch.from_numpy(data["logdet"]))
    assert abs(float(v) - float(data["nll"])) < 1e-5 * max(1.0, abs(float(data["nll"])))


def test_parameter_index_mapping():
    m = bcnf_b200.ParameterIndexMapping(["a", "b"])
    assert len(m) == 2 and m["b"] == 1 and "a" in m and list(m) == ["a", "b"]
    assert m.vectorize({"a": [1, 2], "b": [3, 4]}).shape == (2, 2)
    with pytest.raises(KeyError):
        m.vectorize({"a": 1})


def test_lstm_feature_network_pools_before_the_output_layer_without_changing_the_result():
    """Mean pooling commutes with the affine output layer (reference feature_network.py:168-176 applies the Linear to
    every time step and pools afterwards): same features up to fp32 rounding, 1/seq_len of the FLOPs."""
    import torch
    from bcnf_b200.feature_network import LSTMFeatureNetwork
    torch.manual_seed(0)
    for pool_axis, batch in (("time", 7), ("reference", 30)):
        net = LSTMFeatureNetwork(input_size=3, hidden_size=12, output_size=20, num_layers=2, bidirectional=True,
                                 pooling="mean", pool_axis=pool_axis).eval()
        x = torch.randn(batch, 30, 3)
        with torch.no_grad():
            seq, _ = net.lstm(x)
            ref = net.linear(seq).mean(dim=1 if pool_axis == "time" else 0)
            out = net(x)
        assert out.shape == ref.shape
        assert torch.allclose(out, ref, rtol=1e-5, atol=1e-6)
    net = LSTMFeatureNetwork(input_size=3, hidden_size=12, output_size=20, num_layers=1, pooling="max").eval()
    x = torch.randn(5, 30, 3)
    with torch.no_grad():
        seq, _ = net.lstm(x)
        assert torch.equal(net(x), net.linear(seq).max(dim=1).values)


def test_lstm_step_gemm_layout_reproduces_nn_lstm_on_cpu():
    """The operand layout of the tensor-core LSTM (bcnf_b200/feature_tc.py: gate-interleaved [W_ih | W_hh], inputs and
    h_(t-1) in separate 64-column chunks, per-direction chunks of the layer below) emulated with dense torch ops on the
    CPU: one GEMM + cell update per layer, direction and time step must give nn.LSTM's output."""
    import torch
    from bcnf_b200 import feature_tc
    torch.manual_seed(1)
    B, T, F, H = 5, 7, 3, 6
    lstm = torch.nn.LSTM(F, H, num_layers=2, bidirectional=True, batch_first=True).eval()
    x = torch.randn(B, T, F)
    hc = (H + 63) // 64
    below = None                                   # [direction][t] -> (B, hc*64) padded h of the layer below
    with torch.no_grad():
        ref, _ = lstm(x)
        for layer in range(2):
            outs = []
            for d in range(2):
                wc, bias, in_cols = feature_tc.lstm_cat_weights(lstm, layer, d)
                assert wc.shape == (4 * H, in_cols + hc * 64)
                h = torch.zeros(B, hc * 64)
                c = torch.zeros(B, H)
                per_t = [None] * T
                for t in (range(T) if d == 0 else range(T - 1, -1, -1)):
                    if layer == 0:
                        inp = torch.zeros(B, in_cols)
                        inp[:, :F] = x[:, t]
                    else:
                        inp = torch.cat([below[0][t], below[1][t]], dim=1)
                    gates = torch.cat([inp, h], dim=1) @ wc.t() + bias          # (B, 4H), column n = 4*unit + gate
                    gi, gf, gg, go = (gates[:, k::4] for k in range(4))
                    c = torch.sigmoid(gf) * c + torch.sigmoid(gi) * torch.tanh(gg)
                    hv = torch.sigmoid(go) * torch.tanh(c)
                    h = torch.zeros(B, hc * 64)
                    h[:, :H] = hv
                    per_t[t] = h
                outs.append(per_t)
            below = outs
        seq = torch.stack([torch.cat([below[0][t][:, :H], below[1][t][:, :H]], dim=1) for t in range(T)], dim=1)
    assert torch.allclose(seq, ref, rtol=1e-5, atol=1e-6)


def test_training_plan_splits_the_layer_list_into_networks_and_glue_ops():
    """bcnf_b200/train.py: model.layers -> glue ops before the first coupling + one unit per conditioner network with the
    ActNorm / mixing layers that follow it (reference layer order, cnf.py:392-423)."""
    from bcnf_b200 import _cabi, train
    kinds = ["actnorm", "coupling", "ortho", "actnorm", "coupling", "ortho", "actnorm", "coupling"]
    spec = train._Spec(kinds, n_lin=4, two_way=False, size=19, n_conditions=8, p_drop=0.0, seed=0)
    lead, units = train._plan(spec)
    assert lead == [(_cabi.GLUE_ACTNORM, 0)]
    assert [(u.li, u.net, u.src0, u.din, u.dst0, u.dout) for u in units] == [(1, 0, 0, 10, 10, 9), (4, 0, 0, 10, 10, 9),
                                                                             (7, 0, 0, 10, 10, 9)]
    # parameter offsets: actnorm (2), coupling (2 * n_lin), ortho (1), ...
    assert [u.w0 for u in units] == [2, 13, 24]
    assert units[0].ops == [(_cabi.GLUE_ORTHO, 10), (_cabi.GLUE_ACTNORM, 11)] and units[2].ops == []
    spec2 = train._Spec(["coupling", "ortho", "coupling"], n_lin=3, two_way=True, size=21, n_conditions=4, p_drop=0.0, seed=0)
    lead2, units2 = train._plan(spec2)
    assert lead2 == [] and [(u.li, u.net, u.src0, u.din, u.dst0, u.dout) for u in units2] == [
        (0, 0, 0, 11, 11, 10), (0, 1, 11, 10, 0, 11), (2, 0, 0, 11, 11, 10), (2, 1, 11, 10, 0, 11)]
    assert units2[0].ops == [] and units2[1].ops == [(_cabi.GLUE_ORTHO, 12)]


def test_tensor_core_feature_paths_only_take_what_they_reproduce():
    """bcnf_b200/feature_tc.py host logic: which feature-network modules the tensor-core paths accept."""
    import torch
    from torch import nn
    from bcnf_b200 import feature_tc
    from bcnf_b200.feature_network import FullyConnectedFeatureNetwork, LSTMFeatureNetwork
    assert feature_tc.supported(FullyConnectedFeatureNetwork([90, 310, 310, 1360], dropout=0.1))
    assert feature_tc.supported(FullyConnectedFeatureNetwork([90, 80]))
    assert not feature_tc.supported(FullyConnectedFeatureNetwork([90, 64, 8], activation=nn.ReLU))
    assert not feature_tc.supported(FullyConnectedFeatureNetwork([90, 64, 8], batch_norm=True))
    assert not feature_tc.supported(FullyConnectedFeatureNetwork([90, 64, 8]).double())
    ok = LSTMFeatureNetwork(3, 140, 1360, num_layers=2, bidirectional=True, pooling="mean")
    assert feature_tc.lstm_supported(ok)
    assert not feature_tc.lstm_supported(LSTMFeatureNetwork(3, 140, 1360, num_layers=2, pooling="max"))
    assert not feature_tc.lstm_supported(LSTMFeatureNetwork(3, 140, 1360, num_layers=1, pool_axis="reference"))
    assert not feature_tc.lstm_supported(LSTMFeatureNetwork(3, 141, 64, num_layers=1))        # odd hidden size
    assert not feature_tc.lstm_supported(LSTMFeatureNetwork(3, 512, 64, num_layers=1))        # 4H > 1024
    # CPU tensors, training mode or autograd never reach the tensor-core path
    ok.tc_passes = 3
    x = torch.randn(4, 30, 3)
    with torch.no_grad():
        assert ok.eval()(x).shape == (4, 1360)


def test_image_pool_leases_are_recycled_but_never_shared():
    import gc
    import torch
    from bcnf_b200 import train
    dev = torch.device("cpu")
    shapes = [(8, 70), (8, 130)]
    a = train._PoolLease(dev, shapes)
    b = train._PoolLease(dev, shapes)                       # alive at the same time: different buffers
    assert a.images[0].buf.data_ptr() != b.images[0].buf.data_ptr()
    assert a.images[1].chunks == 4 and a.images[0].rpad == 128 and not a.images[0].buf.any()
    ptr = a.images[0].buf.data_ptr()
    del a
    gc.collect()
    c = train._PoolLease(dev, shapes)                       # the released pool comes back
    assert c.images[0].buf.data_ptr() == ptr
    d = train._PoolLease(dev, [(8, 70)])                    # other shapes: other pool
    assert d.images[0].buf.data_ptr() not in (ptr, b.images[0].buf.data_ptr())


def test_transformer_entry_points_validate_before_they_touch_the_device():
    """bcnf_trf_* (include/bcnf_b200.h): argument checks run without a GPU."""
    lib = _cabi.lib()
    assert lib.bcnf_trf_embed(None, None, None, None, None, 8, 30, 3, 128, None, None, 0, 0, 0, None) == -1
    assert lib.bcnf_trf_embed(16, 16, 16, None, None, 30, 30, 3, 100, 16, 16, 1 << 20, 256, 0, None) == -1     # E % 8
    assert b"multiple of 8" in lib.bcnf_last_error()
    assert lib.bcnf_trf_embed(16, 16, 16, None, None, 300, 30, 3, 128, 16, 16, 1 << 20, 256, 0, None) == -1    # image rows < rows
    assert lib.bcnf_trf_attention(16, 4, 65, 128, 8, None, 16, 1 << 24, 512, 0, None) == -1                    # T > 64
    assert lib.bcnf_trf_attention(16, 4, 30, 128, 5, None, 16, 1 << 24, 512, 0, None) == -1                    # E % heads
    assert lib.bcnf_trf_attention(16, 4, 30, 96, 8, None, 16, 1 << 24, 512, 0, None) == -2                     # head width 12
    assert lib.bcnf_trf_attention_bwd(16, 16, 4, 40, 128, 8, 16, 0, None) == -1                                # T > 32
    assert lib.bcnf_trf_attention_bwd(None, 16, 4, 30, 128, 8, 16, 0, None) == -1
    assert lib.bcnf_trf_add_layernorm(16, 16, None, 16, 16, 1e-5, 30, 128, None, 16, None, 16, 16, 1 << 20, 256, 0, None) == -1   # mean without rstd
    assert lib.bcnf_trf_gelu(None, 30, 128, None, 16, 1 << 20, 256, 0, None) == -1
    assert lib.bcnf_trf_ln_param_grad(16, 16, 16, 16, 30, 128, None, 16, 0, None) == -1
    # empty problems are accepted without a launch
    assert lib.bcnf_trf_gelu(16, 0, 128, None, 16, 1 << 20, 256, 0, None) == 0
    assert lib.bcnf_trf_ln_param_grad(16, 16, 16, 16, 0, 128, 16, 16, 0, None) == 0


def test_transformer_paths_only_take_what_they_reproduce_and_slab_weight_gradient():
    """Host logic of the Transformer paths (feature_tc.transformer_supported, trf_train.usable) and of the slab-batched
    weight gradient the side streams compute (feature_network._wgrad)."""
    import torch
    from torch import nn
    from bcnf_b200 import feature_tc, trf_train
    from bcnf_b200.feature_network import Transformer, _wgrad, OffChain
    ok = Transformer(input_size=3, trf_size=128, n_heads=8, ff_size=128, n_blocks=8, output_size=1360)
    assert feature_tc.transformer_supported(ok)
    assert not feature_tc.transformer_supported(Transformer(3, 100, 4, 64, 1, 8))            # d_model % 8
    assert not feature_tc.transformer_supported(Transformer(3, 96, 8, 64, 1, 8))             # head width 12
    assert not feature_tc.transformer_supported(Transformer(3, 64, 4, 64, 1, 8).double())
    tanh = Transformer(3, 64, 4, 64, 1, 8)
    tanh.layers[0].ffn[1] = nn.GELU(approximate="tanh")
    assert not feature_tc.transformer_supported(tanh)
    assert not trf_train.usable(ok, torch.randn(4, 30, 3))                                    # CPU tensors: PyTorch modules
    # eval mode on the CPU, and training mode outside a Trainer step, run the plain modules
    ok.tc_passes = 3
    with torch.no_grad():
        assert ok.eval()(torch.randn(4, 30, 3)).shape == (4, 1360)
    g = torch.Generator().manual_seed(0)
    for rows in (7680, 4096, 96):          # 32 slabs of 240 / 32 slabs of 128 / too few rows: plain product
        gg, xx = torch.randn(rows, 24, generator=g), torch.randn(rows, 40, generator=g)
        ref = gg.double().t() @ xx.double()
        assert torch.allclose(_wgrad(gg, xx).double(), ref, rtol=1e-5, atol=1e-4)
    # OffChain.accumulate: the first gradient becomes .grad, later ones add to it; frozen parameters are skipped
    p = nn.Parameter(torch.zeros(3))
    OffChain.accumulate(p, torch.ones(3))
    OffChain.accumulate(p, torch.ones(3))
    assert torch.equal(p.grad, torch.full((3,), 2.0))
    frozen = nn.Parameter(torch.zeros(3), requires_grad=False)
    OffChain.accumulate(frozen, torch.ones(3))
    assert frozen.grad is None
