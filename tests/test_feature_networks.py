"""Condition encoders against outputs recorded from the live reference (tests/golden/make_feature_golden.py).

The reference's state_dicts load unchanged into bcnf_b200's modules (same parameter names) and the features h must
match: FullyConnectedFeatureNetwork (feature_network.py:114-145), LSTMFeatureNetwork (:148-178; the reference pools
over the batch axis, reproduced by pool_axis="reference"), Transformer (:263-307).  CPU: plain PyTorch path, 1e-6.
GPU: the same modules on the device, and the tensor-core implementations of the FullyConnected / LSTM / Transformer
encoders (bcnf_b200/feature_tc.py) on a tiled batch large enough to take that path.
"""
import json
import os

import numpy as np
import pytest
import torch

import bcnf_b200
from bcnf_b200 import feature_network as fn
from conftest import GOLDEN_DIR, rel_err

DATA = np.load(os.path.join(GOLDEN_DIR, "feature_networks.npz"))
META = json.loads(str(DATA["meta"]))["cases"]


def _build(name, **extra):
    m = META[name]
    net = getattr(fn, m["class"])(**m["kwargs"], **extra)
    sd = {k[len(name) + 4:]: torch.from_numpy(DATA[k]) for k in DATA.files if k.startswith(name + "/sd/")}
    missing, unexpected = net.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    return net.eval()


@pytest.mark.parametrize("name", sorted(META))
def test_state_dict_loads_and_features_match_the_reference_cpu(name):
    extra = {"pool_axis": "reference"} if name.startswith("lstm") else {}
    net = _build(name, **extra)
    x = torch.from_numpy(DATA[name + "/x"])
    with torch.no_grad():
        h = net(x)
    ref = DATA[name + "/h"]
    assert tuple(h.shape) == ref.shape
    assert rel_err(h.numpy(), ref) < 1e-6, rel_err(h.numpy(), ref)


def test_lstm_default_pooling_is_over_time_and_documented():
    """Deviation: the default pools over the time axis -> one feature row per instance."""
    net = _build("lstm_mean")
    x = torch.from_numpy(DATA["lstm_mean/x"])
    with torch.no_grad():
        h = net(x)
    assert tuple(h.shape) == (x.shape[0], META["lstm_mean"]["kwargs"]["output_size"])
    # with batch == seq_len == 30 both poolings are defined; they differ (different axis), so the flag matters
    assert rel_err(h.numpy(), DATA["lstm_mean/h"]) > 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(META))
def test_features_match_the_reference_on_the_device(name):
    extra = {"pool_axis": "reference"} if name.startswith("lstm") else {}
    net = _build(name, **extra).to("cuda:0")
    x = torch.from_numpy(DATA[name + "/x"]).to("cuda:0")
    with torch.no_grad():
        h = net(x)
    # cuDNN's fp32 LSTM path is at ~7e-5 of the fp64 result; the others are plain ATen kernels
    tol = 2e-4 if name.startswith("lstm") else 2e-6
    assert rel_err(h.cpu().numpy(), DATA[name + "/h"]) < tol


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["fc", "lstm_mean", "transformer", "transformer_pos"])
def test_tensor_core_feature_networks_match_the_reference(name):
    """feature_tc.py (CTA-pair GEMM chain / LSTM-cell epilogue / Transformer encoder kernels) on the fixture's instances
    tiled to 4096 rows."""
    net = _build(name).to("cuda:0")
    x0 = torch.from_numpy(DATA[name + "/x"])
    reps = 4096 // x0.shape[0] + 1
    x = x0.repeat(reps, 1, 1)[:4096].to("cuda:0")
    net.tc_passes = 3
    with torch.no_grad():
        h = net(x)
        net.tc_passes = 0
        h_ref = net(x)
    assert h.shape == h_ref.shape
    assert rel_err(h.cpu().numpy(), h_ref.cpu().numpy()) < 1e-4
    assert not torch.equal(h, h_ref)      # (the tensor-core path did run)
    if not name.startswith("lstm"):       # row i of the tiled batch is instance i % B of the fixture
        ref = np.tile(DATA[name + "/h"], (reps, 1))[:4096]
        assert rel_err(h.cpu().numpy(), ref) < 1e-5


def test_transformer_gradients_match_the_reference_cpu():
    """Backward of the Transformer encoder against gradients recorded from the live reference
    (tests/golden/make_transformer_grad_golden.py: training mode, dropouts at zero, scalar sum(h * w)).  The module here is
    the autograd reference the GPU tests hold the Trainer's hand-written encoder backward against
    (tests/test_gpu_training.py), so this pins that backward to the reference as well."""
    data = np.load(os.path.join(GOLDEN_DIR, "transformer_grads.npz"))
    meta = json.loads(str(data["meta"]))
    net = fn.Transformer(**meta["kwargs"])
    missing, unexpected = net.load_state_dict({k[3:]: torch.from_numpy(data[k]) for k in data.files if k.startswith("sd/")},
                                              strict=True)
    assert not missing and not unexpected
    net.train()
    h = net(torch.from_numpy(data["x"]))
    (h * torch.from_numpy(data["w"])).sum().backward()
    assert rel_err(h.detach().numpy(), data["h"]) < 1e-6
    names = [k[5:] for k in data.files if k.startswith("grad/")]
    assert sorted(names) == sorted(n for n, _ in net.named_parameters())
    for n, p in net.named_parameters():
        ref = data["grad/" + n]
        if n.endswith("k_linear.bias"):          # identically zero in exact arithmetic (softmax shift invariance): noise
            assert np.abs(ref).max() < 1e-5 and np.abs(p.grad.numpy()).max() < 1e-5
            continue
        assert rel_err(p.grad.numpy(), ref) < 1e-5, (n, rel_err(p.grad.numpy(), ref))
