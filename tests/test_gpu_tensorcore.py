"""tcgen05 path (bcnf_b200/csrc/flow_tc2.cuh; flow_tc.cuh with BCNF_FLOW_TC=1) against the CPU oracle, smallest shapes first.

Tolerances (relative to max|ref|, SURVEY.md section 8d "correctness gates"):
  bf16x3 (3-term bf16 split, fp32 accumulate): the fp32 gate of conftest.assert_parity, 1e-5.
  bf16   (single pass):  z, x 1e-2; log-det 3e-2 -- the "stated bf16 tolerance" of north_star.
"""
import numpy as np
import pytest
import torch

import bcnf_b200
from bcnf_b200 import CondRealNVP_v2
from conftest import assert_parity, rel_err
from oracle import flow_oracle as fo

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
BF16_TOL = {"z": 1e-2, "x": 1e-2, "logdet": 3e-2}


@pytest.fixture(autouse=True)
def _no_grad():
    with torch.no_grad():
        yield


def _model(size, nested, n_blocks, n_cond, precision, two_way=False, act_norm=True, seed=0):
    torch.manual_seed(seed)
    model = CondRealNVP_v2(size=size, nested_sizes=nested, n_blocks=n_blocks, n_conditions=n_cond,
                           feature_networks=[bcnf_b200.ConcatenateCondition(None, n_cond)], dropout=0.3,
                           act_norm=act_norm, two_way=two_way, precision=precision)
    g = torch.Generator().manual_seed(seed + 1)
    for layer in model.layers:
        if isinstance(layer, bcnf_b200.ActNorm):
            layer.scale.copy_(0.75 + 0.5 * torch.rand(layer.scale.shape, generator=g))
            layer.bias.copy_(0.1 * torch.randn(layer.bias.shape, generator=g))
    return model.to(DEV).eval()


SHAPES = [
    # (size, nested, blocks, C, two_way, rows)   -- ordered from the simplest MMA structure up
    (19, [64], 1, 8, False, 128),               # one K chunk, one N chunk, one full tile
    (19, [64, 64], 1, 8, False, 128),           # adds a hidden 64x64 layer
    (19, [64, 64], 1, 8, False, 77),            # ragged tile
    (19, [128, 128, 128], 2, 16, False, 300),   # 2 K chunks, several tiles, ActNorm + mixing
    (19, [256] * 5, 2, 128, False, 257),        # sweep corner: N chunk of 256
    (19, [512] * 5, 2, 128, False, 200),        # two N chunks
    (19, [526] * 5, 3, 1360, False, 300),       # trajectory_*_large conditioner
    (21, [175, 175, 175], 3, 107, True, 129),   # D=21, two-way, odd width
    (19, [206, 206, 206], 3, 40, False, 513),
    (19, [1024, 1024, 1024], 2, 64, False, 260),  # four 256-column N chunks (BASELINE config 5 corner)
    (19, [336, 336, 336], 3, 32, False, 700),     # 256 + 80 columns, width a multiple of 16: bias added in the epilogue
]


@pytest.mark.parametrize("precision", ["bf16x3", "bf16"])
@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: f"D{s[0]}_H{s[1][0]}x{len(s[1])}_K{s[2]}_C{s[3]}_tw{int(s[4])}_B{s[5]}")
def test_tensorcore_matches_oracle(shape, precision):
    size, nested, blocks, n_cond, two_way, rows = shape
    model = _model(size, nested, blocks, n_cond, precision, two_way)
    assert model._flow().kernel == "tcgen05"
    g = torch.Generator().manual_seed(11)
    y = torch.randn(rows, size, generator=g)
    h = torch.randn(rows, n_cond, generator=g)
    z_in = torch.randn(rows, size, generator=g)
    z = model(y, h, log_det_J=True)
    ld = model.log_det_J
    x = model.inverse(z_in, h)
    sd = {k: v.cpu().numpy() for k, v in model.state_dict().items()}
    l32 = fo.layers_from_state_dict(sd)
    l64 = fo.layers_from_state_dict(sd, convert=lambda v: np.asarray(v, dtype=np.float64))
    z32, ld32 = fo.stack_forward(l32, y.numpy(), h.numpy())
    x32 = fo.stack_inverse(l32, z_in.numpy(), h.numpy())
    z64, ld64 = fo.stack_forward(l64, y.numpy().astype(np.float64), h.numpy().astype(np.float64))
    x64 = fo.stack_inverse(l64, z_in.numpy().astype(np.float64), h.numpy().astype(np.float64))
    errs = {"z": rel_err(z.cpu().numpy(), z32), "logdet": rel_err(ld.cpu().numpy(), ld32),
            "x": rel_err(x.cpu().numpy(), x32)}
    print(f"\n{precision} {shape}: {errs}")
    if precision == "bf16x3":
        assert_parity(z.cpu().numpy(), z32, z64, what="z")
        assert_parity(ld.cpu().numpy(), ld32, ld64, what="logdet")
        assert_parity(x.cpu().numpy(), x32, x64, what="x")
    else:
        for k, tol in BF16_TOL.items():
            assert errs[k] < tol, (k, errs)


def test_fc_feature_network_on_tensor_cores_matches_pytorch():
    """FullyConnectedFeatureNetwork through the CTA-pair GEMM chain (bcnf_b200/feature_tc.py) vs the nn.Sequential it
    mirrors, fp64 reference: stated tolerance 2e-5 of max|ref| for the 3-pass split (eight chained layers)."""
    from bcnf_b200 import feature_tc
    from bcnf_b200.feature_network import FullyConnectedFeatureNetwork
    torch.manual_seed(3)
    net = FullyConnectedFeatureNetwork([90, 310, 310, 310, 310, 310, 310, 310, 1360], dropout=0.111).to(DEV).eval()
    assert feature_tc.supported(net)
    x = torch.randn(4100, 30, 3, device=DEV)
    with torch.no_grad():
        ref = net.double()(x.double().reshape(4100, -1))
        net.float()
        h_torch = net(x)
        net.tc_passes = 3
        h3 = net(x)
        net.tc_passes = 1
        h1 = net(x)
        # the parameter image cache follows in-place updates
        net.nn[0].weight.mul_(1.5)
        net.tc_passes = 0
        ref2 = net.double()(x.double().reshape(4100, -1))
        net.float()
        net.tc_passes = 3
        h3b = net(x)
    e = lambda a, b: rel_err(a.double().cpu().numpy(), b.cpu().numpy())
    print("fc features: torch fp32", e(h_torch, ref), "bf16x3", e(h3, ref), "bf16", e(h1, ref))
    assert not torch.equal(h3, h_torch)
    assert e(h3, ref) < 2e-5 and e(h3b, ref2) < 2e-5
    assert e(h1, ref) < 2e-2


def test_model_with_tensor_core_features_stays_inside_the_gate(monkeypatch):
    """End to end (trajectory_FC_large-shaped: FullyConnected features -> projection -> stack), the feature MLP on the
    tensor cores vs the PyTorch nn.Sequential: z and log-det move by less than the 1e-5 / 2e-5 gates."""
    from bcnf_b200 import feature_tc
    torch.manual_seed(5)
    fnets = [bcnf_b200.ConcatenateCondition(None, 90),
             bcnf_b200.FullyConnectedFeatureNetwork([90, 310, 310, 310, 256], dropout=0.1)]
    model = CondRealNVP_v2(size=19, nested_sizes=[256] * 3, n_blocks=6, n_conditions=256, feature_networks=fnets,
                           dropout=0.3, act_norm=True, precision="bf16x3").to(DEV).eval()
    g = torch.Generator().manual_seed(6)
    rows = 4096
    y, c = torch.randn(rows, 19, generator=g), torch.randn(rows, 30, 3, generator=g)
    with torch.no_grad():
        z_tc = model(y, c, log_det_J=True)
        ld_tc = model.log_det_J
        assert model.feature_network_stack.feature_networks[1].tc_passes == 3
        monkeypatch.setattr(feature_tc, "MIN_ROWS", 1 << 30)           # PyTorch features
        z_pt = model(y, c, log_det_J=True)
        ld_pt = model.log_det_J
    assert not torch.equal(z_tc, z_pt)
    e_z, e_ld = rel_err(z_tc.cpu().numpy(), z_pt.cpu().numpy()), rel_err(ld_tc.cpu().numpy(), ld_pt.cpu().numpy())
    print("tensor-core features vs PyTorch features: z", e_z, "logdet", e_ld)
    assert e_z < 1e-5 and e_ld < 2e-5


def test_lstm_feature_network_on_tensor_cores_matches_pytorch():
    """LSTMFeatureNetwork (2-layer bi-LSTM -> mean over time -> Linear) as one CTA-pair GEMM per layer, direction and time
    step with the cell update in the epilogue (bcnf_b200/feature_tc.py) vs nn.LSTM in fp64: stated tolerance 5e-5 of
    max|ref| for the 3-pass split through a 30-step recurrence and two layers."""
    from bcnf_b200 import feature_tc
    from bcnf_b200.feature_network import LSTMFeatureNetwork
    torch.manual_seed(4)
    net = LSTMFeatureNetwork(input_size=3, hidden_size=140, output_size=1360, num_layers=2, dropout=0.1,
                             bidirectional=True, pooling="mean").to(DEV).eval()
    assert feature_tc.lstm_supported(net)
    x = torch.randn(2100, 30, 3, device=DEV)
    with torch.no_grad():
        ref = net.double()(x.double())
        net.float()
        h_torch = net(x)
        net.tc_passes = 3
        h3 = net(x)
        h3_again = net(x)
    e = lambda a, b: rel_err(a.double().cpu().numpy(), b.cpu().numpy())
    print("lstm features: torch fp32", e(h_torch, ref), "bf16x3", e(h3, ref))
    assert not torch.equal(h3, h_torch)
    assert torch.equal(h3, h3_again)                  # state buffers are reset between calls
    assert e(h3, ref) < 5e-5


def test_transformer_feature_network_on_tensor_cores_matches_pytorch():
    """Transformer encoder of trajectory_TRF_large (8 post-norm blocks, d_model 128, 8 heads, 30 tokens; reference
    feature_network.py:183-307): Linears on the CTA-pair GEMM, attention / add + LayerNorm / embedding as image-producing
    kernels (csrc/trf.cuh), against the PyTorch module in fp64: stated tolerance 2e-5 of max|ref| for the 3-pass split."""
    from bcnf_b200 import feature_tc
    from bcnf_b200.feature_network import Transformer
    torch.manual_seed(8)
    net = Transformer(input_size=3, trf_size=128, n_heads=8, ff_size=128, n_blocks=8, output_size=1360, trf_dropout=0.1,
                      dropout=0.5).to(DEV).eval()
    assert feature_tc.transformer_supported(net)
    with torch.no_grad():                 # LayerNorm gains / shifts away from their (1, 0) initialisation
        for blk in net.layers:
            for ln in (blk.norm1, blk.norm2):
                ln.weight.uniform_(0.5, 1.5)
                ln.bias.uniform_(-0.3, 0.3)
    x = torch.randn(700, 30, 3, device=DEV)
    with torch.no_grad():
        ref = net.double()(x.double())
        net.float()
        h_torch = net(x)
        net.tc_passes = 3
        h3 = net(x)
        h3_again = net(x)
        net.tc_passes = 1
        h1 = net(x)
        # the image cache follows in-place parameter updates
        net.layers[3].attention.k_linear.weight.mul_(1.25)
        net.tc_passes = 0
        ref2 = net.double()(x.double())
        net.float()
        net.tc_passes = 3
        h3b = net(x)
    e = lambda a, b: rel_err(a.double().cpu().numpy(), b.cpu().numpy())
    print("transformer features: torch fp32", e(h_torch, ref), "bf16x3", e(h3, ref), "bf16", e(h1, ref))
    assert not torch.equal(h3, h_torch)
    assert torch.equal(h3, h3_again)
    assert e(h3, ref) < 2e-5 and e(h3b, ref2) < 2e-5
    assert e(h1, ref) < 3e-2


def _encoder(kind):
    if kind == "fc":
        return bcnf_b200.FullyConnectedFeatureNetwork([90, 310, 310, 200], dropout=0.1), (30, 3)
    if kind == "fc_single":
        return bcnf_b200.FullyConnectedFeatureNetwork([90, 200]), (30, 3)
    if kind == "lstm":
        return bcnf_b200.LSTMFeatureNetwork(input_size=3, hidden_size=70, output_size=200, num_layers=2, dropout=0.1,
                                            bidirectional=True, pooling="mean"), (30, 3)
    return bcnf_b200.Transformer(input_size=3, trf_size=64, n_heads=4, ff_size=96, n_blocks=2, output_size=200), (30, 3)


@pytest.mark.parametrize("kind", ["fc", "fc_single", "lstm", "transformer"])
def test_feature_network_fused_with_the_projection(kind, monkeypatch):
    """the encoder's affine output layer and the condition projection as ONE GEMM (P = u (Wproj W_out)^T +
    Wproj b_out + bproj, feature_tc.fused_projection) against the two-step path (h = output layer, then bcnf_cond_project):
    P within 1e-5 of max|P|, z / log-det of the stack within the 1e-5 / 2e-5 gates; follows in-place parameter updates."""
    from bcnf_b200 import feature_tc
    torch.manual_seed(11)
    enc, cshape = _encoder(kind)
    fnets = [bcnf_b200.ConcatenateCondition(None, 90 if kind.startswith("fc") else 3), enc]
    model = CondRealNVP_v2(size=19, nested_sizes=[128] * 3, n_blocks=4, n_conditions=200, feature_networks=fnets,
                           dropout=0.2, act_norm=True, precision="bf16x3").to(DEV).eval()
    g = torch.Generator().manual_seed(12)
    rows = 4200
    y, c = torch.randn(rows, 19, generator=g), torch.randn(rows, *cshape, generator=g)
    e = lambda a, b: rel_err(a.cpu().numpy(), b.cpu().numpy())

    def both():
        with torch.no_grad():
            n0 = feature_tc.N_LAUNCH[0]
            P_f = model._projection(c)
            assert feature_tc.N_LAUNCH[0] > n0
            z_f = model(y, c, log_det_J=True)
            ld_f = model.log_det_J
            monkeypatch.setattr(feature_tc, "FUSE_PROJECTION", False)
            P_2 = model._projection(c)
            z_2, h = model(y, c, log_det_J=True, return_features=True)
            ld_2 = model.log_det_J
            monkeypatch.setattr(feature_tc, "FUSE_PROJECTION", True)
        assert h.shape == (rows, 200) and P_f.shape == P_2.shape and not torch.equal(P_f, P_2)
        print(kind, "fused vs two-step: P", e(P_f, P_2), "z", e(z_f, z_2), "logdet", e(ld_f, ld_2))
        assert e(P_f, P_2) < 1e-5 and e(z_f, z_2) < 1e-5 and e(ld_f, ld_2) < 2e-5

    both()
    with torch.no_grad():                 # the composite matrix is rebuilt after an update of either factor
        last = model.feature_network_stack.feature_networks[-1]
        out_lin = last.nn[-1] if kind.startswith("fc") else (last.linear if kind == "lstm" else last.output)
        out_lin.weight.mul_(0.8)
        coupling = next(l for l in model.layers if isinstance(l, bcnf_b200.ConditionalAffineCouplingLayer))
        coupling.nn_a.linears()[0].weight.mul_(1.1)
    both()
    # below the row threshold the two-step path runs (PyTorch encoder + bcnf_cond_project)
    with torch.no_grad():
        n0 = feature_tc.N_LAUNCH[0]
        model._projection(c[:16])
        assert feature_tc.N_LAUNCH[0] == n0


def test_tensorcore_agrees_with_fp32_kernels_on_golden_large_batch():
    # same weights through the FMA kernel and the 3-pass tensor-core kernel, many tiles per CTA pair
    m32 = _model(19, [128] * 3, 4, 32, "fp32")
    mtc = _model(19, [128] * 3, 4, 32, "bf16x3")
    g = torch.Generator().manual_seed(5)
    n = 50_000
    y = torch.randn(n, 19, generator=g).to(DEV)
    h = torch.randn(n, 32, generator=g).to(DEV)
    za = m32(y, h, log_det_J=True); la = m32.log_det_J
    zb = mtc(y, h, log_det_J=True); lb = mtc.log_det_J
    assert rel_err(zb.cpu().numpy(), za.cpu().numpy()) < 1e-5
    assert rel_err(lb.cpu().numpy(), la.cpu().numpy()) < 1e-5


@pytest.mark.parametrize("precision", ["bf16x3", "bf16"])
@pytest.mark.parametrize("shape", [(19, [526] * 2, 3, 1360, False, 300), (21, [175, 175], 2, 107, True, 129),
                                   (19, [64], 1, 8, False, 5)], ids=["large", "two_way_odd", "tiny"])
def test_tensorcore_projection_matches_torch(shape, precision, monkeypatch):
    # P = h . W1h^T + b1 for every conditioner network (cnf.py:101-104, feature columns of the first Linear)
    size, nested, blocks, n_cond, two_way, n_inst = shape
    model = _model(size, nested, blocks, n_cond, precision, two_way)
    flow = model._flow()
    g = torch.Generator().manual_seed(31)
    h = torch.randn(n_inst, n_cond, generator=g).to(DEV)
    P = flow.project(h)
    monkeypatch.setenv("BCNF_PROJ_FMA", "1")     # (switches are read once, when a handle is created)
    model._packed = None
    P_fma = model._flow().project(h)             # the fp32 FMA projection kernel on a handle of the same parameters
    monkeypatch.delenv("BCNF_PROJ_FMA")
    hp = -(-nested[0] // 16) * 16
    col, da = 0, (size + 1) // 2
    for layer in model.layers:
        if not isinstance(layer, bcnf_b200.ConditionalAffineCouplingLayer):
            continue
        for net, din in ([(layer.nn_a, da), (layer.nn_b, size - da)] if two_way else [(layer.nn_a, da)]):
            lin = net.linears()[0]
            ref = h.double() @ lin.weight.double()[:, din:].t() + lin.bias.double()
            got = P[:, col: col + nested[0]].double()
            scale = ref.abs().max().item()
            tol = 1e-5 if precision == "bf16x3" else 1e-2
            assert (got - ref).abs().max().item() < tol * scale
            assert (P_fma[:, col: col + nested[0]].double() - ref).abs().max().item() < 1e-5 * scale
            assert torch.all(P[:, col + nested[0]: col + hp] == 0)     # padding columns stay exactly zero
            col += hp
    assert col == flow.proj_width


def test_auto_precision_picks_the_kernel_family():
    assert _model(19, [16] * 3, 2, 8, "auto")._flow().kernel == "rowthread"
    assert _model(19, [128] * 3, 2, 8, "auto")._flow().kernel == "tcgen05"
    # width 1024 in the fp32-class 3-pass mode: the second-generation kernel streams activations through L2, so the
    # shared-memory limit of the first generation (which sent this shape to the fp32 tiled kernel) no longer applies
    assert _model(19, [1024] * 2, 2, 8, "auto")._flow().kernel == "tcgen05"
    assert _model(19, [1024] * 2, 2, 8, "bf16")._flow().kernel == "tcgen05"
    assert _model(19, [1024] * 2, 2, 8, "auto")._flow().info.rows_per_cta == 128


@pytest.mark.parametrize("shape", [s for s in SHAPES if max(s[1]) <= 526],
                         ids=lambda s: f"D{s[0]}_H{s[1][0]}x{len(s[1])}_K{s[2]}_C{s[3]}_tw{int(s[4])}_B{s[5]}")
def test_both_kernel_generations_agree(shape, monkeypatch):
    """flow_tc2.cuh (default: 128 rows per CTA, activations through L2) and flow_tc.cuh (BCNF_FLOW_TC=1: 64 rows per CTA,
    activations in shared memory) run the same arithmetic; they differ only where a bias is folded into the GEMM
    (widths that are not multiples of 16), by the rounding of its bf16 hi / lo split."""
    size, nested, blocks, n_cond, two_way, rows = shape
    new = _model(size, nested, blocks, n_cond, "bf16x3", two_way)
    assert new._flow().info.rows_per_cta == 128
    monkeypatch.setenv("BCNF_FLOW_TC", "1")
    old = _model(size, nested, blocks, n_cond, "bf16x3", two_way)
    assert old._flow().info.rows_per_cta == 64
    monkeypatch.delenv("BCNF_FLOW_TC")
    g = torch.Generator().manual_seed(11)
    y = torch.randn(rows, size, generator=g)
    h = torch.randn(rows, n_cond, generator=g)
    for inverse in (False, True):
        fn = (lambda m: m.inverse(y, h)) if inverse else (lambda m: m(y, h, log_det_J=True))
        a, b = fn(new), fn(old)
        if all(w % 16 == 0 for w in nested):
            assert torch.equal(a, b)
            if not inverse:
                assert torch.equal(new.log_det_J, old.log_det_J)
        else:
            assert rel_err(a.cpu().numpy(), b.cpu().numpy()) < 3e-6
