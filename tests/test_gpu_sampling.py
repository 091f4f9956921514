"""In-kernel latent draw (bcnf_flow_sample) and the fused calibration ranks (bcnf_flow_sample_ranks), every kernel family.

The reference draws z with torch.randn on the CPU generator (cnf.py:566,578,584) and reduces ranks on the host
(calibration.py:44-48).  Here z is a pure function of (seed, row, j) -- Philox4x32-10 + Box-Muller, pinned by
oracle/philox.py against the Random123 known-answer vectors -- and the comparison with y happens in the output stage.
"""
import numpy as np
import pytest
import torch

import bcnf_b200
from bcnf_b200 import CondRealNVP_v2, compute_y_hat_ranks
from conftest import rel_err
from oracle import philox

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
FAMILIES = [("rowthread", [16, 16], "auto"), ("tiled", [40, 40], "fp32"), ("tcgen05", [144, 144], "bf16x3")]


@pytest.fixture(autouse=True)
def _no_grad():
    with torch.no_grad():
        yield


def _model(nested, precision, n_cond=12, blocks=3):
    torch.manual_seed(0)
    m = CondRealNVP_v2(size=19, nested_sizes=nested, n_blocks=blocks, n_conditions=n_cond,
                       feature_networks=[bcnf_b200.ConcatenateCondition(None, n_cond)], dropout=0.1, act_norm=True,
                       precision=precision)
    return m.to(DEV).eval()


@pytest.mark.parametrize("family,nested,precision", FAMILIES, ids=[f[0] for f in FAMILIES])
def test_in_kernel_latent_is_the_philox_field(family, nested, precision):
    model = _model(nested, precision)
    flow = model._flow()
    assert flow.kernel == family
    n_inst, m, seed, sigma = 23, 31, 0x1234_5678_9ABC, 0.7
    h = torch.randn(n_inst, 12, generator=torch.Generator().manual_seed(2)).to(DEV)
    P = flow.project(h)
    rows = m * n_inst
    x = flow.sample(rows, P, seed=seed, sigma=sigma, inst_period=n_inst)
    z = torch.from_numpy(philox.normal_field(seed, rows, 19, sigma))
    x_ref, _ = flow.run(True, z, P, inst_period=n_inst)
    # logf / sincospif on the device vs numpy's float64 Box-Muller: z agrees to ~1e-6, x to the conditioning of the map
    assert rel_err(x.cpu().numpy(), x_ref.cpu().numpy()) < 2e-5
    # a prefix of the rows is the same draw, whatever the launch size; another seed is another draw
    x_short = flow.sample(7 * n_inst, P, seed=seed, sigma=sigma, inst_period=n_inst)
    assert torch.equal(x_short, x[: 7 * n_inst])
    assert not torch.equal(flow.sample(rows, P, seed=seed + 1, sigma=sigma, inst_period=n_inst), x)


@pytest.mark.parametrize("family,nested,precision", FAMILIES, ids=[f[0] for f in FAMILIES])
def test_fused_ranks_equal_the_explicit_reduction(family, nested, precision):
    model = _model(nested, precision)
    flow = model._flow()
    g = torch.Generator().manual_seed(3)
    n_inst, m = 41, 200
    h = torch.randn(n_inst, 12, generator=g).to(DEV)
    y = (0.7 * torch.randn(n_inst, 19, generator=g)).to(DEV)
    P = flow.project(h)
    rows = m * n_inst
    # injected z: exactly sum_m [x < y] of the explicit samples (calibration.py:48)
    z = torch.randn(rows, 19, generator=g).to(DEV)
    x, _ = flow.run(True, z, P, inst_period=n_inst)
    want = (x.view(m, n_inst, 19) < y.unsqueeze(0)).sum(dim=0).to(torch.int32)
    ranks = torch.zeros(n_inst, 19, dtype=torch.int32, device=DEV)
    flow.sample_ranks(rows, P, y, ranks, z=z, inst_period=n_inst)
    assert torch.equal(ranks, want)
    # drawn z: the same Philox stream through the explicit-sample entry point
    xs = flow.sample(rows, P, seed=77, sigma=1.0, inst_period=n_inst)
    want = (xs.view(m, n_inst, 19) < y.unsqueeze(0)).sum(dim=0).to(torch.int32)
    ranks.zero_()
    flow.sample_ranks(rows, P, y, ranks, seed=77, sigma=1.0, inst_period=n_inst)
    assert torch.equal(ranks, want)
    # counters accumulate over calls (chunks of samples)
    flow.sample_ranks(rows, P, y, ranks, seed=78, sigma=1.0, inst_period=n_inst)
    assert int(ranks.max()) <= 2 * m and int(ranks.sum()) > int(want.sum())


def test_compute_y_hat_ranks_is_calibrated_on_its_own_samples():
    """compute_y_hat_ranks (calibration.py:20-48 signature): for y drawn from the model itself the ranks are uniform
    on [0, M]; the fused path never materialises the (M, N, D) samples."""
    model = _model([16, 16], "auto")
    torch.manual_seed(5)
    n_inst, m = 400, 500
    cond = torch.randn(n_inst, 12)
    y = model.sample(1, cond, outer=True)[0]                       # one posterior draw per instance
    torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.max_memory_allocated()
    ranks = compute_y_hat_ranks(model, y, cond, M_samples=m, device=DEV, verbose=False)
    assert ranks.shape == (n_inst, 19) and ranks.dtype == torch.int64
    assert int(ranks.min()) >= 0 and int(ranks.max()) <= m
    u = ranks.double().numpy() / m
    assert abs(u.mean() - 0.5) < 0.02 and abs(u.std() - (1 / 12) ** 0.5) < 0.02
    assert torch.cuda.max_memory_allocated() - base < m * n_inst * 19 * 4     # less than one (M, N, D) fp32 tensor


def test_sample_device_path_uses_the_kernel_draw_and_matches_statistics():
    model = _model([144, 144], "bf16x3")
    cond = torch.randn(50, 12, generator=torch.Generator().manual_seed(1))
    torch.manual_seed(9)
    a = model.sample(64, cond, outer=True)
    torch.manual_seed(9)
    b = model.sample(64, cond, outer=True)
    assert a.shape == (64, 50, 19) and torch.equal(a, b)           # seeded through torch's CPU generator
    model.sample_rng = "reference"
    torch.manual_seed(9)
    c = model.sample(64, cond, outer=True)
    assert abs(float(a.mean() - c.mean())) < 0.2 and abs(float(a.std() / c.std()) - 1) < 0.2
