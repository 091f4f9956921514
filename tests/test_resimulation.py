"""Re-simulation against trajectories recorded from the reference's own physics_ODE_simulation
(scipy odeint; tests/golden/make_resim_golden.py).  Tolerance 2e-7 of max|x|: the reference's LSODA runs at
rtol = atol = 1.49e-8, the fp64 RK4 here (16 substeps) is two orders below that."""
import os
import re

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, rel_err
from oracle import resim_oracle as ro

DATA = np.load(os.path.join(GOLDEN_DIR, "resimulation.npz"))
CASES = [k for k in DATA.files if k.startswith("x_")]


def _args(key):
    T, dt, brk = re.match(r"x_T([0-9.]+)_dt([0-9.]+)_brk([01])", key).groups()
    dt = 1 / 15 if abs(float(dt) - 0.0667) < 1e-3 else float(dt)
    return float(T), dt, bool(int(brk))


@pytest.mark.parametrize("key", CASES)
def test_oracle_matches_the_reference_trajectories(key):
    T, dt, brk = _args(key)
    x = ro.simulate(DATA["params"], T, dt, brk)
    assert x.shape == DATA[key].shape
    assert rel_err(x, DATA[key]) < 2e-7


@pytest.mark.gpu
@pytest.mark.parametrize("key", CASES)
def test_kernel_matches_the_reference_trajectories(key):
    from bcnf_b200.resimulation import physics_ODE_simulation_batch
    T, dt, brk = _args(key)
    p = torch.from_numpy(DATA["params"]).to("cuda:0")
    x = physics_ODE_simulation_batch(p, T=T, dt=dt, break_on_impact=brk)
    assert x.dtype == torch.float64 and tuple(x.shape) == DATA[key].shape
    assert rel_err(x.cpu().numpy(), DATA[key]) < 2e-7
    assert rel_err(x.cpu().numpy(), ro.simulate(DATA["params"], T, dt, brk)) < 1e-12      # same arithmetic as the oracle
    if brk:   # after the impact the object stays where it landed (physics.py:160)
        xr = x.cpu().numpy()
        landed = xr[:, -1, 2] <= 1e-9
        assert landed.sum() == (DATA[key][:, -1, 2] <= 1e-9).sum()
        assert np.all(xr[landed, -1] == xr[landed, -2])


@pytest.mark.gpu
def test_resimulate_has_the_reference_signature_and_layout():
    """resimulate(model, T, dt, data_dict, y_hat) -> (N, M, steps, 3): learned parameters from y_hat, the rest from
    data_dict, exactly as resimulation.py:50 assembles the keyword arguments."""
    import bcnf_b200
    from bcnf_b200.resimulation import PHYSICS_PARAMETERS, resimulate
    learned = ["v0_x", "v0_y", "v0_z", "w_x", "b"]
    torch.manual_seed(0)
    model = bcnf_b200.CondRealNVP_v2(size=len(learned), nested_sizes=[16, 16], n_blocks=2, n_conditions=4,
                                     feature_networks=[bcnf_b200.ConcatenateCondition(None, 4)], act_norm=True,
                                     parameter_index_mapping=bcnf_b200.ParameterIndexMapping(learned)).to("cuda:0").eval()
    P = DATA["params"]
    N, M = 5, 7
    rng = np.random.default_rng(0)
    y_hat = np.stack([np.stack([P[rng.integers(len(P))][[PHYSICS_PARAMETERS.index(n) for n in learned]] for _ in range(N)])
                      for _ in range(M)])                                   # (M, N, D)
    data_dict = {n: list(P[:N, PHYSICS_PARAMETERS.index(n)]) for n in PHYSICS_PARAMETERS}
    X = resimulate(model, 2, 0.1, data_dict, torch.from_numpy(y_hat), break_on_impact=True, verbose=False)
    assert X.shape == (N, M, 20, 3) and X.dtype == np.float64
    for i in range(N):
        for j in range(M):
            row = P[i].copy()
            for d, n in enumerate(learned):
                row[PHYSICS_PARAMETERS.index(n)] = y_hat[j, i, d]
            assert rel_err(X[i, j], ro.simulate(row[None], 2, 0.1, True)[0]) < 1e-12
    with pytest.raises(KeyError):
        resimulate(model, 2, 0.1, {"m": [1.0] * N}, torch.from_numpy(y_hat), verbose=False)
