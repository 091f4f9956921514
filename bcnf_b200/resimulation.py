"""Re-simulation of posterior samples on the GPU (reference src/bcnf/simulation/resimulation.py:21-59).

Adjacent to the hot path: the downstream consumer of ``model.sample``.  The reference maps
``physics_ODE_simulation`` (physics.py:53-165: scipy ``odeint`` on the velocity, explicit position sum, optional stop on
impact) over the M x N sampled parameter sets with a process pool; here the M x N parameter table is built on the
device and ``bcnf_resimulate`` integrates one trajectory per thread (csrc/resim.cuh).  Same signature and the same
(N, M, steps, 3) float64 return value.
"""
from __future__ import annotations

from typing import Any

import numpy as np
import torch

from . import _cabi

__all__ = ["PHYSICS_PARAMETERS", "physics_ODE_simulation_batch", "resimulate"]

# keyword order of physics_ODE_simulation (physics.py:53-72) = column order of bcnf_resimulate's parameter table
PHYSICS_PARAMETERS = ["x0_x", "x0_y", "x0_z", "v0_x", "v0_y", "v0_z", "g_x", "g_y", "g_z", "w_x", "w_y", "w_z",
                      "b", "m", "rho", "r", "a_x", "a_y", "a_z"]


def physics_ODE_simulation_batch(params: torch.Tensor, T: float = 10.0, dt: float = 0.1, break_on_impact: bool = True,
                                 substeps: int = 16) -> torch.Tensor:
    """``physics_ODE_simulation`` for every row of ``params`` (n, 19; columns = PHYSICS_PARAMETERS) on its CUDA device.

    Returns (n, steps, 3) float64 positions, steps = len(np.arange(0, T, dt)).
    """
    if params.ndim != 2 or params.shape[1] != len(PHYSICS_PARAMETERS):
        raise ValueError(f"expected (n, {len(PHYSICS_PARAMETERS)}) parameters, got {tuple(params.shape)}")
    if params.device.type != "cuda":
        raise RuntimeError("bcnf_b200 re-simulates on CUDA devices only; there is no CPU path")
    p = params.to(torch.float64).contiguous()
    n_steps = len(np.arange(0, T, dt))
    out = torch.empty((p.shape[0], n_steps, 3), dtype=torch.float64, device=p.device)
    _cabi.check(_cabi.lib().bcnf_resimulate(p.data_ptr(), p.shape[0], n_steps, float(dt), int(substeps),
                                            int(bool(break_on_impact)), out.data_ptr(), p.device.index or 0,
                                            torch.cuda.current_stream(p.device).cuda_stream), "bcnf_resimulate")
    return out


def resimulate(model: Any, T: int, dt: float, data_dict: dict[str, list], y_hat: torch.Tensor | None = None,
               *conditions: torch.Tensor, m_samples: int = 1000, break_on_impact: bool = False, n_procs: int | None = None,
               batch_size: int = 100, verbose: bool = True) -> np.ndarray:
    """Reference resimulation.py:21-59.  ``n_procs`` is accepted and ignored (there is no process pool).

    y_hat (M, N, D): sampled values of the parameters the model learned (``model.parameter_index_mapping``); every
    other physics parameter of instance i comes from ``data_dict[name][i]``.  Returns (N, M, steps, 3).
    """
    if y_hat is None:
        if len(conditions) != model.feature_network_stack.n_distinct_conditions:
            raise ValueError(f"Expected {model.feature_network_stack.n_distinct_conditions} conditions, got {len(conditions)}")
        y_hat = model.sample(m_samples, *conditions, batch_size=batch_size, verbose=verbose, outer=True,
                             output_device=model.device)
    dev = torch.device(model.device)
    y_hat = torch.as_tensor(y_hat).to(device=dev, dtype=torch.float64)
    M, N = y_hat.shape[0], y_hat.shape[1]
    if verbose:
        print(f"Resimulating {N} trajectories {M} times")
    mapping = model.parameter_index_mapping
    table = torch.empty((N, M, len(PHYSICS_PARAMETERS)), dtype=torch.float64, device=dev)
    for col, name in enumerate(PHYSICS_PARAMETERS):
        if name in mapping:                                   # learned: one value per (sample, instance)
            table[:, :, col] = y_hat[:, :, mapping[name]].t()
        elif name in data_dict:                               # fixed: one value per instance
            fixed = torch.as_tensor(np.asarray(data_dict[name][:N], dtype=np.float64), device=dev)
            table[:, :, col] = fixed[:, None]
        else:
            raise KeyError(f'physics parameter "{name}" is neither learned by the model nor present in data_dict')
    x = physics_ODE_simulation_batch(table.view(N * M, -1), T=T, dt=dt, break_on_impact=break_on_impact)
    return x.view(N, M, x.shape[1], 3).cpu().numpy()
