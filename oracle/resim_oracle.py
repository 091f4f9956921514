"""CPU restatement (numpy, fp64) of the re-simulation kernel csrc/resim.cuh -- test infrastructure.

Follows src/bcnf/simulation/physics.py: ballistic_ODE :12-50 (element-wise cubes over the norm), the position sum
:150-153 and the impact branch :156-162; the velocity comes from classical RK4 with `substeps` steps per interval where
the reference calls scipy.integrate.odeint (:146).  Pinned against outputs of the reference's own
physics_ODE_simulation in tests/golden/resimulation.npz (tests/test_resimulation.py).
"""
from __future__ import annotations

import numpy as np

NAMES = ["x0_x", "x0_y", "x0_z", "v0_x", "v0_y", "v0_z", "g_x", "g_y", "g_z", "w_x", "w_y", "w_z", "b", "m", "rho", "r",
         "a_x", "a_y", "a_z"]


def simulate(params: np.ndarray, T: float, dt: float, break_on_impact: bool, substeps: int = 16) -> np.ndarray:
    p = np.asarray(params, dtype=np.float64)
    n = p.shape[0]
    x0, v, g, w = p[:, 0:3].copy(), p[:, 3:6].copy(), p[:, 6:9], p[:, 9:12]
    b, m, rho, r, a = p[:, 12], p[:, 13], p[:, 14], p[:, 15], p[:, 16:19]
    k = (0.5 * b / m)[:, None]
    with np.errstate(invalid="ignore", divide="ignore"):
        c = g - g * (rho * (4 / 3) * (np.pi * r ** 3) / m)[:, None] + k * (w ** 3 / np.linalg.norm(w, axis=1, keepdims=True)) + a

        def f(v):
            return c - k * (v ** 3 / np.linalg.norm(v, axis=1, keepdims=True))

        n_steps = len(np.arange(0, T, dt))
        out = np.zeros((n, n_steps, 3))
        out[:, 0] = x0
        x = x0
        landed = np.zeros(n, dtype=bool)
        h = dt / substeps
        for s in range(1, n_steps):
            for _ in range(substeps):
                k1 = f(v); k2 = f(v + 0.5 * h * k1); k3 = f(v + 0.5 * h * k2); k4 = f(v + h * k3)
                v = np.where(landed[:, None], v, v + (h / 6.0) * (k1 + 2 * k2 + 2 * k3 + k4))
            xn = x + v * dt
            if break_on_impact:
                hit = (~landed) & (xn[:, 2] < 0)
                ti = -x[:, 2] / v[:, 2]
                xn = np.where(hit[:, None], x + v * ti[:, None], xn)
                xn = np.where(landed[:, None], x, xn)
                landed = landed | hit
            x = xn
            out[:, s] = x
    return out
