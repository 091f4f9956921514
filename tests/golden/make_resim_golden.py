"""Generate tests/golden/resimulation.npz from the LIVE reference (build container only).

    python tests/golden/make_resim_golden.py

Seeded parameter sets in the ranges of the reference's data generator are run through the reference's own
physics_ODE_simulation (src/bcnf/simulation/physics.py:53-165: scipy.integrate.odeint + position sum + impact
branch), with and without break_on_impact; parameters and trajectories are stored.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.ref_shim import import_reference  # noqa: E402
from oracle.resim_oracle import NAMES  # noqa: E402


def main():
    import_reference()
    from bcnf.simulation.physics import physics_ODE_simulation
    rng = np.random.default_rng(11)
    n = 48
    P = np.zeros((n, 19))
    P[:, 0:2] = rng.uniform(-5, 5, (n, 2)); P[:, 2] = rng.uniform(0.5, 2.5, n)          # x0
    P[:, 3:6] = rng.uniform(-6, 6, (n, 3)); P[:, 5] = np.abs(P[:, 5]) + 1.0               # v0 (upwards)
    P[:, 6:8] = 0.0; P[:, 8] = -rng.uniform(1.0, 20.0, n)                                 # g
    P[:, 9:12] = rng.uniform(-15, 15, (n, 3))                                             # w
    P[:, 12] = rng.uniform(0.02, 0.3, n); P[:, 13] = rng.uniform(0.3, 2.0, n)             # b, m
    P[:, 14] = rng.uniform(1.0, 1.4, n); P[:, 15] = rng.uniform(0.05, 0.25, n)            # rho, r
    P[:, 16:19] = rng.uniform(-2, 2, (n, 3))                                              # a
    out = {"params": P}
    for T, dt in ((3.0, 1 / 15), (2.0, 0.1)):
        for brk in (False, True):
            X = np.stack([physics_ODE_simulation(**dict(zip(NAMES, row)), T=T, dt=dt, break_on_impact=brk) for row in P])
            out[f"x_T{T}_dt{dt:.4f}_brk{int(brk)}"] = X
            print(T, dt, brk, X.shape, "landed:", int((X[:, -1, 2] <= 1e-9).sum()) if brk else "-")
    np.savez_compressed(os.path.join(HERE, "resimulation.npz"), **out)


if __name__ == "__main__":
    main()
