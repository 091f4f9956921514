"""Calibration rank statistics on top of posterior sampling (reference src/bcnf/eval/calibration.py).

Adjacent to the hot path: ``compute_y_hat_ranks`` is the main caller of
``model.sample`` in the reference.  There every (M, N, D) sample tensor is copied to the host and
reduced there; here the reduction ``sum_m [y_hat_m < y]`` is the output stage of the inverse kernel
itself (``bcnf_flow_sample_ranks``: Philox z in, int32 counters out), so the samples are never written
anywhere and only the (N, D) ranks ever leave the GPU.  Same signature and return value as the reference.
"""
from __future__ import annotations

from typing import Any

import numpy as np
import torch

__all__ = ["CDF", "brownian_confidence_interval", "compute_y_hat_ranks", "compute_CDF_residuals"]


def CDF(sorted_array_indices: Any, t: np.ndarray, M: int) -> np.ndarray:
    """Empirical CDF of the ranks at the fractions ``t`` (reference calibration.py:7-10)."""
    ranks = np.asarray(sorted_array_indices)
    return (ranks[:, :, None] <= (t * M)[None, None, :]).sum(axis=0) / ranks.shape[0]


def brownian_confidence_interval(t: np.ndarray) -> np.ndarray:
    """Standard deviation of a Brownian bridge at ``t`` (reference calibration.py:13-17)."""
    return np.sqrt(t * (1 - t))


def compute_y_hat_ranks(model: Any, y: torch.Tensor, *conditions: torch.Tensor, M_samples: int = 10_000,
                        batch_size: int = 100, sample_batch_size: int | None = None, device: Any = "cuda",
                        output_device: Any = "cpu", verbose: bool = True) -> torch.Tensor:
    """ranks[n, d] = #{m : y_hat[m, n, d] < y[n, d]} over M posterior samples (reference calibration.py:20-48).

    With ``model.sample_rng == "reference"`` the reference's loops and CPU-generator draws are replayed
    exactly (same ranks as the reference for the same seed); otherwise instances are processed in chunks
    that are sampled, compared and reduced on the device.
    """
    model.to(device).eval()
    if getattr(model, "sample_rng", "device") == "reference":
        y_hat = model.sample(M_samples, *conditions, outer=True, batch_size=batch_size,
                             sample_batch_size=sample_batch_size, output_device=output_device, verbose=verbose)
        y_out = y.to(output_device)
        y_hat_all = torch.cat([y_hat.to(output_device), y_out.unsqueeze(0)], dim=0)
        return torch.sum(y_hat_all < y_out.unsqueeze(0), dim=0)
    dev = torch.device(model.device)
    n = conditions[0].shape[0]
    flow = model._flow()
    ranks = torch.zeros((n, model.size), dtype=torch.int32, device=dev)
    yd = y.to(device=dev, dtype=torch.float32).contiguous()
    # The comparison and the sum over the M samples happen in the output stage of the inverse kernel
    # (bcnf_flow_sample_ranks): z is drawn in the kernel, the (M, chunk, D) samples are never written, only the
    # (N, D) counters exist in memory.  Chunked over instances so that one launch stays below 2^24 rows.
    chunk = max(1, min(n, (1 << 24) // max(M_samples, 1)))
    with torch.no_grad():
        for b in range(0, n, chunk):
            cs = [c[b: b + chunk] for c in conditions]
            nb = cs[0].shape[0]
            P = model._projection(*cs)
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
            flow.sample_ranks(M_samples * nb, P, yd[b: b + nb], ranks[b: b + nb], seed=seed, sigma=1.0, inst_period=nb)
    return ranks.to(device=output_device, dtype=torch.int64)


def compute_CDF_residuals(y_hat_all_sorted_ranks: torch.Tensor, M_samples: int, t_divisions: int = 100,
                          sigma: float = 1) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Scaled residuals of the rank CDF against the uniform CDF (reference calibration.py:51-71)."""
    n = y_hat_all_sorted_ranks.shape[0]
    t = np.linspace(0, 1, t_divisions)
    residuals = CDF(y_hat_all_sorted_ranks.cpu().numpy(), t, M_samples) - t
    return t, residuals * np.sqrt(n) / sigma, brownian_confidence_interval(t)
