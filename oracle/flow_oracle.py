"""CPU oracle for the CondRealNVP_v2 coupling stack -- TEST INFRASTRUCTURE ONLY.

This file restates, on the CPU, the arithmetic of the reference hot path
(psaegert/bcnf, ``src/bcnf/models/cnf.py`` and ``src/bcnf/utils.py``).  It is the
checker that the CUDA path in ``bcnf_b200/csrc`` is compared against.  Nothing in
``bcnf_b200/`` imports it; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may.

Parity pin: the reference ships no golden vectors for this path (its only test
file does not import), so the oracle is pinned against outputs of
the reference itself, generated in the build container by
``tests/golden/make_golden.py`` (live import of ``/root/reference``) and committed
as ``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` checks every function
below against those fixtures.

Two array back ends implement the same statements:

* ``numpy`` (default) -- fp32 or fp64 according to the inputs; erf from
  ``scipy.special``.  This is the checker used by the tests.
* ``torch`` (CPU) -- the same statements on ATen CPU kernels, i.e. the very
  GEMM / erf / tanh / exp implementations the reference's own eager path runs
  on a host.  ``bench.py`` times this one as the CPU baseline.

Parameters are passed in the reference's ``state_dict`` layout (``weight`` is
``(out, in)`` row-major, as ``torch.nn.Linear`` stores it), parsed by
:func:`layers_from_state_dict`.
"""
from __future__ import annotations

import math
import re
from dataclasses import dataclass, field
from typing import Any, Sequence

import numpy as np

__all__ = [
    "ActNormP", "CouplingP", "OrthoP", "layers_from_state_dict",
    "conditioner", "coupling_forward", "coupling_inverse",
    "actnorm_forward", "actnorm_inverse", "ortho_forward", "ortho_inverse",
    "stack_forward", "stack_inverse", "inn_nll", "log_prob", "sample_outer_rows",
    "macs_per_row",
]


# --------------------------------------------------------------------------------------
# array namespace shim (numpy | torch-cpu)
# --------------------------------------------------------------------------------------
class _NP:
    name = "numpy"

    @staticmethod
    def cat(xs, axis):
        return np.concatenate(xs, axis=axis)

    @staticmethod
    def erf(x):
        from scipy.special import erf  # float32 in -> float32 out
        return erf(x)

    tanh = staticmethod(np.tanh)
    exp = staticmethod(np.exp)
    log = staticmethod(np.log)
    abs = staticmethod(np.abs)

    @staticmethod
    def sum_last(x):
        return x.sum(axis=-1, dtype=x.dtype)

    @staticmethod
    def zeros(n, like):
        return np.zeros(n, dtype=like.dtype)

    @staticmethod
    def linear(x, w, b):
        return x @ w.T + b

    @staticmethod
    def gelu(x):
        # nn.GELU() default = exact erf form (factories.py:65-66 -> torch.nn.GELU)
        c = x.dtype.type(0.7071067811865476)
        half = x.dtype.type(0.5)
        one = x.dtype.type(1.0)
        return half * x * (one + _NP.erf(x * c))


class _TH:
    name = "torch"

    @staticmethod
    def cat(xs, axis):
        import torch
        return torch.cat(xs, dim=axis)

    @staticmethod
    def tanh(x):
        return x.tanh()

    @staticmethod
    def exp(x):
        return x.exp()

    @staticmethod
    def log(x):
        return x.log()

    @staticmethod
    def abs(x):
        return x.abs()

    @staticmethod
    def sum_last(x):
        return x.sum(dim=-1)

    @staticmethod
    def zeros(n, like):
        import torch
        return torch.zeros(n, dtype=like.dtype)

    @staticmethod
    def linear(x, w, b):
        import torch
        return torch.addmm(b, x, w.t())

    @staticmethod
    def gelu(x):
        import torch
        return torch.nn.functional.gelu(x)


def _ns(x):
    return _NP if isinstance(x, np.ndarray) else _TH


# --------------------------------------------------------------------------------------
# parameter containers
# --------------------------------------------------------------------------------------
@dataclass
class ActNormP:
    scale: Any
    bias: Any


@dataclass
class OrthoP:
    q: Any  # (D, D); forward is y @ q


@dataclass
class CouplingP:
    nn_a: list = field(default_factory=list)  # [(W(out,in), b(out)), ...]
    nn_b: list | None = None                  # two_way only


def layers_from_state_dict(sd: dict, prefix: str = "layers.", convert=None) -> list:
    """Group ``layers.{i}.*`` entries of a reference state_dict into layer records.

    Key layout (probed on the reference): ``layers.{i}.scale`` /
    ``.bias`` (ActNorm, cnf.py:345-346), ``layers.{i}.nn_a.nn.{j}.weight`` / ``.bias``
    (conditioner Linear modules inside the Sequential, cnf.py:64-85; ``j`` skips
    activation and dropout slots), ``layers.{i}.orthonormal_matrix`` (cnf.py:323).
    """
    if convert is None:
        def convert(v):
            return v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)
    by_idx: dict[int, dict[str, Any]] = {}
    pat = re.compile(re.escape(prefix) + r"(\d+)\.(.+)$")
    for k, v in sd.items():
        m = pat.match(k)
        if m:
            by_idx.setdefault(int(m.group(1)), {})[m.group(2)] = convert(v)
    layers = []
    for i in sorted(by_idx):
        ent = by_idx[i]
        if "orthonormal_matrix" in ent:
            layers.append(OrthoP(ent["orthonormal_matrix"]))
        elif "scale" in ent:
            layers.append(ActNormP(ent["scale"], ent["bias"]))
        else:
            def net(tag):
                js = sorted({int(k.split(".")[2]) for k in ent if k.startswith(tag + ".nn.")})
                return [(ent[f"{tag}.nn.{j}.weight"], ent[f"{tag}.nn.{j}.bias"]) for j in js]
            a = net("nn_a")
            b = net("nn_b")
            layers.append(CouplingP(a, b if b else None))
    return layers


# --------------------------------------------------------------------------------------
# the path, statement by statement
# --------------------------------------------------------------------------------------
def conditioner(y_half, h, lin: Sequence):
    """ConditionalNestedNeuralNetwork.forward, cnf.py:98-107 (eval mode: Dropout is identity).

    u = cat([y_half, h]); hidden = GELU(Linear(.)) for all but the last Linear
    (cnf.py:78-85); t, s = chunk(out, 2); returns (t, tanh(s)).
    """
    xp = _ns(y_half)
    u = xp.cat([y_half, h], 1)                         # cnf.py:101
    for w, b in lin[:-1]:
        u = xp.gelu(xp.linear(u, w, b))                # cnf.py:79-80
    w, b = lin[-1]
    o = xp.linear(u, w, b)                             # cnf.py:85
    half = o.shape[1] // 2
    t, s = o[:, :half], o[:, half:]                    # cnf.py:104 (chunk(2, dim=1))
    return t, xp.tanh(s)                               # cnf.py:107


def _split(y):
    # torch.chunk(2, dim=-1): first chunk has ceil(D/2) columns (cnf.py:175, :200)
    d = y.shape[-1]
    da = (d + 1) // 2
    return y[:, :da], y[:, da:]


def coupling_forward(p: CouplingP, y, h):
    """ConditionalAffineCouplingLayer.forward, cnf.py:165-196. Returns (z, log_det_row)."""
    xp = _ns(y)
    y_a, y_b = _split(y)
    t_a, ls_a = conditioner(y_a, h, p.nn_a)            # cnf.py:178
    z_b = xp.exp(ls_a) * y_b + t_a                     # cnf.py:179
    ld = xp.sum_last(ls_a)                             # cnf.py:190
    if p.nn_b is not None:
        t_b, ls_b = conditioner(z_b, h, p.nn_b)        # cnf.py:183
        z_a = xp.exp(ls_b) * y_a + t_b                 # cnf.py:184
        ld = ld + xp.sum_last(ls_b)                    # cnf.py:193
    else:
        z_a = y_a                                      # cnf.py:186
    return xp.cat([z_a, z_b], -1), ld                  # cnf.py:196


def coupling_inverse(p: CouplingP, z, h):
    """ConditionalAffineCouplingLayer.inverse, cnf.py:198-213.

    Note (reproduced on purpose): for ``two_way`` the reference conditions nn_a on
    ``z_a`` (cnf.py:203) although forward conditioned it on ``y_a``; the two-way layer is
    therefore not an exact inverse of its forward.  One-way layers are exact.
    """
    xp = _ns(z)
    z_a, z_b = _split(z)
    t_a, ls_a = conditioner(z_a, h, p.nn_a)            # cnf.py:203
    y_b = (z_b - t_a) * xp.exp(-ls_a)                  # cnf.py:204
    if p.nn_b is not None:
        t_b, ls_b = conditioner(y_b, h, p.nn_b)        # cnf.py:207
        y_a = (z_a - t_b) * xp.exp(-ls_b)              # cnf.py:208
    else:
        y_a = z_a                                      # cnf.py:210
    return xp.cat([y_a, y_b], -1)                      # cnf.py:213


def actnorm_forward(p: ActNormP, x):
    """ActNorm.forward, cnf.py:348-351. Returns (z, scalar log-det)."""
    xp = _ns(x)
    z = p.scale * x + p.bias                           # cnf.py:349
    ld = xp.log(xp.abs(p.scale)).sum()                 # cnf.py:350
    return z, ld


def actnorm_inverse(p: ActNormP, z):
    return (z - p.bias) / p.scale                      # cnf.py:354


def ortho_forward(p: OrthoP, y):
    return y @ p.q                                     # cnf.py:335


def ortho_inverse(p: OrthoP, z):
    return z @ p.q.T                                   # cnf.py:339


def stack_forward(layers: Sequence, y, h):
    """Layer loop of CondRealNVP_v2.forward, cnf.py:476-488, given features ``h``.

    The log-det accumulator starts at zeros(B) and receives each layer's log-det in
    layer order (cnf.py:477, :487-488); OrthonormalTransformation contributes 0.
    """
    xp = _ns(y)
    ld = xp.zeros(y.shape[0], y)
    for p in layers:
        if isinstance(p, ActNormP):
            y, l = actnorm_forward(p, y)
            ld = ld + l
        elif isinstance(p, CouplingP):
            y, l = coupling_forward(p, y, h)
            ld = ld + l
        elif isinstance(p, OrthoP):
            y = ortho_forward(p, y)
        else:
            raise ValueError(f"unknown layer record {type(p)}")   # cnf.py:485
    return y, ld


def stack_inverse(layers: Sequence, z, h):
    """Layer loop of CondRealNVP_v2.inverse, cnf.py:500-506."""
    for p in reversed(layers):
        if isinstance(p, ActNormP):
            z = actnorm_inverse(p, z)
        elif isinstance(p, CouplingP):
            z = coupling_inverse(p, z, h)
        elif isinstance(p, OrthoP):
            z = ortho_inverse(p, z)
        else:
            raise ValueError(f"unknown layer record {type(p)}")   # cnf.py:506
    return z


def inn_nll(z, log_det, reduction: str = "mean"):
    """inn_nll_loss, utils.py:49-53 (no D/2 log 2 pi term)."""
    per_row = 0.5 * (z ** 2).sum(-1) - log_det
    return per_row.mean() if reduction == "mean" else per_row


def log_prob(z, log_det):
    """New convenience: -nll_row - D/2 log(2 pi)."""
    d = z.shape[-1]
    return -inn_nll(z, log_det, reduction="none") - 0.5 * d * math.log(2.0 * math.pi)


def sample_outer_rows(n_samples: int, n_inst: int):
    """Row -> instance map of ``_sample(outer=True)``, cnf.py:578-582.

    ``c.repeat(m, 1, ...)`` tiles the instance axis m times, so row r belongs to sample
    r // n_inst and instance r % n_inst, and the result is viewed (m, n_inst, D).
    """
    return np.arange(n_samples * n_inst, dtype=np.int64) % n_inst


def macs_per_row(size: int, nested: Sequence[int], n_blocks: int, n_cond: int,
                 hoisted: bool, act_norm: bool = True) -> int:
    """Algorithmic multiply-accumulates per row."""
    da, db = (size + 1) // 2, size // 2
    first = da if hoisted else da + n_cond
    per_block = first * nested[0] + sum(a * b for a, b in zip(nested[:-1], nested[1:])) + nested[-1] * 2 * db
    mix = (n_blocks - 1) * (size * size + (size if act_norm else 0))
    return n_blocks * per_block + mix
