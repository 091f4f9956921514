"""Helpers at the boundary (reference src/bcnf/utils.py): parameter index mapping, the NLL loss
and a YAML run-config loader that needs neither dynaconf nor the reference tree."""
from __future__ import annotations

import os
import re
from typing import Any, Iterator

import numpy as np
import torch

__all__ = ["ParameterIndexMapping", "inn_nll_loss", "load_config", "sub_root_path"]


class ParameterIndexMapping:
    """Name <-> column index of the flow's dimensions (reference utils.py:166-196)."""

    def __init__(self, parameters: list[str]) -> None:
        self.parameters = parameters
        self.map = {name: i for i, name in enumerate(parameters)}

    def __len__(self) -> int:
        return len(self.parameters)

    def vectorize(self, parameter_dict: dict) -> np.ndarray:
        missing = [p for p in self.parameters if p not in parameter_dict]
        if missing:
            raise KeyError(f'Parameter "{missing[0]}" not found in the parameter dictionary. '
                           f"Have available keys: {list(parameter_dict.keys())}")
        return np.array([parameter_dict[p] for p in self.parameters]).T

    def dictify(self, parameter_vector: np.ndarray) -> dict:
        return {p: parameter_vector[i] for i, p in enumerate(self.parameters)}

    def __getitem__(self, key: str) -> int:
        return self.map[key]

    def __iter__(self) -> Iterator[str]:
        return iter(self.parameters)

    def __contains__(self, key: str) -> bool:
        return key in self.map

    def __repr__(self) -> str:
        return str(self.parameters)

    __str__ = __repr__


def inn_nll_loss(z: torch.Tensor, log_det_J: torch.Tensor, reduction: str = "mean") -> torch.Tensor:
    """0.5 * sum z^2 - log|det J|, without the D/2 log(2 pi) constant (reference utils.py:49-53)."""
    per_row = 0.5 * torch.sum(z ** 2, dim=1) - log_det_J
    return torch.mean(per_row) if reduction == "mean" else per_row


def sub_root_path(path: str, root: str | None = None) -> str:
    """Replace ``{{BCNF_ROOT}}`` (reference utils.py:146-163); root defaults to $BCNF_ROOT or cwd."""
    root = root or os.environ.get("BCNF_ROOT", os.getcwd())
    return re.sub(r"{{BCNF_ROOT}}", root, path)


_NUM = re.compile(r"^[+-]?(\d+(_\d+)*)?\.?\d*([eE][+-]?\d+)?$")


def _coerce(v: Any) -> Any:
    # PyYAML (YAML 1.1) reads "2e-4" / "1e-1" as strings where Dynaconf gives floats
    # 
    if isinstance(v, dict):
        return {k: _coerce(x) for k, x in v.items()}
    if isinstance(v, list):
        return [_coerce(x) for x in v]
    if isinstance(v, str) and v and _NUM.match(v) and any(ch.isdigit() for ch in v):
        try:
            return float(v.replace("_", ""))
        except ValueError:
            return v
    return v


def load_config(config_file: str, root: str | None = None) -> dict:
    """Load a ``configs/runs/*.yaml`` file into a plain dict (reference utils.py:13-46).

    Numeric strings are coerced, ``{{BCNF_ROOT}}`` is substituted in the path and in
    ``data.path`` / ``data.config_file``, and ``global.hybrid_weight`` defaults to 0 (the
    reference CLI requires it, __main__.py:66, although the ``old/`` configs omit it).
    """
    import yaml
    config_file = sub_root_path(config_file, root)
    if not os.path.exists(config_file):
        raise FileNotFoundError(f"File '{config_file}' does not exist.")
    with open(config_file) as fh:
        cfg = _coerce(yaml.safe_load(fh))
    data = cfg.get("data")
    if isinstance(data, dict):
        for key in ("path", "config_file"):
            if isinstance(data.get(key), str):
                data[key] = sub_root_path(data[key], root)
    cfg.setdefault("global", {}).setdefault("hybrid_weight", 0)
    return cfg
