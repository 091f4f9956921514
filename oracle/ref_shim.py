"""Live import of the reference (psaegert/bcnf) -- build-container only, test infrastructure.

``/root/reference`` exists in the build container and NOT on the GPU box.  This module is used by
``tests/golden/make_golden.py`` (fixture generation), by not-gpu tests that are skipped when no tree is
present, and by ``bench.py --impl reference`` / its ``cpu_baseline`` leg, which time the reference's own
``CondRealNVP_v2`` from the vendored copy ``oracle/_ref`` (``oracle/build_ref.py``) on the GPU box's host cores.  Recipe from SURVEY.md Appendix A:
``bcnf/__init__.py`` pulls in matplotlib and ``bcnf/utils.py`` imports dynaconf, neither
of which is installed, so the package is registered by hand and dynaconf is stubbed.
"""
from __future__ import annotations

import os
import sys
import types

VENDORED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")   # oracle/build_ref.py: unmodified copy


def _default_root() -> str:
    """/root/reference in the build container; the vendored, byte-identical copy (oracle/_ref) where that is absent
    (the GPU box)."""
    if os.path.isdir(os.path.join("/root/reference", "src", "bcnf")):
        return "/root/reference"
    return VENDORED_ROOT


REFERENCE_ROOT = os.environ.get("BCNF_REFERENCE_ROOT") or _default_root()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "bcnf"))


def import_reference():
    """Return the reference's ``bcnf.models.cnf`` module (unmodified code, imported in place)."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if "bcnf" not in sys.modules or not hasattr(sys.modules["bcnf"], "__path__"):
        pkg = types.ModuleType("bcnf")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "src", "bcnf")]
        sys.modules["bcnf"] = pkg
    if "dynaconf" not in sys.modules:
        dyn = types.ModuleType("dynaconf")
        dyn.Dynaconf = object
        sys.modules["dynaconf"] = dyn
    import bcnf.models.cnf as ref_cnf  # noqa: E402
    return ref_cnf


def reference_config_path(name: str) -> str:
    return os.path.join(REFERENCE_ROOT, "configs", "runs", name)
