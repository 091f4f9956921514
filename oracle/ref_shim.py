"""Live import of the reference (psaegert/bcnf) -- build-container only, test infrastructure.

``/root/reference`` exists in the build container and NOT on the GPU box, so this module is
used only by ``tests/golden/make_golden.py`` (fixture generation) and by not-gpu tests
that are skipped when the tree is absent.  Recipe from SURVEY.md Appendix A:
``bcnf/__init__.py`` pulls in matplotlib and ``bcnf/utils.py`` imports dynaconf, neither
of which is installed, so the package is registered by hand and dynaconf is stubbed.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("BCNF_REFERENCE_ROOT", "/root/reference")


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
