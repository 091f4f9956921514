"""Vendor the UNMODIFIED Python reference into oracle/_ref (build container only; test / baseline infrastructure).

    python oracle/build_ref.py

The reference (psaegert/bcnf) is pure Python: "building" it is copying the package tree byte for byte, from where it
lies under /root/reference, into oracle/_ref/src/bcnf -- the layout oracle/ref_shim.py expects when
BCNF_REFERENCE_ROOT points at oracle/_ref.  oracle/_ref is git-ignored (no reference source ever enters history) but
travels to the GPU box with the snapshot, where `bench.py --impl reference` times the reference's own
CondRealNVP_v2 (src/bcnf/models/cnf.py) on the host cores and reports ``cpu_baseline.kind: "reference"``.
A MANIFEST of sha256 sums is written next to the copy so that "unmodified" can be checked.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC_ROOT = os.environ.get("BCNF_REFERENCE_SRC", "/root/reference")
DST_ROOT = os.path.join(HERE, "_ref")


def build(verbose: bool = True) -> str | None:
    src = os.path.join(SRC_ROOT, "src", "bcnf")
    if not os.path.isdir(src):
        if verbose:
            print(f"oracle/build_ref.py: {src} not present (GPU box?): keeping whatever oracle/_ref holds")
        return None
    dst = os.path.join(DST_ROOT, "src", "bcnf")
    if os.path.isdir(DST_ROOT):
        shutil.rmtree(DST_ROOT)
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    lines = []
    for root, _, files in sorted(os.walk(dst)):
        for name in sorted(files):
            path = os.path.join(root, name)
            rel = os.path.relpath(path, dst)
            a = hashlib.sha256(open(path, "rb").read()).hexdigest()
            b = hashlib.sha256(open(os.path.join(src, rel), "rb").read()).hexdigest()
            assert a == b, rel
            lines.append(f"{a}  src/bcnf/{rel}")
    with open(os.path.join(DST_ROOT, "MANIFEST.sha256"), "w") as f:
        f.write("\n".join(lines) + "\n")
    if verbose:
        print(f"oracle/_ref: {len(lines)} files copied unmodified from {src}")
    return DST_ROOT


if __name__ == "__main__":
    sys.exit(0 if build() or True else 1)
