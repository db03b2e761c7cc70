"""Recipe: makes the UNMODIFIED reference importable on the GPU box.  TEST INFRASTRUCTURE.

The reference is pure Python (+ numba); there is nothing to compile, so the "build" is a verbatim copy of its package
tree from /root/reference/fem into oracle/_ref/fem.  oracle/_ref/ is git-ignored (no reference source enters the
history) but not gpurun-ignored, so it travels to the GPU box with the snapshot, where the `-m gpu` drop-in tests
(tests/test_gpu_dropin_reference.py) run the reference's own Electrodynamics3D.frequency_domain() first on its stock
CPU path (RCM + SuperLU) and then on top of emerge_b200.dropin.install().  The three stub modules the reference needs
in this image (gmsh, numba_progress, pypardiso) are ours and live in oracle/refharness/stubs.

    python oracle/make_ref.py            # copy (idempotent); prints the destination
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference"
DST = os.path.join(HERE, "_ref")


def make_ref(force: bool = False) -> str | None:
    src = os.path.join(SRC, "fem")
    if not os.path.isdir(src):
        return DST if os.path.isdir(os.path.join(DST, "fem")) else None
    dst = os.path.join(DST, "fem")
    if os.path.isdir(dst) and not force:
        same = True
        for root, _, files in os.walk(src):
            for f in files:
                if not f.endswith(".py"):
                    continue
                a = os.path.join(root, f)
                b = os.path.join(dst, os.path.relpath(a, src))
                if not os.path.exists(b) or os.path.getsize(a) != os.path.getsize(b):
                    same = False
        if same:
            return DST
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    os.makedirs(DST, exist_ok=True)
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    for extra in ("LICENSE",):
        if os.path.exists(os.path.join(SRC, extra)):
            shutil.copy2(os.path.join(SRC, extra), os.path.join(DST, extra))
    return DST


if __name__ == "__main__":
    print(make_ref(force="--force" in sys.argv))
