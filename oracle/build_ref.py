"""Recipe for ``oracle/_ref``: a verbatim copy of the hot-path part of the reference tree, made from the sources where
they lie under ``/root/reference`` (read-only there).

TEST / BASELINE INFRASTRUCTURE ONLY.  The reference is pure Python, so "building" it is a copy; the copy is git-ignored
(no reference source enters this repository's history) but NOT gpurun-ignored, so it travels to the GPU box, where
``/root/reference`` does not exist.  There it serves
  * the ``-m gpu`` drop-in tests (the reference's unmodified model.py / main.py driving the CUDA kernels),
  * ``bench.py --impl reference`` and the ``cpu_baseline`` / ``torch_gpu_baseline`` legs (kind "reference").
Nothing in ``insr_pde_b200/`` reads it.

    python oracle/build_ref.py            # (re)create oracle/_ref from /root/reference
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SOURCE = os.environ.get("INSR_REFERENCE_SOURCE", "/root/reference")
# the packages main.py imports for the three PDE families, its entry point and configuration, the example scripts
# (argument lists of the BASELINE configs) and the two meshes shipped with the tree
PARTS = ("base", "advection", "fluid", "elasticity", "scripts", "main.py", "config.py", "recap.py", "README.md")


def build(force: bool = False) -> str | None:
    """copy PARTS of the reference into oracle/_ref; returns the path, or None when the reference tree is absent
    (on the GPU box: the shipped copy is used as it is)"""
    if not os.path.isfile(os.path.join(SOURCE, "base", "networks.py")):
        return DEST if os.path.isfile(os.path.join(DEST, "base", "networks.py")) else None
    stamp = os.path.join(DEST, ".copied_from")
    if not force and os.path.isfile(stamp) and open(stamp).read().strip() == SOURCE:
        return DEST
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    os.makedirs(DEST)
    for part in PARTS:
        src, dst = os.path.join(SOURCE, part), os.path.join(DEST, part)
        if os.path.isdir(src):
            shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        elif os.path.isfile(src):
            shutil.copy(src, dst)
    for root, dirs, files in os.walk(DEST):          # the source tree is read-only; the copy must be removable
        for name in dirs + files:
            os.chmod(os.path.join(root, name), 0o755 if name in dirs else 0o644)
    with open(stamp, "w") as fh:
        fh.write(SOURCE + "\n")
    return DEST


if __name__ == "__main__":
    out = build(force="--force" in sys.argv)
    print(out if out else f"no reference tree at {SOURCE} and no shipped copy at {DEST}")
