#!/usr/bin/env python
"""Puts the UNMODIFIED reference package where bench.py's reference arm and cpu_baseline leg find it:
/root/reference/src/SQUARNA -> baseline/_ref/SQUARNA (git-ignored, shipped to the GPU box by gpurun).

`pip install --no-index --no-build-isolation --target baseline/_ref /root/reference` is what the contract
names first; it fails in this image because the reference's build backend (hatchling, pyproject.toml:2) is not
installed and there is no index to get it from.  The wheel hatchling would build is exactly the pure-Python package
directory src/SQUARNA (pyproject.toml has no build hooks), so copying that directory gives the same files."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/src/SQUARNA"
DST = os.path.join(ROOT, "baseline", "_ref", "SQUARNA")


def install(force=False):
    if not os.path.isdir(SRC):
        return os.path.isdir(DST)                      # the GPU box: use what travelled with the snapshot
    if os.path.isdir(DST) and not force:
        return True
    os.makedirs(os.path.dirname(DST), exist_ok=True)
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    return True


if __name__ == "__main__":
    ok = install(force="--force" in sys.argv)
    print("baseline/_ref/SQUARNA", "ready" if ok else "missing (no /root/reference here)")
