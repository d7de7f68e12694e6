#!/usr/bin/env python
""""Install" the unmodified reference for the CPU arm of bench.py: copy its pure-Python packages
(models/, utils/, config/, datasets/) and its training driver train_nvfi.py from the read-only checkout into baseline/_ref/ (git-ignored; shipped to the
GPU box by gpurun).  The reference has no setup.py / pyproject, so pip has nothing to build:

    python tools/install_reference.py [/root/reference]
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
dst = os.path.join(ROOT, "baseline", "_ref")
if not os.path.isfile(os.path.join(src, "models", "nvfi.py")):
    sys.exit(f"no reference checkout at {src}")
os.makedirs(dst, exist_ok=True)
for d in ("models", "utils", "config", "datasets"):
    t = os.path.join(dst, d)
    if os.path.exists(t):
        shutil.rmtree(t)
    shutil.copytree(os.path.join(src, d), t, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", ".DS_Store"))
for f in ("train_nvfi.py", "test_transfer_vel.py"):      # the drivers tests/test_gpu_dropin.py runs UNMODIFIED
    shutil.copy2(os.path.join(src, f), os.path.join(dst, f))
print("installed", sorted(os.listdir(dst)), "->", dst)
