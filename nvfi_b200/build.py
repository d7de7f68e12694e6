"""Build ``libnvfi_b200.so`` (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m nvfi_b200.build [--force] [--verbose]

The library is linked against the static CUDA runtime, has no torch dependency and is
loaded with ctypes by ``nvfi_b200._lib``.  The built ``.so`` is git-ignored but ships to
the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import argparse
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(ROOT, "include")
BUILD_DIR = os.path.join(HERE, "csrc", "build")
LIB_PATH = os.path.join(HERE, "libnvfi_b200.so")
DEBUG_LIB_PATH = os.path.join(HERE, "libnvfi_b200_debug.so")     # development probes (csrc/debug/), tests only

ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
              "-Xptxas", "-v"]
if os.environ.get("NVFI_TIMELINE"):     # development: phase timeline marks in the tensor-core backward
    NVCC_FLAGS.append("-DNVFI_TIMELINE")


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the nvfi_b200 CUDA library cannot be built")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def debug_sources():
    d = os.path.join(CSRC, "debug")
    return sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cu"))


def _deps_mtime() -> float:
    files = sources() + debug_sources() + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    files += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)]
    return max(os.path.getmtime(f) for f in files)


def needs_build() -> bool:
    t = _deps_mtime()
    return any(not os.path.exists(p) or os.path.getmtime(p) < t for p in (LIB_PATH, DEBUG_LIB_PATH))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    nvcc = find_nvcc()
    os.makedirs(BUILD_DIR, exist_ok=True)
    log_path = os.path.join(BUILD_DIR, "ptxas.log")

    def compile_one(src):
        obj = os.path.join(BUILD_DIR, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc, *ARCH_FLAGS, *NVCC_FLAGS, "-I", INCLUDE, "-I", CSRC, "-c", src, "-o", obj]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{p.stdout}\n{p.stderr}")
        return obj, p.stderr

    n_prod = len(sources())
    with ThreadPoolExecutor(max_workers=min(8, n_prod)) as ex:
        results = list(ex.map(compile_one, sources() + debug_sources()))
    objs = [o for o, _ in results[:n_prod]]
    debug_objs = [o for o, _ in results[n_prod:]]
    with open(log_path, "w") as f:
        for _, err in results:
            f.write(err)
    if verbose:
        for _, err in results:
            sys.stderr.write(err)
    for out, oo in ((LIB_PATH, objs), (DEBUG_LIB_PATH, debug_objs)):
        cmd = [nvcc, *ARCH_FLAGS, "-shared", "-o", out, *oo, "-cudart", "static", "-ldl"]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError(f"link failed:\n{p.stdout}\n{p.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(force=a.force, verbose=a.verbose))
