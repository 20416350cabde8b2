"""In-tree nvcc build of libbbbp_b200.so (sm_100a only).

The shared library is the product's C ABI (include/bbbp_b200.h).  It is built next to this
file so that it travels with the source tree; nothing is installed into site-packages.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libbbbp_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "--threads", "0",
]


def sources() -> list[str]:
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        glob.glob(os.path.join(PKG_DIR, "..", "include", "*.h"))
    return any(os.path.getmtime(p) > built for p in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu into libbbbp_b200.so; returns the library path."""
    if not force and not is_stale():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libbbbp_b200.so")
    tmp = LIB_PATH + ".tmp"
    # BBBP_NVCC_FLAGS: extra flags for measurement builds (e.g. -DBBBP_CONV1_PF_MERGED=3, the kernels' tuning macros)
    cmd = [nvcc, *NVCC_FLAGS, *os.environ.get("BBBP_NVCC_FLAGS", "").split(), "-o", tmp, *sources(), "-ldl"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + proc.stdout + proc.stderr)
    os.replace(tmp, LIB_PATH)
    if verbose:
        print(proc.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
