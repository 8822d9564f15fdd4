"""In-tree build of ``libgpras_b200.so`` (nvcc, sm_100a only).

The shared library is the product's only compute path; it is built next to this file so that the
snapshot shipped to a GPU box carries it and ``/proc/self/maps`` shows an in-tree ``.so``.
"""

from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libgpras_b200.so"
SOURCES = ["gpras_abi.cu", "sgpr.cu"]
NVCC_FLAGS = [
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-O3",
    "-lineinfo",
    "-std=c++17",
    "-Xcompiler",
    "-fPIC",
    "-shared",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: gpras_b200 has no CPU fallback and cannot be built without CUDA")
    return exe


def needs_build() -> bool:
    if not LIB_PATH.exists():
        return True
    newest = max(p.stat().st_mtime for p in list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [PKG_DIR.parent / "include" / "gpras_b200.h"])
    return newest > LIB_PATH.stat().st_mtime


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source for sm_100a into ``libgpras_b200.so``; returns its path."""
    if not force and not needs_build():
        return LIB_PATH
    srcs = [str(CSRC / s) for s in SOURCES if (CSRC / s).exists()]
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", str(LIB_PATH), *srcs]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
