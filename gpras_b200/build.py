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
SOURCES = ["gpras_abi.cu", "sgpr.cu", "pre.cu", "kmeans.cu"]
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


def _compile(src: Path, obj: Path, verbose: bool) -> str:
    cmd = [_nvcc(), *[f for f in NVCC_FLAGS if f != "-shared"], "-c", "-o", str(obj), str(src)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    return res.stderr


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source for sm_100a (one object per translation unit, in parallel) and link
    ``libgpras_b200.so``; returns its path."""
    if not force and not needs_build():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor

    srcs = [CSRC / s for s in SOURCES if (CSRC / s).exists()]
    objs = [CSRC / (s.stem + ".o") for s in srcs]
    with ThreadPoolExecutor(max_workers=len(srcs)) as ex:
        logs = list(ex.map(lambda so: _compile(so[0], so[1], verbose), zip(srcs, objs)))
    cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB_PATH), *map(str, objs)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print("\n".join(logs))
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
