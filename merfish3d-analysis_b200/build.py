"""Build recipe for libm3d_b200.so (hand-written sm_100a kernels + C ABI).

``python -m merfish3d_analysis_b200.build`` or ``__graft_entry__.build()``.
Objects are compiled in parallel with nvcc and linked in-tree next to this file so the
shared library travels with the repository snapshot.
"""

from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
BUILD_DIR = PKG_DIR / "build"
LIB_PATH = PKG_DIR / "libm3d_b200.so"
SOURCES = ["api.cu", "decode.cu", "lowpass.cu", "extract.cu", "table.cu", "centroid.cu", "upload.cu", "zarrio.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    # the decode arithmetic must not be contracted into FMAs (voxel_math.cuh)
    "-fmad=false",
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; libm3d_b200.so cannot be built")


def _extra_flags() -> list[str]:
    """Experiment hook: extra nvcc flags (e.g. -DM3D_GATE_MINB=4) from $M3D_NVCC_EXTRA."""
    return os.environ.get("M3D_NVCC_EXTRA", "").split()


def _digest() -> str:
    h = hashlib.sha256()
    for p in sorted(CSRC.glob("*.cu*")) + [PKG_DIR.parent / "include" / "m3d_b200.h"]:
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS + _extra_flags()).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = True) -> Path:
    stamp = BUILD_DIR / "digest.txt"
    digest = _digest()
    if not force and LIB_PATH.exists() and stamp.exists() and stamp.read_text() == digest:
        return LIB_PATH
    nvcc = _nvcc()
    BUILD_DIR.mkdir(exist_ok=True)

    def compile_one(src: str) -> Path:
        obj = BUILD_DIR / (Path(src).stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *_extra_flags(), "-c", str(CSRC / src), "-o", str(obj)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", str(LIB_PATH), *map(str, objs), "-lcudart"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    stamp.write_text(digest)
    if verbose:
        print(f"built {LIB_PATH}", file=sys.stderr)
    return LIB_PATH


if __name__ == "__main__":
    build(force="--force" in sys.argv)
