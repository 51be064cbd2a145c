"""In-tree build of the CUDA C-ABI library (``doper_b200/_lib/libdoper_b200.so``).

``nvcc`` cross-compiles for sm_100a without a GPU; the built ``.so`` is git-ignored but
travels to the GPU box with the ``gpurun`` snapshot.  Flags that matter for parity:
``-fmad=false`` (no implicit contraction: every fused multiply-add in the kernels is an
explicit ``__fmaf_rn``), IEEE division/sqrt, denormals kept.
"""
from __future__ import annotations

import hashlib
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
OUT_DIR = ROOT / "_lib"
LIB = OUT_DIR / "libdoper_b200.so"
STAMP = OUT_DIR / "build.stamp"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "shared",
]


def _sources():
    return sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + [ROOT.parent / "include" / "doper_b200.h"]


def _digest() -> str:
    h = hashlib.sha256()
    for s in _sources():
        h.update(s.name.encode())
        h.update(s.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    OUT_DIR.mkdir(exist_ok=True)
    dig = _digest()
    if LIB.exists() and STAMP.exists() and STAMP.read_text() == dig and not force:
        return LIB
    cmd = ["nvcc", *NVCC_FLAGS, "-o", str(LIB), str(CSRC / "kernels.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    STAMP.write_text(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
