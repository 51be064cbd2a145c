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
    inc = ROOT.parent / "include"
    return sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + [inc / "doper_b200.h", inc / "doper_fp_policy.h"]


def _digest(extra=()) -> str:
    h = hashlib.sha256()
    for s in _sources():
        h.update(s.name.encode())
        h.update(s.read_bytes())
    h.update(" ".join([*NVCC_FLAGS, *extra]).encode())
    return h.hexdigest()


def lib_path(policy: int | None = None) -> Path:
    """``policy``: floating-point contraction mask (include/doper_fp_policy.h); None = the header's default.
    Other masks build side by side as ``libdoper_b200_p<mask>.so`` (selected with DOPER_B200_LIB)."""
    return LIB if policy is None else OUT_DIR / f"libdoper_b200_p{int(policy)}.so"


def is_current(policy: int | None = None) -> bool:
    """The library of this policy exists and was built from the sources as they are now."""
    extra = [] if policy is None else [f"-DDOPER_FP_POLICY={int(policy)}"]
    stamp = STAMP if policy is None else OUT_DIR / f"build_p{int(policy)}.stamp"
    return lib_path(policy).exists() and stamp.exists() and stamp.read_text() == _digest(extra)


def build(force: bool = False, verbose: bool = False, policy: int | None = None, wait: bool = True):
    OUT_DIR.mkdir(exist_ok=True)
    extra = [] if policy is None else [f"-DDOPER_FP_POLICY={int(policy)}"]
    lib = lib_path(policy)
    stamp = STAMP if policy is None else OUT_DIR / f"build_p{int(policy)}.stamp"
    dig = _digest(extra)
    if lib.exists() and stamp.exists() and stamp.read_text() == dig and not force:
        return lib if wait else (lambda: lib)
    cmd = ["nvcc", *NVCC_FLAGS, *extra, "-o", str(lib), str(CSRC / "kernels.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    if not wait:  # caller overlaps several builds: returns a function that waits for this one
        proc = subprocess.Popen(cmd)

        def finish():
            if proc.wait() != 0:
                raise subprocess.CalledProcessError(proc.returncode, cmd)
            stamp.write_text(dig)
            return lib
        return finish
    subprocess.run(cmd, check=True)
    stamp.write_text(dig)
    return lib


# the non-default contraction mask that is built next to the product library, to show (tests/test_fp_policy_gpu.py)
# that switching the policy is a flag on both sides: POS | LERP | DOT | FORCE
ALT_POLICY = 1 | 4 | 8 | 32


def build_all(force: bool = False):
    """The product library and the alternative-policy build, compiled concurrently."""
    alt = build(force=force, policy=ALT_POLICY, wait=False)
    lib = build(force=force)
    return lib, alt()


if __name__ == "__main__":
    if "--all" in sys.argv:
        print(build_all(force="--force" in sys.argv))
    else:
        print(build(force="--force" in sys.argv, verbose=True))
