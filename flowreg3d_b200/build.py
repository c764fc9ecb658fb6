"""Build libfr3d.so (hand-written CUDA kernels + C ABI) for sm_100a, in tree.

    python -m flowreg3d_b200.build            # nvcc cross-compiles without a GPU

The shared library lands next to this file (flowreg3d_b200/libfr3d.so); it is git-ignored but
travels to the GPU box with the working tree.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
SRC = HERE / "csrc"
OUT = HERE / "libfr3d.so"

NVCC_FLAGS = [
    "-std=c++17", "-O3",
    "-fmad=false",  # rounding points must match the reference: no implicit FMA contraction
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def sources():
    return sorted(SRC.glob("*.cu"))


def needs_build() -> bool:
    if not OUT.exists():
        return True
    t = OUT.stat().st_mtime
    deps = list(SRC.glob("*")) + [HERE.parent / "include" / "fr3d.h"]
    return any(p.stat().st_mtime > t for p in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return OUT
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", str(OUT), *map(str, sources())]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return OUT


def ensure_built(local_rank: int = 0, timeout_s: float = 600.0) -> Path:
    """Build the library if (and only if) it is MISSING -- e.g. a fresh checkout on a box with nvcc.  Safe under a
    one-process-per-GPU launcher: local rank 0 compiles into a temporary name and renames, the others wait for
    the file.  This builds the CUDA extension; it is not a fallback (without nvcc it raises)."""
    import time
    if OUT.exists():
        return OUT
    if local_rank == 0:
        tmp = OUT.with_suffix(".so.tmp%d" % os.getpid())
        cmd = [_nvcc(), *NVCC_FLAGS, "-o", str(tmp), *map(str, sources())]
        subprocess.check_call(cmd)
        os.replace(tmp, OUT)
        return OUT
    t0 = time.time()
    while not OUT.exists():
        if time.time() - t0 > timeout_s:
            raise RuntimeError(f"{OUT} did not appear within {timeout_s:.0f} s (local rank 0 builds it)")
        time.sleep(0.5)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(OUT)
