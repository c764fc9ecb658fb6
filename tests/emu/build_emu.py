"""Builds the kernel-logic EMULATOR of libfr3d (test infrastructure, never shipped).

The CUDA kernels of flowreg3d_b200/csrc are functors executed one item per thread; compiled with
-DFR3D_EMU by g++ the very same functors run as serial loops on the host.  The not-gpu tests use it
to check kernel logic and the host-side driver against the oracle in containers without a GPU.
The product never loads it: flowreg3d_b200/_lib.py loads flowreg3d_b200/libfr3d.so and nothing else; the emulator
is installed by test code calling _lib._select_for_tests() in its own process (no environment variable).
"""
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
OUT = Path(__file__).resolve().parent / "_build" / "libfr3d_emu.so"


def build(force=False):
    src = ROOT / "flowreg3d_b200" / "csrc"
    deps = list(src.glob("*")) + [ROOT / "include" / "fr3d.h"]
    if not force and OUT.exists() and all(d.stat().st_mtime <= OUT.stat().st_mtime for d in deps):
        return OUT
    OUT.parent.mkdir(exist_ok=True)
    subprocess.check_call(["g++", "-x", "c++", "-std=c++17", "-DFR3D_EMU", "-O2", "-ffp-contract=off",
                           "-Wall", "-Wno-unknown-pragmas", "-fPIC", "-shared", "-o", str(OUT),
                           str(src / "fr3d_api.cu")])
    return OUT


if __name__ == "__main__":
    print(build(force=True))
