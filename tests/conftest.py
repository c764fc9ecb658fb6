"""pytest configuration.

Markers:
  gpu -- needs a CUDA device (run on the B200 box: ``pytest -m gpu``); everything else runs on CPU.

The parity tests are written once and parametrised over a ``backend`` fixture:
  "cuda" (marked gpu)  -- the product: libfr3d.so on a B200, called through the C ABI;
  "emu"  (CPU)         -- tests/emu: the same kernel functors compiled by g++ and run serially, so
                          kernel logic and the host driver are checked against the oracle without
                          a GPU.  The emulator is test infrastructure and is never loaded by the
                          package on its own.
"""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (B200)")


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = np.load(GOLDEN / f"{name}.npz")
        return cache[name]
    return load


_current = {"backend": None}


def _select(backend):
    if _current["backend"] == backend:
        return
    from flowreg3d_b200 import _lib
    if backend == "emu":
        from emu.build_emu import build
        _lib._select_for_tests(build())
    else:
        import torch
        assert torch.cuda.is_available(), "gpu-marked test running without a CUDA device"
        from flowreg3d_b200 import build as fr3d_build
        fr3d_build.ensure_built()                 # fresh checkout: compile once (needs nvcc); no-op otherwise
        _lib._select_for_tests(None)
        lib = _lib.load()
        assert not _lib.is_emulator() and str(_lib.library_path()).endswith("libfr3d.so")
        assert lib is not None
    _current["backend"] = backend


@pytest.fixture(params=["emu", pytest.param("cuda", marks=pytest.mark.gpu)])
def backend(request):
    _select(request.param)
    return request.param


@pytest.fixture
def emu_backend():
    _select("emu")
    return "emu"


@pytest.fixture
def cuda_backend():
    _select("cuda")
    return "cuda"


def epe_stats(a, b):
    """Endpoint-error statistics between two (...,3) flow fields."""
    e = np.sqrt(((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2).sum(-1))
    return float(e.mean()), float(e.max())


def rel_l2(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def ulp_diff(a, b):
    a = np.ascontiguousarray(a, np.float32).view(np.int32).astype(np.int64)
    b = np.ascontiguousarray(b, np.float32).view(np.int32).astype(np.int64)
    return np.abs(a - b)
