"""Device-memory plumbing: PyTorch tensors own the buffers, libfr3d gets raw pointers.

torch is used for allocation, host<->device copies and streams only -- no torch op touches the
data.  Under the test-suite's kernel-logic emulator (installed by tests/conftest.py through
_lib._select_for_tests; CPU-only containers) the same host code runs with CPU tensors.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

_TORCH_DT = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64,
             np.dtype(np.uint8): torch.uint8, np.dtype(np.uint16): torch.uint16,
             np.dtype(np.int16): torch.int16, np.dtype(np.int32): torch.int32,
             np.dtype(np.int64): torch.int64}


def torch_dtype(dt):
    return _TORCH_DT[np.dtype(dt)]


def default_device(index=None) -> torch.device:
    if _lib.is_emulator():
        return torch.device("cpu")
    if not torch.cuda.is_available():
        raise RuntimeError("flowreg3d_b200 needs a CUDA device (B200, sm_100a); none is visible and "
                           "there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device() if index is None else index)


def to_device(a: np.ndarray, device: torch.device, pin: bool = False) -> torch.Tensor:
    a = np.ascontiguousarray(a)
    if not a.flags.writeable:
        a = a.copy()
    t = torch.from_numpy(a)
    if device.type == "cpu":
        return t
    if pin:
        t = t.pin_memory()
    return t.to(device, non_blocking=pin)


def empty(shape, dtype, device: torch.device) -> torch.Tensor:
    return torch.empty(tuple(int(s) for s in shape), dtype=torch_dtype(dtype), device=device)


def to_host(t: torch.Tensor) -> np.ndarray:
    return t.cpu().numpy() if t.device.type != "cpu" else t.numpy()


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def bind_host_to_gpu(device: torch.device) -> list:
    """Pin the calling process to the CPU cores NVML reports as local to `device` (same NUMA node / PCIe root),
    so that pinned staging buffers allocated afterwards are first-touched next to the GPU.  One process per GPU
    launchers (torchrun) do not do this.  Returns the cores chosen ([] when nothing was changed)."""
    import os
    if device.type != "cuda" or not hasattr(os, "sched_setaffinity"):
        return []
    try:
        import pynvml
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(device).uuid)
        h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid if not uuid.startswith("GPU-") else uuid).encode())
        ncpu = os.cpu_count() or 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        local = {i for i in range(ncpu) if (int(mask[i // 64]) >> (i % 64)) & 1}
        cpus = sorted(local & set(os.sched_getaffinity(0)))
        if cpus and len(cpus) < len(os.sched_getaffinity(0)):
            os.sched_setaffinity(0, cpus)
            return cpus
    except Exception:
        pass
    return []
