"""Host-side costs that bound the one-call array API (compensate_arr_3D with pageable numpy in / out):
pinned allocation, first touch of fresh pageable memory, pinned -> pageable memcpy, float32 -> float64 astype,
pageable vs pinned H2D / D2H.  Prints GB/s for each."""
import time
import numpy as np
import torch

GB = 1 << 30
n = 2 * GB


def t(f, nbytes, name):
    t0 = time.perf_counter()
    r = f()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{name:55s} {nbytes / GB / dt:7.2f} GB/s  ({dt * 1e3:8.1f} ms for {nbytes / GB:.1f} GB)", flush=True)
    return r


torch.cuda.init()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
pin = t(lambda: torch.empty(n, dtype=torch.uint8, pin_memory=True), n, "cudaHostAlloc (torch.empty(pin_memory=True))")
pin2 = t(lambda: torch.empty(n, dtype=torch.uint8).pin_memory(), n, "torch.empty().pin_memory() (alloc + copy)")
fresh = np.empty(n, np.uint8)
t(lambda: fresh.fill(1), n, "first touch of fresh pageable memory (fill)")
t(lambda: fresh.fill(2), n, "second touch (fill)")
fresh2 = np.empty(n, np.uint8)
t(lambda: np.copyto(fresh2, pin.numpy()), n, "memcpy pinned -> fresh pageable numpy")
t(lambda: np.copyto(fresh2, pin.numpy()), n, "memcpy pinned -> touched pageable numpy")
f32 = np.ones(n // 8, np.float32)
t(lambda: f32.astype(np.float64), n // 2 + n // 4, "astype float32 -> float64 (bytes read + written)")
t(lambda: d.copy_(pin, non_blocking=True), n, "H2D from pinned")
t(lambda: d.copy_(torch.from_numpy(fresh2)), n, "H2D from pageable")
t(lambda: pin.copy_(d, non_blocking=True), n, "D2H to pinned")
out = torch.from_numpy(fresh)
t(lambda: out.copy_(d), n, "D2H to pageable")
reg = np.empty(n, np.uint8)
t(lambda: torch.cuda.cudart().cudaHostRegister(reg.ctypes.data, n, 0), n, "cudaHostRegister of fresh pageable memory")
t(lambda: torch.from_numpy(reg).copy_(d, non_blocking=True), n, "D2H to registered memory")
torch.cuda.cudart().cudaHostUnregister(reg.ctypes.data)
