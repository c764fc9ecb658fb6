"""Host-side plan: pyramid schedule, resampling tap tables and pre-filter kernels.

Everything here is O(levels * axis length) bookkeeping; it mirrors the reference's own host code so
that the tables handed to the GPU are the ones the reference would build
(paths relative to /root/reference/src/flowreg3d/):
    level schedule      core/optical_flow_3d.py:77-85, 389-408, 485-490, 517
    tap tables          util/resize_util_3D.py:98-131
    Gaussian pre-filter util/image_processing_3D.py:95-162 (scipy.ndimage.gaussian_filter,
                        mode="reflect", truncate=4)
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib

SWEEP_LEXICOGRAPHIC = 0
SWEEP_REDBLACK = 1


def warping_depth(eta: float, levels: int, p: int, m: int, n: int) -> int:
    """core/optical_flow_3d.py:77-85."""
    min_dim = min(p, m, n)
    depth = 0
    for _ in range(levels):
        depth += 1
        min_dim *= eta
        if round(min_dim) < 10:
            break
    return depth


def level_schedule(shape: Sequence[int], eta: float, levels: int, min_level: int):
    """Level index and grid size per level, coarse -> fine, and the clamped min_level
    (core/optical_flow_3d.py:389-408; python round() = round-half-even, as in the reference)."""
    p, m, n = (int(s) for s in shape)
    mz = warping_depth(eta, levels, p, m, n)
    my = warping_depth(eta, levels, m, n, p)
    mx = warping_depth(eta, levels, n, p, m)
    cap = min(mx, my, mz) * 4
    mz, my, mx = min(mz, cap), min(my, cap), min(mx, cap)
    top = max(mx, my, mz)
    if top <= min_level:
        min_level = top - 1
    if min_level < 0:
        min_level = 0
    sched = []
    for i in range(top, min_level - 1, -1):
        sched.append((i, (int(round(p * eta ** min(i, mz))),
                          int(round(m * eta ** min(i, my))),
                          int(round(n * eta ** min(i, mx))))))
    return sched, min_level


def gauss_taps(sigma: float) -> Tuple[int, np.ndarray]:
    """util/resize_util_3D.py:100-107: the float32 Gaussian of the fused resize -- the same numpy
    expression as the reference, so the taps are bit-identical to its."""
    if sigma <= 0.0:
        return 0, np.array([1.0], dtype=np.float32)
    R = int(np.ceil(2.0 * sigma))
    x = np.arange(-R, R + 1, dtype=np.float32)
    g = np.exp(-0.5 * (x / sigma) ** 2).astype(np.float32)
    g /= g.sum()
    return R, g


def resize_sigma(in_shape, out_shape, sigma_coeff: float = 0.6) -> float:
    """util/resize_util_3D.py:117-131 (per_axis=False)."""
    s = min(out_shape[a] / in_shape[a] for a in (2, 1, 0))
    return (sigma_coeff / s) if s < 1.0 else 0.0


@dataclass
class TableSet:
    """(x, y, z) resampling tables between two grids; keeps the numpy buffers alive."""
    in_shape: Tuple[int, int, int]
    out_shape: Tuple[int, int, int]
    idx: List[np.ndarray] = field(default_factory=list)
    wt: List[np.ndarray] = field(default_factory=list)

    def c_tables(self):
        arr = (_lib.AxisTable * 3)()
        self.fill(arr)
        return arr

    def fill(self, arr):
        for q, ax in enumerate((2, 1, 0)):  # x, y, z
            arr[q].in_len = self.in_shape[ax]
            arr[q].out_len = self.out_shape[ax]
            arr[q].P = self.idx[q].shape[1]
            arr[q].idx = self.idx[q].ctypes.data
            arr[q].wt = self.wt[q].ctypes.data


def make_tables(in_shape, out_shape) -> TableSet:
    lib = _lib.load()
    in_shape = tuple(int(v) for v in in_shape)
    out_shape = tuple(int(v) for v in out_shape)
    R, g = gauss_taps(resize_sigma(in_shape, out_shape))
    ts = TableSet(in_shape, out_shape)
    for ax in (2, 1, 0):
        P = 2 * R + 4
        idx = np.empty((out_shape[ax], P), np.int32)
        wt = np.empty((out_shape[ax], P), np.float32)
        rc = lib.fr3d_fill_resize_table(in_shape[ax], out_shape[ax], g.ctypes.data, R,
                                        idx.ctypes.data, wt.ctypes.data)
        if rc != 0:
            raise ValueError(f"fr3d_fill_resize_table({in_shape[ax]}->{out_shape[ax]}) failed: {rc}")
        ts.idx.append(idx)
        ts.wt.append(wt)
    return ts


def gaussian_half_kernel(sigma: float, truncate: float = 4.0) -> np.ndarray:
    """scipy.ndimage.gaussian_filter1d's kernel (scipy/ndimage/_filters.py _gaussian_kernel1d):
    radius int(truncate*sigma + 0.5); returns w[0..r], w[0] the centre tap.  Axes with
    sigma <= 1e-15 are skipped by scipy -> identity."""
    sigma = float(sigma)
    if sigma <= 1e-15:
        return np.array([1.0])
    r = int(truncate * sigma + 0.5)
    x = np.arange(-r, r + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    phi = phi / phi.sum()
    return np.ascontiguousarray(phi[r:])


def sigma_zyx(sigma, C: int) -> np.ndarray:
    """OFOptions.sigma rows are [sx, sy, sz, st]; the filter wants (sz, sy, sx) per channel
    (util/image_processing_3D.py:117-130, 140-146)."""
    sig = np.asarray(sigma, dtype=float)
    if sig.ndim == 1:
        sig = sig[None]
    out = np.zeros((C, 3))
    for c in range(C):
        row = sig[min(c, len(sig) - 1)]
        out[c] = (row[2], row[1], row[0])
    return out


def sigma_t(sigma, C: int) -> np.ndarray:
    """Temporal sigma per channel (0 when the rows have no 4th entry): 5-D batches are filtered in 4-D
    (util/image_processing_3D.py:140-156)."""
    sig = np.asarray(sigma, dtype=float)
    if sig.ndim == 1:
        sig = sig[None]
    out = np.zeros(C)
    for c in range(C):
        row = sig[min(c, len(sig) - 1)]
        out[c] = float(row[3]) if len(row) >= 4 else 0.0
    return out


@dataclass
class FlowParams:
    """The reference's flow_params dict (compensate_recording_3D.py:301-315) plus the get_displacement
    defaults (core/optical_flow_3d.py:319-333)."""
    alpha: Tuple[float, float, float] = (2.0, 2.0, 2.0)
    update_lag: int = 10
    iterations: int = 20
    min_level: int = 0
    levels: int = 50
    eta: float = 0.8
    a_smooth: float = 0.5
    a_data: object = 0.45


class PlanHolder:
    """Owns the ctypes fr3d_plan and every host buffer it points to."""

    def __init__(self, shape, C_: int, fp: FlowParams, max_batch: int = 1, interp: int = 3,
                 sigma=None, sweep: int = SWEEP_LEXICOGRAPHIC, state_dtype=np.float64):
        Z, Y, X = (int(s) for s in shape)
        self.shape = (Z, Y, X)
        self.C = int(C_)
        if not 1 <= self.C <= _lib.MAX_CHANNELS:
            raise ValueError(f"number of channels must be 1..{_lib.MAX_CHANNELS}, got {self.C}")
        alpha = tuple(float(a) for a in fp.alpha)
        if len(alpha) != 3:
            raise ValueError("alpha must have 3 entries (x, y, z)")
        self.sched, self.min_level = level_schedule(self.shape, fp.eta, fp.levels, fp.min_level)
        self._keep = []
        n = len(self.sched)
        self.levels = (_lib.Level * n)()
        prev = None
        for li, (i, size) in enumerate(self.sched):
            L = self.levels[li]
            for q in range(3):
                L.size[q] = size[q]
                L.h[q] = float(self.shape[q]) / size[q]
            scal = 1 if i == self.min_level else fp.eta ** (-0.5 * i)
            for q in range(3):
                L.alpha[q] = scal * alpha[q]
            L.median = 1 if min(size) > 5 else 0
            t = make_tables(self.shape, size)
            t.fill(L.from_full)
            self._keep.append(t)
            if prev is not None:
                t = make_tables(prev, size)
                t.fill(L.from_prev)
                self._keep.append(t)
            prev = size
        self.plan = _lib.Plan()
        P = self.plan
        P.abi_version = _lib.ABI_VERSION
        P.Z, P.Y, P.X, P.C = Z, Y, X, self.C
        P.max_batch = int(max_batch)
        P.n_levels = n
        P.levels = C.cast(self.levels, C.POINTER(_lib.Level))
        if self.min_level > 0:
            t = make_tables(prev, self.shape)
            t.fill(P.to_full)
            self._keep.append(t)
        P.iterations = int(fp.iterations)
        P.update_lag = int(fp.update_lag)
        ad = np.asarray(fp.a_data, dtype=float).ravel()
        for c in range(_lib.MAX_CHANNELS):
            P.a_data[c] = float(ad[min(c, len(ad) - 1)])
        P.a_smooth = float(fp.a_smooth)
        P.sweep = int(sweep)
        P.interp = int(interp)
        if np.dtype(state_dtype) not in (np.dtype(np.float32), np.dtype(np.float64)):
            raise ValueError("state_dtype must be float32 or float64")
        P.state_dtype = _lib.dtype_code(state_dtype)
        sz = sigma_zyx(sigma, self.C) if sigma is not None else np.zeros((self.C, 3))
        for c in range(self.C):
            for a in range(3):
                w = gaussian_half_kernel(sz[c, a])
                self._keep.append(w)
                P.gauss_radius[c][a] = len(w) - 1
                P.gauss_w[c][a] = w.ctypes.data
        st = sigma_t(sigma, self.C) if sigma is not None else np.zeros(self.C)
        self.temporal = False
        for c in range(self.C):
            w = gaussian_half_kernel(st[c])
            self._keep.append(w)
            P.gauss_radius_t[c] = len(w) - 1
            P.gauss_w_t[c] = w.ctypes.data
            self.temporal = self.temporal or len(w) > 1
