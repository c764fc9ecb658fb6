"""Rigid cross-correlation pre-alignment on the B200 (OFOptions.cc_initialization).

Host side of `estimate_rigid_xcorr_3d` (util/xcorr_prealignment.py:15-99) for B frames at a time:
mean projections, 2-D fused Gauss-cubic resize, whitening + Hann window and every transform run in
the kernels of csrc/fr3d_xcorr.h through the C ABI; this module holds the frame-invariant tables
(DFT matrices, windows, resize taps, the reference planes' spectra) and the scalar bookkeeping that
the reference delegates to `skimage.registration.phase_cross_correlation(upsample_factor=up,
normalization="phase", disambiguate=True)`: integer peak -> wrap to signed shift -> matrix-multiply
DFT of the ceil(1.5 up)^2 up-sampled neighbourhood -> the 2^2 wrap candidates ranked by the Pearson
correlation of overlapping tiles.  Arithmetic differences from the reference: the DFTs are float64
dense products (scipy.fft works in complex64 for the reference's float32 planes), so estimates can
differ only where two correlation peaks tie to ~1e-6; everything up to the windowed planes follows
the reference's float32 rounding points.

Like the reference pipeline, this runs on single-channel recordings only: BatchMotionCorrector hands
the full (Z,Y,X,C) weight array to estimate_rigid_xcorr_3d, whose tensordot raises for C > 1.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib, device as dev
from .core import Context, _check, bare_context
from .plan import TableSet, gauss_taps


def _dft_matrix(n: int, inverse: bool = False) -> np.ndarray:
    k = np.arange(n)
    w = np.exp((2j if inverse else -2j) * np.pi * np.outer(k, k) / n)
    return (w / n) if inverse else w


def _tables_2d(in_hw, out_hw, sigma_coeff: float = 0.6) -> TableSet:
    """imresize2d_gauss_cubic (util/resize_util_3D.py:159-166): the fused resize of a (1,H,W) volume with one
    sigma per axis (per_axis=True, :120-123)."""
    lib = _lib.load()
    ts = TableSet((1, int(in_hw[0]), int(in_hw[1])), (1, int(out_hw[0]), int(out_hw[1])))
    for ax in (2, 1, 0):
        i, o = ts.in_shape[ax], ts.out_shape[ax]
        s = o / i
        R, g = gauss_taps(sigma_coeff / s if s < 1.0 else 0.0)
        idx = np.empty((o, 2 * R + 4), np.int32)
        wt = np.empty((o, 2 * R + 4), np.float32)
        rc = lib.fr3d_fill_resize_table(i, o, g.ctypes.data, R, idx.ctypes.data, wt.ctypes.data)
        if rc != 0:
            raise ValueError(f"fr3d_fill_resize_table({i}->{o}) failed: {rc}")
        ts.idx.append(idx)
        ts.wt.append(wt)
    return ts


class _Plane:
    """One projection plane (XY or XZ): sizes, tables, DFT matrices, the reference spectrum."""

    def __init__(self, ctx: Context, in_hw: Tuple[int, int], out_hw: Tuple[int, int]):
        self.ctx = ctx
        d = ctx.device
        self.in_hw = (int(in_hw[0]), int(in_hw[1]))
        self.H, self.W = int(out_hw[0]), int(out_hw[1])
        self.tables = None if self.in_hw == (self.H, self.W) else _tables_2d(self.in_hw, (self.H, self.W))
        cplx = lambda a: dev.to_device(np.ascontiguousarray(a.astype(np.complex128)).view(np.float64), d)
        self.WH, self.WW = cplx(_dft_matrix(self.H)), cplx(_dft_matrix(self.W))
        self.iWH, self.iWW = cplx(_dft_matrix(self.H, True)), cplx(_dft_matrix(self.W, True))
        self.hy = dev.to_device(np.hanning(self.H).astype(np.float32), d)
        self.hx = dev.to_device(np.hanning(self.W).astype(np.float32), d)
        self.ref_c = None    # windowed reference plane, complex (H,W)
        self.ref_F = None    # its 2-D DFT

    # -- device steps ---------------------------------------------------------------------------
    def resize(self, p: torch.Tensor) -> torch.Tensor:
        """(B,h,w) float32 -> (B,H,W) float32."""
        if self.tables is None:
            return p
        B = p.shape[0]
        out = dev.empty((B, self.H, self.W), np.float32, p.device)
        lib, h = self.ctx.lib, self.ctx.h
        _check(h, lib.fr3d_resize3d(h, dev.ptr(p), B, 1, self.in_hw[0], self.in_hw[1], self.tables.c_tables(),
                                    dev.ptr(out)))
        return out

    def window(self, p: torch.Tensor) -> torch.Tensor:
        B = p.shape[0]
        out = dev.empty((B, self.H, self.W, 2), np.float64, p.device)
        _check(self.ctx.h, self.ctx.lib.fr3d_cc_window(self.ctx.h, dev.ptr(p), B, self.H, self.W, dev.ptr(self.hy),
                                                       dev.ptr(self.hx), dev.ptr(out)))
        return out

    def gemm(self, A, a_stride, Bm, b_stride, M, N, K, nbatch) -> torch.Tensor:
        out = dev.empty((nbatch, M, N, 2), np.float64, self.ctx.device)
        _check(self.ctx.h, self.ctx.lib.fr3d_cc_cgemm(self.ctx.h, dev.ptr(A), int(a_stride), dev.ptr(Bm), int(b_stride),
                                                      dev.ptr(out), M, N, K, nbatch))
        return out

    def dft2(self, x_c: torch.Tensor, inverse: bool = False) -> torch.Tensor:
        """2-D DFT of B complex planes: W_H (x W_W)."""
        B, H, W = x_c.shape[0], self.H, self.W
        wh, ww = (self.iWH, self.iWW) if inverse else (self.WH, self.WW)
        t = self.gemm(x_c, H * W, ww, 0, H, W, W, B)
        return self.gemm(wh, 0, t, H * W, H, W, H, B)

    def argmax(self, cc: torch.Tensor) -> np.ndarray:
        B = cc.shape[0]
        n = cc[0].numel() // 2
        idx = dev.empty((B,), np.int64, cc.device)
        _check(self.ctx.h, self.ctx.lib.fr3d_cc_abs_argmax(self.ctx.h, dev.ptr(cc), n, B, dev.ptr(idx)))
        self.ctx.sync()
        return dev.to_host(idx)

    def set_reference(self, p_ref: torch.Tensor):
        self.ref_c = self.window(self.resize(p_ref))
        self.ref_F = self.dft2(self.ref_c)

    # -- phase_cross_correlation for B moving planes ----------------------------------------------
    def shifts(self, p_mov: torch.Tensor, up: int, normalize: bool = True, disambiguate: bool = True) -> np.ndarray:
        """(B,2) float32: the (row, column) shift registering each moving plane with the reference plane."""
        lib, h = self.ctx.lib, self.ctx.h
        H, W = self.H, self.W
        mov_c = self.window(self.resize(p_mov))
        B = mov_c.shape[0]
        Fm = self.dft2(mov_c)
        P = dev.empty((B, H, W, 2), np.float64, mov_c.device)
        _check(h, lib.fr3d_cc_cross_power(h, dev.ptr(self.ref_F), dev.ptr(Fm), dev.ptr(P), H * W, B,
                                          1 if normalize else 0))
        cc = self.dft2(P, inverse=True)
        peak = self.argmax(cc)
        shape = np.array([H, W])
        midpoint = np.fix(shape / 2)
        shift = np.stack(np.unravel_index(peak, (H, W)), -1).astype(np.float32)      # float_dtype of complex64
        for b in range(B):
            over = shift[b] > midpoint
            shift[b][over] -= shape[over]
        if up > 1:
            upf = np.float32(up)
            shift = np.round(shift * upf) / upf
            S = int(np.ceil(upf * 1.5))
            dftshift = np.fix(S / 2.0)
            off = dftshift - shift * upf                                              # (B,2)
            # conj(_upsampled_dft(conj(P))) = conj(K_row) P conj(K_col)^T
            ar = np.arange(S)[None, :, None]
            KR = np.exp(2j * np.pi * (ar - off[:, 0][:, None, None]) * np.fft.fftfreq(H, up)[None, None, :])
            KC = np.exp(2j * np.pi * (ar - off[:, 1][:, None, None]) * np.fft.fftfreq(W, up)[None, None, :])
            cplx = lambda a: dev.to_device(np.ascontiguousarray(a.astype(np.complex128)).view(np.float64), mov_c.device)
            kr = cplx(KR)                                                             # (B,S,H)
            kc = cplx(np.transpose(KC, (0, 2, 1)))                                    # (B,W,S)
            u = self.gemm(kr, S * H, P, H * W, S, W, H, B)
            v = self.gemm(u, S * W, kc, W * S, S, S, W, B)
            pk = self.argmax(v)
            maxima = np.stack(np.unravel_index(pk, (S, S)), -1).astype(np.float32) - np.float32(dftshift)
            shift = shift + maxima / upf
        shift = shift.astype(np.float32)
        for dim in range(2):
            if shape[dim] == 1:
                shift[:, dim] = 0
        if not disambiguate:
            return shift
        # skimage _disambiguate_shift: which of the 4 wrap candidates overlaps best (float32 scalars, as there)
        pos = np.mod(shift, shape[None, :].astype(np.float32))
        neg = pos - shape[None, :].astype(np.float32)
        work = dev.empty((B, H, W), np.float64, mov_c.device)
        shifted = dev.empty((B, H, W), np.float64, mov_c.device)
        sh = np.ascontiguousarray(shift, np.float64)
        _check(h, lib.fr3d_cc_wrap_shift(h, dev.ptr(mov_c), B, H, W, sh.ctypes.data, dev.ptr(work), dev.ptr(shifted)))
        split = np.ascontiguousarray(np.round(pos).astype(np.int32))
        sums = dev.empty((B, 4, 6), np.float64, mov_c.device)
        _check(h, lib.fr3d_cc_tile_sums(h, dev.ptr(self.ref_c), dev.ptr(shifted), B, H, W, split.ctypes.data,
                                        dev.ptr(sums)))
        self.ctx.sync()
        s = dev.to_host(sums)
        out = np.empty_like(shift)
        for b in range(B):
            best, best_tile = -1.0, None
            for tile in range(4):                       # itertools.product order: (y part, x part)
                n, sa, sb, saa, sbb, sab = s[b, tile]
                corr = -1.0
                if n > 2:
                    va, vb = saa - sa * sa / n, sbb - sb * sb / n
                    cov = sab - sa * sb / n
                    corr = cov / np.sqrt(va * vb) if va > 0 and vb > 0 else np.nan
                if corr > best:
                    best, best_tile = corr, tile
            if best_tile is None:
                out[b] = shift[b]
            else:
                out[b, 0] = neg[b, 0] if (best_tile >> 1) == 0 else pos[b, 0]     # first slice (stop = i) -> negative
                out[b, 1] = neg[b, 1] if (best_tile & 1) == 0 else pos[b, 1]
        return out


class RigidXCorr:
    """estimate_rigid_xcorr_3d against one fixed single-channel volume, B moving volumes per call."""

    def __init__(self, shape, target_hw=(256, 256), target_z: Optional[int] = None, up: int = 10,
                 device: Optional[torch.device] = None, ctx: Optional[Context] = None):
        self.ctx = ctx if ctx is not None else bare_context(device)
        Z, H, W = (int(s) for s in shape)
        self.shape = (Z, H, W)
        if isinstance(target_hw, int):
            target_hw = (target_hw, target_hw)
        Th = H if target_hw is None else min(H, int(target_hw[0]))
        Tw = W if target_hw is None else min(W, int(target_hw[1]))
        Tz = Z if target_z is None else min(Z, int(target_z))
        self.sy, self.sx, self.sz = H / Th, W / Tw, Z / Tz
        self.up = int(up)
        self.xy = _Plane(self.ctx, (H, W), (Th, Tw))
        self.xz = _Plane(self.ctx, (Z, W), (Tz, Tw))

    def _project(self, vol: torch.Tensor, acc64: bool):
        Z, H, W = self.shape
        B = vol.shape[0]
        pxy = dev.empty((B, H, W), np.float32, vol.device)
        pxz = dev.empty((B, Z, W), np.float32, vol.device)
        _check(self.ctx.h, self.ctx.lib.fr3d_cc_project(self.ctx.h, dev.ptr(vol), B, Z, H, W, 1 if acc64 else 0,
                                                        dev.ptr(pxy), dev.ptr(pxz)))
        return pxy, pxz

    def set_reference(self, ref_vol, float64_mean: bool = True):
        """ref_vol: (Z,Y,X) float32 device tensor or array (the pre-processed fixed volume, single channel).
        float64_mean: the reference volume is float64 in the reference pipeline (numpy then averages in float64);
        False reproduces numpy's float32 mean of a float32 volume."""
        v = ref_vol if isinstance(ref_vol, torch.Tensor) else dev.to_device(np.asarray(ref_vol, np.float32),
                                                                             self.ctx.device)
        v = v.reshape((1,) + self.shape).contiguous()
        pxy, pxz = self._project(v, acc64=bool(float64_mean))
        self.xy.set_reference(pxy)
        self.xz.set_reference(pxz)

    def estimate(self, mov) -> np.ndarray:
        """mov: (B,Z,Y,X) float32 (device tensor or array) -> (B,3) float32 = -[dx, dy, dz] per frame."""
        v = mov if isinstance(mov, torch.Tensor) else dev.to_device(np.asarray(mov, np.float32), self.ctx.device)
        v = v.reshape((-1,) + self.shape).contiguous()
        pxy, pxz = self._project(v, acc64=False)
        s_xy = self.xy.shifts(pxy, self.up)
        s_xz = self.xz.shifts(pxz, self.up)
        dy = s_xy[:, 0].astype(np.float64) * self.sy
        dx = s_xy[:, 1].astype(np.float64) * self.sx
        dz = s_xz[:, 0].astype(np.float64) * self.sz
        return -np.stack([dx, dy, dz], -1).astype(np.float32)


def estimate_rigid_xcorr_3d(ref_vol, mov_vol, target_hw=(256, 256), target_z=None, up=10, weight=None):
    """Drop-in for flowreg3d.util.xcorr_prealignment.estimate_rigid_xcorr_3d (single pair).  Multi-channel inputs are
    reduced to one channel on the host exactly as the reference does (:25-36) before the device path."""
    ref_vol = np.asarray(ref_vol)
    mov_vol = np.asarray(mov_vol)
    if ref_vol.ndim == 4 and ref_vol.shape[3] > 1:
        if weight is not None:
            w = np.asarray(weight).reshape(-1).astype(np.float32)
            w = w / w.sum()
            ref_vol = np.tensordot(ref_vol, w, axes=([3], [0]))     # raises ValueError like the reference for a full array
            mov_vol = np.tensordot(mov_vol, w, axes=([3], [0]))
        else:
            ref_vol = ref_vol.mean(axis=3)
            mov_vol = mov_vol.mean(axis=3)
    elif ref_vol.ndim == 4:
        ref_vol = ref_vol[..., 0]
        mov_vol = mov_vol[..., 0]
    r = RigidXCorr(ref_vol.shape, target_hw, target_z, up)
    r.set_reference(np.ascontiguousarray(ref_vol, np.float32), float64_mean=ref_vol.dtype == np.float64)
    return r.estimate(np.ascontiguousarray(mov_vol, np.float32)[None])[0]
