"""Per-pair entry points with the reference's signatures (drop-in boundary B2):

    get_displacement(fixed, moving, alpha, update_lag, iterations, min_level, levels, eta, a_smooth,
                     a_data, const_assumption, uvw, weight) -> (Z,Y,X,3) float64
    imregister_wrapper(f2_level, u, v, w, f1_level, interpolation_method) -> float32

(reference: src/flowreg3d/core/optical_flow_3d.py:319-333 and :22-74), plus the Registration
context they are built on.  All per-voxel work happens in libfr3d's CUDA kernels.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib, device as dev
from .plan import FlowParams, PlanHolder, SWEEP_LEXICOGRAPHIC, make_tables


# Storage precision of the solver state (du,dv,dw and the constant Laplacian term) used when a caller does not
# choose: numpy.float64 | numpy.float32 | "auto".  The arithmetic and the system matrix are float64 either way.
#   float64 (DEFAULT) reproduces the reference to float64 rounding: 0.0 EPE against the live reference at config 2.
#   float32 is the storage SURVEY 7.3-D calls parity-safe and 8(d) budgets (108 B / voxel / sweep); it moves 24 % fewer
#   bytes (solver 53 -> 44 ms per 25 config-2 frames).  Measured on a B200 against the reference (round 2,
#   tests/test_gpu_published_sizes.py, tests/test_gpu_full_size.py): config 2 (min_level 5) mean 2.3e-6 / max 2.6e-3
#   voxel, corrected volume 1.7e-6 relative L2; config 4 reduced (min_level 2) 7.2e-5 / 4.5e-3; min_level 0 at
#   32x512x512 max 0.048 against the float64 state -- but the config-3 style sequence (expansion / recoil + jitter,
#   min_level 2, 40 sweeps, w_init chained over batches) reaches 2.2e-3 / 0.076, OUTSIDE the 0.05 max-EPE tolerance.
#   The rounding of the increments is amplified by the non-converged omega = 1.95 iteration and the psi non-linearity
#   in a data-dependent way, so float32 stays an explicit opt-in (state_dtype=numpy.float32, bench.py --state f32).
#   "auto" = float32 when the effective min_level is >= AUTO_F32_FROM_MIN_LEVEL, float64 below (opt-in policy).
STATE_DTYPE = np.float64
AUTO_F32_FROM_MIN_LEVEL = 2


def resolve_state_dtype(choice, shape, params: FlowParams):
    """numpy dtype of the solver state for `choice` (None = the module default)."""
    if choice is None:
        choice = STATE_DTYPE
    if isinstance(choice, str) and choice == "auto":
        from .plan import level_schedule
        _, eff_min_level = level_schedule(tuple(int(s) for s in shape), params.eta, params.levels, params.min_level)
        return np.float32 if eff_min_level >= AUTO_F32_FROM_MIN_LEVEL else np.float64
    return np.dtype(choice).type


# Sweep order of the level solver used when a caller does not choose: SWEEP_LEXICOGRAPHIC (wavefront
# schedule, reproduces the reference) or plan.SWEEP_REDBLACK (checkerboard; opt-in, not reference-exact).
SWEEP = SWEEP_LEXICOGRAPHIC


def _check(ctx, rc):
    if rc != 0:
        msg = _lib.load().fr3d_last_error(ctx)
        raise _lib.Fr3dError(rc, msg.decode() if msg else "?")


def _stream_handle(device: torch.device) -> int:
    if device.type != "cuda":
        return 0
    return torch.cuda.current_stream(device).cuda_stream


class Context:
    """RAII wrapper of fr3d_ctx."""

    def __init__(self, plan: Optional[PlanHolder] = None, device: Optional[torch.device] = None):
        self.lib = _lib.load()
        self.device = device if device is not None else dev.default_device()
        self.plan = plan
        h = C.c_void_p()
        idx = self.device.index if self.device.type == "cuda" else 0
        self.stream_handle = _stream_handle(self.device)   # the library launches everything on this stream
        rc = self.lib.fr3d_create(C.byref(h), idx or 0, C.byref(plan.plan) if plan is not None else None,
                                  self.stream_handle)
        if rc != 0:
            msg = self.lib.fr3d_last_error(None)
            raise _lib.Fr3dError(rc, msg.decode() if msg else "?")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            for fn in getattr(self, "_on_close", []):     # e.g. unmap the z-neighbours' buffers (multigpu._SlabPeers)
                try:
                    fn()
                except Exception:
                    pass
            self._on_close = []
            self.lib.fr3d_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        _check(self.h, self.lib.fr3d_synchronize(self.h))

    def order_with_torch(self):
        """Make work that torch (NCCL collectives, copies) enqueues next see the library's kernels, and vice versa.
        The library launches on the stream that was current when the context was created; when that is still torch's
        current stream, stream order already does it and nothing happens -- otherwise fall back to a synchronisation
        of both streams."""
        if self.device.type != "cuda":
            return
        cur = torch.cuda.current_stream(self.device)
        if cur.cuda_stream != self.stream_handle:
            cur.synchronize()
            self.sync()

    def profile(self, on: bool):
        _check(self.h, self.lib.fr3d_profile_enable(self.h, 1 if on else 0))

    def profile_report(self) -> dict:
        """{kernel name: (launches, total_ms)} since the last report; synchronises."""
        buf = C.create_string_buffer(1 << 16)
        n = self.lib.fr3d_profile_report(self.h, buf, len(buf))
        if n < 0:
            _check(self.h, int(n))
        out = {}
        for line in buf.value.decode().splitlines():
            name, cnt, ms = line.rsplit("\t", 2)
            out[name] = (int(cnt), float(ms))
        return out

    @property
    def launches(self) -> int:
        return int(self.lib.fr3d_launch_count(self.h))

    @property
    def device_bytes(self) -> int:
        return int(self.lib.fr3d_device_bytes(self.h))


import threading

_bare: dict = {}
_CACHE_LOCK = threading.RLock()     # guards _bare and _PAIR_CACHE (the per-pair functions are called from worker threads
                                    # by the reference's ThreadingExecutor3D, threading_3d.py:211-225)


def bare_context(device: Optional[torch.device] = None) -> Context:
    """Plan-less context serving the stage entry points: one per (device, calling thread) -- a context is not
    re-entrant, so threads never share one."""
    device = device if device is not None else dev.default_device()
    key = (str(device), threading.get_ident())
    with _CACHE_LOCK:
        if key not in _bare:
            _bare[key] = Context(None, device)
        return _bare[key]


class Registration:
    """B frames against one fixed volume: pre-process, flow, compensation -- device resident.

    Mirrors what BatchMotionCorrector asks of an executor (compensate_recording_3D.py:285-340):
    the fixed volume's pyramid is built once (set_reference) and reused for every frame.
    """

    def __init__(self, shape, n_channels: int, params: FlowParams, max_batch: int = 1,
                 interpolation_method: str = "cubic", sigma=None, sweep: Optional[int] = None,
                 device: Optional[torch.device] = None, state_dtype=None):
        meth = str(getattr(interpolation_method, "value", interpolation_method)).lower()
        if meth not in ("cubic", "linear"):
            raise ValueError("Unsupported interpolation method. Use 'linear' or 'cubic'.")
        self.shape = tuple(int(s) for s in shape)
        self.C = int(n_channels)
        self.max_batch = int(max_batch)
        self.plan = PlanHolder(self.shape, self.C, params, max_batch=max_batch,
                               interp=3 if meth == "cubic" else 1, sigma=sigma,
                               sweep=SWEEP if sweep is None else sweep,
                               state_dtype=resolve_state_dtype(state_dtype, self.shape, params))
        self.ctx = Context(self.plan, device)
        self.device = self.ctx.device
        self._ref_raw = None
        self._keep = []

    # -- reference ------------------------------------------------------------------------
    def set_reference(self, ref_proc, weight=None, ref_raw=None):
        """ref_proc (Z,Y,X,C) pre-processed fixed volume; weight: None (1/C), 1-D per-channel,
        (Z,Y,X) or (Z,Y,X,C) (core/optical_flow_3d.py:351-381); ref_raw: raw fixed volume used as
        the out-of-volume fill of the compensation warp."""
        Z, Y, X = self.shape
        rp = self._as_dev(ref_proc, np.float32, (Z, Y, X, self.C))
        wdev, wconst = None, None
        if weight is None:
            wconst = np.full(self.C, 1.0 / self.C)
        else:
            w = weight.cpu().numpy() if isinstance(weight, torch.Tensor) else np.asarray(weight)
            w = w.astype(np.float64)
            if w.ndim == 1:
                if len(w) < self.C:
                    we = np.full(self.C, 1.0 / self.C)
                    we[:len(w)] = w
                    w = we
                elif len(w) > self.C:
                    w = w[:self.C]
                wconst = w / w.sum()
            elif w.ndim == 3:
                w = np.ones((Z, Y, X, self.C)) * w[..., None]
            if wconst is None:
                flat = w.reshape(-1, self.C)
                if np.all(flat == flat[0]):
                    wconst = flat[0].copy()  # spatially constant: broadcast on the device instead of uploading
                else:
                    wdev = self._as_dev(w, np.float32, (Z, Y, X, self.C))
        wc = None
        if wconst is not None:
            # the reference resizes float32(weight); FillK rounds the float64 constant to float32 the same way
            wc = (C.c_double * _lib.MAX_CHANNELS)(*([float(v) for v in wconst] + [0.0] * (_lib.MAX_CHANNELS - self.C)))
        _check(self.ctx.h, self.ctx.lib.fr3d_set_reference(self.ctx.h, dev.ptr(rp), dev.ptr(wdev), wc))
        self._ref_proc = rp      # out-of-volume fill of the pre-alignment warps (cc_initialization)
        self._xcorr = None       # (key, RigidXCorr) for this reference
        if ref_raw is not None:
            rr = ref_raw if isinstance(ref_raw, torch.Tensor) else np.asarray(ref_raw)
            if not isinstance(rr, torch.Tensor) and rr.dtype not in _lib._DTYPES:
                rr = rr.astype(np.float64)
            self._ref_raw = self._as_dev(rr, None, (Z, Y, X, self.C))
        self.ctx.sync()  # rp / wdev may be temporaries

    # -- stages ---------------------------------------------------------------------------
    def preprocess(self, raw, lo, den, out: Optional[torch.Tensor] = None, temporal: bool = True,
                   out64: Optional[torch.Tensor] = None) -> torch.Tensor:
        """(raw - lo)/den then the Gaussian pre-filter; raw (B,Z,Y,X,C) any supported dtype, device
        tensor or ndarray; returns device float32 (B,Z,Y,X,C).  With a temporal sigma (>= 0.125) the B frames
        are one batch of the reference and are filtered across frames first; temporal=False is the
        reference's 4-D input case (fixed volume).  out64: optional float64 tensor of the same shape that
        receives the result before its rounding to float32."""
        raw_t = self._as_dev(raw, None, None)
        if raw_t.dim() == 4:
            raw_t = raw_t[None]
        B = raw_t.shape[0]
        if tuple(raw_t.shape[1:]) != self.shape + (self.C,):
            raise ValueError(f"raw frames have shape {tuple(raw_t.shape)}, expected (B,{self.shape},{self.C})")
        if out is None:
            out = dev.empty((B,) + self.shape + (self.C,), np.float32, self.device)
        lo = np.broadcast_to(np.asarray(lo, float), (self.C,)).copy()
        den = np.broadcast_to(np.asarray(den, float), (self.C,)).copy()
        _check(self.ctx.h, self.ctx.lib.fr3d_preprocess(self.ctx.h, dev.ptr(raw_t), self._code(raw_t), B,
                                                        lo.ctypes.data, den.ctypes.data, 1 if temporal else 0,
                                                        dev.ptr(out), dev.ptr(out64)))
        self._keep = [raw_t]
        return out

    def get_displacement(self, moving_proc, uvw=None, out_dtype=np.float32,
                         out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """moving_proc (B,Z,Y,X,C) float32 (device tensor or ndarray); uvw (Z,Y,X,3) shared initial
        flow or None.  Returns device (B,Z,Y,X,3)."""
        mv = self._as_dev(moving_proc, np.float32, None)
        if mv.dim() == 4:
            mv = mv[None]
        B = mv.shape[0]
        if tuple(mv.shape[1:]) != self.shape + (self.C,):
            raise ValueError(f"moving frames have shape {tuple(mv.shape)}, expected (B,{self.shape},{self.C})")
        uv = None if uvw is None else self._as_dev(uvw, np.float32, self.shape + (3,))
        if out is None:
            out = dev.empty((B,) + self.shape + (3,), out_dtype, self.device)
        _check(self.ctx.h, self.ctx.lib.fr3d_get_displacement(
            self.ctx.h, dev.ptr(mv), dev.ptr(uv), B, dev.ptr(out), self._code(out)))
        self._keep = [mv, uv]
        return out

    def get_displacement_cc(self, moving_proc, w_init, cc_hw=256, cc_up: int = 10,
                            ref_proc64=None) -> Tuple[torch.Tensor, np.ndarray]:
        """Flows of B frames with the rigid cross-correlation pre-alignment (cc_initialization=True), the six steps
        of parallelization/sequential_3d.py:89-145 for all frames of the batch at once:
        (1) warp every frame by w_init ("linear"), (2) rigid residual per frame by phase correlation of the mean
        projections (flowreg3d_b200.xcorr), (3) w_init + rigid, (4) warp the ORIGINAL frame by the combined field
        ("linear"), (5) non-rigid residual flow from zero, (6) total = combined + residual, rounded once to float32.
        moving_proc: (B,Z,Y,X,1) pre-processed frames, float64 (as the reference warps them) or float32.
        ref_proc64: the float64 pre-processed fixed volume if the caller has it (the reference averages its
        projections in float64); default: the float32 volume given to set_reference.
        Single channel only, like the reference pipeline (see flowreg3d_b200/xcorr.py).
        Returns (flows (B,Z,Y,X,3) float32 on the device, rigid (B,3) float32 host)."""
        from .xcorr import RigidXCorr
        if self.C != 1:
            # what the reference pipeline does for C > 1: estimate_rigid_xcorr_3d contracts the full (Z,Y,X,C)
            # weight array with the channel axis (xcorr_prealignment.py:26-30)
            raise ValueError("shape-mismatch for sum (cc_initialization supports single-channel recordings only, "
                             "as in the reference)")
        Z, Y, X = self.shape
        mv = self._as_dev(moving_proc, None, None)
        if mv.dim() == 4:
            mv = mv[None]
        B = mv.shape[0]
        if tuple(mv.shape[1:]) != (Z, Y, X, 1):
            raise ValueError(f"moving frames have shape {tuple(mv.shape)}, expected (B,{self.shape},1)")
        wi = self._as_dev(w_init, np.float32, (Z, Y, X, 3))
        lib, h = self.ctx.lib, self.ctx.h
        key = (cc_hw if isinstance(cc_hw, int) else tuple(cc_hw), int(cc_up))
        if self._xcorr is None or self._xcorr[0] != key:
            xc = RigidXCorr(self.shape, target_hw=cc_hw, up=int(cc_up), ctx=self.ctx)
            if ref_proc64 is not None:
                r64 = self._as_dev(ref_proc64, np.float64, (Z, Y, X))
                xc.set_reference(r64.to(torch.float32), float64_mean=True)   # projections of float32(ref): <= 1 ulp
            else:
                xc.set_reference(self._ref_proc.reshape(Z, Y, X), float64_mean=True)
            self._xcorr = (key, xc)
        xc = self._xcorr[1]
        nvox = Z * Y * X

        def rigid_field(rigid):
            out = dev.empty((B, Z, Y, X, 3), np.float32, self.device)
            r = np.ascontiguousarray(rigid, np.float32)
            _check(h, lib.fr3d_rigid_flow(h, dev.ptr(wi), r.ctypes.data, B, nvox, dev.ptr(out)))
            return out

        def warp_linear(flow):
            out = dev.empty((B, Z, Y, X, 1), np.float32, self.device)
            _check(h, lib.fr3d_warp_flow(h, dev.ptr(mv), self._code(mv), dev.ptr(flow), dev.ptr(self._ref_proc),
                                         _lib.F32, B, Z, Y, X, 1, 1, dev.ptr(out)))
            return out

        part = warp_linear(rigid_field(np.zeros((B, 3), np.float32)))            # (1)
        rigid = xc.estimate(part.reshape(B, Z, Y, X))                             # (2)
        comb = rigid_field(rigid)                                                 # (3)
        aligned = warp_linear(comb)                                               # (4)
        flows = dev.empty((B, Z, Y, X, 3), np.float32, self.device)
        for t0 in range(0, B, self.max_batch):                                    # (5), (6)
            t1 = min(B, t0 + self.max_batch)
            resid = self.get_displacement(aligned[t0:t1], uvw=None, out_dtype=np.float64)
            _check(h, lib.fr3d_add_flow(h, dev.ptr(comb[t0:t1]), dev.ptr(resid), (t1 - t0) * nvox * 3,
                                        dev.ptr(flows[t0:t1])))
        self._keep = [mv, wi, comb, aligned, part]
        return flows, rigid

    def compensate(self, raw, flow, ref_raw=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Warp B raw frames with their flows (float32), out-of-volume voxels from ref_raw."""
        raw_t = self._as_dev(raw, None, None)
        if raw_t.dim() == 4:
            raw_t = raw_t[None]
        B = raw_t.shape[0]
        fl = self._as_dev(flow, np.float32, (B,) + self.shape + (3,))
        rr = self._ref_raw if ref_raw is None else self._as_dev(ref_raw, None, self.shape + (self.C,))
        if rr is None:
            raise ValueError("no raw reference: pass ref_raw here or to set_reference")
        if out is None:
            out = dev.empty((B,) + self.shape + (self.C,), np.float32, self.device)
        _check(self.ctx.h, self.ctx.lib.fr3d_compensate(
            self.ctx.h, dev.ptr(raw_t), self._code(raw_t), dev.ptr(fl), dev.ptr(rr), self._code(rr), B,
            dev.ptr(out)))
        self._keep = [raw_t, fl, rr]
        return out

    def mean_frames(self, frames: torch.Tensor) -> torch.Tensor:
        """numpy.mean(frames, axis=0) for float32 device frames (T, ...) -> (...)."""
        T = frames.shape[0]
        n = frames[0].numel()
        out = dev.empty(tuple(frames.shape[1:]), np.float32, self.device)
        fr = frames.contiguous()
        _check(self.ctx.h, self.ctx.lib.fr3d_mean_frames(self.ctx.h, dev.ptr(fr), T, n, dev.ptr(out)))
        self._keep = [fr]
        return out

    def mean_frames_f64(self, frames: torch.Tensor) -> torch.Tensor:
        """numpy.mean(axis=0) of float32 device frames (T, ...) accumulated in float64 -> float64 (...)."""
        T = frames.shape[0]
        n = frames[0].numel()
        out = dev.empty(tuple(frames.shape[1:]), np.float64, self.device)
        fr = frames.contiguous()
        _check(self.ctx.h, self.ctx.lib.fr3d_mean_frames_f64(self.ctx.h, dev.ptr(fr), T, n, dev.ptr(out)))
        self._keep = [fr]
        return out

    def flow_stats(self, flows: torch.Tensor) -> torch.Tensor:
        """Per-frame statistics of BatchMotionCorrector (compensate_recording_3D.py:488-508) for float32 device
        flows (B,Z,Y,X,3): (B,4) float64 = mean |w|, max |w|, mean divergence, |mean translation|."""
        fl = flows.contiguous()
        B = fl.shape[0]
        Z, Y, X = self.shape
        out = dev.empty((B, 4), np.float64, self.device)
        _check(self.ctx.h, self.ctx.lib.fr3d_flow_stats(self.ctx.h, dev.ptr(fl), B, Z, Y, X, dev.ptr(out)))
        self._keep = [fl]
        return out

    def sync(self):
        self.ctx.sync()

    # -- helpers --------------------------------------------------------------------------
    def _as_dev(self, a, dtype, shape):
        if isinstance(a, torch.Tensor):
            t = a
            if dtype is not None and t.dtype != dev.torch_dtype(dtype):
                t = t.to(dev.torch_dtype(dtype))
            if t.device != self.device:
                t = t.to(self.device)
            t = t.contiguous()
        else:
            a = np.asarray(a)
            if dtype is not None:
                a = a.astype(dtype, copy=False)
            elif a.dtype not in _lib._DTYPES:
                a = a.astype(np.float64)
            t = dev.to_device(a, self.device)
        if shape is not None:
            t = t.reshape(tuple(shape))
        return t

    @staticmethod
    def _code(t: torch.Tensor) -> int:
        return _lib.dtype_code(str(t.dtype).replace("torch.", ""))


class SplitRegistration:
    """Registration whose batch is cut into `n_streams` contiguous parts, each owned by its own
    fr3d context on its own CUDA stream.  Frames are independent in every stage, so the parts only
    meet at the fork (inputs ready) and the join (outputs ready) of each call.  Purpose: the
    persistent wavefront solver of one part spends a third of its time in grid barriers and in
    ramp-up / ramp-down waves; kernels of the other part (ALU-bound median, fp64-bound warp, or its
    own solver) fill those holes.  Each solver launch claims one CTA per SM so that two can be
    resident together."""

    def __init__(self, shape, n_channels: int, params: FlowParams, max_batch: int = 2, n_streams: int = 2,
                 device: Optional[torch.device] = None, **kw):
        self.device = device if device is not None else dev.default_device()
        self.n = max(1, min(int(n_streams), int(max_batch)))
        self.max_batch = int(max_batch)
        per = -(-self.max_batch // self.n)
        self.parts: List[Registration] = []
        self.streams = []
        for _ in range(self.n):
            if self.device.type == "cuda":
                st = torch.cuda.Stream(self.device)
                with torch.cuda.stream(st):
                    r = Registration(shape, n_channels, params, max_batch=per, device=self.device, **kw)
            else:
                st = None
                r = Registration(shape, n_channels, params, max_batch=per, device=self.device, **kw)
            _check(r.ctx.h, r.ctx.lib.fr3d_set_option(r.ctx.h, _lib.OPT_SOR_CTAS_PER_SM, 1))
            self.parts.append(r)
            self.streams.append(st)
        first = self.parts[0]
        self.shape, self.C, self.plan, self.ctx = first.shape, first.C, first.plan, first.ctx

    # -- fork / join --------------------------------------------------------------------
    def _bounds(self, B):
        per = -(-B // self.n)
        return [(i * per, min(B, (i + 1) * per)) for i in range(self.n) if i * per < B]

    def _run(self, B, fn):
        """fn(part, lo, hi) on every part's stream; joins back into the caller's stream."""
        cur = torch.cuda.current_stream(self.device) if self.device.type == "cuda" else None
        used = []
        for i, (lo, hi) in enumerate(self._bounds(B)):
            st = self.streams[i]
            if st is None:
                fn(self.parts[i], lo, hi)
                continue
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                fn(self.parts[i], lo, hi)
            used.append(st)
        for st in used:
            cur.wait_stream(st)

    def _all(self, fn):
        cur = torch.cuda.current_stream(self.device) if self.device.type == "cuda" else None
        for r, st in zip(self.parts, self.streams):
            if st is None:
                fn(r)
                continue
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                fn(r)
            cur.wait_stream(st)

    # -- Registration interface ---------------------------------------------------------
    def set_reference(self, ref_proc, weight=None, ref_raw=None):
        self._all(lambda r: r.set_reference(ref_proc, weight=weight, ref_raw=ref_raw))

    def _as_dev(self, a, dtype, shape):
        return self.parts[0]._as_dev(a, dtype, shape)

    def preprocess(self, raw, lo, den, out: Optional[torch.Tensor] = None, temporal: bool = True,
                   out64: Optional[torch.Tensor] = None) -> torch.Tensor:
        raw_t = self._as_dev(raw, None, None)
        if raw_t.dim() == 4:
            raw_t = raw_t[None]
        B = raw_t.shape[0]
        if out is None:
            out = dev.empty((B,) + self.shape + (self.C,), np.float32, self.device)
        if temporal and self.plan.temporal:
            # the temporal filter couples the frames of the batch: one part takes all of it
            self._run(1, lambda r, a, b: r.preprocess(raw_t, lo, den, out=out, out64=out64))
        else:
            self._run(B, lambda r, a, b: r.preprocess(raw_t[a:b], lo, den, out=out[a:b], temporal=temporal,
                                                      out64=None if out64 is None else out64[a:b]))
        return out

    def get_displacement(self, moving_proc, uvw=None, out_dtype=np.float32,
                         out: Optional[torch.Tensor] = None) -> torch.Tensor:
        mv = self._as_dev(moving_proc, np.float32, None)
        if mv.dim() == 4:
            mv = mv[None]
        B = mv.shape[0]
        uv = None if uvw is None else self._as_dev(uvw, np.float32, self.shape + (3,))
        if out is None:
            out = dev.empty((B,) + self.shape + (3,), out_dtype, self.device)
        self._run(B, lambda r, a, b: r.get_displacement(mv[a:b], uvw=uv, out=out[a:b]))
        return out

    def compensate(self, raw, flow, ref_raw=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        raw_t = self._as_dev(raw, None, None)
        if raw_t.dim() == 4:
            raw_t = raw_t[None]
        B = raw_t.shape[0]
        fl = self._as_dev(flow, np.float32, (B,) + self.shape + (3,))
        rr = None if ref_raw is None else self._as_dev(ref_raw, None, self.shape + (self.C,))
        if out is None:
            out = dev.empty((B,) + self.shape + (self.C,), np.float32, self.device)
        self._run(B, lambda r, a, b: r.compensate(raw_t[a:b], fl[a:b], ref_raw=rr, out=out[a:b]))
        return out

    def mean_frames(self, frames: torch.Tensor) -> torch.Tensor:
        res = []
        self._run(1, lambda r, a, b: res.append(r.mean_frames(frames)))
        return res[0]

    def flow_stats(self, flows: torch.Tensor) -> torch.Tensor:
        res = []
        self._run(1, lambda r, a, b: res.append(r.flow_stats(flows)))
        return res[0]

    def sync(self):
        for r in self.parts:
            r.sync()
        if self.device.type == "cuda":
            torch.cuda.current_stream(self.device).synchronize()

    def close(self):
        for r in self.parts:
            r.ctx.close()

    @property
    def launches(self) -> int:
        return sum(r.ctx.launches for r in self.parts)


# --------------------------------------------------------------------------------------------
# Reference-signature functions
# --------------------------------------------------------------------------------------------
def get_displacement(fixed, moving, alpha=(2, 2, 2), update_lag=10, iterations=20, min_level=0, levels=50,
                     eta=0.8, a_smooth=0.5, a_data=0.45, const_assumption="gc", uvw=None, weight=None):
    """Drop-in for flowreg3d.core.optical_flow_3d.get_displacement (:319-542): dense 3-D flow of
    `moving` towards `fixed`, (Z,Y,X[,C]) arrays in, (Z,Y,X,3) float64 out, [...,0]=dx.
    `const_assumption` is accepted and ignored exactly as in the reference (gradient constancy is
    hard-wired, :457).  a_smooth != 1 selects the nonlinear smoothness term (psi_s recomputed every sweep)."""
    fixed = np.asarray(fixed)
    moving = np.asarray(moving)
    if fixed.ndim == 3:
        fixed = fixed[..., None]
        moving = moving[..., None]
    if fixed.ndim != 4 or fixed.shape != moving.shape:
        raise ValueError(f"fixed/moving must be (Z,Y,X) or (Z,Y,X,C) of equal shape, got {fixed.shape} / {moving.shape}")
    if isinstance(alpha, (int, float)):
        alpha = (alpha,) * 3
    Z, Y, X, Cn = fixed.shape
    fp = FlowParams(alpha=tuple(alpha), update_lag=update_lag, iterations=iterations, min_level=min_level,
                    levels=levels, eta=eta, a_smooth=a_smooth, a_data=a_data)
    reg = _pair_registration((Z, Y, X), Cn, fp)
    # the fixed volume's pyramid is rebuilt only when the fixed volume (or the weight) changed since this thread's
    # previous call with the same shape and parameters -- the reference's executors call once per frame against the
    # same fixed volume
    token = (_content_token(fixed), _content_token(weight))
    if getattr(reg, "_pair_token", None) != token:
        reg.set_reference(fixed.astype(np.float32), weight=weight)
        reg._pair_token = token
    uv = None if uvw is None else np.asarray(uvw).astype(np.float32)
    # With min_level > 0 the flow is the output of the final resize, i.e. float32-exact values (the reference's
    # imresize casts to float32 too): fetch 12 B/voxel and widen on the host.  At min_level == 0 it is the
    # float64 accumulation of the level increments.
    odt = np.float32 if reg.plan.min_level > 0 else np.float64
    out = reg.get_displacement(moving.astype(np.float32)[None], uvw=uv, out_dtype=odt)
    if reg.device.type == "cuda":
        # pinned staging buffer (kept with the cached context): one asynchronous copy instead of a pageable one
        stage = getattr(reg, "_pair_stage", None)
        if stage is None or stage.shape != out.shape or stage.dtype != out.dtype:
            stage = reg._pair_stage = torch.empty(out.shape, dtype=out.dtype).pin_memory()
        stage.copy_(out, non_blocking=True)
        torch.cuda.current_stream(reg.device).synchronize()
        return stage.numpy()[0].astype(np.float64)
    reg.sync()
    return dev.to_host(out)[0].astype(np.float64)


def _content_token(a):
    """Cheap content fingerprint of an array (None passes through): shape, dtype, full sum and a strided sample."""
    if a is None:
        return None
    a = np.asarray(a)
    f = a.reshape(-1)
    return (a.shape, str(a.dtype), float(f.sum(dtype=np.float64)), float(f[:: max(1, f.size // 4096)].sum(dtype=np.float64)))


# The reference's executors call get_displacement once per frame with the same shape and parameters
# (sequential_3d.py:89-173): keep the last few contexts (plan tables, solver geometry, workspaces) alive
# instead of rebuilding them on every call.
_PAIR_CACHE: "dict" = {}
_PAIR_CACHE_MAX = 4       # per calling thread


def _pair_registration(shape, Cn, fp: FlowParams) -> Registration:
    """The cached Registration of the calling thread for these parameters (contexts are not re-entrant: every
    thread gets its own; least recently used ones of the thread are closed beyond _PAIR_CACHE_MAX)."""
    device = dev.default_device()
    tid = threading.get_ident()
    key = (tid, shape, Cn, tuple(float(a) for a in fp.alpha), int(fp.update_lag), int(fp.iterations), int(fp.min_level),
           int(fp.levels), float(fp.eta), float(fp.a_smooth), tuple(np.asarray(fp.a_data, float).ravel().tolist()),
           np.dtype(resolve_state_dtype(None, shape, fp)).str, int(SWEEP), str(device), str(_lib.library_path()))
    with _CACHE_LOCK:
        reg = _PAIR_CACHE.pop(key, None)
        if reg is None or reg.ctx.h is None:
            reg = Registration(shape, Cn, fp, max_batch=1, device=device)
        _PAIR_CACHE[key] = reg                      # most recently used last
        mine = [k for k in _PAIR_CACHE if k[0] == tid]
        for k in mine[:-_PAIR_CACHE_MAX]:
            _PAIR_CACHE.pop(k).ctx.close()
    return reg


def imregister_wrapper(f2_level, u, v, w, f1_level, interpolation_method="cubic"):
    """Drop-in for flowreg3d.core.optical_flow_3d.imregister_wrapper (:22-74): backward warp
    warped(x) = f2(x + (u,v,w)); voxels leaving the volume take f1's value; float32 result with the
    channel axis squeezed for single-channel input."""
    meth = str(interpolation_method).lower()
    if meth == "cubic":
        order = 3
    elif meth == "linear":
        order = 1
    else:
        raise ValueError("Unsupported interpolation method. Use 'linear' or 'cubic'.")
    f2 = np.asarray(f2_level)
    f1 = np.asarray(f1_level)
    if f2.ndim == 3:
        f2 = f2[..., None]
        f1 = f1[..., None]
    Z, Y, X, Cn = f2.shape
    ctx = bare_context()
    d = ctx.device

    def up(a, dt=None):
        a = np.asarray(a)
        if dt is not None:
            a = a.astype(dt, copy=False)
        elif a.dtype not in _lib._DTYPES:
            a = a.astype(np.float64)
        return dev.to_device(a, d)

    f2t = up(f2)
    f1t = f2t if f1 is f2 else up(np.broadcast_to(f1, f2.shape))
    out = dev.empty((Z, Y, X, Cn), np.float32, d)
    code = lambda t: _lib.dtype_code(str(t.dtype).replace("torch.", ""))
    flow3 = _interleaved_flow(u, v, w, (Z, Y, X))
    if flow3 is not None:
        # the usual call of the reference's executors: u, v, w are the three component views flow[..., q] of ONE
        # float32 (Z,Y,X,3) array (sequential_3d.py:148-160).  Upload that array once (12 B / voxel instead of three
        # float64 planes) and take the per-frame-flow entry point: the coordinates are float32 sums either way.
        ft = up(flow3)
        _check(ctx.h, ctx.lib.fr3d_warp_flow(ctx.h, dev.ptr(f2t), code(f2t), dev.ptr(ft), dev.ptr(f1t), code(f1t), 1, Z, Y,
                                             X, Cn, order, dev.ptr(out)))
    else:
        ut, vt, wt = (up(np.broadcast_to(np.asarray(a), (Z, Y, X)), np.float64) for a in (u, v, w))
        _check(ctx.h, ctx.lib.fr3d_warp(ctx.h, dev.ptr(f2t), code(f2t), dev.ptr(ut), dev.ptr(vt), dev.ptr(wt),
                                        dev.ptr(f1t), code(f1t), Z, Y, X, Cn, order, dev.ptr(out)))
    ctx.sync()
    res = dev.to_host(out).copy()
    return res[..., 0] if Cn == 1 else res


def _interleaved_flow(u, v, w, shape):
    """The float32 (Z,Y,X,3) array that u, v, w are the component views of, or None."""
    if not all(isinstance(a, np.ndarray) and a.dtype == np.float32 and a.shape == tuple(shape) for a in (u, v, w)):
        return None
    Z, Y, X = shape
    want = (Y * X * 12, X * 12, 12)
    if not all(a.strides == want for a in (u, v, w)):
        return None
    p0, p1, p2 = (a.__array_interface__["data"][0] for a in (u, v, w))
    if p1 != p0 + 4 or p2 != p0 + 8:
        return None
    return np.lib.stride_tricks.as_strided(u, shape=(Z, Y, X, 3), strides=want + (4,), writeable=False)


# --------------------------------------------------------------------------------------------
# Stage wrappers (numpy in / numpy out) used by the stage-wise parity tests
# --------------------------------------------------------------------------------------------
def resize(img, size):
    """imresize_fused_gauss_cubic3D (util/resize_util_3D.py:114-156) for float images."""
    img = np.asarray(img)
    x = img.astype(np.float32)
    squeeze = x.ndim == 3
    if squeeze:
        x = x[..., None]
    D, H, W, Cn = x.shape
    size = tuple(int(s) for s in size[:3])
    ts = make_tables((D, H, W), size)
    ctx = bare_context()
    src = dev.to_device(np.ascontiguousarray(np.moveaxis(x, -1, 0)), ctx.device)
    dst = dev.empty((Cn,) + size, np.float32, ctx.device)
    _check(ctx.h, ctx.lib.fr3d_resize3d(ctx.h, dev.ptr(src), Cn, D, H, W, ts.c_tables(), dev.ptr(dst)))
    ctx.sync()
    y = np.moveaxis(dev.to_host(dst), 0, -1)
    y = y[..., 0] if squeeze else y
    return np.ascontiguousarray(y).astype(img.dtype if np.issubdtype(img.dtype, np.floating) else np.float32)


def motion_tensor(f1, f2, hz, hy, hx, kind: str = "gc"):
    """get_motion_tensor_gc (core/optical_flow_3d.py:92-152) without its zero ring: (10,p,m,n).
    kind "gray" / "cs": get_motion_tensor_gray (:218-259) / get_motion_tensor_cs (:155-215) for float64 images --
    stage functions only; like the reference's driver, get_displacement always uses "gc"."""
    f1 = np.asarray(f1)
    f2 = np.asarray(f2)
    if kind != "gc":
        if kind not in ("gray", "cs"):
            raise ValueError("kind must be 'gc', 'gray' or 'cs'")
        p, m, n = f1.shape
        ctx = bare_context()
        a = dev.to_device(f1.astype(np.float64), ctx.device)
        b = dev.to_device(f2.astype(np.float64), ctx.device)
        J = dev.empty((10, p, m, n), np.float64, ctx.device)
        _check(ctx.h, ctx.lib.fr3d_motion_tensor_alt(ctx.h, 1 if kind == "gray" else 2, dev.ptr(a), dev.ptr(b), p, m, n,
                                                     float(hz), float(hy), float(hx), dev.ptr(J)))
        ctx.sync()
        return dev.to_host(J).copy()
    f2f32 = 1 if f2.dtype == np.float32 else 0
    p, m, n = f1.shape
    ctx = bare_context()
    a = dev.to_device(f1.astype(np.float32), ctx.device)
    b = dev.to_device(f2.astype(np.float32), ctx.device)
    J = dev.empty((10, p, m, n), np.float64, ctx.device)
    _check(ctx.h, ctx.lib.fr3d_motion_tensor(ctx.h, dev.ptr(a), dev.ptr(b), p, m, n, float(hz), float(hy),
                                             float(hx), f2f32, dev.ptr(J)))
    ctx.sync()
    return dev.to_host(J).copy()


def sor_level(J, weight, uvw, alpha, h, iterations, update_lag, a_data, a_smooth=1.0, sweep=SWEEP_LEXICOGRAPHIC,
              state_dtype=np.float64):
    """compute_flow_3d (core/level_solver_3d.py:314-546) on interior arrays:
    J (C,10,p,m,n), weight (C,p,m,n), uvw (3,p,m,n), alpha (x,y,z), h (hz,hy,hx) -> (3,p,m,n)."""
    J = np.ascontiguousarray(J, np.float64)
    Cn, _, p, m, n = J.shape
    ctx = bare_context()
    Jt = dev.to_device(J, ctx.device)
    wt = dev.to_device(np.ascontiguousarray(weight, np.float64), ctx.device)
    ut = dev.to_device(np.ascontiguousarray(uvw, np.float64), ctx.device)
    out = dev.empty((3, p, m, n), np.float64, ctx.device)
    al = np.asarray(alpha, float)
    ad = np.broadcast_to(np.asarray(a_data, float), (Cn,)).copy()
    _check(ctx.h, ctx.lib.fr3d_sor_level(ctx.h, dev.ptr(Jt), dev.ptr(wt), dev.ptr(ut), p, m, n, Cn,
                                         al.ctypes.data, float(h[0]), float(h[1]), float(h[2]), int(iterations),
                                         int(update_lag), ad.ctypes.data, float(a_smooth), int(sweep),
                                         _lib.dtype_code(state_dtype), dev.ptr(out)))
    ctx.sync()
    return dev.to_host(out).copy()


def median5(vols):
    """scipy.ndimage.median_filter(size=(5,5,5), mode="mirror") on (nvol,p,m,n) float64."""
    v = np.ascontiguousarray(vols, np.float64)
    if v.ndim == 3:
        v = v[None]
    nv, p, m, n = v.shape
    ctx = bare_context()
    s = dev.to_device(v, ctx.device)
    o = dev.empty(v.shape, np.float64, ctx.device)
    _check(ctx.h, ctx.lib.fr3d_median5(ctx.h, dev.ptr(s), nv, p, m, n, dev.ptr(o)))
    ctx.sync()
    return dev.to_host(o).copy()
