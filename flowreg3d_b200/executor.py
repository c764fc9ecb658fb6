"""Batch executor -- the plugin seam of the reference (drop-in boundary B1).

``B200Executor3D.process_batch`` has the signature and return contract of
flowreg3d.motion_correction.parallelization.base_3d.BaseExecutor3D.process_batch (base_3d.py:38-71)
and the per-frame semantics of SequentialExecutor3D (sequential_3d.py:89-173):

    flow_t = get_displacement(reference_proc, batch_proc[t], uvw=w_init, **flow_params).astype(f32)
    reg_t  = imregister_wrapper(batch[t], flow_t[...,0..2], reference_raw, interpolation_method)

but runs all frames of the batch concurrently on the GPU, with the fixed volume's pyramid cached
across calls.  When the reference package is importable, ``register()`` adds it to its
RuntimeContext so ``RegistrationConfig(parallelization="b200")`` selects it (the reference appends
"3d" to the requested name, compensate_recording_3D.py:88-91).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np
import torch

from . import device as dev
from .core import Registration
from .plan import FlowParams

_CC_KEYS = ("cc_initialization", "cc_hw", "cc_up")

try:  # subclass the reference's ABC when it is installed, so isinstance checks hold
    from flowreg3d.motion_correction.parallelization.base_3d import BaseExecutor3D as _Base  # type: ignore
except Exception:  # pragma: no cover - the reference is absent on the GPU box
    class _Base:  # minimal mirror of base_3d.py:11-117
        def __init__(self, n_workers: Optional[int] = None):
            self.n_workers = n_workers or 1
            self.name = self.__class__.__name__.replace("Executor", "").lower()

        def __enter__(self):
            self.setup()
            return self

        def __exit__(self, exc_type, exc_val, exc_tb):
            self.cleanup()
            return False

        def setup(self):
            pass

        def cleanup(self):
            pass

        def get_info(self):
            return {"name": self.name, "type": self.__class__.__name__, "n_workers": self.n_workers}


def flow_params_from_dict(d: dict) -> Tuple[FlowParams, object]:
    """Split the reference's flow_params dict into solver parameters and the weight array."""
    d = {k: v for k, v in d.items() if k not in _CC_KEYS}
    weight = d.pop("weight", None)
    d.pop("const_assumption", None)  # accepted and ignored, as in the reference (optical_flow_3d.py:457)
    d.pop("uvw", None)
    alpha = d.get("alpha", (2, 2, 2))
    if isinstance(alpha, (int, float)):
        alpha = (alpha,) * 3
    fp = FlowParams(alpha=tuple(float(a) for a in alpha), update_lag=int(d.get("update_lag", 10)),
                    iterations=int(d.get("iterations", 20)), min_level=int(d.get("min_level", 0)),
                    levels=int(d.get("levels", 50)), eta=float(d.get("eta", 0.8)),
                    a_smooth=float(d.get("a_smooth", 0.5)), a_data=d.get("a_data", 0.45))
    return fp, weight


class B200Executor3D(_Base):
    """GPU batch executor (one process per GPU; frames of a batch run concurrently)."""

    def __init__(self, n_workers: Optional[int] = None, max_batch: int = 16, device: Optional[torch.device] = None,
                 state_dtype=np.float64):
        """state_dtype: storage of the solver increments.  float64 (the package default): an executor whose name ends
        in "3d" is held to the reference's cross-executor consistency test
        (tests/motion_correction/test_parallelization.py:152-198, rtol 1e-5 against sequential3d), which float32
        increments do not meet."""
        super().__init__(n_workers=1)
        self.name = "b2003d"
        self.max_batch = int(max_batch)
        self.device = device
        self.state_dtype = state_dtype
        self._reg: Optional[Registration] = None
        self._key = None
        self._ref_token = None

    # -- BaseExecutor3D API -------------------------------------------------------------
    def process_batch(self, batch: np.ndarray, batch_proc: np.ndarray, reference_raw: np.ndarray,
                      reference_proc: np.ndarray, w_init: np.ndarray, get_displacement_func: Callable = None,
                      imregister_func: Callable = None, interpolation_method: str = "cubic",
                      progress_callback: Optional[Callable[[int], None]] = None, **kwargs):
        batch = np.asarray(batch)
        if batch.ndim != 5:
            raise ValueError(f"batch must be (T,Z,Y,X,C), got shape {batch.shape}")
        T, Z, Y, X, Cn = batch.shape
        fp_all = kwargs.get("flow_params", {})
        use_cc = bool(fp_all.get("cc_initialization", False))            # sequential_3d.py:62-72
        cc_hw = fp_all.get("cc_hw", 256)
        cc_up = int(fp_all.get("cc_up", 10))
        fp, weight = flow_params_from_dict(fp_all)
        reg = self._registration((Z, Y, X), Cn, fp, interpolation_method)
        self._ensure_reference(reg, reference_proc, reference_raw, weight)
        registered = np.empty_like(batch)
        flows = np.empty((T, Z, Y, X, 3), np.float32)
        uv = None if w_init is None else np.asarray(w_init).astype(np.float32)
        uvt = None if uv is None else dev.to_device(uv, reg.device)
        for t0 in range(0, T, reg.max_batch):
            t1 = min(T, t0 + reg.max_batch)
            if use_cc:
                w0 = uvt if uvt is not None else np.zeros((Z, Y, X, 3), np.float32)
                flow, _ = reg.get_displacement_cc(np.asarray(batch_proc[t0:t1], np.float64), w0, cc_hw=cc_hw,
                                                  cc_up=cc_up, ref_proc64=np.asarray(reference_proc, np.float64)
                                                  .reshape(Z, Y, X, -1)[..., 0])
            else:
                mv = np.asarray(batch_proc[t0:t1]).astype(np.float32)
                flow = reg.get_displacement(mv, uvw=uvt)
            out = reg.compensate(batch[t0:t1], flow)
            reg.ctx.order_with_torch()            # the device -> host copies below are ordered by the stream
            flows[t0:t1] = dev.to_host(flow)
            registered[t0:t1] = dev.to_host(out)  # numpy cast to the batch dtype, as sequential_3d.py:163-169
            if progress_callback is not None:
                progress_callback(t1 - t0)
        return registered, flows

    def cleanup(self):
        if self._reg is not None:
            self._reg.ctx.close()
        self._reg = None
        self._key = None
        self._ref_token = None

    def get_info(self) -> dict:
        info = super().get_info()
        info.update({"parallel": True, "description": "B200 (sm_100a) CUDA executor, batch-concurrent frames",
                     "max_batch": self.max_batch})
        return info

    @classmethod
    def register(cls) -> bool:
        """Register with the reference's RuntimeContext (no-op when flowreg3d is not installed)."""
        try:
            from flowreg3d._runtime import RuntimeContext  # type: ignore
        except Exception:
            return False
        RuntimeContext.register_parallelization_executor("b2003d", cls)
        return True

    def invalidate_reference(self):
        """Forget the cached fixed volume: the next process_batch uploads it and rebuilds its pyramid."""
        self._ref_token = None

    # -- internals ------------------------------------------------------------------------
    def _registration(self, shape, Cn, fp: FlowParams, interpolation_method) -> Registration:
        meth = str(getattr(interpolation_method, "value", interpolation_method)).lower()
        key = (tuple(shape), Cn, fp.alpha, fp.update_lag, fp.iterations, fp.min_level, fp.levels, fp.eta,
               fp.a_smooth, tuple(np.asarray(fp.a_data, float).ravel().tolist()), meth)
        if self._reg is None or key != self._key:
            self.cleanup()
            self._reg = Registration(shape, Cn, fp, max_batch=self.max_batch, interpolation_method=meth,
                                     device=self.device, state_dtype=self.state_dtype)
            self._key = key
        return self._reg

    def _ensure_reference(self, reg: Registration, reference_proc, reference_raw, weight):
        rp = np.asarray(reference_proc)
        rr = np.asarray(reference_raw)
        w = None if weight is None else np.asarray(weight)
        # content tokens: the full sum (one pass over the array, ~10 ms per 100 MB) plus a strided sample -- an
        # in-place edit of the reference, or a new array that numpy placed at a freed address, changes them; callers
        # that rewrite the reference in a way sums cannot see can call invalidate_reference()

        def tok(a):
            if a is None:
                return None
            f = a.reshape(-1)
            return (a.shape, str(a.dtype), float(f.sum(dtype=np.float64)),
                    float(f[:: max(1, f.size // 65536)].sum(dtype=np.float64)))
        token = (tok(rp), tok(rr), tok(w))
        if token != self._ref_token:
            if rp.ndim == 3:
                rp = rp[..., None]
                rr = rr[..., None]
            reg.set_reference(rp.astype(np.float32), weight=w, ref_raw=rr)
            self._ref_token = token
