"""Sequence entry points (drop-in boundary L4 of the reference).

``compensate_arr_3D(c1, c_ref, options)`` has the signature, shape canonicalisation and return
contract of flowreg3d.motion_correction.compensate_arr_3D.compensate_arr_3D
(compensate_arr_3D.py:13-143) and the batch semantics of BatchMotionCorrector.run
(compensate_recording_3D.py:431-557):

  * reference pre-processing (normalise against its own range, Gaussian pre-filter)   :198-254
  * frames pre-processed against the reference's range                                :459-463
  * first batch: w_init = mean flow of its first min(22, T) frames solved from zero   :342-393
  * per batch: flow from w_init, compensation warp, w_init <- mean of the last <= 20 flows :476-485

Unlike the reference, where the executor receives host arrays pre-processed by scipy, everything
between the raw frames and (registered, flow) stays on the GPU: normalise + pre-filter, pyramid,
level solves, flow up-sampling, compensation warp and the w_init averages.

Multi-GPU (one process per GPU, torch.distributed/NCCL): frames of each batch are split into
contiguous shards, the fixed volume is replicated, and the only exchange is one all-reduce of the
partial w_init sums per batch.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import numpy as np
import torch

from . import _lib, device as dev
from .core import Registration, SplitRegistration
from .options import OFOptions
from .plan import FlowParams

_OUT_TYPES = {"single": np.float32, "double": np.float64, "uint8": np.uint8, "uint16": np.uint16,
              "int16": np.int16, "int32": np.int32}


def flow_params_from_options(options) -> FlowParams:
    """compensate_recording_3D.py:301-315."""
    return FlowParams(alpha=tuple(float(a) for a in options.alpha), update_lag=int(options.update_lag),
                      iterations=int(options.iterations),
                      min_level=int(getattr(options, "effective_min_level", getattr(options, "min_level", 0))),
                      levels=int(options.levels), eta=float(options.eta), a_smooth=float(options.a_smooth),
                      a_data=options.a_data)


def normalization_range(reference_raw: np.ndarray, channel_normalization) -> Tuple[np.ndarray, np.ndarray]:
    """(lo, den) per channel such that normalised = (x - lo)/den  (util/image_processing_3D.py:12-92).
    NB: the reference compares the enum against the string "separate"; JOINT ("joint") therefore takes
    the global branch with eps = 1e-8 in the denominator."""
    mode = str(getattr(channel_normalization, "value", channel_normalization))
    Cn = reference_raw.shape[-1]
    if mode == "separate":
        lo = np.array([reference_raw[..., c].min() for c in range(Cn)], dtype=np.float64)
        hi = np.array([reference_raw[..., c].max() for c in range(Cn)], dtype=np.float64)
        rng = hi - lo
        return lo, np.where(rng > 0, rng, 1.0)
    lo = np.float64(reference_raw.min())
    hi = np.float64(reference_raw.max())
    return np.full(Cn, lo), np.full(Cn, hi - lo + 1e-8)


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous split of n frames over `world` ranks (first n % world ranks get one more)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class SequenceCorrector:
    """Stateful batch processor: owns the Registration, the fixed volume and the w_init chain."""

    def __init__(self, reference_raw: np.ndarray, options, max_batch: Optional[int] = None,
                 device: Optional[torch.device] = None, group=None, streams: int = 1, statistics: bool = False,
                 cc_prealign: bool = False):
        self.options = options
        ref = np.asarray(reference_raw)
        if ref.ndim == 3:
            ref = ref[..., None]
        self.reference_raw = ref.astype(np.float64)  # compensate_recording_3D.py:201-205
        Z, Y, X, Cn = self.reference_raw.shape
        self.shape, self.C = (Z, Y, X), Cn
        self.group = group
        self.world = 1
        self.rank = 0
        if group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(group)
            self.rank = torch.distributed.get_rank(group)
        fp = flow_params_from_options(options)
        mb = int(max_batch or options.buffer_size)
        kw = dict(interpolation_method=getattr(options, "interpolation_method", "cubic"), sigma=options.sigma)
        cc_wanted = bool(getattr(options, "cc_initialization", False)) and bool(cc_prealign)
        if cc_wanted or bool(getattr(options, "update_reference", False)):
            streams = 1     # both couple the frames of a batch through one context: one stream per GPU
        if int(streams) > 1 and mb > 1:
            self.reg = SplitRegistration(self.shape, Cn, fp, max_batch=mb, n_streams=int(streams), device=device, **kw)
        else:
            self.reg = Registration(self.shape, Cn, fp, max_batch=mb, device=device, **kw)
        self.device = self.reg.device
        # OFOptions.cc_initialization / cc_hw / cc_up.  What the REFERENCE pipeline does with the option (checked
        # against the live reference, tests/golden/xcorr_sequence.npz): BatchMotionCorrector starts the w_init chain
        # from a zero field instead of the bootstrap solve (compensate_recording_3D.py:346-356) -- and that is all,
        # because _process_batch_parallel builds flow_params WITHOUT the cc_* keys (:301-315), so its executors never
        # see cc_initialization (sequential_3d.py:63) and run the plain flow.  That is the default here too
        # (cc_zero_start).  cc_prealign=True additionally runs the rigid pre-alignment the executors implement
        # (sequential_3d.py:89-145) for every frame, i.e. what the option is documented to do.
        self.cc_zero_start = bool(getattr(options, "cc_initialization", False))
        self.cc = self.cc_zero_start and bool(cc_prealign)
        self.cc_hw = getattr(options, "cc_hw", 256)
        self.cc_up = int(getattr(options, "cc_up", 10))
        if self.cc and Cn != 1:
            # the reference pipeline fails the same way at its first batch (xcorr_prealignment.py:26-30 contracts the
            # full (Z,Y,X,C) weight array with the channel axis)
            raise ValueError("shape-mismatch for sum: cc_initialization supports single-channel recordings only, "
                             "as in the reference")
        # weights (compensate_recording_3D.py:211-224)
        wvec = [options.get_weight_at(c, Cn) for c in range(Cn)]
        if all(np.ndim(w) == 0 for w in wvec):
            weight = np.ones((1, 1, 1, Cn)) * np.asarray(wvec, float).reshape(1, 1, 1, Cn)
            weight = np.broadcast_to(weight, (Z, Y, X, Cn))
        else:
            weight = np.ones((Z, Y, X, Cn))
            for c in range(Cn):
                weight[..., c] = wvec[c]
        self.lo, self.den = normalization_range(self.reference_raw,
                                                getattr(options, "channel_normalization", "together"))
        ref_dev = dev.to_device(self.reference_raw, self.device)
        # the reference normalises the fixed volume against ITS OWN range (normalization_ref=None);
        # that is the same (lo, den) as above
        self.update_reference = bool(getattr(options, "update_reference", False))
        # frames of temporal-filter reach around a shard (sigma_t >= 0.125 couples the frames of a batch)
        self.temporal_halo = int(max(self.reg.plan.plan.gauss_radius_t[c] for c in range(Cn))) if self.reg.plan.temporal else 0
        self._weight = weight
        self._ref_proc64 = (dev.empty((1, Z, Y, X, Cn), np.float64, self.device)
                            if (self.update_reference or self.cc) else None)
        ref_proc = self.reg.preprocess(ref_dev[None], self.lo, self.den, temporal=False, out64=self._ref_proc64)
        self.reg.set_reference(ref_proc[0], weight=weight, ref_raw=ref_dev)
        self.w_init: Optional[torch.Tensor] = None  # (Z,Y,X,3) float32 on device
        self.collect_statistics = bool(statistics)
        self._stats: List[torch.Tensor] = []        # per batch (t,4) device tensors, see statistics()

    # -- helpers ------------------------------------------------------------------------
    def _flows(self, proc: torch.Tensor, uvw: Optional[torch.Tensor], proc64: Optional[torch.Tensor] = None) -> torch.Tensor:
        if self.cc:
            # sequential_3d.py:89-145; the warps read the float64 pre-processed frames as there
            Z, Y, X = self.shape
            w0 = uvw if uvw is not None else torch.zeros((Z, Y, X, 3), dtype=torch.float32, device=self.device)
            src = proc64 if proc64 is not None else proc
            flows, _ = self.reg.get_displacement_cc(src, w0, cc_hw=self.cc_hw, cc_up=self.cc_up,
                                                    ref_proc64=self._ref_proc64[0, ..., 0])
            return flows
        outs = []
        for t0 in range(0, proc.shape[0], self.reg.max_batch):
            outs.append(self.reg.get_displacement(proc[t0:t0 + self.reg.max_batch], uvw=uvw))
        return outs[0] if len(outs) == 1 else torch.cat(outs, 0)

    def _mean_of(self, flows: Optional[torch.Tensor], count: int) -> torch.Tensor:
        """Mean over `count` frames of the GLOBAL batch, `flows` being this rank's share."""
        Z, Y, X = self.shape
        if self.world == 1:
            return self.reg.mean_frames(flows)
        part = torch.zeros((Z, Y, X, 3), dtype=torch.float32, device=self.device)
        if flows is not None and flows.shape[0] > 0:
            part = self.reg.mean_frames(flows) * float(flows.shape[0])
            if isinstance(self.reg, SplitRegistration):
                self.reg.sync()
            else:
                self.reg.ctx.order_with_torch()      # stream order, no host synchronisation
        torch.distributed.all_reduce(part, op=torch.distributed.ReduceOp.SUM, group=self.group)
        return part / float(count)

    # -- one batch ------------------------------------------------------------------------
    def process_batch(self, raw_local, global_size: Optional[int] = None, local_offset: int = 0,
                      compensate: bool = True, halo_before=None, halo_after=None):
        """raw_local: this rank's frames (t,Z,Y,X,C) of the current batch (ndarray or device tensor);
        global_size / local_offset place them in the global batch (defaults: single process).
        halo_before / halo_after: with a temporal pre-filter (sigma_t >= 0.125) and a batch sharded over ranks, the up
        to `self.temporal_halo` frames of the SAME batch that precede / follow this rank's frames (the filter couples
        the frames of a batch, image_processing_3D.py:140-156; at the ends of the batch scipy's reflection applies).
        Returns device tensors (registered float32 (t,Z,Y,X,C), flows float32 (t,Z,Y,X,3))."""
        raw = self.reg._as_dev(raw_local, None, None)
        t = raw.shape[0]
        G = t if global_size is None else int(global_size)
        want64 = (self.update_reference or self.cc) and t > 0
        nb = 0 if halo_before is None else int(np.shape(halo_before)[0])
        na = 0 if halo_after is None else int(np.shape(halo_after)[0])
        if t > 0 and (nb or na):
            parts = ([self.reg._as_dev(halo_before, None, None)] if nb else []) + [raw] + \
                    ([self.reg._as_dev(halo_after, None, None)] if na else [])
            ext = torch.cat([p_.to(raw.dtype) for p_ in parts], 0)
            e64 = dev.empty(tuple(ext.shape), np.float64, self.device) if want64 else None
            proc = self.reg.preprocess(ext, self.lo, self.den, out64=e64)[nb:nb + t]
            proc64 = e64[nb:nb + t] if want64 else None
        else:
            proc64 = dev.empty(tuple(raw.shape), np.float64, self.device) if want64 else None
            proc = self.reg.preprocess(raw, self.lo, self.den, out64=proc64) if t > 0 else None
        if self.w_init is None and self.cc_zero_start:
            # compensate_recording_3D.py:346-356: with cc_initialization the chain starts from a ZERO field and no
            # bootstrap frames are solved (the rigid estimate of every frame replaces the bootstrap)
            Z, Y, X = self.shape
            self.w_init = torch.zeros((Z, Y, X, 3), dtype=torch.float32, device=self.device)
        if self.w_init is None:
            # bootstrap (compensate_recording_3D.py:359-388): first min(22, G) frames from zero flow
            n_init = min(22, G)
            a, b = max(0, -local_offset), max(0, min(t, n_init - local_offset))
            w0 = self._flows(proc[a:b], None, None if proc64 is None else proc64[a:b]) if b > a else None
            self.w_init = self._mean_of(w0, n_init)
        use_chain = bool(getattr(self.options, "update_initialization_w", True))
        uvw = self.w_init if use_chain else torch.zeros_like(self.w_init)
        flows = self._flows(proc, uvw, proc64) if t > 0 else None
        if use_chain:
            # (:481-485) mean of the last <= 20 flows of the global batch
            first = G - 20 if G > 20 else 0
            a, b = max(0, first - local_offset), t
            self.w_init = self._mean_of(flows[a:b] if (flows is not None and b > a) else None, G - first)
        if t > 0 and self.collect_statistics:
            self._stats.append(self.reg.flow_stats(flows))      # stays on the device until statistics() is read
        # flows are final here: the host pipeline may start fetching them while the compensation warp runs
        self._flows_ready = (torch.cuda.current_stream(self.device).record_event()
                             if self.device.type == "cuda" else None)
        reg = None
        if compensate and t > 0:
            outs = []
            for t0 in range(0, t, self.reg.max_batch):
                outs.append(self.reg.compensate(raw[t0:t0 + self.reg.max_batch], flows[t0:t0 + self.reg.max_batch]))
            reg = outs[0] if len(outs) == 1 else torch.cat(outs, 0)
        if self.update_reference and self.world == 1 and t > 0:
            # (compensate_recording_3D.py:395-429) new fixed volume = mean of the last <= 100 pre-processed frames of
            # the batch warped by their flows (float64 frames in, float32 warps, float64 mean); used from the next
            # batch on.  Out-of-volume voxels take the current pre-processed reference.
            n_ref = min(100, t)
            comp = self.reg.compensate(proc64[t - n_ref:], flows[t - n_ref:], ref_raw=self._ref_proc64[0])
            new64 = self.reg.mean_frames_f64(comp)
            self._ref_proc64 = new64[None]
            self.reg.set_reference(new64.to(torch.float32), weight=self._weight)
        elif self.update_reference:
            # the same over a batch sharded across ranks: every rank warps ITS frames of the global window (the last
            # <= 100 of the batch), the float64 partial sums are all-reduced
            n_ref = min(100, G)
            a, b = max(0, G - n_ref - local_offset), t
            Z, Y, X = self.shape
            part = torch.zeros((Z, Y, X, self.C), dtype=torch.float64, device=self.device)
            if t > 0 and b > a:
                comp = self.reg.compensate(proc64[a:b], flows[a:b], ref_raw=self._ref_proc64[0])
                part = self.reg.mean_frames_f64(comp) * float(b - a)
                self.reg.ctx.order_with_torch()
            torch.distributed.all_reduce(part, op=torch.distributed.ReduceOp.SUM, group=self.group)
            new64 = part / float(n_ref)
            self._ref_proc64 = new64[None]
            self.reg.set_reference(new64.to(torch.float32), weight=self._weight)
        return reg, flows

    # -- host-resident recordings: copy / compute / copy pipeline --------------------------
    def run_pipelined(self, host_batches, out_reg: Optional[list] = None, out_flow: Optional[list] = None,
                      global_sizes: Optional[list] = None, local_offsets: Optional[list] = None,
                      sink: Optional[Callable] = None):
        """Process a sequence of HOST batches (tensors or arrays (t,Z,Y,X,C); pinned memory makes the
        copies asynchronous) with the host->device copy of batch k+1 and the device->host copy of
        batch k-1 running on their own streams under the compute of batch k.  Results land in
        out_reg[k] (t,Z,Y,X,C) float32 and out_flow[k] (t,Z,Y,X,3) float32 host tensors (allocated
        pinned when not given).  With `sink`, sink(k, reg_host, flow_host) is called in order as soon
        as batch k has arrived on the host and only two result buffers are cycled (consume or copy the
        arrays inside the callback).  Returns (out_reg, out_flow) after everything has completed."""
        K = len(host_batches)
        def as_tensor(b):
            if isinstance(b, torch.Tensor):
                return b
            b = np.asarray(b)
            if b.dtype not in _lib._DTYPES:
                b = b.astype(np.float64)
            b = np.ascontiguousarray(b)
            return torch.from_numpy(b if b.flags.writeable else b.copy())
        hb = [as_tensor(b) for b in host_batches]
        Z, Y, X = self.shape
        if self.device.type != "cuda":  # kernel-logic emulator (tests): plain loop
            out_reg = list(out_reg) if out_reg is not None else [None] * K
            out_flow = list(out_flow) if out_flow is not None else [None] * K
            for k in range(K):
                g = None if global_sizes is None else global_sizes[k]
                o = 0 if local_offsets is None else local_offsets[k]
                reg, fl = self.process_batch(hb[k], global_size=g, local_offset=o)
                self.reg.sync()
                t = hb[k].shape[0]
                r_ = reg if reg is not None else torch.empty((0, Z, Y, X, self.C), dtype=torch.float32)
                f_ = fl if fl is not None else torch.empty((0, Z, Y, X, 3), dtype=torch.float32)
                if out_reg[k] is None:
                    out_reg[k] = r_.clone()
                else:
                    out_reg[k][:t].copy_(r_)
                if out_flow[k] is None:
                    out_flow[k] = f_.clone()
                else:
                    out_flow[k][:t].copy_(f_)
                if sink is not None:
                    sink(k, out_reg[k][:t], out_flow[k][:t])
            return out_reg, out_flow
        main = torch.cuda.current_stream(self.device)
        if not hasattr(self, "_s_in"):
            self._s_in, self._s_out = torch.cuda.Stream(self.device), torch.cuda.Stream(self.device)
            self._in_bufs = [None, None]
        s_in, s_out = self._s_in, self._s_out
        out_reg = list(out_reg) if out_reg is not None else [None] * K
        out_flow = list(out_flow) if out_flow is not None else [None] * K
        tmax = max([b.shape[0] for b in hb], default=0)
        # sink mode: two cycling pinned result buffers per kind, kept on the corrector -- run_stream calls this method once
        # per window of the stream, and page-locking fresh host memory costs more than the copies it serves
        # (cudaHostAlloc 2.6 GB/s on the box, results/r02_host_costs.txt)
        cyc = self.__dict__.setdefault("_cyc_out", {})

        def cycling(kind, slot, last):
            buf = cyc.get((kind, slot))
            if buf is None or buf.shape[0] < tmax or tuple(buf.shape[1:]) != (Z, Y, X, last):
                buf = cyc[(kind, slot)] = torch.empty((tmax, Z, Y, X, last), dtype=torch.float32, pin_memory=True)
            return buf

        for k in range(K):
            t = hb[k].shape[0]
            if out_reg[k] is None:
                if sink is not None:
                    out_reg[k] = cycling("reg", k % 2, self.C)
                else:
                    out_reg[k] = torch.empty((t, Z, Y, X, self.C), dtype=torch.float32, pin_memory=True)
            if out_flow[k] is None:
                if sink is not None:
                    out_flow[k] = cycling("flow", k % 2, 3)
                else:
                    out_flow[k] = torch.empty((t, Z, Y, X, 3), dtype=torch.float32, pin_memory=True)
        ev_in, ev_done, ev_out = [None] * K, [None] * K, [None] * K

        def drain(k):
            if sink is not None and k >= 0:
                if ev_out[k] is not None:
                    ev_out[k].synchronize()
                t_ = hb[k].shape[0]
                sink(k, out_reg[k][:t_], out_flow[k][:t_])

        def stage(k):
            with torch.cuda.stream(s_in):
                if k >= 2:
                    s_in.wait_event(ev_done[k - 2])      # buffer k % 2 is free once batch k-2 was consumed
                else:
                    s_in.wait_stream(main)
                buf = self._in_bufs[k % 2]
                if buf is None or buf.shape[1:] != hb[k].shape[1:] or buf.shape[0] < hb[k].shape[0] \
                        or buf.dtype != hb[k].dtype:
                    buf = self._in_bufs[k % 2] = torch.empty(tuple(hb[k].shape), dtype=hb[k].dtype, device=self.device)
                buf[:hb[k].shape[0]].copy_(hb[k], non_blocking=True)
                ev_in[k] = s_in.record_event()

        if K > 0:
            stage(0)
        for k in range(K):
            if k + 1 < K:
                stage(k + 1)
            main.wait_event(ev_in[k])
            t = hb[k].shape[0]
            g = None if global_sizes is None else global_sizes[k]
            o = 0 if local_offsets is None else local_offsets[k]
            reg, fl = self.process_batch(self._in_bufs[k % 2][:t], global_size=g, local_offset=o)
            ev_done[k] = main.record_event()
            # host buffer k % 2 must have been consumed (batch k-2 drained below) before it is refilled
            if t > 0:
                with torch.cuda.stream(s_out):
                    s_out.wait_event(self._flows_ready)        # the flow fields go first, under the compensation warp
                    out_flow[k][:t].copy_(fl, non_blocking=True)
                    s_out.wait_event(ev_done[k])
                    out_reg[k][:t].copy_(reg, non_blocking=True)
                    reg.record_stream(s_out)
                    fl.record_stream(s_out)
                    ev_out[k] = s_out.record_event()
            drain(k - 1)       # blocks the host until batch k-1 is on the host; the GPU keeps computing batch k
        s_out.synchronize()
        main.synchronize()
        drain(K - 1)
        return out_reg, out_flow

    def run_stream(self, batches, sink: Callable, lookahead: int = 1):
        """Streaming form of run_pipelined for recordings that do not fit in host memory: `batches` is any iterable
        of host batches (t,Z,Y,X,C) -- e.g. a generator around the reference's reader,
        `(b for b in iter_batches(reader))` with `while reader.has_batch(): yield reader.read_batch()`
        (util/io/_base_3d.py:230-256); a `None` item ends the stream -- and `sink(k, registered, flows)` receives the
        host tensors of batch k in order as soon as they have arrived (e.g. `writer.write_frames(registered.numpy())`,
        `:339-345`).  Batches are pulled `lookahead` ahead of the one being computed; the copy / compute / copy overlap
        is that of run_pipelined, which is called on a sliding window of the stream."""
        it = iter(batches)
        window, k0 = [], 0

        def pull():
            try:
                b = next(it)
            except StopIteration:
                return False
            if b is None:
                return False
            window.append(b)
            return True

        while len(window) < 1 + lookahead and pull():
            pass
        while window:
            # process everything pulled so far as one pipelined run (H2D of the later batches overlaps the compute of
            # the earlier ones), then refill the window
            chunk, window = window, []
            base = k0
            self.run_pipelined(chunk, sink=lambda k, r, f, base=base: sink(base + k, r, f))
            k0 += len(chunk)
            while len(window) < 1 + lookahead and pull():
                pass

    def statistics(self) -> dict:
        """The per-frame lists BatchMotionCorrector keeps (compensate_recording_3D.py:488-508) for this rank's
        frames so far: mean_disp, max_disp, mean_div, mean_translation (computed on the device from the flows)."""
        if not self._stats:
            return {"mean_disp": [], "max_disp": [], "mean_div": [], "mean_translation": []}
        a = dev.to_host(torch.cat(self._stats, 0))
        return {"mean_disp": a[:, 0].tolist(), "max_disp": a[:, 1].tolist(), "mean_div": a[:, 2].tolist(),
                "mean_translation": a[:, 3].tolist()}

    def close(self):
        self.__dict__.pop("_cyc_out", None)
        if isinstance(self.reg, SplitRegistration):
            self.reg.close()
        else:
            self.reg.ctx.close()


def _prefault(a: np.ndarray, threads: int = 8):
    """Touch every page of a fresh array from several threads (numpy releases the GIL in the strided store), so that
    the device -> host copies that follow do not pay the first-touch page faults one page at a time."""
    flat = a.reshape(-1).view(np.uint8)
    n = flat.size
    if n < (64 << 20):
        return
    from concurrent.futures import ThreadPoolExecutor
    import os
    threads = max(1, min(threads, os.cpu_count() or 1))
    step = -(-n // threads)
    step += (-step) % 4096

    def touch(i):
        flat[i * step:min(n, (i + 1) * step):4096] = 0
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(touch, range(threads)))


def compensate_arr_3D(c1: np.ndarray, c_ref: np.ndarray, options=None,
                      progress_callback: Optional[Callable[[int, int], None]] = None,
                      device: Optional[torch.device] = None, cc_prealign: bool = False):
    """Drop-in for flowreg3d.motion_correction.compensate_arr_3D: returns (registered, flow) with
    registered shaped like c1 and flow (T,Z,Y,X,3) float32.  cc_prealign: see SequenceCorrector (not a reference
    argument; False reproduces what the reference pipeline computes for OFOptions.cc_initialization)."""
    c1 = np.asarray(c1)
    if c_ref is None:
        # not in the reference's signature (c_ref is required there): the fixed volume the options describe, i.e. the
        # mean of the frames listed in options.reference_frames (OF_options_3D.py:496-503)
        c_ref = (OFOptions() if options is None else options).get_reference_frame(c1)
    c_ref = np.asarray(c_ref)
    squeezed = False
    original_shape = c1.shape
    if c1.size == 0:
        raise ValueError("Input array cannot be empty")
    if c1.ndim == 4 and c_ref.ndim == 3:
        c1 = c1[..., None]
        c_ref = c_ref[..., None]
        squeezed = True
    elif c1.ndim == 3:
        c1 = c1[None, :, :, :, None]
        if c_ref.ndim == 3:
            c_ref = c_ref[..., None]
        squeezed = True
    options = OFOptions() if options is None else options.copy()
    T = c1.shape[0]
    seq = SequenceCorrector(c_ref, options, device=device, cc_prealign=cc_prealign)
    bs = int(options.buffer_size)
    bounds = [(b0, min(T, b0 + bs)) for b0 in range(0, T, bs)]
    tn = getattr(options, "output_typename", None)
    out_dt = np.dtype(_OUT_TYPES[tn]) if (tn and tn in _OUT_TYPES) else c1.dtype
    # The reference assigns the float32 result into an array of the INPUT dtype (numpy cast) and converts that to
    # output_typename at the end (sequential_3d.py:163-169, compensate_arr_3D.py).  For floating-point inputs both
    # casts are exact widenings / roundings of the float32 values, so the final array is filled directly (the widening
    # happens on the device, one D2H copy into the final array); integer inputs keep the two-step host path.
    direct = c1.dtype in (np.float32, np.float64) and out_dt in (np.dtype(np.float32), np.dtype(np.float64))
    registered = np.empty(c1.shape, out_dt if direct else c1.dtype)
    w = np.empty((T,) + seq.shape + (3,), np.float32)
    # The host side of this call is bound by fresh pageable memory: measured on the B200 box (tools/host_costs.py)
    # first touch 5.5 GB/s per thread, cudaHostAlloc 2.6 GB/s, D2H into touched pageable memory 18 GB/s, into pinned
    # 53 GB/s.  So: no pinned staging (allocating it costs more than it saves for a one-shot call), the result arrays
    # are pre-faulted by several threads, and the copies go straight from / to the caller's arrays.
    _prefault(registered)
    _prefault(w)
    done = 0
    try:
        for k, (b0, b1) in enumerate(bounds):
            reg_d, fl_d = seq.process_batch(c1[b0:b1])      # pageable H2D inside (device.to_device)
            seq.reg.ctx.order_with_torch()
            torch.from_numpy(w[b0:b1]).copy_(fl_d)
            if direct:
                torch.from_numpy(registered[b0:b1]).copy_(reg_d if registered.dtype == np.float32
                                                          else reg_d.to(torch.float64))
            else:
                registered[b0:b1] = dev.to_host(reg_d)      # numpy cast to the input dtype
            done += b1 - b0
            if progress_callback is not None:
                try:
                    progress_callback(done, T)
                except Exception as e:  # compensate_recording_3D.py:158-162
                    import warnings
                    warnings.warn(f"Progress callback error: {e}")
    finally:
        seq.close()
    if not direct and tn and tn in _OUT_TYPES:
        registered = registered.astype(_OUT_TYPES[tn])
    if squeezed:
        if len(original_shape) == 3:
            registered = np.squeeze(registered)
            w = np.squeeze(w, axis=0)
        elif len(original_shape) == 4:
            registered = np.squeeze(registered, axis=-1)
    return registered, w


def compensate_arr_3D_sharded(c1: np.ndarray, c_ref: np.ndarray, options=None, group=None,
                              device: Optional[torch.device] = None, cc_prealign: bool = False):
    """Multi-GPU variant: every rank passes the same (T,Z,Y,X,C) array (or a memory map of it) and
    gets back ITS shard: (registered, flow, frame_indices).  Results are identical in layout to
    compensate_arr_3D restricted to frame_indices."""
    c1 = np.asarray(c1)
    c_ref = np.asarray(c_ref)
    if c1.ndim == 4:
        c1 = c1[..., None]
    if c_ref.ndim == 3:
        c_ref = c_ref[..., None]
    options = OFOptions() if options is None else options.copy()
    seq = SequenceCorrector(c_ref, options, device=device, group=group, cc_prealign=cc_prealign)
    regs, flows, idx = [], [], []
    T = c1.shape[0]
    try:
        for b0 in range(0, T, int(options.buffer_size)):
            b1 = min(T, b0 + int(options.buffer_size))
            lo, hi = shard_bounds(b1 - b0, seq.world, seq.rank)
            hr = seq.temporal_halo
            before = c1[b0 + max(0, lo - hr):b0 + lo] if (hr and hi > lo and lo > 0) else None
            after = c1[b0 + hi:min(b1, b0 + hi + hr)] if (hr and hi > lo and hi < b1 - b0) else None
            reg, fl = seq.process_batch(c1[b0 + lo:b0 + hi], global_size=b1 - b0, local_offset=lo,
                                        halo_before=before, halo_after=after)
            seq.reg.sync()
            if hi > lo:
                regs.append(dev.to_host(reg).copy())
                flows.append(dev.to_host(fl).copy())
                idx.extend(range(b0 + lo, b0 + hi))
    finally:
        seq.close()
    Z, Y, X = seq.shape
    if regs:
        return np.concatenate(regs, 0), np.concatenate(flows, 0), np.asarray(idx)
    return (np.empty((0, Z, Y, X, seq.C), np.float32), np.empty((0, Z, Y, X, 3), np.float32), np.asarray(idx, int))
