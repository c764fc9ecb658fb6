"""Single large volume on several GPUs: the sweep-pipelined level solve (get_displacement_pipelined) and the
z-slab decomposition with halo exchange (get_displacement_zslab, at the end of the file).

Frames of a recording shard trivially (compensate.py).  ONE volume that is too slow on one GPU does
not: the reference's solver is a lexicographic Gauss-Seidel sweep, so a z-slab decomposition has to
exchange the slab faces after EVERY wave of the wavefront schedule, in both directions.
The dependency structure offers something cheaper: sweep t+1 of a hyperplane needs nothing but sweep t
of that hyperplane and of its two neighbours.  So the T sweeps are split over the ranks instead of the
space: rank r runs sweeps [t_r, t_{r+1}) of the global schedule  q = (k+j+i) + 2t  over the WHOLE level,
and the increments stream rank r -> r+1 as contiguous hyperplane blocks (solver storage is
hyperplane-major) while both keep sweeping -- a one-directional NVLink pipeline with a handful of
NCCL send/recv per level, no per-wave synchronisation between GPUs, and exactly the reference's
update order (the result is bit-identical to the single-GPU solve).

The 5x5x5 median of the increments (a fifth of the time of a large level) is z-slab parallel: every rank
holds the finished increments, filters its own planes and the flow slabs are exchanged.  Everything else
around the solve (pyramid, warp, assembly) is computed redundantly on every rank from the same inputs -- it
is deterministic, so all ranks stay identical without communication; only levels large enough to be
bandwidth-bound are pipelined, small ones are solved redundantly too.

    reg  = Registration(shape, C, params, max_batch=1)        # on every rank, same arguments
    reg.set_reference(fixed_proc)
    flow = get_displacement_pipelined(reg, moving_proc)        # (B,Z,Y,X,3) device tensor on every rank
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, device as dev
from .core import Registration, _check


def sweep_partition(T: int, lag: int, world: int) -> List[Tuple[int, int]]:
    """Split sweeps [0,T) into <= world contiguous ranges whose starts are multiples of `lag`
    (a range must begin with a psi refresh).  Ranks beyond the number of lag blocks get (T, T)."""
    blocks = -(-T // lag)
    base, rem = divmod(blocks, world)
    out, at = [], 0
    for r in range(world):
        nb = base + (1 if r < rem else 0)
        t0, t1 = min(T, at * lag), min(T, (at + nb) * lag)
        out.append((t0, t1))
        at += nb
    return out


def pipeline_schedule(S: int, T: int, parts: List[Tuple[int, int]], n_chunks: int):
    """Stages of every active rank: list (per rank) of (recv (h0,h1) | None, q_begin, q_end, send (h0,h1) | None),
    hyperplane ranges.  Rank r's receives are rank r-1's sends, in order."""
    active = [r for r, (a, b) in enumerate(parts) if b > a]
    sched = {}
    prev_sends: Optional[list] = None
    for idx, r in enumerate(active):
        t0, t1 = parts[r]
        last = idx == len(active) - 1
        q_first, q_last_end = 2 * t0, (S - 1) + 2 * (t1 - 1) + 1
        stages, q_done, sent = [], q_first, 0
        if prev_sends is None:
            ends = sorted({max(1, min(S, round(c * S / n_chunks))) for c in range(1, n_chunks + 1)})
            inputs = [(None, h) for h in ends]           # no receive; stage boundary = hyperplanes final after it
        else:
            inputs = [((a, b), b) for (a, b) in prev_sends]
        for k, (rcv, h_in) in enumerate(inputs):
            final_stage = k == len(inputs) - 1
            if prev_sends is None:
                q_end = h_in - 1 + 2 * (t1 - 1) + 1
            else:
                q_end = h_in + 2 * t0 - 1              # waves that only touch hyperplanes < h_in (and read < h_in)
            q_end = q_last_end if final_stage else max(q_done, min(q_last_end, q_end))
            fin = max(0, min(S, q_end - 2 * (t1 - 1)))   # hyperplanes final after waves < q_end
            if final_stage:
                fin = S
            snd = None
            if not last and fin > sent:
                snd = (sent, fin)
                sent = fin
            stages.append((rcv, q_done, q_end, snd))
            q_done = q_end
        sched[r] = stages
        prev_sends = [s[3] for s in stages if s[3] is not None]
    return active, sched


def _level_info(reg: Registration, li: int):
    lib, h = reg.ctx.lib, reg.ctx.h
    size = (C.c_int32 * 3)()
    S = C.c_int32()
    nslots = C.c_int64()
    _check(h, lib.fr3d_level_info(h, li, size, C.byref(S), C.byref(nslots), None))
    start = np.empty(S.value + 1, np.int32)
    _check(h, lib.fr3d_level_info(h, li, size, C.byref(S), C.byref(nslots), start.ctypes.data))
    return tuple(size), int(S.value), int(nslots.value), start


def _require_ctx_stream(reg: Registration):
    """The NCCL calls below are ordered against the library's kernels by STREAM ORDER: both must run on the stream the
    Registration was created under."""
    if reg.device.type == "cuda" and torch.cuda.current_stream(reg.device).cuda_stream != reg.ctx.stream_handle:
        raise RuntimeError("call the multi-GPU solves under the CUDA stream the Registration was created on "
                           "(torch.cuda.current_stream() differs from the context's stream)")


def get_displacement_pipelined(reg: Registration, moving_proc, uvw=None, group=None, out_dtype=np.float32,
                               min_slots: int = 1 << 21, n_chunks: int = 8,
                               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """get_displacement for B frames with the level solves pipelined over the ranks of `group`.
    Every rank must call with identical arguments; every rank returns the full result.
    min_slots: levels with fewer solver slots (x frames) are solved redundantly without communication."""
    _require_ctx_stream(reg)
    lib, h = reg.ctx.lib, reg.ctx.h
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    mv = reg._as_dev(moving_proc, np.float32, None)
    if mv.dim() == 4:
        mv = mv[None]
    B = mv.shape[0]
    uv = None if uvw is None else reg._as_dev(uvw, np.float32, reg.shape + (3,))
    if out is None:
        out = dev.empty((B,) + reg.shape + (3,), out_dtype, reg.device)
    T, lag = int(reg.plan.plan.iterations), int(reg.plan.plan.update_lag)
    state_dt = np.float64 if reg.plan.plan.state_dtype == _lib.F64 else np.float32
    vlen = 3 if state_dt == np.float64 else 4      # components per state vector (float32 vectors are padded to 16 bytes)
    pipelinable = reg.plan.plan.sweep == 0 and float(reg.plan.plan.a_smooth) == 1.0
    nl = lib.fr3d_level_count(h)
    keep = []
    if world > 1 and pipelinable:
        _pair_group(group, world, 0, 1)         # collective: every rank builds the pair communicators
    for li in range(nl):
        _check(h, lib.fr3d_level_begin(h, li, dev.ptr(mv), dev.ptr(uv), B))
        _, S, nslots, start = _level_info(reg, li)
        parts = sweep_partition(T, lag, world)
        active = [r for r, (a, b) in enumerate(parts) if b > a]
        if world == 1 or not pipelinable or len(active) < 2 or nslots * B < min_slots:
            _check(h, lib.fr3d_level_sweeps(h, li, -1, -1, -1, -1))      # redundant on every rank, no exchange
        else:
            active, sched = pipeline_schedule(S, T, parts, n_chunks)
            if rank in sched:
                t0, t1 = parts[rank]
                pos = active.index(rank)
                src = active[pos - 1] if pos > 0 else None
                dst = active[pos + 1] if pos + 1 < len(active) else None
                # one communicator per neighbouring pair: the receives from rank-1 and the sends to rank+1 then
                # progress independently (on one communicator NCCL would serialise them into a lock-step chain);
                # all receives of the level are posted up front
                g_in = _pair_group(group, world, src, rank) if src is not None else None
                g_out = _pair_group(group, world, rank, dst) if dst is not None else None
                pending = []
                for rcv, _, _, _ in sched[rank]:
                    if rcv is not None:
                        a, b = int(start[rcv[0]]), int(start[rcv[1]])
                        buf = dev.empty((B, b - a, vlen), state_dt, reg.device)
                        pending.append((buf, dist.irecv(buf, src=_global_rank(group, src), group=g_in)))
                        keep.append(buf)
                for rcv, q0, q1, snd in sched[rank]:
                    if rcv is not None:
                        a, b = int(start[rcv[0]]), int(start[rcv[1]])
                        buf, work = pending.pop(0)
                        work.wait()
                        _check(h, lib.fr3d_level_state(h, li, 1, dev.ptr(buf), a, b))
                    _check(h, lib.fr3d_level_sweeps(h, li, t0, t1, q0, q1))
                    if snd is not None:
                        a, b = int(start[snd[0]]), int(start[snd[1]])
                        buf = dev.empty((B, b - a, vlen), state_dt, reg.device)
                        _check(h, lib.fr3d_level_state(h, li, 0, dev.ptr(buf), a, b))
                        keep.append((buf, dist.isend(buf, dst=_global_rank(group, dst), group=g_out)))
            # the last active rank holds the finished increments: hand them to everybody
            full = dev.empty((B, nslots, vlen), state_dt, reg.device)
            if rank == active[-1]:
                _check(h, lib.fr3d_level_state(h, li, 0, dev.ptr(full), 0, nslots))
            dist.broadcast(full, src=_global_rank(group, active[-1]), group=group)
            if rank != active[-1]:
                _check(h, lib.fr3d_level_state(h, li, 1, dev.ptr(full), 0, nslots))
            keep.append(full)
            # z-slab parallel median / accumulation: own planes, then one broadcast per rank's slab
            (pz, py, px) = _level_info(reg, li)[0]
            bounds = [_split(pz, world, r) for r in range(world)]
            z0, z1 = bounds[rank]
            _check(h, lib.fr3d_level_end_range(h, li, z0, z1))
            for r, (a, b) in enumerate(bounds):
                if b <= a:
                    continue
                slab = dev.empty((B, 3, b - a, py, px), np.float64, reg.device)
                if r == rank:
                    _check(h, lib.fr3d_flow_slab(h, li, 0, dev.ptr(slab), a, b))
                dist.broadcast(slab, src=_global_rank(group, r), group=group)
                if r != rank:
                    _check(h, lib.fr3d_flow_slab(h, li, 1, dev.ptr(slab), a, b))
                keep.append(slab)
            continue
        _check(h, lib.fr3d_level_end(h, li))
    _check(h, lib.fr3d_flow_finish(h, dev.ptr(out), reg._code(out)))
    reg._keep = [mv, uv, keep]
    return out


_PAIR_GROUPS: dict = {}


def _pair_group(group, world: int, a: int, b: int):
    """Process group of the neighbouring ranks (a, b).  torch.distributed.new_group is collective over the DEFAULT
    group: the groups of ALL neighbouring pairs are created together, once, in the same order, and EVERY rank of the
    default group must make the first call for a given `group` (a proper subgroup used by only some ranks would hang
    there -- create the Registration's pair groups from all ranks first).  Cached per set of member ranks."""
    key = tuple(_global_rank(group, r) for r in range(world))
    if key not in _PAIR_GROUPS:
        pairs = {}
        for r in range(world - 1):
            ranks = [_global_rank(group, r), _global_rank(group, r + 1)]
            pairs[(r, r + 1)] = dist.new_group(ranks=ranks)
        _PAIR_GROUPS[key] = pairs
    return _PAIR_GROUPS[key][(a, b)]


def _split(n: int, world: int, rank: int) -> Tuple[int, int]:
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _global_rank(group, r: int) -> int:
    return r if group is None else dist.get_global_rank(group, r)


# ------------------------------------------------------------------------------------------------------------------
# z-slab decomposition with halo exchange (the decomposition SURVEY.md 8(e) / the north star names).
# ------------------------------------------------------------------------------------------------------------------
class _SlabPeers:
    """CUDA-IPC plumbing of the device-side halo exchange, set up once per Registration: this rank's increment array
    (reserved for the largest level, so that it never moves) and its flag words are exported, the handles all-gathered,
    and the two z-neighbours' buffers mapped (peer access over NVLink).  The mappings live as long as the
    Registration."""

    def __init__(self, reg: Registration, group):
        self.reg, self.group = reg, group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.flag_base = 0
        self.flags = self._exchange(1)            # neighbour rank -> mapped pointer
        self.incr = self._exchange(0)
        self._token = torch.zeros(1, dtype=torch.int32, device=reg.device)
        # the mappings are released when the context is closed (before the library frees this rank's own buffers)
        reg.ctx.__dict__.setdefault("_on_close", []).append(self.close)

    def _exchange(self, which: int) -> dict:
        """Export this rank's buffer `which`, all-gather the handles, map the two z-neighbours'."""
        lib, h = self.reg.ctx.lib, self.reg.ctx.h
        mine = (C.c_ubyte * 64)()
        _check(h, lib.fr3d_ipc_export(h, which, mine))
        t = torch.tensor(list(mine), dtype=torch.uint8, device=self.reg.device)
        allh = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(allh, t, group=self.group)
        out = {}
        for r in (self.rank - 1, self.rank + 1):
            if 0 <= r < self.world:
                raw = (C.c_ubyte * 64)(*allh[r].cpu().tolist())
                ptr = C.c_void_p()
                _check(h, lib.fr3d_ipc_open(h, raw, C.byref(ptr)))
                out[r] = ptr.value
        return out

    def stream_rendezvous(self):
        """Stream-ordered rendezvous (no host synchronisation): kernels enqueued after it start only when every
        rank's stream has passed this point, i.e. when every rank has finished preparing (zeroing) its arrays."""
        dist.all_reduce(self._token, group=self.group)

    def close(self):
        lib, h = self.reg.ctx.lib, self.reg.ctx.h
        for p in list(self.flags.values()) + list(self.incr.values()):
            lib.fr3d_ipc_close(h, C.c_void_p(p))
        self.flags, self.incr = {}, {}


def get_displacement_zslab(reg: Registration, moving_proc, uvw=None, group=None, out_dtype=np.float32,
                           min_slots: int = 1 << 21, out: Optional[torch.Tensor] = None,
                           p2p: Optional[bool] = None, stats: Optional[dict] = None) -> torch.Tensor:
    """get_displacement for B frames with every large level solved on z-slabs: rank r owns the planes
    [z_r, z_{r+1}) of the level and runs, wave by wave, the sweeps of its planes only
    (`fr3d_level_sweeps_slab`).  A voxel of wave q = (k+j+i) + 2t reads nothing newer than wave q-1, so after every
    wave the ranks trade, with both z-neighbours, the cells of their two boundary planes that the wave updated (one
    anti-diagonal per sweep in flight, `fr3d_level_wave_cells`; send/recv over NCCL, gloo in the CPU tests) --
    S + 2(T-1) exchanges per level, both directions.  At the end of
    the level the slabs are gathered, the 5^3 median runs on the same z-slabs and the flow slabs are exchanged as in
    the sweep-pipelined solve.  The update order is the reference's: the result is bit-identical to one GPU.

    NOTE on memory: every rank still allocates and assembles the WHOLE level (J, system, increments) and gathers all
    increments before the median -- the decomposition divides the solver's time, not its memory; a slab-sized
    assembly is not built.  Every rank must call with identical arguments; every rank returns the full result.

    p2p (default: on CUDA devices): the halo exchange runs INSIDE the persistent solver kernel -- the ranks map each
    other's increment arrays through CUDA IPC, a boundary-plane voxel is stored into the z-neighbour's memory as it
    is produced (NVLink peer store) and one flag word per neighbour and wave replaces the message pair
    (`fr3d_level_sweeps_slab_p2p`): one launch per level.  p2p=False is the host-driven exchange described above
    (the only one the CPU emulator / gloo tests can run).
    stats: optional dict that receives wall-clock milliseconds per phase (begin = pyramid level, warp, assembly --
    replicated on every rank; sweeps; gather = slabs of the increments to every rank; end = median on own planes +
    flow slabs); measuring synchronises after every phase."""
    _require_ctx_stream(reg)
    import time as _time

    def _mark(name, t0):
        if stats is not None:
            reg.sync()
            stats[name] = stats.get(name, 0.0) + (_time.perf_counter() - t0) * 1e3
        return _time.perf_counter()
    lib, h = reg.ctx.lib, reg.ctx.h
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    mv = reg._as_dev(moving_proc, np.float32, None)
    if mv.dim() == 4:
        mv = mv[None]
    B = mv.shape[0]
    uv = None if uvw is None else reg._as_dev(uvw, np.float32, reg.shape + (3,))
    if out is None:
        out = dev.empty((B,) + reg.shape + (3,), out_dtype, reg.device)
    T = int(reg.plan.plan.iterations)
    state_dt = np.float64 if reg.plan.plan.state_dtype == _lib.F64 else np.float32
    vlen = 3 if state_dt == np.float64 else 4      # components per state vector (float32 vectors are padded to 16 bytes)
    slabbable = reg.plan.plan.sweep == 0 and float(reg.plan.plan.a_smooth) == 1.0
    nl = lib.fr3d_level_count(h)
    keep = []
    if p2p is None:
        p2p = reg.device.type == "cuda"
    peers = None
    if p2p and world > 1 and slabbable:
        peers = getattr(reg, "_slab_peers", None)
        if peers is None or peers.group is not group:
            peers = reg._slab_peers = _SlabPeers(reg, group)
    for li in range(nl):
        tm = _time.perf_counter()
        _check(h, lib.fr3d_level_begin(h, li, dev.ptr(mv), dev.ptr(uv), B))
        (pz, py, px), S, nslots, _ = _level_info(reg, li)
        tm = _mark("begin", tm)
        if world == 1 or not slabbable or pz < world or nslots * B < min_slots:
            _check(h, lib.fr3d_level_sweeps(h, li, -1, -1, -1, -1))      # redundant on every rank, no exchange
            tm = _mark("sweeps_replicated", tm)
            _check(h, lib.fr3d_level_end(h, li))
            tm = _mark("end_replicated", tm)
            continue
        bounds = [_split(pz, world, r) for r in range(world)]
        z0, z1 = bounds[rank]
        plane = py * px

        def planes_out(k0, k1):
            buf = dev.empty((B, (k1 - k0) * plane, vlen), state_dt, reg.device)
            _check(h, lib.fr3d_level_planes(h, li, 0, dev.ptr(buf), k0, k1))
            reg.ctx.order_with_torch()                       # the buffer leaves through torch.distributed
            return buf

        def planes_in(buf, k0, k1):
            _check(h, lib.fr3d_level_planes(h, li, 1, dev.ptr(buf), k0, k1))

        def cells_out(k, q):
            buf = torch.zeros((B, T * py, vlen), dtype=dev.torch_dtype(state_dt), device=reg.device)
            _check(h, lib.fr3d_level_wave_cells(h, li, 0, dev.ptr(buf), k, q))
            reg.ctx.order_with_torch()                       # the buffer leaves through torch.distributed
            return buf

        lo = _global_rank(group, rank - 1) if rank > 0 else None
        hi = _global_rank(group, rank + 1) if rank + 1 < world else None
        if peers is not None:
            # device-side exchange: one launch for all waves of the level
            vp = C.c_void_p
            peers.stream_rendezvous()
            _check(h, lib.fr3d_level_sweeps_slab_p2p(
                h, li, z0, z1, vp(peers.incr.get(rank - 1)), vp(peers.incr.get(rank + 1)),
                vp(peers.flags.get(rank - 1)), vp(peers.flags.get(rank + 1)), peers.flag_base))
            peers.flag_base += S + 2 * (T - 1)
        for q in range(S + 2 * (T - 1) if peers is None else 0):
            _check(h, lib.fr3d_level_sweeps_slab(h, li, q, q + 1, z0, z1))
            # halo: only the cells of the boundary planes that this wave updated (one anti-diagonal per sweep in
            # flight: <= T*py cells instead of py*px)
            ops, recvs = [], []
            if lo is not None:
                sb = cells_out(z0, q)
                rb = dev.empty((B, T * py, vlen), state_dt, reg.device)
                ops += [dist.P2POp(dist.isend, sb, lo, group), dist.P2POp(dist.irecv, rb, lo, group)]
                recvs.append((rb, z0 - 1))
                keep.append(sb)
            if hi is not None:
                sb = cells_out(z1 - 1, q)
                rb = dev.empty((B, T * py, vlen), state_dt, reg.device)
                ops += [dist.P2POp(dist.isend, sb, hi, group), dist.P2POp(dist.irecv, rb, hi, group)]
                recvs.append((rb, z1))
                keep.append(sb)
            for w in dist.batch_isend_irecv(ops):
                w.wait()
            for rb, k in recvs:
                _check(h, lib.fr3d_level_wave_cells(h, li, 1, dev.ptr(rb), k, q))
            keep = keep[-8:]
        tm = _mark("sweeps_slab", tm)
        # gather the slabs of the finished increments: every rank needs them around its planes for the median
        for r, (a, b) in enumerate(bounds):
            buf = planes_out(a, b) if r == rank else dev.empty((B, (b - a) * plane, vlen), state_dt, reg.device)
            dist.broadcast(buf, src=_global_rank(group, r), group=group)
            if r != rank:
                planes_in(buf, a, b)
            keep.append(buf)
        tm = _mark("gather", tm)
        _check(h, lib.fr3d_level_end_range(h, li, z0, z1))
        for r, (a, b) in enumerate(bounds):
            slab = dev.empty((B, 3, b - a, py, px), np.float64, reg.device)
            if r == rank:
                _check(h, lib.fr3d_flow_slab(h, li, 0, dev.ptr(slab), a, b))
                reg.ctx.order_with_torch()
            dist.broadcast(slab, src=_global_rank(group, r), group=group)
            if r != rank:
                _check(h, lib.fr3d_flow_slab(h, li, 1, dev.ptr(slab), a, b))
            keep.append(slab)
        tm = _mark("end_slab", tm)
    tm = _time.perf_counter()
    _check(h, lib.fr3d_flow_finish(h, dev.ptr(out), reg._code(out)))
    _mark("finish", tm)
    reg._keep = [mv, uv, keep]
    return out
