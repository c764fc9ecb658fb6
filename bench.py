#!/usr/bin/env python
"""bench.py -- throughput of the flowreg3D hot path on B200 (BASELINE.json metric: volumes/s on a
2-channel 32x512x512 recording, fixed reference volume, frames sharded over the GPUs of one box).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one batch of B frames per GPU through the whole path: normalise + Gaussian pre-filter,
pyramid, per-level cubic warp / motion tensor / wavefront SOR / 5^3 median, flow up-sampling and the
cubic compensation warp of both channels, with the reference's w_init chaining between batches
(OFOptions defaults: alpha 0.25, 100 iterations, update_lag 5, min_level 5, eta 0.8).

  value  frames/s with the raw frames already resident in HBM (kernel-side throughput)
  e2e    frames/s through the public API with pinned HOST buffers: H2D of the raw frames and D2H of
         the registered frames and flow fields inside the timed region
  roofline / cpu_baseline / clocks / gpu_launches: see DESIGN.md "Measurement".

--impl reference times the reference's CPU implementation of the same path (its restatement in
oracle/, the reference being Python and unable to travel) on the host cores, one frame per worker
process per step.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

SHAPE = (32, 512, 512)
CHANNELS = 2
WORKLOAD = "config2: 2-channel 32x512x512 recording, fixed reference, OFOptions defaults (min_level 5 -> 2 levels, 100 it)"


# ------------------------------------------------------------------------------------------------
# synthetic inputs (numpy/scipy only; SURVEY.md 8(d) recipe)
# ------------------------------------------------------------------------------------------------
def make_reference():
    from tests_inputs import synth_volume
    return np.stack([synth_volume(SHAPE, 10 + c) for c in range(CHANNELS)], -1)


def lowres_flow(seed, magnitude=2.0):
    rng = np.random.default_rng(seed)
    f = rng.standard_normal((3, 4, 8, 8)).astype(np.float32)
    return f * (magnitude / np.abs(f).max())


FRAMES = ("frame_t = backwarp(R, -g_t, linear) + 0.01 N(0,1); g_t = a 4x8x8 random field (<= 2 voxels, numpy "
          "default_rng(1000 + t)) brought to 32x512x512 by the reference's fused Gauss-cubic resize; noise from the torch "
          "CPU generator seeded 3000 + t -- the same recipe, seeds and frame indices in both arms")


def frame_noise(seed):
    """0.01 N(0,1) of one frame, identical in both arms (torch CPU generator)."""
    import torch
    g = torch.Generator(device="cpu")
    g.manual_seed(3000 + seed)
    return 0.01 * torch.randn(SHAPE + (CHANNELS,), generator=g, dtype=torch.float32)


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path, one frame per worker process
# ------------------------------------------------------------------------------------------------
_CPU = {}


def _cpu_init(ref, tmpdir):
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    import torch
    torch.set_num_threads(1)
    from oracle import oracle as O
    O.build()
    _CPU["O"] = O
    _CPU["ref"] = ref
    _CPU["tmp"] = tmpdir
    sigma = np.array([[1.0, 1.0, 1.0, 0.1]] * CHANNELS)
    _CPU["sigma"] = sigma
    _CPU["ref_proc"] = O.preprocess(ref, sigma)          # once per recording, as the reference does


def _cpu_prepare(t):
    """Frame t of the workload (untimed): the recipe of FRAMES with the oracle's resize and linear warp."""
    O, ref = _CPU["O"], _CPU["ref"]
    lr = lowres_flow(1000 + t)
    g = [O.imresize_fused_gauss_cubic3D(lr[q].astype(np.float64), SHAPE) for q in range(3)]
    frame = O.imregister_wrapper(ref, -g[0], -g[1], -g[2], ref, "linear")
    frame = (np.asarray(frame, np.float32) + frame_noise(1000 + t).numpy()).astype(np.float32)
    np.save(os.path.join(_CPU["tmp"], f"frame{t}.npy"), frame)
    return t


def _cpu_frame(t):
    O = _CPU["O"]
    ref = _CPU["ref"]
    frame = np.load(os.path.join(_CPU["tmp"], f"frame{t}.npy"))
    mp_ = O.preprocess(frame[None], _CPU["sigma"], ref)[0]
    flow = O.get_displacement(_CPU["ref_proc"], mp_, alpha=(0.25,) * 3, update_lag=5, iterations=100,
                              min_level=5, levels=100, eta=0.8, a_smooth=1.0, a_data=0.45,
                              weight=np.full(ref.shape, 0.5)).astype(np.float32)
    reg = O.imregister_wrapper(frame, flow[..., 0], flow[..., 1], flow[..., 2], ref, "cubic")
    return float(reg[3, 5, 7, 0])


def cpu_cores():
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        import psutil
        n = min(n, max(1, int(psutil.virtual_memory().available // (6 << 30))))  # ~6 GB per worker
    except Exception:
        pass
    return max(1, min(n, 64))


def run_cpu(steps, warmup, cores=None):
    """Returns (frames/s, cores, seconds per step).  Each step: frames 0 .. cores-1 of the workload, one per worker."""
    import shutil
    import tempfile
    cores = cores or cpu_cores()
    ref = make_reference().astype(np.float64)
    tmp = tempfile.mkdtemp(prefix="fr3d_bench_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    ctx = mp.get_context("fork")
    try:
        with ctx.Pool(cores, initializer=_cpu_init, initargs=(ref, tmp)) as pool:
            pool.map(_cpu_prepare, range(cores))             # synthetic inputs, not timed
            for _ in range(warmup):
                pool.map(_cpu_frame, range(cores))
            t0 = time.perf_counter()
            for _ in range(steps):
                pool.map(_cpu_frame, range(cores))
            dt = time.perf_counter() - t0
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return cores * steps / dt, cores, dt / max(steps, 1)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def algorithmic_bytes(name, plan_levels, B, C, iters, lag):
    """ALGORITHMIC HBM bytes of ONE launch of kernel `name`."""
    Z, Y, X = SHAPE
    NF = Z * Y * X
    if name in ("fr3d_sor_wavefront", "fr3d_sor_staged"):
        # parity-safe storage: 108 B / voxel / sweep + (12C+84) B per psi refresh; mean over the levels
        per = [B * n * (108 * iters + (12 * C + 84) * -(-iters // lag)) for n in plan_levels]
        return float(np.mean(per))
    if "WarpGather" in name:
        # cubic compensation warp (the full-resolution launch dominates): flow 12 B + C*(s_in + 12) + C*s_out
        return float(B * NF * (12 + C * (4 + 12) + C * 4))
    if "Spline" in name:
        return float(B * C * NF * (4 + 8) if "SplineZ" in name else B * C * NF * 16)
    if "PreZ" in name:
        return float(B * C * NF * (4 + 8))
    if "PreY" in name:
        return float(B * C * NF * 16)
    if "PreX" in name:
        return float(B * C * NF * (8 + 4))
    if "ResizePass" in name:
        return float(B * C * NF * 4)
    return None


def measured_traffic(kernel, state, B, level_n):
    """ncu `dram__bytes_read.sum + dram__bytes_write.sum` per launch of the dominant kernel, from the summary of an
    `ncu --set full` capture of THIS build (profiles/r02_sor_traffic.json, written by tools/ncu_traffic.py from the
    .ncu-rep): bytes per frame and level voxel for the kernel / solver state it was captured with; None if the
    profile does not match what is being benched."""
    f = ROOT / "profiles" / "r02_sor_traffic.json"
    if not f.exists():
        return None, None
    t = json.loads(f.read_text()).get(state)
    if not t or t.get("kernel") != kernel:
        return None, None
    return round(t["dram_bytes_per_frame_voxel"] * B * float(np.mean(level_n))), t.get("source")


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from flowreg3d_b200 import build as fr3d_build
    fr3d_build.ensure_built(int(os.environ.get("LOCAL_RANK", "0")))     # no-op when libfr3d.so is present
    import flowreg3d_b200 as F
    from flowreg3d_b200 import core, device as dev

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world and world == 1 and args.gpus > 1:
        raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        # first thing, before any CUDA call in this process: the worker pool forks
        v, cores, sec = run_cpu(steps=1, warmup=0)
        cpu = {"value": round(v, 4), "unit": "volumes/s", "cores": cores, "kind": "port",
               "sample": f"{cores} frames of the same workload (one per worker process, {sec:.1f} s): oracle "
                         "port of the reference CPU path (pre-filter + get_displacement + cubic warp)"}

    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    bound = dev.bind_host_to_gpu(device) if world > 1 else []   # NUMA-local pinned buffers for the e2e path
    if world > 1:
        import datetime
        # NCCL_DEBUG=VERSION makes NCCL print its version banner on STDOUT, ahead of the one JSON line this
        # script owes the driver; keep warnings, drop the banner (INFO etc. are left alone when asked for)
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=device, timeout=datetime.timedelta(minutes=20))

    B = args.batch
    Z, Y, X = SHAPE
    C = CHANNELS
    core.STATE_DTYPE = {"f32": np.float32, "f64": np.float64, "auto": "auto"}[args.state]
    core.SWEEP = 1 if args.sweep == "redblack" else 0
    ref = make_reference()
    opts = F.OFOptions(buffer_size=B)        # defaults: alpha .25, 100 it, lag 5, min_level 5, cubic, weight [.5,.5]
    seq = F.SequenceCorrector(ref, opts, max_batch=B, device=device, group=None, streams=args.streams)
    reg = seq.reg
    ctxs = [p_.ctx for p_ in reg.parts] if hasattr(reg, "parts") else [reg.ctx]

    def launches_now():
        return sum(c_.launches for c_ in ctxs)

    def profile_all(on):
        for c_ in ctxs:
            c_.profile(on)

    def profile_report_all():
        merged = {}
        for c_ in ctxs:
            for name, (cnt, tot) in c_.profile_report().items():
                a_, b_ = merged.get(name, (0, 0.0))
                merged[name] = (a_ + cnt, b_ + tot)
        return merged

    # synthetic frames (not timed), the recipe of FRAMES: the resize and the linear warp run through the library's own
    # kernels (bit-equal / <= 1 ulp to the oracle's, which the CPU arm uses), the noise comes from the same torch CPU
    # generator in both arms; frame index t = rank * 7919 + set * B + b
    ctxb = core.bare_context(device)
    ref_dev = dev.to_device(ref, device)
    n_sets = 2
    sets = []
    for s in range(n_sets):
        frames = torch.empty((B, Z, Y, X, C), dtype=torch.float32, device=device)
        for b in range(B):
            t_idx = 7919 * rank + s * B + b
            lr = lowres_flow(1000 + t_idx)
            g = np.stack([core.resize(lr[q].astype(np.float64), SHAPE) for q in range(3)], 0).astype(np.float64)
            u, v, w = (dev.to_device(-g[q], device) for q in range(3))
            out = frames[b]
            core._check(ctxb.h, ctxb.lib.fr3d_warp(ctxb.h, dev.ptr(ref_dev), 0, dev.ptr(u), dev.ptr(v), dev.ptr(w),
                                                   dev.ptr(ref_dev), 0, Z, Y, X, C, 1, dev.ptr(out)))
            ctxb.sync()
            frames[b] += frame_noise(1000 + t_idx).to(device)
        sets.append(frames)
    host_sets = [s.cpu().pin_memory() for s in sets]
    out_reg = [torch.empty((B, Z, Y, X, C), dtype=torch.float32).pin_memory() for _ in range(2)]
    out_flow = [torch.empty((B, Z, Y, X, 3), dtype=torch.float32).pin_memory() for _ in range(2)]

    def barrier():
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(device)

    def step_resident(i):
        return seq.process_batch(sets[i % n_sets], global_size=B * world, local_offset=B * rank)

    def run_e2e(first, count):
        # the public streaming call: pinned host batches in, pinned host results out; H2D of batch k+1 and
        # D2H of batch k-1 overlap the compute of batch k
        hb = [host_sets[(first + i) % n_sets] for i in range(count)]
        seq.run_pipelined(hb, out_reg=[out_reg[i % 2] for i in range(count)],
                          out_flow=[out_flow[i % 2] for i in range(count)],
                          global_sizes=[B * world] * count, local_offsets=[B * rank] * count)

    # ---- device-resident throughput ("value") ----
    for i in range(args.warmup):
        step_resident(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    profile_all(True)
    l0 = launches_now()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step_resident(args.warmup + i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = launches_now() - l0
    prof = profile_report_all()
    profile_all(False)
    clocks = sampler.stop() if rank == 0 else None

    # ---- end-to-end through the public API with host buffers ----
    run_e2e(0, min(args.warmup, 2))
    barrier()
    e0.record()
    run_e2e(args.warmup, args.steps)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)

    # ---- copy ceiling: the same pinned buffers and byte counts, both directions concurrently, NO compute --------
    # (what e2e could reach at best on this box; on multi-GPU nodes the host memory / PCIe path is shared)
    s_ci, s_co = torch.cuda.Stream(device), torch.cuda.Stream(device)
    d_reg = torch.empty((B, Z, Y, X, C), dtype=torch.float32, device=device)
    d_flow = torch.empty((B, Z, Y, X, 3), dtype=torch.float32, device=device)
    d_in = torch.empty((B, Z, Y, X, C), dtype=torch.float32, device=device)

    def copy_steps(n):
        main_s = torch.cuda.current_stream(device)
        s_ci.wait_stream(main_s)
        s_co.wait_stream(main_s)
        for i in range(n):
            with torch.cuda.stream(s_ci):
                d_in.copy_(host_sets[i % n_sets], non_blocking=True)
            with torch.cuda.stream(s_co):
                out_flow[i % 2].copy_(d_flow, non_blocking=True)
                out_reg[i % 2].copy_(d_reg, non_blocking=True)
        main_s.wait_stream(s_ci)
        main_s.wait_stream(s_co)

    copy_steps(1)
    barrier()
    e0.record()
    copy_steps(args.steps)
    e1.record()
    barrier()
    ms_copy = e0.elapsed_time(e1)
    del d_reg, d_flow, d_in

    # ---- the one-call public entry point with PAGEABLE numpy in / out (single GPU only: host memory) ------------
    arr_api = None
    if world == 1 and not args.no_arr_api:
        video = np.concatenate([h.numpy() for h in host_sets], 0)          # (2B, Z, Y, X, C) pageable copy
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        reg_np, w_np = F.compensate_arr_3D(video, ref, F.OFOptions(buffer_size=B))
        dt = time.perf_counter() - t0
        arr_api = {"value": round(video.shape[0] / dt, 3), "unit": "volumes/s", "frames": int(video.shape[0]),
                   "what": "flowreg3d_b200.compensate_arr_3D(video, reference, OFOptions(buffer_size=B)) -- the reference's "
                           "one-call signature -- with pageable numpy in and fresh numpy arrays out (registered as "
                           "float64, the reference's output_typename default): wall clock of the whole call, i.e. "
                           "context + reference pyramid + w_init bootstrap + page faults of 0.23 GB of fresh result "
                           "memory per frame + pageable copies"}
        del reg_np, w_np, video

    t = torch.tensor([ms, ms_e2e, float(launches), ms_copy], dtype=torch.float64, device=device)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, ms_e2e, launches, ms_copy = float(tmax[0]), float(tmax[1]), int(tsum[2]), float(tmax[3])
    frames_total = B * world * args.steps

    if rank == 0:
        peaks = {}
        pk = ROOT / "MEASURED_PEAKS.json"
        if pk.exists():
            peaks = json.loads(pk.read_text())
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        level_n = [int(np.prod(s)) for _, s in reg.plan.sched]
        table = []
        for name, (cnt, tot) in prof.items():
            ab = algorithmic_bytes(name, level_n, B, C, opts.iterations, opts.update_lag)
            table.append({"kernel": name.replace("fr3d::", ""), "launches": cnt, "ms_total": round(tot, 3),
                          "share": round(tot / ms, 4),
                          "gbs": None if ab is None else round(ab / (tot / cnt * 1e-3) / 1e9, 1)})
        table.sort(key=lambda r: -r["ms_total"])
        top = table[0]
        top_name = [n for n in prof if n.replace("fr3d::", "") == top["kernel"]][0]
        ab = algorithmic_bytes(top_name, level_n, B, C, opts.iterations, opts.update_lag)
        ach = None if ab is None else ab / (prof[top_name][1] / prof[top_name][0] * 1e-3) / 1e9
        state_used = "f64" if reg.plan.plan.state_dtype == 1 else "f32"
        traffic, traffic_src = measured_traffic(top["kernel"], state_used, B, level_n)
        line = {
            "metric": "volumes/s (32x512x512, 2 channels)", "value": round(frames_total / (ms * 1e-3), 3),
            "unit": "volumes/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_gpu_per_step": B, "global_batch": B * world,
                       "sharding": f"frames x{world}, one all-reduce of w_init per batch" if world > 1 else "single GPU",
                       "solver_sweep": "lexicographic (wavefront schedule, reference order)" if args.sweep == "lexicographic"
                                       else "red-black (opt-in; NOT the reference order, outside its parity tolerance)",
                       "solver_state": f"{state_used} increments ({args.state}), f64 system matrix and arithmetic",
                       "frames": FRAMES,
                       "streams": f"{len(ctxs)} (batch split into {len(ctxs)} concurrent parts; per-kernel times in "
                                  "'kernels' are event-bracketed on each part's stream and include time shared with the other part)"
                                  if len(ctxs) > 1 else "1",
                       "arithmetic": "f64 solver/spline/pre-filter math on f32 images (the reference's rounding points)",
                       "l2": f"inputs ({B * Z * Y * X * C * 4 / 2 ** 30:.2f} GiB per step) exceed the 126 MB L2",
                       "host_affinity": f"rank 0 bound to {len(bound)} GPU-local cores" if bound else "unchanged"},
            "e2e": {"value": round(frames_total / (ms_e2e * 1e-3), 3), "unit": "volumes/s",
                    "h2d_bytes_per_step": int(B * Z * Y * X * C * 4),
                    "d2h_bytes_per_step": int(B * Z * Y * X * (C + 3) * 4),
                    "api": "SequenceCorrector.run_pipelined (the streaming entry point for host-resident recordings): pinned "
                           "host batches in, pinned host results out, copies inside the timed region",
                    "copy_ceiling_volumes_per_s": round(frames_total / (ms_copy * 1e-3), 3),
                    "frac_of_copy_ceiling": round(ms_copy / ms_e2e, 4),
                    "copy_ceiling": "the same buffers and bytes per step copied H2D and D2H concurrently with no "
                                    "compute, max over ranks (the node's host path is shared by all GPUs)",
                    "compensate_arr_3D_pageable": arr_api},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": top["kernel"], "achieved": None if ach is None else round(ach, 1),
                         "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                         "frac": None if ach is None else round(ach / peak, 4),
                         "frac_of_nominal_8TBs": None if ach is None else round(ach / 8000.0, 4),
                         "traffic": traffic, "traffic_source": traffic_src,
                         "share_of_step": top["share"]},
            "kernels": table[:8],
            "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    v, cores, sec = run_cpu(steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": "volumes/s (32x512x512, 2 channels)", "value": round(v, 4),
        "unit": "volumes/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(sec * 1e3, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step": cores, "frames": FRAMES},
        "cpu_baseline": {"value": round(v, 4), "unit": "volumes/s", "cores": cores, "kind": "port",
                         "sample": f"each step = {cores} frames of the workload, one per worker process "
                                   "(oracle port of the reference CPU path: pre-filter + get_displacement + cubic warp)"},
        "e2e": {"value": round(v, 4), "unit": "volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=25,
                    help="frames per GPU per step (OFOptions.buffer_size; 25 = the 200-frame recording of config 2 "
                         "in 8 batches, or one batch per GPU at 8 GPUs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-arr-api", action="store_true", help="skip the pageable compensate_arr_3D measurement")
    ap.add_argument("--streams", type=int, default=1, help="concurrent half-batch pipelines per GPU")
    ap.add_argument("--sweep", default="lexicographic", choices=["lexicographic", "redblack"],
                    help="solver sweep order; only lexicographic reproduces the reference")
    ap.add_argument("--state", default="f64", choices=["f64", "f32", "auto"],
                    help="storage precision of the solver increments: f64 (the package default: the reference to "
                         "float64 rounding), f32 (opt-in: 24 %% fewer solver bytes, inside the tolerance on config 2 "
                         "but not on every workload, see flowreg3d_b200/core.py), auto (f32 when min_level >= 2)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
