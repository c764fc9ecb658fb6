"""One large volume on N GPUs: sweep-pipelined level solve (flowreg3d_b200/multigpu.py) vs one GPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_pipelined.py \
        [--shape 32 512 512] [--channels 2] [--min-level 0] [--iterations 100]

Every rank computes the flow alone (reference timing, identical on all ranks) and then together; rank 0
prints one JSON line with both device times and whether the results are bit-identical.
"""
import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import flowreg3d_b200 as F  # noqa: E402
from flowreg3d_b200.multigpu import get_displacement_pipelined  # noqa: E402
from tests_inputs import synth_volume  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", type=int, nargs=3, default=[32, 512, 512])
ap.add_argument("--channels", type=int, default=2)
ap.add_argument("--min-level", type=int, default=0)
ap.add_argument("--iterations", type=int, default=100)
ap.add_argument("--chunks", type=int, default=8)
ap.add_argument("--zslab", action="store_true", help="z-slab decomposition with per-wave halo exchange instead of the "
                "sweep pipeline (device-side exchange over peer memory unless --host-exchange)")
ap.add_argument("--host-exchange", action="store_true", help="z-slabs with the host-driven NCCL exchange per wave")
ap.add_argument("--state", default="f64", choices=["f64", "f32"])
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
device = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=device)
shape, C = tuple(args.shape), args.channels


def cached_volume(shape, seed):
    """synth_volume with a per-box file cache (the 128x1024x1024 volume costs a minute of scipy per rank otherwise);
    rank 0 generates, the others wait for the file."""
    import time
    path = Path("/dev/shm" if os.path.isdir("/dev/shm") else "/tmp") / f"fr3d_synth_{'x'.join(map(str, shape))}_{seed}.npy"
    if rank == 0 and not path.exists():
        tmp = path.with_suffix(".tmp.npy")
        np.save(tmp, synth_volume(shape, seed))
        os.replace(tmp, path)
    while not path.exists():
        time.sleep(0.2)
    return np.load(path)


ref = np.stack([cached_volume(shape, 10 + c) for c in range(C)], -1)
mov = np.roll(ref, (0, 2, -3), (0, 1, 2)) + 0.01 * np.random.default_rng(0).standard_normal(ref.shape).astype(np.float32)
fp = F.FlowParams(alpha=(0.25,) * 3, update_lag=5, iterations=args.iterations, min_level=args.min_level, levels=100,
                  eta=0.8, a_smooth=1.0, a_data=0.45)
reg = F.Registration(shape, C, fp, max_batch=1, device=device,
                     state_dtype=np.float32 if args.state == "f32" else np.float64)
reg.set_reference(ref.astype(np.float32))
mv = torch.from_numpy(mov.astype(np.float32)[None]).to(device)


def timed(fn, reps):
    fn()
    torch.cuda.synchronize(device)
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize(device)
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t), out


t1, single = timed(lambda: reg.get_displacement(mv), 2)
if args.zslab:
    from flowreg3d_b200.multigpu import get_displacement_zslab
    tn, piped = timed(lambda: get_displacement_zslab(reg, mv, p2p=not args.host_exchange), 1 if args.host_exchange else 2)
else:
    tn, piped = timed(lambda: get_displacement_pipelined(reg, mv, n_chunks=args.chunks), 2)
same = bool(torch.equal(single, piped))
stats = {}
if args.zslab:
    get_displacement_zslab(reg, mv, p2p=not args.host_exchange, stats=stats)   # one more pass with phase timers

if rank == 0:
    print(json.dumps({"case": f"single volume {shape}x{C}, min_level {args.min_level}, {args.iterations} sweeps",
                      "levels": [list(s) for _, s in reg.plan.sched], "n_gpus": world,
                      "mode": ("zslab, host-driven exchange" if args.host_exchange else "zslab, device-side exchange (peer stores + flags)") if args.zslab else "sweep-pipelined", "state": args.state, "ms_one_gpu": round(t1, 2), "ms_pipelined": round(tn, 2), "speedup": round(t1 / tn, 3),
                      "bit_identical": same, "phase_ms_rank0": {k: round(v, 2) for k, v in stats.items()}}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
