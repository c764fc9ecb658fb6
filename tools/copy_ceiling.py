"""Host <-> device copy ceiling of a node for the bench's end-to-end traffic pattern: every rank copies, per step, one
batch of raw frames host -> device and the registered frames + flow fields device -> host, both directions
concurrently on their own streams, with NO compute.  This is what bench.py's `e2e` can reach at best on the same box.

    python tools/copy_ceiling.py [--batch 25] [--steps 4]                       # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/copy_ceiling.py

Prints one JSON line (rank 0): aggregate GB/s per direction and the frames/s ("volumes/s") they allow.
"""
import argparse
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def measure(device, B, steps, warmup=1, shape=(32, 512, 512), C=2, in_bytes=4, world=1):
    """Returns (ms for `steps` steps on this rank, h2d bytes per step, d2h bytes per step)."""
    Z, Y, X = shape
    n_in = B * Z * Y * X * C * in_bytes
    n_reg = B * Z * Y * X * C * 4
    n_flow = B * Z * Y * X * 3 * 4
    h_in = [torch.empty(n_in, dtype=torch.uint8).pin_memory() for _ in range(2)]
    h_reg = [torch.empty(n_reg, dtype=torch.uint8).pin_memory() for _ in range(2)]
    h_flow = [torch.empty(n_flow, dtype=torch.uint8).pin_memory() for _ in range(2)]
    d_in = torch.empty(n_in, dtype=torch.uint8, device=device)
    d_reg = torch.zeros(n_reg, dtype=torch.uint8, device=device)
    d_flow = torch.zeros(n_flow, dtype=torch.uint8, device=device)
    s_in, s_out = torch.cuda.Stream(device), torch.cuda.Stream(device)

    def step(i):
        with torch.cuda.stream(s_in):
            d_in.copy_(h_in[i % 2], non_blocking=True)
        with torch.cuda.stream(s_out):
            h_flow[i % 2].copy_(d_flow, non_blocking=True)
            h_reg[i % 2].copy_(d_reg, non_blocking=True)

    def barrier():
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(device)

    for i in range(warmup):
        step(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream(device)
    e0.record(main)
    s_in.wait_stream(main)
    s_out.wait_stream(main)
    for i in range(steps):
        step(i)
    main.wait_stream(s_in)
    main.wait_stream(s_out)
    e1.record(main)
    barrier()
    return e0.elapsed_time(e1), n_in, n_reg + n_flow


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=25)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--in-bytes", type=int, default=4, help="bytes per raw sample (4 float32, 2 uint16)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    ms, bi, bo = measure(device, args.batch, args.steps, in_bytes=args.in_bytes, world=world)
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    if rank == 0:
        sec = ms * 1e-3
        print(json.dumps({"n_gpus": world, "batch": args.batch, "steps": args.steps, "ms_per_step": round(ms / args.steps, 2),
                          "h2d_gbs_aggregate": round(world * bi * args.steps / sec / 1e9, 2),
                          "d2h_gbs_aggregate": round(world * bo * args.steps / sec / 1e9, 2),
                          "total_gbs_aggregate": round(world * (bi + bo) * args.steps / sec / 1e9, 2),
                          "copy_ceiling_volumes_per_s": round(world * args.batch * args.steps / sec, 2),
                          "bytes_per_volume": {"h2d": bi // args.batch, "d2h": bo // args.batch}}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
