"""Timing of the other BASELINE.json configurations on one GPU (device-resident inputs, CUDA events).

    python tools/bench_configs.py [--quick]

config 1: pair 64x128x128x1, OFOptions defaults             (get_displacement + compensation warp)
config 3: 64x256x256x2 sequence, batches of 16               (throughput stress; frames/s)
config 4: single 128x1024x1024x1 volume, min_level 5 and 2   (one GPU: no z-slab decomposition)
config 5: 32x512x512x2, min_level in {5, 4, 3, 0}, B = 4     (solver sweep; SOR share)
Writes one JSON line per case to stdout.
"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import flowreg3d_b200 as F  # noqa: E402
from tests_inputs import synth_volume  # noqa: E402


def frames_like(ref, B, rng):
    fr = np.stack([np.roll(ref, (0, b + 1, -b - 1), (0, 1, 2)) for b in range(B)], 0)
    return (fr + 0.01 * rng.standard_normal(fr.shape)).astype(np.float32)


def run(name, shape, C, B, steps=3, **opt):
    rng = np.random.default_rng(0)
    ref = np.stack([synth_volume(shape, 10 + c) for c in range(C)], -1)
    opts = F.OFOptions(buffer_size=B, weight=[1.0 / C] * C, sigma=[[1.0, 1.0, 1.0, 0.1]] * C, **opt)
    seq = F.SequenceCorrector(ref, opts, max_batch=B)
    dev_frames = torch.from_numpy(frames_like(ref, B, rng)).cuda()
    ctx = seq.reg.ctx
    for _ in range(2):
        seq.process_batch(dev_frames)
    seq.reg.sync()
    ctx.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        seq.process_batch(dev_frames)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    prof = ctx.profile_report()
    ctx.profile(False)
    top = sorted(((v[1] / steps, k) for k, v in prof.items()), reverse=True)[:4]
    out = {"case": name, "shape": list(shape), "channels": C, "frames_per_step": B, "levels": [list(s) for _, s in seq.reg.plan.sched],
           "ms_per_step": round(ms, 3), "volumes_per_s": round(B / ms * 1e3, 2),
           "device_GB": round(ctx.device_bytes / 2 ** 30, 2),
           "top_kernels_ms": [[k.replace("fr3d::", ""), round(t, 3)] for t, k in top]}
    print(json.dumps(out), flush=True)
    seq.close()
    del seq, dev_frames
    torch.cuda.empty_cache()


def run_pair_api():
    """config 1 through the per-pair drop-in functions (host arrays in and out, wall clock)."""
    import time
    shape = (64, 128, 128)
    fixed = synth_volume(shape, 1)
    moving = np.roll(fixed, (1, 2, -2), (0, 1, 2))
    kw = dict(alpha=(0.25,) * 3, update_lag=5, iterations=100, min_level=5, levels=100, eta=0.8, a_smooth=1.0,
              a_data=0.45)
    t = []
    for _ in range(4):
        t0 = time.perf_counter()
        flow = F.get_displacement(fixed, moving, **kw)
        f32 = flow.astype(np.float32)
        reg = F.imregister_wrapper(moving, f32[..., 0], f32[..., 1], f32[..., 2], fixed, "cubic")
        t.append((time.perf_counter() - t0) * 1e3)
    print(json.dumps({"case": "config1 per-pair API get_displacement + imregister_wrapper (host in/out, wall clock)",
                      "ms_first_call": round(t[0], 2), "ms_later_calls": [round(x, 2) for x in t[1:]]}), flush=True)


if __name__ == "__main__":
    quick = "--quick" in sys.argv
    run_pair_api()
    run("config1 pair 64x128x128x1 defaults", (64, 128, 128), 1, 1)
    run("config1 batch of 16", (64, 128, 128), 1, 16)
    run("config3 64x256x256x2 B=16", (64, 256, 256), 2, 16)
    run("config5 32x512x512x2 min_level 4", (32, 512, 512), 2, 4, min_level=4)
    run("config5 32x512x512x2 min_level 3", (32, 512, 512), 2, 4, min_level=3)
    if not quick:
        run("config5 32x512x512x2 min_level 0 (HBM-bound SOR)", (32, 512, 512), 2, 2, min_level=0, steps=1)
        run("config4 128x1024x1024x1 min_level 5, one GPU", (128, 1024, 1024), 1, 1)
        run("config4 128x1024x1024x1 min_level 2, one GPU", (128, 1024, 1024), 1, 1, min_level=2, steps=1)
