"""A/B of the level-solver kernels on one GPU: direct-load wavefront kernel vs the staged (cp.async.bulk + mbarrier)
kernel, float64 / float32 state, stage counts.  Prints the solver's milliseconds per get_displacement call, its
achieved algorithmic GB/s (108 B / voxel / sweep + (12C+84) B per psi refresh) and whether the flows
of every variant are bit-identical to the direct kernel's (same arithmetic, same order).

    python tools/sor_ab.py [--batch 25] [--min-level 5] [--frames-shape 32 512 512] [--stages 2 3 4] [--reps 3]
"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import flowreg3d_b200 as F  # noqa: E402
from flowreg3d_b200 import _lib, core  # noqa: E402
from tests_inputs import synth_volume  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=25)
ap.add_argument("--min-level", type=int, default=5)
ap.add_argument("--shape", type=int, nargs=3, default=[32, 512, 512])
ap.add_argument("--channels", type=int, default=2)
ap.add_argument("--stages", type=int, nargs="*", default=[0])
ap.add_argument("--states", nargs="*", default=["f64", "f32"])
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--iterations", type=int, default=100)
ap.add_argument("--ctas", type=int, nargs="*", default=[0])
ap.add_argument("--kernels", type=int, nargs="*", default=[0, 2], help="0 direct wavefront, 1 staged wavefront, 2 tiles")
ap.add_argument("--tiles", nargs="*", default=["0"], help="FR3D_OPT_SOR_TILE values Tb,K,J,I (e.g. 5,8,8,8) for kernel 2")
ap.add_argument("--sched", type=int, nargs="*", default=[0], help="FR3D_OPT_SOR_SCHED values for kernel 0")
ap.add_argument("--frames-per-item", type=int, default=0, help="FR3D_OPT_SOR_FRAMES_PER_ITEM (0 = default 2)")
ap.add_argument("--lag", type=int, default=5, help="update_lag (psi refresh every lag sweeps)")
ap.add_argument("--library", default=None, help="a tuning build of libfr3d.so (CUDA) instead of the in-tree one")
args = ap.parse_args()
if args.library:
    _lib._select_for_tests(args.library, emulator=False)

B, shape, Cn = args.batch, tuple(args.shape), args.channels
ref = np.stack([synth_volume(shape, 10 + c) for c in range(Cn)], -1)
rng = np.random.default_rng(0)
frames = np.stack([np.roll(ref, (0, (b % 3) + 1, -(b % 4) - 1), (0, 1, 2)) for b in range(B)], 0)
frames = (frames + 0.01 * rng.standard_normal(frames.shape)).astype(np.float32)
dev_frames = torch.from_numpy(frames).cuda()
fp = F.FlowParams(min_level=args.min_level, a_smooth=1.0, iterations=args.iterations, alpha=(0.25,) * 3, update_lag=args.lag,
                  levels=100, eta=0.8, a_data=0.45)
lag = args.lag


def tile_code(spec):
    if spec in ("0", 0):
        return 0
    tb, k, j, i = (int(x) for x in spec.split(","))
    return tb | (k << 8) | (j << 16) | (i << 24)


def run(state, kernel, stages, ctas, tile="0", sched=0):
    reg = F.Registration(shape, Cn, fp, max_batch=B, state_dtype=np.float32 if state == "f32" else np.float64)
    reg.set_reference(ref, weight=np.full(Cn, 1.0 / Cn))
    h, lib = reg.ctx.h, reg.ctx.lib
    core._check(h, lib.fr3d_set_option(h, _lib.OPT_SOR_KERNEL, kernel))
    core._check(h, lib.fr3d_set_option(h, _lib.OPT_SOR_STAGES, stages))
    core._check(h, lib.fr3d_set_option(h, _lib.OPT_SOR_CTAS_PER_SM, ctas))
    core._check(h, lib.fr3d_set_option(h, _lib.OPT_SOR_TILE, tile_code(tile)))
    core._check(h, lib.fr3d_set_option(h, _lib.OPT_SOR_SCHED, sched))
    core._check(h, lib.fr3d_set_option(h, _lib.OPT_SOR_FRAMES_PER_ITEM, args.frames_per_item))
    flow = reg.get_displacement(dev_frames)
    reg.sync()
    reg.ctx.profile(True)
    for _ in range(args.reps):
        flow = reg.get_displacement(dev_frames)
    rep = reg.ctx.profile_report()
    reg.ctx.profile(False)
    sor = {k: v for k, v in rep.items() if "sor" in k}
    ms = sum(v[1] for v in sor.values()) / args.reps
    total = sum(v[1] for v in rep.values()) / args.reps
    level_n = [int(np.prod(s)) for _, s in reg.plan.sched]
    nref = -(-args.iterations // lag)
    alg = sum(B * n * (108 * args.iterations + (12 * Cn + 84) * nref) for n in level_n)
    out = flow.cpu().numpy()
    del reg
    torch.cuda.empty_cache()
    return ms, total, alg / (ms * 1e-3) / 1e9, out, list(sor), level_n


base = {}
for state in args.states:
    for kernel in args.kernels:
        variants = ([(st, "0", 0) for st in args.stages] if kernel == 1 else [(0, tl, 0) for tl in args.tiles] if kernel == 2
                    else [(0, "0", sc) for sc in args.sched])
        for stages, tile, sched in variants:
            for ctas in args.ctas:
                try:
                    ms, total, gbs, flow, names, level_n = run(state, kernel, stages, ctas, tile, sched)
                except Exception as e:  # report and go on: a failing variant must not hide the others
                    print(json.dumps({"state": state, "kernel": kernel, "stages": stages, "ctas": ctas, "tile": tile,
                                      "error": str(e)}), flush=True)
                    continue
                if kernel == 0 and state not in base:
                    base[state] = flow
                same = bool(np.array_equal(flow, base[state])) if state in base else None
                dmax = float(np.abs(flow - base[state]).max()) if state in base else None
                print(json.dumps({"state": state, "kernel": kernel, "stages": stages, "ctas": ctas, "tile": tile, "sched": sched,
                                  "sor_ms": round(ms, 3), "all_kernels_ms": round(total, 3),
                                  "alg_gbs": round(gbs, 1), "frac_of_6453": round(gbs / 6453.1, 4),
                                  "bit_identical_to_direct": same, "max_abs_diff": dmax, "names": names,
                                  "levels": level_n, "B": B}), flush=True)
if "f64" in base and "f32" in base:
    e = np.sqrt(((base["f64"].astype(np.float64) - base["f32"]) ** 2).sum(-1))
    print(json.dumps({"f32_vs_f64_state_epe_mean": float(e.mean()), "max": float(e.max())}))
