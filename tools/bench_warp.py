"""Kernel times of the compensation warp alone (config 2 frames, B = 16): python tools/bench_warp.py [--factored]"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import flowreg3d_b200 as F  # noqa: E402
from tests_inputs import synth_volume  # noqa: E402

B, shape = 16, (32, 512, 512)
ref = np.stack([synth_volume(shape, 10 + c) for c in range(2)], -1)
reg = F.Registration(shape, 2, F.FlowParams(min_level=5, a_smooth=1.0, iterations=2), max_batch=B)
frames = torch.from_numpy(np.stack([np.roll(ref, b, 2) for b in range(B)], 0)).cuda()
gen = torch.Generator(device="cuda").manual_seed(0)
flow = (torch.rand((B,) + shape + (3,), device="cuda", generator=gen) * 0.1 +
        torch.linspace(-2, 2, shape[2], device="cuda")[None, None, None, :, None]).float().contiguous()
refd = torch.from_numpy(ref).cuda()
if "--factored" in sys.argv:   # FR3D_OPT_WARP_FACTORED (<= 1 float32 ulp from the default; not yet timed on a B200)
    from flowreg3d_b200 import _lib, core
    core._check(reg.ctx.h, reg.ctx.lib.fr3d_set_option(reg.ctx.h, _lib.OPT_WARP_FACTORED, 1))
for _ in range(2):
    reg.compensate(frames, flow, ref_raw=refd)
reg.sync()
reg.ctx.profile(True)
for _ in range(3):
    reg.compensate(frames, flow, ref_raw=refd)
rep = reg.ctx.profile_report()
print({k.replace("fr3d::", ""): round(v[1] / 3, 3) for k, v in rep.items()}, "total", round(sum(v[1] for v in rep.values()) / 3, 3))
