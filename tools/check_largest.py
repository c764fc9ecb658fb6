"""Index-range sanity at the largest configuration: one 128x1024x1024 volume solved at FULL resolution
(min_level 0: 134 M solver slots, 13 levels) on one GPU; few sweeps to keep it short.  Properties only
(the oracle would take hours): identical inputs give zero flow, a shifted copy gives a flow of the right sign."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import flowreg3d_b200 as F  # noqa: E402
from flowreg3d_b200 import device as dev  # noqa: E402
from tests_inputs import synth_volume  # noqa: E402

shape = (128, 1024, 1024)
ref = synth_volume(shape, 4)[..., None]
fp = F.FlowParams(alpha=(0.25,) * 3, update_lag=5, iterations=10, min_level=0, levels=100, eta=0.8, a_smooth=1.0, a_data=0.45)
reg = F.Registration(shape, 1, fp, max_batch=1)
reg.set_reference(ref)
mov = np.roll(ref, 2, 2)                       # moving(x) = ref(x - 2)  ->  flow u ~ +2
t0 = time.time()
flow = dev.to_host(reg.get_displacement(mov[None]))[0]
dt = time.time() - t0
core = flow[16:-16, 64:-64, 64:-64]
print("levels", len(reg.plan.sched), "device GB", round(reg.ctx.device_bytes / 2 ** 30, 1), "seconds", round(dt, 2))
print("median u", float(np.median(core[..., 0])), "median |v|,|w|", float(np.median(np.abs(core[..., 1]))), float(np.median(np.abs(core[..., 2]))))
assert 1.0 < np.median(core[..., 0]) < 3.0
zero = dev.to_host(reg.get_displacement(ref[None]))[0]
assert np.abs(zero).max() <= 1e-6, np.abs(zero).max()
print("ok")
