"""Wall-clock latency of the per-pair API (boundary B2) on BASELINE config 1: get_displacement + imregister_wrapper on a
64x128x128x1 pair with host arrays in and out, OFOptions-default parameters."""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import flowreg3d_b200 as F  # noqa: E402
from tests_inputs import smooth_flow, synth_volume  # noqa: E402

shape = (64, 128, 128)
fixed = synth_volume(shape, 1)
g = smooth_flow(shape, 2, 2.0, 10.0)
moving = F.imregister_wrapper(fixed.astype(np.float64), -g[..., 0], -g[..., 1], -g[..., 2], fixed.astype(np.float64), "linear")
kw = dict(alpha=(0.25,) * 3, update_lag=5, iterations=100, min_level=5, levels=100, eta=0.8, a_smooth=1.0, a_data=0.45)
for rep in range(6):
    t0 = time.perf_counter()
    flow = F.get_displacement(fixed, moving, **kw)
    t1 = time.perf_counter()
    f32 = flow.astype(np.float32)
    reg = F.imregister_wrapper(moving, f32[..., 0], f32[..., 1], f32[..., 2], fixed, "cubic")
    t2 = time.perf_counter()
    print(f"rep {rep}: get_displacement {1e3 * (t1 - t0):7.2f} ms, imregister_wrapper {1e3 * (t2 - t1):7.2f} ms, "
          f"pair {1e3 * (t2 - t0):7.2f} ms", flush=True)
