"""Small end-to-end run for compute-sanitizer (one tool per call):  compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import flowreg3d_b200 as F  # noqa: E402
from tests_inputs import synth_volume  # noqa: E402

shape = (11, 37, 45)            # odd sizes: ragged tiles, odd median pairs, short spline segments
ref = np.stack([synth_volume(shape, 3 + c) for c in range(2)], -1)
video = np.stack([np.roll(ref, (0, t + 1, -t), (0, 1, 2)) for t in range(3)], 0).astype(np.float32)
for kw in (dict(min_level=0, iterations=6, update_lag=3), dict(min_level=1, iterations=5, update_lag=2, a_smooth=0.5)):
    opts = F.OFOptions(buffer_size=2, weight=[0.5, 0.5], **kw)
    reg, w = F.compensate_arr_3D(video, ref, opts)
    print(kw, float(np.abs(w).max()), float(reg.mean()))
big = np.stack([synth_volume((6, 70, 300), 9)], -1)   # long x lines: 8 spline segments
out = F.imregister_wrapper(big[..., 0].astype(np.float64), np.full(big.shape[:3], 1.3), np.full(big.shape[:3], -0.7),
                           np.full(big.shape[:3], 0.4), big[..., 0].astype(np.float64), "cubic")
print("warp", float(out.mean()))
