"""Config 2 frames: endpoint error of the optional solver modes against the default (float64 state,
lexicographic sweep), GPU vs GPU at full size.  python tools/compare_modes.py"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import flowreg3d_b200 as F  # noqa: E402
from flowreg3d_b200 import core, device as dev  # noqa: E402
from tests_inputs import smooth_flow, synth_volume  # noqa: E402

shape, B = (32, 512, 512), 4
ref = np.stack([synth_volume(shape, 10 + c) for c in range(2)], -1)
r64 = ref.astype(np.float64)
frames = []
for b in range(B):
    g = smooth_flow(shape, 1000 + b, 2.0, 12.0)
    frames.append(F.imregister_wrapper(r64, -g[..., 0], -g[..., 1], -g[..., 2], r64, "linear"))
frames = np.stack(frames, 0).astype(np.float32)
rng = np.random.default_rng(0)
frames += 0.01 * rng.standard_normal(frames.shape).astype(np.float32)


def run(state, sweep):
    core.STATE_DTYPE, core.SWEEP = state, sweep
    seq = F.SequenceCorrector(ref, F.OFOptions(buffer_size=B), max_batch=B)
    reg, fl = seq.process_batch(frames)
    seq.reg.sync()
    out = dev.to_host(reg).astype(np.float64), dev.to_host(fl).astype(np.float64)
    seq.close()
    return out


base_reg, base_fl = run(np.float64, 0)
for name, st, sw in (("float32 state, lexicographic", np.float32, 0), ("float64 state, red-black", np.float64, 1)):
    reg, fl = run(st, sw)
    e = np.sqrt(((fl - base_fl) ** 2).sum(-1))
    print(json.dumps({"mode": name, "vs": "float64 state, lexicographic (default)", "frames": B,
                      "epe_mean": float(e.mean()), "epe_p99": float(np.percentile(e, 99)), "epe_max": float(e.max()),
                      "registered_rel_l2": float(np.linalg.norm(reg - base_reg) / np.linalg.norm(base_reg))}))
