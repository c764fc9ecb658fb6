"""Small driver for ncu: B frames of the config-2 workload through SequenceCorrector.process_batch.
    python tools/profile_step.py [B] [steps]
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import flowreg3d_b200 as F  # noqa: E402
from tests_inputs import synth_volume  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
shape = (32, 512, 512)
ref = np.stack([synth_volume(shape, 10 + c) for c in range(2)], -1)
rng = np.random.default_rng(0)
frames = np.stack([np.roll(ref, (0, b + 1, -b - 1), (0, 1, 2)) for b in range(B)], 0)
frames = (frames + 0.01 * rng.standard_normal(frames.shape)).astype(np.float32)
seq = F.SequenceCorrector(ref, F.OFOptions(buffer_size=B), max_batch=B)
dev_frames = torch.from_numpy(frames).cuda()
for _ in range(steps):
    reg, flow = seq.process_batch(dev_frames)
seq.reg.sync()
print("ok", float(flow.abs().max()))
