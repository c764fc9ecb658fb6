"""Cost of the rigid cross-correlation pre-alignment: one batch of single-channel 32x512x512 frames through
SequenceCorrector.process_batch with and without OFOptions.cc_initialization.   python tools/bench_cc.py [B] [--block-scans]"""
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import flowreg3d_b200 as F  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
BLOCK_SCANS = "--block-scans" in sys.argv      # FR3D_OPT_CC_BLOCK_SCANS (not yet timed on a B200)
B = int(args[0]) if args else 10
shape = (32, 512, 512)
rng = np.random.default_rng(0)
z, y, x = np.ogrid[:shape[0], :shape[1], :shape[2]]
ref = (0.3 * rng.random(shape) + np.exp(-((z - 16) ** 2 + (y - 250) ** 2 + (x - 260) ** 2) / 4000.0)).astype(np.float32)
frames = np.stack([np.roll(ref, (b % 3, 2 * b + 1, -b - 2), (0, 1, 2)) for b in range(B)], 0)[..., None]
dev_frames = torch.from_numpy(frames).cuda()
for cc in (False, True):
    seq = F.SequenceCorrector(ref, F.OFOptions(buffer_size=B, cc_initialization=cc), max_batch=B)
    if BLOCK_SCANS:
        from flowreg3d_b200 import _lib, core
        core._check(seq.reg.ctx.h, seq.reg.ctx.lib.fr3d_set_option(seq.reg.ctx.h, _lib.OPT_CC_BLOCK_SCANS, 1))
    seq.process_batch(dev_frames)          # bootstrap + first batch (tables, spectra of the reference)
    seq.reg.sync()
    seq.reg.ctx.profile(True)
    t0 = time.perf_counter()
    reg, flow = seq.process_batch(dev_frames)
    seq.reg.sync()
    dt = time.perf_counter() - t0
    rep = seq.reg.ctx.profile_report()
    ccms = sum(v[1] for k, v in rep.items() if "Cc" in k or "StridedCopy" in k)
    print(f"cc_initialization={cc}: {dt * 1e3:.1f} ms per batch of {B} ({B / dt:.1f} volumes/s); "
          f"pre-alignment kernels {ccms:.2f} ms; mean |flow| {float(flow.abs().mean()):.3f}")
    top = sorted(((v[1], k) for k, v in rep.items() if "Cc" in k), reverse=True)[:4]
    if top:
        print("   ", [(k.replace("fr3d::", ""), round(ms, 2)) for ms, k in top])
    seq.close()
