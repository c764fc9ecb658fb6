#!/bin/bash
# A/B measurements that are prepared (emulator-verified) but have NOT been run on a B200 yet -- the round's GPU
# budget ran out first.  Run under gpurun; each line prints per-kernel milliseconds.
set -x
python tools/bench_cc.py 10                  # shipped: one thread per plane for arg-max / tile sums / plane mean (29.9 ms)
python tools/bench_cc.py 10 --block-scans    # FR3D_OPT_CC_BLOCK_SCANS: one CTA per plane; expect the three scans < 1 ms
python tools/bench_warp.py                   # shipped gather: 7.6 ms per 16 frames, fp64 pipe 61 %
python tools/bench_warp.py --factored        # FR3D_OPT_WARP_FACTORED: 3x fewer float64 operations (<= 1 float32 ulp)
# 2 GPUs (gpurun --gpus 2): the z-slab solve has never run on GPUs; start small (host-driven: ~2 400 launches + messages per level)
# python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/bench_pipelined.py --shape 32 256 256 --min-level 2 --iterations 20 --zslab
