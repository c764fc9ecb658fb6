#!/bin/bash
# 2-GPU session d: flag wait relaxed to the boundary items (NVLink latency under interior work)
O=gpurun_out/m2d; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/pytest_multi.log 2>&1; echo "pytest multi rc $?" | tee -a $O/rc.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR tools/bench_pipelined.py --shape 128 1024 1024 --channels 1 --min-level 5 --zslab > $O/zslab_p2p_c4_ml5.log 2>&1; echo "zslab c4 ml5 rc $?" | tee -a $O/rc.txt
timeout 400 $TR tools/bench_pipelined.py --shape 128 1024 1024 --channels 1 --min-level 2 --zslab > $O/zslab_p2p_c4_ml2.log 2>&1; echo "zslab c4 ml2 rc $?" | tee -a $O/rc.txt
timeout 400 $TR tools/bench_pipelined.py --shape 32 512 512 --channels 2 --min-level 0 --zslab > $O/zslab_p2p_c5_ml0.log 2>&1; echo "zslab c5 ml0 rc $?" | tee -a $O/rc.txt
tail -3 $O/pytest_multi.log; for f in zslab_p2p_c4_ml5 zslab_p2p_c4_ml2 zslab_p2p_c5_ml0; do echo "== $f"; tail -1 $O/$f.log | cut -c1-1000; done
