#!/bin/bash
# 2-GPU session: multi-GPU tests on real devices (NCCL) + z-slab / pipelined timings
O=gpurun_out/m2; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/pytest_multi.log 2>&1; echo "pytest multi rc $?" | tee -a $O/rc.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/bench_pipelined.py --shape 64 256 256 --channels 1 --min-level 2 --zslab > $O/zslab_p2p_small.log 2>&1; echo "zslab small rc $?" | tee -a $O/rc.txt
timeout 300 $TR tools/bench_pipelined.py --shape 64 256 256 --channels 1 --min-level 2 --zslab --host-exchange > $O/zslab_host_small.log 2>&1; echo "zslab host small rc $?" | tee -a $O/rc.txt
timeout 400 $TR tools/bench_pipelined.py --shape 128 1024 1024 --channels 1 --min-level 5 --zslab > $O/zslab_p2p_c4_ml5.log 2>&1; echo "zslab c4 ml5 rc $?" | tee -a $O/rc.txt
timeout 400 $TR tools/bench_pipelined.py --shape 128 1024 1024 --channels 1 --min-level 2 --zslab > $O/zslab_p2p_c4_ml2.log 2>&1; echo "zslab c4 ml2 rc $?" | tee -a $O/rc.txt
timeout 400 $TR tools/bench_pipelined.py --shape 128 1024 1024 --channels 1 --min-level 2 > $O/piped_c4_ml2.log 2>&1; echo "piped c4 ml2 rc $?" | tee -a $O/rc.txt
timeout 200 $TR tools/copy_ceiling.py > $O/copy2.log 2>&1
timeout 200 python tools/copy_ceiling.py > $O/copy1.log 2>&1
tail -5 $O/pytest_multi.log; for f in zslab_p2p_small zslab_host_small zslab_p2p_c4_ml5 zslab_p2p_c4_ml2 piped_c4_ml2 copy2 copy1; do echo "== $f"; tail -2 $O/$f.log | cut -c1-500; done
