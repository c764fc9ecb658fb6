#!/bin/bash
# 2-GPU session f: final build -- multi-GPU tests (NCCL: z-slab p2p / host exchange, pipelined, sharded) and the driver's
# 2-GPU bench invocation (both arms)
O=gpurun_out/m2f; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/pytest_multi.log 2>&1; echo "pytest multi rc $?" | tee -a $O/rc.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench2.log 2> $O/bench2.err; echo "bench2 rc $?" | tee -a $O/rc.txt
timeout 300 $TR bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > $O/bench2_ref.log 2> $O/bench2_ref.err; echo "bench2 ref rc $?" | tee -a $O/rc.txt
tail -2 $O/pytest_multi.log; cut -c1-260 $O/bench2.log | tail -1; cut -c1-200 $O/bench2_ref.log | tail -1
