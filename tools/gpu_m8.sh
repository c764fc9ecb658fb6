#!/bin/bash
# 8-GPU session: z-slabs on 4 and 8 GPUs (config 4), node copy ceiling, bench at N = 8
O=gpurun_out/m8; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR --nproc-per-node 4 tools/bench_pipelined.py --shape 128 1024 1024 --channels 1 --min-level 2 --zslab > $O/zslab_g4_ml2.log 2>&1; echo "zslab g4 ml2 rc $?" | tee -a $O/rc.txt
timeout 400 $TR --nproc-per-node 8 tools/bench_pipelined.py --shape 128 1024 1024 --channels 1 --min-level 2 --zslab > $O/zslab_g8_ml2.log 2>&1; echo "zslab g8 ml2 rc $?" | tee -a $O/rc.txt
timeout 300 $TR --nproc-per-node 4 tools/bench_pipelined.py --shape 128 1024 1024 --channels 1 --min-level 5 --zslab > $O/zslab_g4_ml5.log 2>&1; echo "zslab g4 ml5 rc $?" | tee -a $O/rc.txt
timeout 200 $TR --nproc-per-node 8 tools/copy_ceiling.py > $O/copy8.log 2>&1
timeout 200 $TR --nproc-per-node 4 tools/copy_ceiling.py > $O/copy4.log 2>&1
timeout 600 $TR --nproc-per-node 8 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_g8.log 2> $O/bench_g8.err; echo "bench g8 rc $?" | tee -a $O/rc.txt
for f in zslab_g4_ml2 zslab_g8_ml2 zslab_g4_ml5 copy4 copy8; do echo "== $f"; grep "^{" $O/$f.log | tail -1 | cut -c1-1000; done; grep "^{" $O/bench_g8.log | tail -1 | cut -c1-1800
