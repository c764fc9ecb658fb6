#!/bin/bash
# 2-GPU session e: 24-byte float64 state vectors -- solver timing, full GPU suite, multi-GPU tests, bench
O=gpurun_out/m2e; mkdir -p $O
timeout 600 python tools/sor_ab.py --states f64 f32 --kernels 0 1 2 --stages 2 > $O/sor_ab.log 2>&1; echo "sor_ab rc $?" | tee -a $O/rc.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc $?" | tee -a $O/rc.txt
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench.log 2> $O/bench.err; echo "bench rc $?" | tee -a $O/rc.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR tools/bench_pipelined.py --shape 128 1024 1024 --channels 1 --min-level 2 --zslab > $O/zslab_ml2.log 2>&1; echo "zslab rc $?" | tee -a $O/rc.txt
cut -c1-200 $O/sor_ab.log; tail -3 $O/pytest.log; grep "^{" $O/zslab_ml2.log | tail -1 | cut -c1-700; python - <<'P'
import json
d=json.loads(open("gpurun_out/m2e/bench.log").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["roofline"])
for k in d["kernels"][:4]: print(k)
P
