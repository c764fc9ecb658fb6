#!/bin/bash
# GPU session 7 (single GPU): median fast recovery, new compensate_arr_3D, per-pair latency, full tests + bench
O=gpurun_out/s7; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc $?" | tee -a $O/rc.txt
timeout 300 python tools/pair_latency.py > $O/pair_latency.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench.log 2> $O/bench.err; echo "bench rc $?" | tee -a $O/rc.txt
tail -3 $O/pytest.log; cat $O/pair_latency.log; python - <<'P'
import json
d=json.loads(open("gpurun_out/s7/bench.log").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["e2e"]["compensate_arr_3D_pageable"], d["roofline"])
for k in d["kernels"]: print(k)
P
