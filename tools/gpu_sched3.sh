#!/bin/bash
mkdir -p gpurun_out/sched
timeout 600 python -m pytest tests/test_parity_stages.py -q -m gpu -k "item_scheduling or wavefront_reproduces" 2>&1 | tail -3
timeout 500 python tools/sor_ab.py --kernels 0 --states f64 f32 --reps 3 --sched 0 128 138 148 158 20 > gpurun_out/sched/bal.jsonl 2> gpurun_out/sched/bal.err
echo "rc $?"; tail -2 gpurun_out/sched/bal.err
python - <<'PY'
import json
for l in open('gpurun_out/sched/bal.jsonl'):
    d=json.loads(l)
    if 'sor_ms' in d: print(d['state'], d['sched'], d['sor_ms'], d['frac_of_6453'], d['bit_identical_to_direct'])
    elif 'error' in d: print(d)
PY
