#!/bin/bash
mkdir -p gpurun_out/sched
timeout 400 python tools/sor_ab.py --kernels 0 --states f64 f32 --reps 3 --sched 0 128 384 640 896 > gpurun_out/sched/split2.jsonl 2> gpurun_out/sched/split2.err
echo "rc $?"
for v in bar3_fences_only bar4_no_ivall; do
  timeout 200 python tools/sor_ab.py --library $PWD/build/variants/$v.so --kernels 0 --states f64 f32 --reps 3 --sched 0 > gpurun_out/sched/$v.jsonl 2> gpurun_out/sched/$v.err
  echo "$v rc $?"
done
python - <<'PY'
import json,glob
for f in ['split2','bar3_fences_only','bar4_no_ivall']:
    print('==',f)
    for l in open(f'gpurun_out/sched/{f}.jsonl'):
        d=json.loads(l)
        if 'sor_ms' in d: print(d['state'], d['sched'], d['sor_ms'], d['bit_identical_to_direct'], d['max_abs_diff'])
        elif 'error' in d: print(d)
PY
