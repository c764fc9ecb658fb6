#!/bin/bash
# Split-phase waves and barrier diagnostics (results/r02_sor_sched.md).  The split-phase kernel is compiled only with
# -DFR3D_SOR_SPLIT_EXPERIMENT (build/variants/split.so below); FR3D_OPT_SOR_SCHED = 1024 + (mode << 8) selects it
# (when this was measured it was part of the default build under the values 128 / 384 / 640 / 896).
mkdir -p gpurun_out/sched
timeout 400 python tools/sor_ab.py --library $PWD/build/variants/split.so --kernels 0 --states f64 f32 --reps 3 --sched 0 1024 1280 1536 1792 > gpurun_out/sched/split2.jsonl 2> gpurun_out/sched/split2.err
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
