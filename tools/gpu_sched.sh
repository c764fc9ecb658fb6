#!/bin/bash
# Item scheduling of the wavefront kernel (FR3D_OPT_SOR_SCHED = percent of each wave dealt by ticket).  The first run of
# this script also carried the early-locate / first-item-prefetch switches (values 256 / 768), which were removed
# after they measured no gain (results/r02_sor_sched.md).
mkdir -p gpurun_out/sched
timeout 900 python tools/sor_ab.py --kernels 0 --states f64 f32 --reps 3 \
    --sched 0 10 20 30 40 60 100 > gpurun_out/sched/ab.jsonl 2> gpurun_out/sched/ab.err
echo "ab rc $?" > gpurun_out/sched/rc.txt
tail -3 gpurun_out/sched/ab.err
python - <<'PY'
import json
for l in open('gpurun_out/sched/ab.jsonl'):
    d = json.loads(l)
    if 'sor_ms' in d:
        print(d['state'], d['sched'], d['sor_ms'], d['frac_of_6453'], d['bit_identical_to_direct'])
    else:
        print(d)
PY
