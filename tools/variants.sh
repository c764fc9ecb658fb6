#!/bin/bash
# usage: tools/variants.sh  -- time config 1 (B = 1, wave-latency bound) and config 2 for every library build
# variant under build/variants (tuning aid)
for so in base build/variants/*.so; do
  if [ "$so" = base ]; then unset FR3D_LIBRARY_VARIANT; else export FR3D_LIBRARY_VARIANT=$PWD/$so; fi
  echo "== $so"
  python tools/bench_configs.py --quick 2>/dev/null | python -c "
import json,sys
for ln in sys.stdin:
    d=json.loads(ln)
    if 'ms_per_step' in d and ('config1' in d['case']): print('  ',d['case'], d['ms_per_step'], d['top_kernels_ms'][0])
"
  python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=[x for x in d['kernels'] if 'sor' in x['kernel']][0]
print('   config2 fps',d['value'],'sor_ms_per_step',round(k['ms_total']/d['steps'],2))"
done
