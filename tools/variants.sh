#!/bin/bash
# usage: tools/variants.sh  -- bench.py (config 2, both solver state dtypes) for every library build variant
# under build/variants (tuning aid)
for so in base build/variants/*.so; do
  if [ "$so" = base ]; then unset FR3D_LIBRARY_VARIANT; else export FR3D_LIBRARY_VARIANT=$PWD/$so; fi
  for st in f64 f32; do
  python bench.py --steps 3 --warmup 2 --no-cpu-baseline --state $st 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=[x for x in d['kernels'] if 'sor' in x['kernel']][0]
print('$so $st config2 fps',d['value'],'sor_ms_per_step',round(k['ms_total']/d['steps'],2))"
  done
done
