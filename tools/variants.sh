#!/bin/bash
# usage: tools/variants.sh  -- bench every SOR build variant under build/variants (tuning aid)
for so in base build/variants/*.so; do
  for st in f64 f32; do
    if [ "$so" = base ]; then unset FR3D_LIBRARY_VARIANT; else export FR3D_LIBRARY_VARIANT=$PWD/$so; fi
    python bench.py --steps 3 --warmup 2 --no-cpu-baseline --state $st 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=[x for x in d['kernels'] if 'sor' in x['kernel']][0]
print('$so','$st','fps',d['value'],'sor_ms_per_step',round(k['ms_total']/d['steps'],2))"
  done
done
