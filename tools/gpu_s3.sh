#!/bin/bash
# GPU session 3: reworked tile solver (prepass refresh, smem row bases, L2 prefetch): phase timing + A/B
O=gpurun_out/s3; mkdir -p $O
timeout 600 python tools/sor_ab.py --states f64 --kernels 2 --tiles 0 --reps 1 --library build/variants/libfr3d_timing.so > $O/timing_c2.log 2>&1; echo "timing rc $?" | tee -a $O/rc.txt
timeout 900 python tools/sor_ab.py --states f64 --kernels 0 2 --tiles 0 5,8,8,16 5,8,16,8 5,8,6,6 5,8,12,12 > $O/sor_tiles_c2_f64.log 2>&1; echo "tiles c2 f64 rc $?" | tee -a $O/rc.txt
timeout 600 python tools/sor_ab.py --states f32 --kernels 0 2 --tiles 0 5,8,8,16 5,8,16,16 > $O/sor_tiles_c2_f32.log 2>&1; echo "tiles c2 f32 rc $?" | tee -a $O/rc.txt
for v in t256 t192 t64; do
  timeout 600 python tools/sor_ab.py --states f64 --kernels 2 --tiles 0 5,8,8,16 --library build/variants/libfr3d_$v.so > $O/sor_tiles_c2_$v.log 2>&1; echo "variant $v rc $?" | tee -a $O/rc.txt
done
timeout 900 python tools/sor_ab.py --states f64 --min-level 0 --batch 2 --reps 2 --kernels 0 2 --tiles 0 > $O/sor_tiles_ml0.log 2>&1; echo "tiles ml0 rc $?" | tee -a $O/rc.txt
timeout 600 python tools/sor_ab.py --states f64 --min-level 0 --batch 2 --reps 1 --kernels 2 --tiles 0 --library build/variants/libfr3d_timing.so > $O/timing_ml0.log 2>&1
timeout 900 python tools/sor_ab.py --states f64 --shape 64 128 128 --channels 1 --batch 1 --kernels 0 2 --tiles 0 > $O/sor_tiles_c1.log 2>&1; echo "tiles c1 rc $?" | tee -a $O/rc.txt
grep "tile timing" $O/timing_c2.log | tail -4; grep "tile timing" $O/timing_ml0.log | tail -8; cat $O/sor_tiles_*.log | cut -c1-250
