#!/bin/bash
# GPU session 4: wavefront kernel with L2 prefetch of the next item (vs without), occupancy variants
O=gpurun_out/s4; mkdir -p $O
timeout 600 python tools/sor_ab.py --states f64 f32 --kernels 0 > $O/base_c2.log 2>&1; echo "base rc $?" | tee -a $O/rc.txt
for v in nopf f32m4 f64m3; do
  timeout 600 python tools/sor_ab.py --states f64 f32 --kernels 0 --library build/variants/libfr3d_$v.so > $O/${v}_c2.log 2>&1; echo "variant $v rc $?" | tee -a $O/rc.txt
done
timeout 600 python tools/sor_ab.py --states f64 f32 --kernels 0 --min-level 0 --batch 2 --reps 2 > $O/base_ml0.log 2>&1
timeout 600 python tools/sor_ab.py --states f64 f32 --kernels 0 --min-level 0 --batch 2 --reps 2 --library build/variants/libfr3d_nopf.so > $O/nopf_ml0.log 2>&1
timeout 600 python -m pytest tests -m gpu -x -q -k "sor or config2 or parity_stages or smoke or mid_size" > $O/pytest.log 2>&1; echo "pytest rc $?" | tee -a $O/rc.txt
for f in base_c2 nopf_c2 f32m4_c2 f64m3_c2 base_ml0 nopf_ml0; do echo "== $f"; cut -c1-200 $O/$f.log; done; tail -3 $O/pytest.log
