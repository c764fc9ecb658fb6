#!/bin/bash
# GPU session 2 (round 2): the time-blocked tile solver -- bit-identity with the wavefront kernel, timings over tile
# geometries / block sizes, then the whole GPU test-suite (incl. the published-size parity tests) and the bench.
O=gpurun_out/s2; mkdir -p $O
timeout 900 python tools/sor_ab.py --states f64 --kernels 0 2 --tiles 0 5,8,8,16 4,8,8,8 8,8,8,8 5,8,6,6 5,8,16,16 > $O/sor_tiles_c2_f64.log 2>&1; echo "tiles c2 f64 rc $?" | tee -a $O/rc.txt
timeout 600 python tools/sor_ab.py --states f32 --kernels 0 2 --tiles 0 5,8,8,16 8,8,8,8 5,8,16,16 > $O/sor_tiles_c2_f32.log 2>&1; echo "tiles c2 f32 rc $?" | tee -a $O/rc.txt
for v in t256 t192 t96; do
  timeout 600 python tools/sor_ab.py --states f64 --kernels 2 --tiles 0 5,8,8,16 --library build/variants/libfr3d_$v.so > $O/sor_tiles_c2_$v.log 2>&1; echo "variant $v rc $?" | tee -a $O/rc.txt
done
timeout 900 python tools/sor_ab.py --states f64 f32 --min-level 0 --batch 2 --reps 2 --kernels 0 2 --tiles 0 5,8,8,16 > $O/sor_tiles_ml0.log 2>&1; echo "tiles ml0 rc $?" | tee -a $O/rc.txt
timeout 900 python tools/sor_ab.py --states f64 --shape 64 128 128 --channels 1 --batch 1 --kernels 0 2 --tiles 0 > $O/sor_tiles_c1.log 2>&1; echo "tiles c1 rc $?" | tee -a $O/rc.txt
timeout 1500 python -m pytest tests -m gpu -x -q -s > $O/pytest.log 2>&1; echo "pytest rc $?" | tee -a $O/rc.txt
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench.log 2> $O/bench.err; echo "bench rc $?" | tee -a $O/rc.txt
cat $O/sor_tiles_*.log | cut -c1-300; grep -E "config|passed|failed|error" $O/pytest.log | tail -20; cut -c1-900 $O/bench.log
