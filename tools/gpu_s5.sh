#!/bin/bash
# GPU session 5 (single GPU): frames-per-item sweep of the wavefront kernel, full test-suite, bench (both arms), ncu
O=gpurun_out/s5; mkdir -p $O
for fg in 1 2 3 4 5 8; do
  timeout 300 python tools/sor_ab.py --frames-per-item $fg --states f64 f32 --kernels 0 > $O/fg$fg.log 2>&1; echo "fg $fg rc $?" | tee -a $O/rc.txt
done
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc $?" | tee -a $O/rc.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc $?" | tee -a $O/rc.txt
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_ref.log 2> $O/bench_ref.err; echo "bench ref rc $?" | tee -a $O/rc.txt
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench.log 2> $O/bench.err; echo "bench rc $?" | tee -a $O/rc.txt
timeout 600 python bench.py --steps 10 --warmup 3 --state f64 --no-cpu-baseline --no-arr-api > $O/bench_f64.log 2> $O/bench_f64.err; echo "bench f64 rc $?" | tee -a $O/rc.txt
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:fr3d_sor_wavefront -s 4 -c 2 \
   -o $O/prof_sor -f python tools/profile_step.py 25 2 > $O/ncu_full.log 2>&1; echo "ncu full rc $?" | tee -a $O/rc.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -c 1600 --csv \
   --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-arr-api > $O/ncu_launch.log 2>&1; echo "ncu launches rc $?" | tee -a $O/rc.txt
for fg in 1 2 3 4 5 8; do echo "== fg $fg"; cut -c1-140 $O/fg$fg.log | head -2; done; tail -3 $O/pytest.log; cat $O/smoke.log | tail -2; cut -c1-300 $O/bench_ref.log; cut -c1-2500 $O/bench.log
