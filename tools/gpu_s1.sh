#!/bin/bash
# GPU session 1 (round 2): validate the staged solver kernel + the opt-in kernels written after round 1's GPU budget
# was spent; A/B timings; ncu of the staged kernel.  Every step has its own timeout.
O=gpurun_out/s1; mkdir -p $O
nvidia-smi -L > $O/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc $?" | tee -a $O/rc.txt
for st in f64 f32; do
  timeout 600 python tools/sor_ab.py --states $st --stages 2 3 4 6 > $O/sor_ab_c2_$st.log 2>&1; echo "sor_ab c2 $st rc $?" | tee -a $O/rc.txt
done
timeout 600 python tools/sor_ab.py --states f64 f32 --min-level 0 --batch 2 --stages 3 6 --reps 2 > $O/sor_ab_ml0.log 2>&1; echo "sor_ab ml0 rc $?" | tee -a $O/rc.txt
timeout 300 python tools/bench_warp.py > $O/warp.log 2>&1
timeout 300 python tools/bench_warp.py --factored > $O/warp_factored.log 2>&1
timeout 300 python tools/bench_cc.py 10 > $O/cc.log 2>&1
timeout 300 python tools/bench_cc.py 10 --block-scans > $O/cc_block.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench.log 2> $O/bench.err; echo "bench rc $?" | tee -a $O/rc.txt
timeout 600 python bench.py --steps 10 --warmup 3 --state f32 --no-cpu-baseline > $O/bench_f32.log 2> $O/bench_f32.err; echo "bench f32 rc $?" | tee -a $O/rc.txt
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:fr3d_sor_staged -s 2 -c 2 \
   -o $O/prof_sor_staged -f python tools/profile_step.py 25 2 > $O/ncu_full.log 2>&1; echo "ncu full rc $?" | tee -a $O/rc.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -c 1400 --csv \
   --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_launch.log 2>&1; echo "ncu launches rc $?" | tee -a $O/rc.txt
tail -3 $O/pytest.log; cat $O/sor_ab_c2_f64.log $O/sor_ab_c2_f32.log $O/sor_ab_ml0.log | cut -c1-330; cat $O/warp.log $O/warp_factored.log $O/cc.log $O/cc_block.log | tail -12; cut -c1-600 $O/bench.log
