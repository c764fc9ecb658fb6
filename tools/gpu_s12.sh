#!/bin/bash
# GPU session 12 (single GPU): final records of the round -- ncu traffic + summary of the solver with the 24-byte state,
# launch list of the bench, bench (both arms), smoke
O=gpurun_out/s12; mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc $?" | tee -a $O/rc.txt
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.log 2> $O/bench.err; echo "bench rc $?" | tee -a $O/rc.txt
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.log 2> $O/bench_ref.err; echo "bench ref rc $?" | tee -a $O/rc.txt
timeout 600 python bench.py --steps 10 --warmup 3 --state f32 --no-cpu-baseline --no-arr-api > $O/bench_f32.log 2> $O/bench_f32.err; echo "bench f32 rc $?" | tee -a $O/rc.txt
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:fr3d_sor_wavefront -s 4 -c 2 \
   -o /tmp/prof_sor_f64 -f python tools/profile_step.py 25 2 > $O/ncu_full.log 2>&1; echo "ncu full rc $?" | tee -a $O/rc.txt
ncu -i /tmp/prof_sor_f64.ncu-rep --page raw --csv > $O/prof_sor_f64_raw.csv 2>/dev/null
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -c 1600 --csv \
   --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-arr-api > $O/ncu_launch.log 2>&1; echo "ncu launches rc $?" | tee -a $O/rc.txt
cat $O/smoke.log | tail -1; cut -c1-300 $O/bench_ref.log; python - <<'P'
import json
for f in ("bench", "bench_f32"):
    d=json.loads(open(f"gpurun_out/s12/{f}.log").read().strip().splitlines()[-1])
    print(f, d["value"], d["e2e"]["value"], d["e2e"].get("compensate_arr_3D_pageable"), d["roofline"]["frac"], d["cpu_baseline"])
P
