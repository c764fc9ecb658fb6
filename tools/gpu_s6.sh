#!/bin/bash
# GPU session 6 (single GPU): TMA-staged spline X pass, host-cost microbench, bench with the f64 default, ncu traffic f64
O=gpurun_out/s6; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc $?" | tee -a $O/rc.txt
timeout 300 python tools/bench_warp.py > $O/warp_tma.log 2>&1
timeout 300 python - > $O/warp_notma.log 2>&1 <<'P'
import sys, re
src = open("tools/bench_warp.py").read().replace('if "--factored" in sys.argv:', 'if True:\n    from flowreg3d_b200 import _lib, core\n    core._check(reg.ctx.h, reg.ctx.lib.fr3d_set_option(reg.ctx.h, _lib.OPT_SPLINE_TMA, 0))\nif "--factored" in sys.argv:')
exec(compile(src, "bench_warp_notma", "exec"))
P
timeout 300 python tools/host_costs.py > $O/host_costs.log 2>&1; echo "host costs rc $?" | tee -a $O/rc.txt
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench.log 2> $O/bench.err; echo "bench rc $?" | tee -a $O/rc.txt
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:fr3d_sor_wavefront -s 4 -c 2 \
   -o $O/prof_sor_f64 -f python tools/profile_step.py 25 2 > $O/ncu_full.log 2>&1; echo "ncu full rc $?" | tee -a $O/rc.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -c 1600 --csv \
   --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-arr-api > $O/ncu_launch.log 2>&1; echo "ncu launches rc $?" | tee -a $O/rc.txt
tail -3 $O/pytest.log; cat $O/warp_tma.log $O/warp_notma.log | tail -2; cat $O/host_costs.log; cut -c1-400 $O/bench.log
