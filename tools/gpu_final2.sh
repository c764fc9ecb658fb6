#!/bin/bash
# final check of the round on one GPU (ticket-dealt waves at 384 threads x 168 registers as the float64 default): the whole GPU suite,
# smoke(), the driver's bench invocation (both arms), ncu capture + launch list of the new default solver kernel
O=gpurun_out/final3; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc $?" | tee -a $O/rc.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc $?" | tee -a $O/rc.txt
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench.log 2> $O/bench.err; echo "bench rc $?" | tee -a $O/rc.txt
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:fr3d_sor_wavefront -s 4 -c 2 \
   -o /tmp/prof_sor_f64 -f python tools/profile_step.py 25 2 > $O/ncu_full.log 2>&1; echo "ncu full rc $?" | tee -a $O/rc.txt
ncu -i /tmp/prof_sor_f64.ncu-rep --page raw --csv > $O/prof_sor_f64_raw.csv 2>/dev/null
ncu -i /tmp/prof_sor_f64.ncu-rep --page details > $O/prof_sor_f64_details.txt 2>/dev/null
cp profiles/r02_sor_traffic.json $O/traffic_before.json
python tools/ncu_traffic.py $O/prof_sor_f64_raw.csv f64 25 > $O/traffic.log 2>&1; cp profiles/r02_sor_traffic.json $O/r02_sor_traffic.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -c 1600 --csv \
   --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-arr-api > $O/ncu_launch.log 2>&1; echo "ncu launches rc $?" | tee -a $O/rc.txt
tail -2 $O/pytest.log; tail -1 $O/smoke.log; cut -c1-200 $O/bench.log; tail -2 $O/traffic.log | cut -c1-400
