#!/bin/bash
# GPU session 8 (single GPU): ncu --set full of the non-solver kernels of a step (one launch each), per-pair latency
O=gpurun_out/s8; mkdir -p $O
timeout 300 python tools/pair_latency.py > $O/pair_latency.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
   -k regex:"Median5PairK|WarpGatherLeanK|ResizePassK|PreYXWinK|PreZWinK|SplineTileK|SplineZK|AssembleK" -s 60 -c 44 \
   -o $O/prof_others -f python tools/profile_step.py 25 2 > $O/ncu_full.log 2>&1; echo "ncu full rc $?" | tee -a $O/rc.txt
cat $O/pair_latency.log; tail -3 $O/ncu_full.log
