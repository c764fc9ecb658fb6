#!/bin/bash
# GPU session 9: source-level ncu of the non-solver kernels, summarised ON the box (the report itself is too large to bring back)
O=gpurun_out/s9; mkdir -p $O
timeout 1200 ncu --section SourceCounters --section SpeedOfLight --section WarpStateStats --section InstructionStats --section LaunchStats --section Occupancy --section MemoryWorkloadAnalysis \
   --clock-control none --import-source on --kernel-name-base demangled \
   -k regex:"Median5PairK|WarpGatherLeanK|ResizePassK|PreYXWinK|PreZWinK|SplineTileK|SplineZK|AssembleK" -s 60 -c 44 \
   -o /tmp/prof_others -f python tools/profile_step.py 25 2 > $O/ncu.log 2>&1; echo "ncu rc $?" | tee -a $O/rc.txt
python tools/ncu_hotspots.py /tmp/prof_others.ncu-rep $O > $O/hot.log 2>&1
ls -la $O /tmp/prof_others.ncu-rep; tail -5 $O/hot.log
