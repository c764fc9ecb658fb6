#!/bin/bash
# final check of the round on one GPU: the whole GPU suite, smoke(), the driver's bench invocation
O=gpurun_out/final; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc $?" | tee -a $O/rc.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc $?" | tee -a $O/rc.txt
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench.log 2> $O/bench.err; echo "bench rc $?" | tee -a $O/rc.txt
tail -2 $O/pytest.log; tail -1 $O/smoke.log; cut -c1-200 $O/bench.log
