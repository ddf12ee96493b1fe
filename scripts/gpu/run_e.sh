#!/bin/bash
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_tests.log); tail -4 gpurun_out/r2e_tests.log
run() { echo "== $*"; env "$@" timeout 300 python scripts/scale_probe.py 1000000 13333 3 2>&1 | grep -E "^outer" | cut -c1-75; }
run SC_X=0 > gpurun_out/r2e_sweep.log
run SC_HUB_U=4 >> gpurun_out/r2e_sweep.log
run SC_R0_SEED=0 >> gpurun_out/r2e_sweep.log
run SC_R0_SEED=8 >> gpurun_out/r2e_sweep.log
run SC_R0_SEED=32 >> gpurun_out/r2e_sweep.log
run SC_HUB_MAX=32 >> gpurun_out/r2e_sweep.log
run SC_HUB_MAX=96 >> gpurun_out/r2e_sweep.log
cat gpurun_out/r2e_sweep.log
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2e_launches_1m.csv python scripts/ncu_probe.py 1000000 > gpurun_out/r2e_ncu.log 2>&1; tail -1 gpurun_out/r2e_ncu.log | cut -c1-200
