#!/bin/bash
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_tests.log); tail -4 gpurun_out/r2h_tests.log
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2h_launches_1m.csv python scripts/ncu_probe.py 1000000 > gpurun_out/r2h_ncu.log 2>&1; tail -1 gpurun_out/r2h_ncu.log | cut -c1-200
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; tail -3 gpurun_out/r2h_bench.err; cat gpurun_out/r2h_bench.json
