#!/bin/bash
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_tests.log); tail -15 gpurun_out/r2d_tests.log
timeout 300 python scripts/scale_probe.py 1000000 13333 3 > gpurun_out/r2d_scale_1m.log 2>&1; grep -E "outer|init|kernel build|mem|setup" gpurun_out/r2d_scale_1m.log | cut -c1-100
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2d_launches_1m.csv python scripts/ncu_probe.py 1000000 > gpurun_out/r2d_ncu.log 2>&1; tail -2 gpurun_out/r2d_ncu.log | cut -c1-300
