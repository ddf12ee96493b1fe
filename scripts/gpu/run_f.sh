#!/bin/bash
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_tests.log); tail -4 gpurun_out/r2f_tests.log
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2f_launches_1m.csv python scripts/ncu_probe.py 1000000 > gpurun_out/r2f_ncu.log 2>&1; tail -1 gpurun_out/r2f_ncu.log | cut -c1-200
SC_HUB_U=4 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2f_launches_1m_u4.csv python scripts/ncu_probe.py 1000000 > gpurun_out/r2f_ncu_u4.log 2>&1
