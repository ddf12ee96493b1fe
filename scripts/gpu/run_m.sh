#!/bin/bash
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_tests.log); tail -4 gpurun_out/r2m_tests.log
timeout 300 python scripts/fit_sweep.py 1000000 "" > gpurun_out/r2m_sweep.log 2>&1; cat gpurun_out/r2m_sweep.log
timeout 120 python scripts/fit_sweep.py 10000 "" > gpurun_out/r2m_sweep_c1.log 2>&1; cat gpurun_out/r2m_sweep_c1.log
timeout 120 python scripts/fit_sweep.py 100000 "" > gpurun_out/r2m_sweep_c2.log 2>&1; cat gpurun_out/r2m_sweep_c2.log
