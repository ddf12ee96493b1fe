#!/bin/bash
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2o_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o_tests.log); tail -4 gpurun_out/r2o_tests.log
timeout 400 python scripts/fit_sweep.py 1000000 "" "SC_NO_OVERLAP=1" > gpurun_out/r2o_sweep.log 2>&1; cat gpurun_out/r2o_sweep.log
