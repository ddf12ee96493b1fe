#!/bin/bash
# GPU session B: tests after the explicit-K / in-library sharding rewrite, stage profile at 1M
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_tests.log); tail -25 gpurun_out/r2b_tests.log
SC_PROFILE=1 timeout 300 python scripts/scale_probe.py 1000000 13333 3 > gpurun_out/r2b_profile_1m.log 2>&1; tail -7 gpurun_out/r2b_profile_1m.log
timeout 300 python scripts/scale_probe.py 1000000 13333 3 > gpurun_out/r2b_scale_1m.log 2>&1; grep -E "outer|init|kernel build|mem" gpurun_out/r2b_scale_1m.log | cut -c1-120
