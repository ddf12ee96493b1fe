#!/bin/bash
# last session of the round: full GPU test-suite on the final build, then the 1M fit under the two summation switches
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2v_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2v_tests.log); tail -3 gpurun_out/r2v_tests.log
FIT_SWEEP_REPS=1 timeout 200 python scripts/fit_sweep.py 1000000 "" "SC_FIT_RESUM_A=1" "SC_FIT_FRESH_KB=1" > gpurun_out/r2v_variants_1m.log 2>&1; cut -c1-175 gpurun_out/r2v_variants_1m.log
