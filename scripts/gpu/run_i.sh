#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/fit_sweep.py 1000000 "" "SC_R0_SEED=0" "SC_R0_SEED=32" "SC_R0_SEED=64" "SC_HUB_MAX=96" "SC_HUB_MAX=128" > gpurun_out/r2i_sweep.log 2>&1; cat gpurun_out/r2i_sweep.log
