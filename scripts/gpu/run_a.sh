#!/bin/bash
# GPU session A: tests, stage profile at 1M, SpMM probe, tile-stream microbench, launch list of one outer iteration
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_tests.log); tail -5 gpurun_out/r2a_tests.log
SC_PROFILE=1 timeout 300 python scripts/scale_probe.py 1000000 13333 3 > gpurun_out/r2a_profile_1m.log 2>&1; tail -12 gpurun_out/r2a_profile_1m.log
timeout 300 python scripts/scale_probe.py 1000000 13333 3 > gpurun_out/r2a_scale_1m.log 2>&1; tail -8 gpurun_out/r2a_scale_1m.log
timeout 120 python scripts/spmm_probe.py 100000 1333 > gpurun_out/r2a_spmm.log 2>&1; cat gpurun_out/r2a_spmm.log
timeout 60 scripts/microbench/tile_stream > gpurun_out/r2a_tile_stream.txt 2>&1; cat gpurun_out/r2a_tile_stream.txt
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2a_launches_1m.csv python scripts/ncu_probe.py 1000000 > gpurun_out/r2a_ncu.log 2>&1; tail -2 gpurun_out/r2a_ncu.log
