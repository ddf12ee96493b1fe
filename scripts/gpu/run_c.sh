#!/bin/bash
# GPU session C: launch list of one outer iteration + full captures of the per-step kernels
mkdir -p gpurun_out
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2c_launches_1m.csv python scripts/ncu_probe.py 1000000 > gpurun_out/r2c_ncu.log 2>&1; tail -2 gpurun_out/r2c_ncu.log
for K in k_updateB_hub k_kb_merge k_updateB_eval k_push_emit k_rows_from_ckeys; do
  timeout 400 ncu --set full --clock-control none --import-source on --kernel-name regex:$K --launch-skip 12 --launch-count 2 -f -o gpurun_out/r2c_$K python scripts/ncu_probe.py 1000000 > gpurun_out/r2c_ncu_$K.log 2>&1; tail -1 gpurun_out/r2c_ncu_$K.log
done
