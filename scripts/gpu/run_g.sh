#!/bin/bash
mkdir -p gpurun_out
for K in k_updateB_hub k_push_gather k_kb_merge_staged; do
  timeout 400 ncu --set full --clock-control none --import-source on --kernel-name regex:$K --launch-skip 12 --launch-count 1 -f -o gpurun_out/r2g_$K python scripts/ncu_probe.py 1000000 > gpurun_out/r2g_ncu_$K.log 2>&1; tail -1 gpurun_out/r2g_ncu_$K.log | cut -c1-100
done
