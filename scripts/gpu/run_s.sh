#!/bin/bash
# final single-GPU capture: smoke, bench line, launch list, ncu --set full of the dominant kernel
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2s_bench_1gpu.json 2> gpurun_out/r2s_bench_1gpu.err; tail -1 gpurun_out/r2s_bench_1gpu.err | cut -c1-200; cut -c1-300 gpurun_out/r2s_bench_1gpu.json
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2s_launches_1m.csv python scripts/ncu_probe.py 1000000 > gpurun_out/r2s_ncu.log 2>&1; tail -1 gpurun_out/r2s_ncu.log | cut -c1-120
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name regex:k_updateB_hub --launch-skip 12 --launch-count 1 -f -o gpurun_out/r2s_k_updateB_hub python scripts/ncu_probe.py 1000000 > gpurun_out/r2s_ncu_hub.log 2>&1; tail -1 gpurun_out/r2s_ncu_hub.log | cut -c1-100
