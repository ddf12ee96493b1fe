#!/bin/bash
# capture session: launch list of one outer iteration, ncu --set full of the dominant kernels, bench lines (ours + reference arm)
mkdir -p gpurun_out
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2p_launches_1m.csv python scripts/ncu_probe.py 1000000 > gpurun_out/r2p_ncu.log 2>&1; tail -1 gpurun_out/r2p_ncu.log | cut -c1-120
for K in k_updateB_hub k_updateB_eval k_kb_merge_staged; do
  timeout 400 ncu --set full --clock-control none --import-source on --kernel-name regex:$K --launch-skip 12 --launch-count 2 -f -o gpurun_out/r2p_$K python scripts/ncu_probe.py 1000000 > gpurun_out/r2p_ncu_$K.log 2>&1; tail -1 gpurun_out/r2p_ncu_$K.log | cut -c1-100
done
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name regex:k_spmm_csr_panel --launch-skip 3 --launch-count 1 -f -o gpurun_out/r2p_k_spmm_csr_panel python scripts/spmm_probe.py 100000 1333 5 > gpurun_out/r2p_ncu_spmm.log 2>&1; tail -1 gpurun_out/r2p_ncu_spmm.log | cut -c1-200
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2p_bench_1gpu.json 2> gpurun_out/r2p_bench_1gpu.err; tail -2 gpurun_out/r2p_bench_1gpu.err | cut -c1-200; cut -c1-400 gpurun_out/r2p_bench_1gpu.json
timeout 1500 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2p_bench_reference.json 2> gpurun_out/r2p_bench_reference.err; cut -c1-600 gpurun_out/r2p_bench_reference.json
