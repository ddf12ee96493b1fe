#!/bin/bash
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2l_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_tests.log); tail -4 gpurun_out/r2l_tests.log
timeout 300 python scripts/fit_sweep.py 1000000 "" > gpurun_out/r2l_sweep.log 2>&1; cat gpurun_out/r2l_sweep.log
timeout 400 python scripts/knn_sweep.py > gpurun_out/r2l_knn_sweep_1gpu.jsonl 2> gpurun_out/r2l_knn_sweep.err; cat gpurun_out/r2l_knn_sweep_1gpu.jsonl; tail -2 gpurun_out/r2l_knn_sweep.err
timeout 300 python bench.py --cells 100000 --steps 3 --warmup 3 --no-extras > gpurun_out/r2l_bench_c2_100k.json 2>/dev/null; cut -c1-900 gpurun_out/r2l_bench_c2_100k.json
timeout 300 python bench.py --cells 200000 --pcs 30 --kind atac --steps 3 --warmup 3 --no-extras > gpurun_out/r2l_bench_c3_200k_atac.json 2>/dev/null; cut -c1-900 gpurun_out/r2l_bench_c3_200k_atac.json
