#!/bin/bash
# N-GPU session (N = number of visible GPUs): final bench line and kNN sweep
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2r_bench_${N}gpu.json 2> gpurun_out/r2r_bench_${N}gpu.err; tail -2 gpurun_out/r2r_bench_${N}gpu.err | cut -c1-300; cut -c1-330 gpurun_out/r2r_bench_${N}gpu.json; python -c "
import json;d=json.load(open('gpurun_out/r2r_bench_${N}gpu.json'));print(d['config']['fit_seconds'], d['config']['construct_kernel_matrix_s'], d.get('parity_vs_1gpu'))"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 scripts/knn_sweep.py > gpurun_out/r2r_knn_sweep_${N}gpu.jsonl 2>/dev/null; cut -c1-200 gpurun_out/r2r_knn_sweep_${N}gpu.jsonl
