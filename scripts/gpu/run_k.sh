#!/bin/bash
# 8-GPU session: sharded probe with stage profile, bench at N = 8 and N = 4
mkdir -p gpurun_out
SC_PROFILE=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 scripts/scale_probe_sharded.py 1000000 3 > gpurun_out/r2k_probe8.log 2>&1; grep -E "^world|^outer" gpurun_out/r2k_probe8.log | cut -c1-420
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2k_bench8.json 2> gpurun_out/r2k_bench8.err; tail -2 gpurun_out/r2k_bench8.err | cut -c1-300; cut -c1-1200 gpurun_out/r2k_bench8.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/r2k_bench4.json 2> gpurun_out/r2k_bench4.err; tail -2 gpurun_out/r2k_bench4.err | cut -c1-300; cut -c1-700 gpurun_out/r2k_bench4.json
