#!/bin/bash
# 2-GPU session: full tests (NCCL + two-device tests active), sharded probe, bench at N=2, single-GPU sweep on GPU 0
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2j_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_tests.log); tail -6 gpurun_out/r2j_tests.log
SC_PROFILE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/scale_probe_sharded.py 1000000 3 > gpurun_out/r2j_probe2.log 2>&1; grep -E "^world|^outer" gpurun_out/r2j_probe2.log | cut -c1-420
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 2 > gpurun_out/r2j_bench2.json 2> gpurun_out/r2j_bench2.err; tail -3 gpurun_out/r2j_bench2.err | cut -c1-300; cat gpurun_out/r2j_bench2.json | cut -c1-1500
timeout 300 python scripts/fit_sweep.py 1000000 "" > gpurun_out/r2j_sweep.log 2>&1; cat gpurun_out/r2j_sweep.log
