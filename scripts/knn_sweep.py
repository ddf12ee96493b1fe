"""BASELINE config 5: kNN / kernel-construction-only sweep (n x 50 PCs), one JSON line per size.
    python scripts/knn_sweep.py [sizes...]                 single GPU
    torchrun --nproc-per-node N scripts/knn_sweep.py ...   query slabs sharded over N GPUs (in-library all-gather)
Reports construct_kernel_matrix seconds (sc_kernel_build: kNN + sigma + symmetrise + Gaussian + CSR), the device time of the
tcgen05 shortlist kernel, algorithmic TFLOP/s = 2 n_q n d / t and its fraction of the measured bf16 peak."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch
from seacells_b200 import _lib, build_kernel
from seacells_b200.synth import synth_embedding

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
ctx = _lib.default_context(local)
if world > 1:
    import torch.distributed as dist
    from seacells_b200.parallel import init_nccl
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    init_nccl(ctx)
sizes = [int(float(a)) for a in sys.argv[1:]] or [100_000, 250_000, 500_000, 1_000_000, 2_000_000, 4_000_000]
peak = 1375.4
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["bf16_tflops_sustained"])
except Exception:
    pass
build_kernel(synth_embedding(4096, 50, 0), 15, "union", _lib.Context(local)).close()
for n in sizes:
    X = torch.from_numpy(synth_embedding(n, 50, 0)).cuda()   # embedding resident in HBM: the hand-off is a device pointer
    ts = []
    for rep in range(2):
        ctx.sync(); torch.cuda.synchronize()
        if world > 1: dist.barrier()
        t = time.perf_counter(); km = build_kernel(X, 15, "union", ctx); ctx.sync(); ts.append(time.perf_counter() - t)
        nnz = km.nnz; km.close()
    p4 = (C.c_double * 4)(); ctx.lib.sc_knn_profile(ctx.h, p4)
    nq = (n + world - 1) // world
    filt = p4[0] / 1e3
    if rank == 0:
        print(json.dumps({"n": n, "d": 50, "n_gpus": world, "construct_kernel_matrix_s": min(ts), "knn_filter_ms": p4[0],
                          "filter_kind": int(p4[1]), "fallback_rows": int(p4[3]), "nnz_per_row": nnz / n,
                          "algorithmic_tflops_per_gpu": 2.0 * nq * n * 50 / filt / 1e12 if filt > 0 else None,
                          "frac_of_bf16_peak": 2.0 * nq * n * 50 / filt / 1e12 / peak if filt > 0 else None}), flush=True)
    del X; torch.cuda.empty_cache()
if world > 1:
    dist.destroy_process_group()
