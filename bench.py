#!/usr/bin/env python
"""bench.py - fit() throughput of the B200 engine on BASELINE.json's headline configuration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--cells n]

One "step" = one complete SEACells fit() (initialize + Frank-Wolfe outer iterations until the reference's convergence
rule fires, min_iter=10) on the synthetic 1M-cell x 50-PC workload (n_SEACells = n/75 = 13333, n_neighbors=15, union
graph); metric = cells/s = n / fit seconds.  Kernel-matrix construction is set-up, reported separately
(`config.construct_kernel_matrix_s`), exactly as the reference's fit() excludes it.

Each timed step is bracketed by CUDA events on the library's stream:
    e0 -H2D (kernel matrix CSR, archetypes, initial assignments, pinned host)-> e1 -fit-> e2 -D2H (labels, RSS)-> e3
`value` uses e1..e2 (inputs resident in HBM), `e2e` uses e0..e3 (host buffers in, host results out).
Inputs (>= 0.3 GB kernel matrix, >= 1 GB of lists) are far larger than the 126 MB L2, so no explicit L2 flush is needed.

`--impl reference` times the CPU restatement of the reference (oracle/, pinned bit-exact against the unmodified
reference by tests/golden) on a bounded sample of the same workload; /root/reference is never read at run time.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "fit_cells_per_s"
UNIT = "cells/s"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def _knn_roofline(n, d, world, filter_s, kind, kp, fallback_rows):
    """Tensor-pipe roofline of the kNN shortlist kernel (hot path 1; not part of fit()): the distance contraction of this
    rank's query slab against all n cells.  algorithmic = 2 n_q n d; executed = 3 split products over the padded K."""
    if filter_s <= 0 or kind == 0:
        return None
    nq = (n + world - 1) // world
    alg = 2.0 * nq * n * d / filter_s / 1e12
    out = {"kernel": {1: "k_knn_filter (fp32 FFMA tiles)",
                      2: "k_knn_filter_tc (tcgen05 kind::f16, 2xFP16 split = 22-bit operands, fp32 accumulate)"}[kind],
           "filter_ms": filter_s * 1e3, "queries": nq, "algorithmic_tflops": alg, "fallback_rows": fallback_rows}
    if kind == 2:
        ex = 3.0 * 2.0 * nq * n * kp / filter_s / 1e12
        peak, src = 1375.4, "fallback"
        p = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(p):
            try:
                peak, src = float(json.load(open(p))["bf16_tflops_sustained"]), "measured bf16_tflops_sustained"
            except Exception:
                pass
        out.update({"bound": "tensor", "executed_f16_tflops": ex, "peak": peak, "unit": "TFLOP/s", "frac": ex / peak,
                    "peak_source": src + " (fp16 and bf16 MMAs run at the same rate)"})
    return out


def _traffic(n, world):
    """DRAM bytes per launch group from the committed ncu --set full capture (1M cells, 1 GPU only)."""
    p = os.path.join(ROOT, "profiles", "r1_updB_group_traffic.json")
    if n == 1_000_000 and world == 1 and os.path.exists(p):
        try:
            return float(json.load(open(p))["dram_bytes_per_launch_group"])
        except Exception:
            return None
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------- CPU arm / baseline
def cpu_fit_sample(n, seed=0):
    """One full fit() of the oracle (CPU restatement of cpu.py) on an n-cell sample of the workload; returns seconds."""
    from oracle import seacells_oracle as O
    from seacells_b200.synth import synth_archetypes, synth_assignments, synth_embedding
    k = max(2, n // 75)
    X = synth_embedding(n, 50, seed).astype(np.float64)
    M, *_ = O.build_kernel(X, 15)  # set-up, untimed (fit() excludes construct_kernel_matrix)
    arch, A0 = synth_archetypes(n, k), synth_assignments(n, k)
    t = time.perf_counter()
    r = O.fit(M, arch, A0, max_iter=100, min_iter=10, convergence_epsilon=1e-3, max_FW_iter=50)
    dt = time.perf_counter() - t
    return dt, r["n_iter"], k


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    total = args.steps + args.warmup
    n = 4000 if total <= 4 else (2500 if total <= 8 else 1600)
    for _ in range(args.warmup):
        cpu_fit_sample(n)
    ts, iters = [], 0
    for _ in range(args.steps):
        dt, iters, k = cpu_fit_sample(n)
        ts.append(dt)
    sec = float(np.mean(ts))
    val = n / sec
    sample = (f"full fit() (initialize + {iters} outer iterations, T=50) of the oracle restatement of cpu.SEACellsCPU on a "
              f"{n}-cell x 50-PC sample (k={k}) of the workload; scipy sparse, single thread")
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": workload_name(args.cells), "sample_cells": n},
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample,
                            "host_cores_available": os.cpu_count()},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(out)


def workload_name(n, pcs=50):
    return (f"synthetic {n} cells x {pcs} PCs (Gaussian mixture, decaying spectrum), n_SEACells={n // 75}, n_neighbors=15, "
            f"union graph, fit(min_iter=10, max_iter=100, eps=1e-3, max_FW_iter=50), given archetypes + sparse A0")


# --------------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from sklearn.preprocessing import normalize

    from seacells_b200 import FitEngine, _lib, build_kernel
    from seacells_b200.synth import synth_archetypes, synth_assignments, synth_embedding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = _lib.default_context(local)
    lib_stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))

    # N > 1: ONE n-cell fit, cells row-sharded over the ranks (strong scaling; seacells_b200/parallel.py).
    # Every rank holds the same embedding; the kNN queries are sharded too, the kernel matrix is replicated.
    from seacells_b200.parallel import ShardedFit, sharded_knn
    n = args.cells
    k = n // 75
    X = synth_embedding(n, args.pcs, seed=0, kind=args.kind)
    build_kernel(X[:4096], 15, "union", ctx).close()  # untimed: CUDA context, module load, first allocations
    ctx.sync()
    t0 = time.perf_counter()
    knn_graph = sharded_knn(X, 15, ctx) if world > 1 else None
    km = build_kernel(X, 15, "union", ctx, knn_graph=knn_graph)
    ctx.sync()
    t_build = time.perf_counter() - t0
    import ctypes as C
    kp4 = (C.c_double * 4)()
    ctx.lib.sc_knn_profile(ctx.h, kp4)
    knn_info = _knn_roofline(n, args.pcs, world, kp4[0] / 1e3, int(kp4[1]), int(kp4[2]), int(kp4[3]))
    M, _, _, _ = km.export(want_knn=False)
    km.close()
    # pinned host buffers: these are what every timed step copies to the device
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h_indptr, h_indices, h_data = pin(M.indptr.astype(np.int32)), pin(M.indices.astype(np.int32)), pin(M.data)
    arch = synth_archetypes(n, k)
    A0 = normalize(synth_assignments(n, k), axis=0, norm="l1").tocsc()
    A0.sort_indices()
    h_colptr, h_rows, h_vals = pin(A0.indptr.astype(np.int64)), pin(A0.indices.astype(np.int32)), pin(A0.data)
    h_labels = torch.empty(n, dtype=torch.int32).pin_memory()
    h2d = sum(t.numel() * t.element_size() for t in (h_indptr, h_indices, h_data, h_colptr, h_rows, h_vals)) + 8 * k
    d2h = 4 * n + 8 * 102

    eng = FitEngine(n, k, 50, ctx)
    sharded = ShardedFit(eng) if world > 1 else None
    lib = ctx.lib
    ptr = _lib.ptr

    def step():
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record(lib_stream)
        ctx.check(lib.sc_fit_set_kernel(eng.h, ptr(h_indptr), ptr(h_indices), ptr(h_data), int(M.nnz)), "set_kernel")
        eng.set_archetypes(arch)
        ctx.check(lib.sc_fit_set_A_csc(eng.h, ptr(h_colptr), ptr(h_rows), ptr(h_vals)), "set_A")
        ev[1].record(lib_stream)
        if sharded is not None:
            rss, n_iter, conv, thr = sharded.run(max_iter=100, min_iter=10, eps=1e-3)
        else:
            rss, n_iter, conv, thr = eng.run(max_iter=100, min_iter=10, eps=1e-3)
        ev[2].record(lib_stream)
        ctx.check(lib.sc_fit_hard_assignments(eng.h, ptr(h_labels)), "labels")
        ev[3].record(lib_stream)
        ev[3].synchronize()
        return ev[1].elapsed_time(ev[2]) / 1e3, ev[0].elapsed_time(ev[3]) / 1e3, n_iter, conv, rss

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(args.warmup):
        step()
    st0 = eng.stats()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launches
    dev_s, e2e_s = [], []
    n_iter = conv = rss = None
    for _ in range(args.steps):
        a, b, n_iter, conv, rss = step()
        dev_s.append(a)
        e2e_s.append(b)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = ctx.launches - launches0
    st1 = eng.stats()

    t_dev, t_e2e = float(np.sum(dev_s)), float(np.sum(e2e_s))
    if world > 1:
        tt = torch.tensor([t_dev, t_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev, t_e2e = float(tt[0]), float(tt[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = n * args.steps / t_dev      # whole job: one n-cell fit per step, however many GPUs share it
    e2e = n * args.steps / t_e2e
    # roofline of the dominant kernel: gradient evaluation of _updateB (k_row_max + k_updateB_eval x2 per FW step)
    steps_timed = st1["updB_steps_timed"] - st0["updB_steps_timed"]
    eval_s = (st1["updB_eval_ns_total"] - st0["updB_eval_ns_total"]) / 1e9
    per_launch_s = eval_s / max(steps_timed, 1)
    alg_bytes = 12.0 * st1["t2_nnz_last_update_B"] + 12.0 * st1["kb_nnz_last"] + 16.0 * n / world  # rank 0's shard
    peak, peak_src = _peaks()
    achieved = alg_bytes / per_launch_s / 1e9 if per_launch_s > 0 else 0.0
    roofline = {"bound": "hbm",
                "kernel": "gradient of _updateB on the support of t2: k_row_max + k_updateB_eval (2 rounds) + "
                          "k_updateB_hub (+reduce), one launch group per Frank-Wolfe step",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": _traffic(n, world),
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                "avg_launch_ms": per_launch_s * 1e3, "launches_timed": steps_timed,
                "share_of_step": eval_s / t_dev if t_dev > 0 else None,
                "bytes_definition": "12 B x nnz(t2) candidate (cell, t2) pairs + 12 B x nnz(K B) row entries + 16 B x n "
                                    "row maxima, each moved once per Frank-Wolfe step (rank 0's shard when N > 1)",
                "traffic_note": "ncu dram bytes of the two largest kernels of the group are in profiles/*.summary.txt"}
    cpu_n = 2500
    cpu_dt, cpu_iters, cpu_k = cpu_fit_sample(cpu_n)
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True,
           "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic",
           "config": {"workload": workload_name(n, args.pcs), "cells_per_gpu": n // world if world > 1 else n, "outer_iterations": n_iter,
                      "converged": bool(conv), "rss_first_last": [rss[0], rss[-1]],
                      "fit_seconds": t_dev / args.steps, "construct_kernel_matrix_s": t_build,
                      "l2_policy": "inputs (>= 1 GB of lists per step) exceed the 126 MB L2; no explicit flush",
                      "sharding": ("single GPU" if world == 1 else
                                   f"cells row-sharded over {world} GPUs; per outer iteration one all-gather of the "
                                   "compressed A and 50 all-gathers of k x 4 fp64 argmin partials (NCCL)")},
           "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
           "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "knn": knn_info,
           "cpu_baseline": {"value": cpu_n / cpu_dt, "unit": UNIT, "cores": 1, "kind": "port",
                            "sample": f"full fit() ({cpu_iters} outer iterations) of the oracle restatement of "
                                      f"cpu.SEACellsCPU on a {cpu_n}-cell x 50-PC sample (k={cpu_k}); scipy sparse, "
                                      f"single thread; {os.cpu_count()} host cores available"}}
    _emit(out)
    if world > 1:
        dist.destroy_process_group()


def _emit(obj):
    """The one JSON line goes to the real stdout; everything else (NCCL banners, library prints) was sent to stderr."""
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)  # NCCL prints its version banner on fd 1; keep the contract's stdout to exactly one JSON line
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", type=int, default=1_000_000)
    ap.add_argument("--pcs", type=int, default=50, help="embedding dimensions (30 for the scATAC-like config 3)")
    ap.add_argument("--kind", default="rna", choices=["rna", "atac"], help="noise spectrum of the synthetic embedding")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
