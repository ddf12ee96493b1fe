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
        out.update({"bound": "tensor", "executed_f16_tflops": ex, "peak": peak, "unit": "TFLOP/s",
                    "frac_algorithmic": alg / peak, "frac_executed": ex / peak, "frac": alg / peak,
                    "frac_note": "frac = algorithmic 2 n_q n d flops / peak (SURVEY.md 8d); frac_executed counts the 3 "
                                 "split products over the padded K the tensor pipe actually ran",
                    "peak_source": src + " (fp16 and bf16 MMAs run at the same rate)"})
    return out


def _traffic(n, world):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (1M cells, 1 GPU only)."""
    p = os.path.join(ROOT, "profiles", "r2_hub_kernel_traffic.json")
    if n == 1_000_000 and world == 1 and os.path.exists(p):
        try:
            return float(json.load(open(p))["dram_bytes_per_launch"])
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


def _cpu_replica(args):
    n, seed, reps = args
    os.environ["OMP_NUM_THREADS"] = os.environ["OPENBLAS_NUM_THREADS"] = os.environ["MKL_NUM_THREADS"] = "1"
    from oracle import seacells_oracle as O
    from seacells_b200.synth import synth_archetypes, synth_assignments, synth_embedding
    k = max(2, n // 75)
    X = synth_embedding(n, 50, seed).astype(np.float64)
    M, *_ = O.build_kernel(X, 15)
    arch, A0 = synth_archetypes(n, k), synth_assignments(n, k)
    ts, it = [], 0
    for _ in range(reps):
        t = time.perf_counter()
        r = O.fit(M, arch, A0, max_iter=100, min_iter=10, convergence_epsilon=1e-3, max_FW_iter=50)
        ts.append(time.perf_counter() - t)
        it = r["n_iter"]
    return ts, it, k


def cpu_fit_replicas(n, procs, reps):
    """`procs` independent fits of the same n-cell sample side by side (the reference's loop is single-threaded scipy, so
    replicas are how it uses every host core); returns per-rep wall seconds (max over the replicas), iterations, k."""
    import multiprocessing as mp
    with mp.get_context("spawn").Pool(procs) as pool:
        res = pool.map_async(_cpu_replica, [(n, 0, reps)] * procs).get(timeout=1700)  # never hang the arm
    per_rep = [max(r[0][i] for r in res) for i in range(reps)]
    return per_rep, res[0][1], res[0][2]


C1_CELLS, C1_K = 10_000, 133  # BASELINE.json configs[0]: the reference's own CPU-runnable case


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 32))
    total = args.steps + args.warmup
    n = 4000 if total <= 4 else (2500 if total <= 8 else 1600)
    per_rep, iters, k = cpu_fit_replicas(n, procs, total)
    sec = float(np.mean(per_rep[args.warmup:]))
    val = procs * n / sec
    sample = (f"{procs} independent replicas (one per host core, {cores} available) of a full fit() (initialize + {iters} "
              f"outer iterations, T=50) of the oracle restatement of cpu.SEACellsCPU on a {n}-cell x 50-PC sample (k={k}) "
              f"of the workload; scipy sparse, each replica single-threaded as the reference's loop is; value = "
              f"replicas x cells / wall seconds")
    # same-config pair (BASELINE config 1): ONE full 10k-cell fit, alone on the box (the latency a user of the reference
    # sees; its Frank-Wolfe loop is single-threaded scipy), kernel construction untimed as in the GPU arm
    same = None
    if not args.no_same_config:
        c1_rep, c1_iters, c1_k = cpu_fit_replicas(C1_CELLS, 1, 1)
        same = {"config": f"config 1: synthetic {C1_CELLS} cells x 50 PCs, n_SEACells={c1_k}, full fit() "
                          f"(initialize + {c1_iters} outer iterations)", "cpu_fit_s": c1_rep[0], "cores": 1,
                "cpu_cells_per_s": C1_CELLS / c1_rep[0], "kind": "port",
                "gpu_side": "`same_config.gpu_fit_s` of the default (`--impl ours`) line of the same run"}
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": workload_name(args.cells), "sample_cells": n, "same_config": False,
                      "note": "the reference cannot run the 1M-cell configuration (SURVEY.md 6); see `same_config` for "
                              "a pair measured on one configuration"},
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample,
                            "host_cores_available": cores},
           "same_config": same,
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

    # N > 1: ONE n-cell fit, cells row-sharded over the ranks (strong scaling).  The context joins the job (the library's
    # own NCCL communicator, bootstrapped through torch.distributed); from then on sc_kernel_build / sc_fit_run exchange
    # what the other ranks need inside the library.  Every rank holds the same embedding.
    from seacells_b200.parallel import init_nccl
    init_nccl(ctx)
    n = args.cells
    k = n // 75
    X = synth_embedding(n, args.pcs, seed=0, kind=args.kind)
    build_kernel(X[:4096], 15, "union", ctx).close()  # untimed: CUDA context, module load, first allocations
    ctx.sync()
    t0 = time.perf_counter()
    km = build_kernel(X, 15, "union", ctx)
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
    A0_raw = synth_assignments(n, k)
    A0 = normalize(A0_raw, axis=0, norm="l1").tocsc()
    A0.sort_indices()
    h_colptr, h_rows, h_vals = pin(A0.indptr.astype(np.int64)), pin(A0.indices.astype(np.int32)), pin(A0.data)
    h_labels = torch.empty(n, dtype=torch.int32).pin_memory()
    h2d = sum(t.numel() * t.element_size() for t in (h_indptr, h_indices, h_data, h_colptr, h_rows, h_vals)) + 8 * k
    d2h = 4 * n + 8 * 102

    eng = FitEngine(n, k, 50, ctx)
    lib = ctx.lib
    ptr = _lib.ptr

    def step():
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record(lib_stream)
        ctx.check(lib.sc_fit_set_kernel(eng.h, ptr(h_indptr), ptr(h_indices), ptr(h_data), int(M.nnz)), "set_kernel")
        eng.set_archetypes(arch)
        ctx.check(lib.sc_fit_set_A_csc(eng.h, ptr(h_colptr), ptr(h_rows), ptr(h_vals)), "set_A")
        ev[1].record(lib_stream)
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

    import hashlib
    labels_sha = hashlib.sha256(h_labels.numpy().tobytes()).hexdigest()
    t_dev, t_e2e = float(np.sum(dev_s)), float(np.sum(e2e_s))
    if world > 1:
        tt = torch.tensor([t_dev, t_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev, t_e2e = float(tt[0]), float(tt[1])
    if rank != 0:
        if world > 1:
            dist.barrier()  # rank 0 runs the single-GPU extras below
            dist.destroy_process_group()
        return

    extras = {} if args.no_extras else _extras(args, ctx, lib_stream, X, M, arch, A0, A0_raw, world, eng, rss, labels_sha)
    value = n * args.steps / t_dev      # whole job: one n-cell fit per step, however many GPUs share it
    e2e = n * args.steps / t_e2e
    # roofline of the dominant kernel.  Largest single kernel of the fit by time: k_updateB_hub, the cell-major gradient of
    # the hub archetypes, one launch per Frank-Wolfe step (CUDA events on the library's stream around that launch alone).
    # Algorithmic bytes per launch: 24 B x cells x hub archetypes (running product read + written, t2 read)
    # + 12 B x delta entries of the step.  The launch group it belongs to (row maxima, both rounds of k_updateB_eval, hub
    # kernel + reduction) is reported next to it with the byte count of round 1.
    peak, peak_src = _peaks()
    hub_steps = st1["hub_steps_timed"] - st0["hub_steps_timed"]
    hub_s = (st1["hub_ns_total"] - st0["hub_ns_total"]) / 1e9
    hub_launch_s = hub_s / max(hub_steps, 1)
    hub_bytes = float(st1["hub_bytes_per_launch"])
    hub_gbs = hub_bytes / hub_launch_s / 1e9 if hub_launch_s > 0 else 0.0
    steps_timed = st1["updB_steps_timed"] - st0["updB_steps_timed"]
    eval_s = (st1["updB_eval_ns_total"] - st0["updB_eval_ns_total"]) / 1e9
    per_launch_s = eval_s / max(steps_timed, 1)
    alg_bytes = 12.0 * st1["t2_nnz_last_update_B"] + 12.0 * st1["kb_nnz_last"] + 16.0 * n / world  # rank 0's shard
    grp_gbs = alg_bytes / per_launch_s / 1e9 if per_launch_s > 0 else 0.0
    roofline = {"bound": "hbm",
                "kernel": "k_updateB_hub (gradient of _updateB for the hub archetypes, cell-major; incremental launches)",
                "achieved": hub_gbs, "peak": peak, "unit": "GB/s", "frac": hub_gbs / peak, "traffic": _traffic(n, world),
                "peak_source": peak_src, "algorithmic_bytes_per_launch": hub_bytes,
                "avg_launch_ms": hub_launch_s * 1e3, "launches_timed": hub_steps,
                "share_of_step": hub_s / t_dev if t_dev > 0 else None,
                "bytes_definition": "24 B x cells x hub archetypes (HH read + write, t2 read; 8 B each) + 12 B x delta "
                                    "entries of the step (rank 0's shard when N > 1)",
                "launch_group": {"what": "gradient of _updateB, all archetypes: k_row_max (steps 0-1) + k_updateB_eval x 2 + "
                                         "k_updateB_hub + reduction, one group per Frank-Wolfe step",
                                 "avg_group_ms": per_launch_s * 1e3, "groups_timed": steps_timed,
                                 "share_of_step": eval_s / t_dev if t_dev > 0 else None,
                                 "round1_definition_bytes": alg_bytes, "round1_definition_GBps": grp_gbs,
                                 "round1_definition_frac": grp_gbs / peak,
                                 "note": "round-1 definition = 12 B x nnz(t2) + 12 B x nnz(K B) + 16 B x n per group; since "
                                         "the sorted-list bound discards most 1024-candidate chunks unread "
                                         f"({st1['chunks_pruned_unread_last_update_B']} of "
                                         f"{st1['chunk_launches_last_update_B']} in the last _updateB) those bytes are "
                                         "no longer all moved - kept for comparison with round 1 only"}}
    cpu_n = 4000
    cpu_dt, cpu_iters, cpu_k = cpu_fit_sample(cpu_n)
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True,
           "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic",
           "config": {"workload": workload_name(n, args.pcs), "cells_per_gpu": n // world if world > 1 else n, "outer_iterations": n_iter,
                      "converged": bool(conv), "rss_first_last": [rss[0], rss[-1]],
                      "fit_seconds": t_dev / args.steps, "construct_kernel_matrix_s": t_build,
                      "labels_sha256": labels_sha,
                      "l2_policy": "inputs (>= 1 GB of lists per step) exceed the 126 MB L2; no explicit flush",
                      "sharding": ("single GPU" if world == 1 else
                                   f"cells row-sharded over {world} GPUs; in-library NCCL all-gathers: per outer "
                                   "iteration the compressed A (x1), per-cell RSS terms (x1), candidate counts (x1) and "
                                   "50 x k x 4 fp64 argmin partials")},
           "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
           "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "knn": knn_info,
           "cpu_baseline": {"value": cpu_n / cpu_dt, "unit": UNIT, "cores": 1, "kind": "port",
                            "sample": f"full fit() ({cpu_iters} outer iterations) of the oracle restatement of "
                                      f"cpu.SEACellsCPU on a {cpu_n}-cell x 50-PC sample (k={cpu_k}); scipy sparse, "
                                      f"single thread; {os.cpu_count()} host cores available"}}
    out.update(extras)
    _emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()



def _extras(args, ctx, lib_stream, X, M, arch, A0, A0_raw, world, eng, rss, labels_sha):
    """Single-GPU measurements that ride along on rank 0 after the timed region (none of them enters `value`):
    parity of the sharded run against an unsharded one, the CSR SpMM roofline probe, the same-config (config 1) pair,
    and the fit through the reference-facing class (`e2e_api`)."""
    import hashlib

    import torch
    from seacells_b200 import FitEngine, SEACells, _lib, build_kernel
    from seacells_b200.synth import MiniAnnData, synth_archetypes, synth_assignments, synth_embedding
    lib, ptr = ctx.lib, _lib.ptr
    n, k = X.shape[0], len(arch)
    out = {}

    def timed(fn, reps=1):
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(lib_stream)
            fn()
            e1.record(lib_stream)
            e1.synchronize()
            ts.append(e0.elapsed_time(e1) / 1e3)
        return ts

    eng.close()  # the timed engine's 85 GB of device state is no longer needed
    torch.cuda.empty_cache()
    # ---- N > 1: the sharded fit must reproduce the single-GPU fit (labels and RSS trajectory to the last bit)
    if world > 1:
        solo = _lib.Context(ctx.device)   # a context that is not part of the job: plain single-GPU fit
        e1 = FitEngine(n, k, 50, solo)
        e1.set_kernel(M)
        e1.set_archetypes(arch)
        e1.set_A(A0)
        rss1, *_ = e1.run(max_iter=100, min_iter=10, eps=1e-3)
        sha1 = hashlib.sha256(e1.hard_assignments().astype(np.int32).tobytes()).hexdigest()
        e1.close()
        solo.close()
        out["parity_vs_1gpu"] = {"labels_identical": sha1 == labels_sha, "rss_identical": list(rss1) == list(rss),
                                 "labels_sha256_1gpu": sha1, "rss_last_1gpu": rss1[-1], "rss_last": rss[-1]}
        if sha1 != labels_sha:
            raise SystemExit(f"bench.py: {world}-GPU hard assignments differ from the 1-GPU fit ({labels_sha} vs {sha1})")
        return out  # the remaining probes are single-GPU measurements: they ride on the N = 1 line only

    # ---- CSR SpMM (BASELINE metric "SpMM HBM GB/s vs peak"; cpu_dense.py:353,389 / gpu.py:429,477) at config-2 shape:
    #      Y[r x c] = M[r x r] X[r x c], r = 100k cells, c = 1333, operands resident in HBM
    try:
        r_, c_ = min(100_000, n), min(100_000, n) // 75
        km2 = build_kernel(X[:r_], 15, "union", ctx)
        M2, *_ = km2.export(want_knn=False)
        km2.close()
        dev = torch.device("cuda", ctx.device)
        ip = torch.from_numpy(M2.indptr.astype(np.int32)).to(dev)
        ix = torch.from_numpy(M2.indices.astype(np.int32)).to(dev)
        dv = torch.from_numpy(M2.data).to(dev)
        Xd = torch.rand((r_, c_), dtype=torch.float64, device=dev)
        Yd = torch.empty((r_, c_), dtype=torch.float64, device=dev)
        torch.cuda.synchronize()
        call = lambda: ctx.check(lib.sc_spmm_csr_f64(ctx.h, r_, r_, c_, ptr(ip), ptr(ix), ptr(dv), ptr(Xd), ptr(Yd)), "spmm")
        timed(call, 3)
        ts = timed(call, 10)
        alg = 12.0 * M2.nnz + 4.0 * (r_ + 1) + 16.0 * c_ * r_
        peak, src = _peaks()
        ms = float(np.median(ts)) * 1e3
        out["spmm"] = {"kernel": "k_spmm_csr_panel (sc_spmm_csr_f64)", "shape": {"r": r_, "m": r_, "c": c_, "nnz": int(M2.nnz)},
                       "ms": ms, "algorithmic_bytes": alg, "achieved": alg / ms / 1e6, "unit": "GB/s", "peak": peak,
                       "frac": alg / ms / 1e6 / peak, "peak_source": src, "traffic": _spmm_traffic(),
                       "bytes_definition": "12 nnz + 4 (r+1) + 8 c m + 8 c r (SURVEY.md 8d): every operand once; the "
                                           "nnz/r = 24 re-reads of each X row through L2 are not counted"}
        del Xd, Yd, ip, ix, dv
        torch.cuda.empty_cache()
    except Exception as e:  # the probe must never take the headline down
        out["spmm"] = {"error": repr(e)}

    # ---- same-config pair, GPU side (BASELINE config 1: 10k cells x 50 PCs, k = 133, full fit); the CPU side is the
    #      `same_config` object of `bench.py --impl reference`
    if not args.no_same_config:
        X1 = synth_embedding(C1_CELLS, 50, 0)
        km1 = build_kernel(X1, 15, "union", ctx)
        a1, A01 = synth_archetypes(C1_CELLS, C1_K), synth_assignments(C1_CELLS, C1_K)
        from sklearn.preprocessing import normalize
        A01n = normalize(A01, axis=0, norm="l1")
        e1 = FitEngine(C1_CELLS, C1_K, 50, ctx)
        e1.set_kernel(km1)
        res = {}

        def fit1():
            e1.set_archetypes(a1)
            e1.set_A(A01n)
            res["r"] = e1.run(max_iter=100, min_iter=10, eps=1e-3)

        timed(fit1, 1)
        ts = timed(fit1, 3)
        ad1 = MiniAnnData(X1)
        m1 = SEACells(ad1, "X_pca", C1_K, verbose=False, use_sparse=True)
        m1.construct_kernel_matrix()
        t0 = time.perf_counter()
        try:
            m1.fit(max_iter=100, min_iter=10, initial_archetypes=a1, initial_assignments=A01)
        except RuntimeWarning:
            pass
        m1.get_hard_assignments()
        t_api = time.perf_counter() - t0
        out["same_config"] = {"config": f"config 1: synthetic {C1_CELLS} cells x 50 PCs, n_SEACells={C1_K}, full fit() "
                                        f"(initialize + {res['r'][1]} outer iterations)",
                              "gpu_fit_s": float(np.median(ts)), "gpu_cells_per_s": C1_CELLS / float(np.median(ts)),
                              "gpu_api_fit_s": t_api, "rss_last": res["r"][0][-1],
                              "cpu_side": "`same_config.cpu_fit_s` of the `--impl reference` line of the same run"}
        e1.close()
        km1.close()

    # ---- the headline configuration through the reference-facing class: SEACells(...).fit() + get_hard_assignments()
    if world == 1:
        try:
            ad = MiniAnnData(X)
            model = SEACells(ad, "X_pca", k, verbose=False, use_sparse=True)
            model.add_precomputed_kernel_matrix(M)  # fit() excludes construct_kernel_matrix, as in the reference
            t0 = time.perf_counter()
            try:
                model.fit(max_iter=100, min_iter=10, initial_archetypes=arch, initial_assignments=A0_raw)
            except RuntimeWarning:
                pass
            lab = model.get_hard_assignments()
            t_api = time.perf_counter() - t0
            lab_i = np.array([int(x[8:]) for x in lab["SEACell"].values], dtype=np.int32)
            out["e2e_api"] = {"value": n / t_api, "unit": UNIT, "seconds": t_api,
                              "call": "seacells_b200.SEACells(ad, 'X_pca', k, use_sparse=True).fit(initial_archetypes=, "
                                      "initial_assignments=) + get_hard_assignments(): host scipy/numpy inputs, pandas "
                                      "labels out (includes A0 normalisation, CSR->slot conversion, 1M label strings)",
                              "labels_identical_to_engine_run": hashlib.sha256(lab_i.tobytes()).hexdigest() == labels_sha}
        except Exception as e:
            out["e2e_api"] = {"error": repr(e)}
    return out


def _spmm_traffic():
    p = os.path.join(ROOT, "profiles", "r2_spmm_traffic.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["dram_bytes_per_launch"])
        except Exception:
            return None
    return None


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
    ap.add_argument("--no-same-config", action="store_true", help="skip the config-1 (10k cells) same-config pair")
    ap.add_argument("--no-extras", action="store_true", help="skip the spmm / e2e_api / same_config / parity extras")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
