"""Row-sharded fit across the GPUs of one box: one process per GPU, torch.distributed (NCCL) for the exchange.

Cells are sharded by contiguous, equally sized slabs.  Per outer iteration the ranks exchange
  * the compressed A (all-gather of the slot arrays, once after `_updateA`), and
  * per Frank-Wolfe step of `_updateB`, one k x 4 fp64 partial per rank (all-gather) holding each archetype's local
    candidate minimum and local first-zero / minimum outside the support of t2 - the (value, index) argmin pairs of the
    north_star; the merge is a lexicographic minimum, so every rank takes the identical simplex step.
M, B, Y = M B, the Gram A A^T and the RSS are replicated (small or cheap); K B rows, candidate lists and `_updateA`
are computed for the rank's slab only.  The library itself never communicates (include/seacells_b200.h, multi-GPU
section); everything here is host-side plumbing and is what the gloo tests on CPU exercise.
"""
from __future__ import annotations

import ctypes as C

import numpy as np


def slab_size(n: int, world: int) -> int:
    return (n + world - 1) // world


def slab_range(rank: int, world: int, n: int):
    """Cells [begin, end) owned by `rank`; slabs are equal (the last one may be short or empty)."""
    c = slab_size(n, world)
    b = min(n, rank * c)
    return b, min(n, b + c)


def merge_partials(gathered: np.ndarray) -> np.ndarray:
    """Reference (numpy) statement of the merge done by k_updateB_step.  gathered: [world, k, 4] fp64 holding
    {candidate min G, its cell, non-candidate result G, its cell}.  Returns the argmin cell per archetype."""
    g = np.asarray(gathered, dtype=np.float64)
    world, k, _ = g.shape
    out = np.zeros(k, dtype=np.int64)
    for a in range(k):
        cand = min(((g[r, a, 0], int(g[r, a, 1])) for r in range(world)))
        nonc = min(((g[r, a, 2], int(g[r, a, 3])) for r in range(world)))
        if cand[0] < 0.0:
            out[a] = cand[1]
        else:
            best = min(cand, nonc)
            out[a] = best[1] if best[1] != 0x7fffffff else 0
    return out


class _DevArray:
    """Expose a raw device pointer to torch through __cuda_array_interface__ (zero-copy view of library memory)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def device_view(ptr: int, shape, dtype):
    import torch
    typestr = {torch.int32: "<i4", torch.float64: "<f8"}[dtype]
    return torch.as_tensor(_DevArray(ptr, shape, typestr), device="cuda")


class ShardedFit:
    """Drives one FitEngine per rank.  All ranks must hold the same M, archetypes and A0 and call the same methods."""

    def __init__(self, engine, group=None):
        import torch
        import torch.distributed as dist
        self.eng, self.dist, self.group = engine, dist, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n, self.k, self.S, self.T = engine.n, engine.k, engine.S, engine.T
        self.chunk = slab_size(self.n, self.world)
        if self.chunk * self.world > self.n + 64:
            raise ValueError("too many ranks for this number of cells")
        self.begin, self.end = slab_range(self.rank, self.world, self.n)
        ctx = engine.ctx
        ctx.check(ctx.lib.sc_fit_set_shard(engine.h, self.begin, self.end), "sc_fit_set_shard")
        dev = torch.device("cuda", ctx.device)
        self.stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
        self.partial = torch.zeros(self.k * 4, dtype=torch.float64, device=dev)
        self.gathered = torch.zeros(self.world * self.k * 4, dtype=torch.float64, device=dev)

    # -- exchanges ---------------------------------------------------------------------------------------------
    def _allgather_A(self):
        if self.world == 1:
            return
        import torch
        lib, ctx = self.eng.ctx.lib, self.eng.ctx
        pi, pv, pc, rows = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_longlong()
        ctx.check(lib.sc_fit_A_slots_device(self.eng.h, C.byref(pi), C.byref(pv), C.byref(pc), C.byref(rows)),
                  "sc_fit_A_slots_device")
        R, S, c, r = int(rows.value), self.S, self.chunk, self.rank
        with torch.cuda.stream(self.stream):
            for ptr, dtype, width in ((pi.value, torch.int32, S), (pv.value, torch.float64, S), (pc.value, torch.int32, 1)):
                full = device_view(ptr, (R * width,), dtype)
                out = full[: c * self.world * width]
                mine = out[r * c * width:(r + 1) * c * width].clone()
                self.dist.all_gather_into_tensor(out, mine, group=self.group)
        ctx.sync()

    # -- the hot path --------------------------------------------------------------------------------------------
    def update_A(self, l2_penalty=0.0):
        self.eng.update_A(l2_penalty)
        self._allgather_A()

    def update_B(self):
        import torch
        lib, ctx, h = self.eng.ctx.lib, self.eng.ctx, self.eng.h
        ctx.check(lib.sc_fit_update_B_begin(h), "sc_fit_update_B_begin")
        for t in range(self.T):
            ctx.check(lib.sc_fit_update_B_eval(h, t, C.c_void_p(self.partial.data_ptr())), "sc_fit_update_B_eval")
            if self.world > 1:
                with torch.cuda.stream(self.stream):
                    self.dist.all_gather_into_tensor(self.gathered, self.partial, group=self.group)
                src = self.gathered
            else:
                src = self.partial
            ctx.check(lib.sc_fit_update_B_step(h, t, C.c_void_p(src.data_ptr()), self.world), "sc_fit_update_B_step")
        ctx.check(lib.sc_fit_update_B_end(h), "sc_fit_update_B_end")

    def rss(self) -> float:
        return self.eng.rss()  # replicated: every rank evaluates the full trace identity

    def run(self, max_iter=100, min_iter=10, eps=1e-3, l2_penalty=0.0, threshold=None):
        """initialize + _fit loop (cpu.py:204-285, 595-651); returns (rss_iters, n_iter, converged, threshold)."""
        if max_iter < min_iter:
            raise ValueError("The maximum number of iterations specified is lower than the minimum number of "
                             "iterations specified.")
        self.update_A(l2_penalty)
        rss = [self.rss()]
        thr = eps * rss[0] if threshold is None else threshold
        converged, it = False, 0
        while (not converged and it < max_iter) or it < min_iter:
            it += 1
            self.update_A(l2_penalty)
            self.update_B()
            rss.append(self.rss())
            if abs(rss[-2] - rss[-1]) < thr:
                converged = True
        return rss, it, converged, thr


def sharded_knn(X, n_neighbors, ctx, group=None):
    """Exact kNN with the query rows sharded over the ranks, all-gathered to every rank (numpy in, numpy out)."""
    import torch
    import torch.distributed as dist
    from .core import knn
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = X.shape[0]
    kk = n_neighbors - 1
    c = slab_size(n, world)
    b, e = slab_range(rank, world, n)
    idx, d2 = knn(X, n_neighbors, ctx, b, e)
    if world == 1:
        return idx, d2
    dev = torch.device("cuda", ctx.device) if torch.cuda.is_available() else torch.device("cpu")
    li = torch.zeros((c, kk), dtype=torch.int32, device=dev)
    ld = torch.zeros((c, kk), dtype=torch.float64, device=dev)
    li[: e - b] = torch.from_numpy(idx).to(dev)
    ld[: e - b] = torch.from_numpy(d2).to(dev)
    gi = torch.empty((world * c, kk), dtype=torch.int32, device=dev)
    gd = torch.empty((world * c, kk), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(gi, li, group=group)
    dist.all_gather_into_tensor(gd, ld, group=group)
    return gi[:n].cpu().numpy(), gd[:n].cpu().numpy()
