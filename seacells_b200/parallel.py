"""Row-sharded jobs across the GPUs of one box.

The exchange itself lives in the library: a context that has joined a job (`init_nccl` / `attach_local`) owns a contiguous
slab of cells, and `sc_kernel_build`, `sc_fit_update_A`, `sc_fit_update_B`, `sc_fit_rss`, `sc_fit_run` all-gather what the
other ranks need on the library's stream (compressed A, per-cell RSS terms, per-archetype argmin partials, kNN rows).  This
module only bootstraps the communicator and restates the partial/merge protocol in numpy for the CPU tests:
  * `init_nccl(ctx)`     one process per GPU (torchrun): rank 0 creates the NCCL id, torch.distributed broadcasts it
  * `LocalGroup(world)`  ranks are threads of one process (tests on a single GPU; one process driving several GPUs)
Every rank must hold the same M, archetypes and A0 and make the same calls in the same order.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import sys

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


def _nccl_library_path():
    """libnccl.so.2 that ships with torch (nvidia-nccl wheel); the library also tries the plain SONAME itself."""
    for base in sys.path:
        hits = glob.glob(os.path.join(base, "nvidia", "nccl", "lib", "libnccl.so*"))
        if hits:
            return sorted(hits)[0]
    return None


def init_nccl(ctx, group=None):
    """Join the job spanned by the torch.distributed process group (one process per GPU).  Rank 0 creates the NCCL
    unique id, the existing process group (any backend) broadcasts its 128 bytes, every rank initialises its rank of
    the library's own communicator.  No-op for a single-process run."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return ctx
    if ctx.world > 1:
        return ctx
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    path = _nccl_library_path()
    ctx.check(ctx.lib.sc_nccl_load_library(path.encode() if path else None), "sc_nccl_load_library")
    uid = np.zeros(128, np.uint8)
    if rank == 0:
        ctx.check(ctx.lib.sc_nccl_unique_id(C.c_void_p(uid.ctypes.data)), "sc_nccl_unique_id")
    t = torch.from_numpy(uid)
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    dist.broadcast(t, src=0, group=group)
    uid = t.cpu().numpy().copy()
    ctx.check(ctx.lib.sc_ctx_init_nccl(ctx.h, rank, world, C.c_void_p(uid.ctypes.data)), "sc_ctx_init_nccl")
    return ctx


class LocalGroup:
    """`world` ranks as threads of this process: `contexts[r]` is rank r's context (its own stream; `devices[r]` or
    device 0).  `run(fn)` calls fn(rank, ctx) on one thread per rank and returns the results in rank order - the
    library calls block inside the exchange until every rank has arrived, so they must run concurrently."""

    def __init__(self, world: int, devices=None):
        from . import _lib
        self.world = world
        lib = _lib.load()
        h = C.c_void_p()
        if lib.sc_local_group_new(world, C.byref(h)) != 0:
            raise _lib.SeacellsB200Error("sc_local_group_new failed")
        self._lib, self.h = lib, h
        self.contexts = []
        for r in range(world):
            ctx = _lib.Context(devices[r] if devices else 0)
            ctx.check(lib.sc_ctx_attach_local(ctx.h, h, r), "sc_ctx_attach_local")
            self.contexts.append(ctx)

    def run(self, fn):
        import threading
        out, err = [None] * self.world, [None] * self.world

        def work(r):
            try:
                out[r] = fn(r, self.contexts[r])
            except BaseException as e:  # surfaced on the caller's thread below
                err[r] = e
                self._lib.sc_local_group_abort(self.h)  # peers waiting in an exchange fail instead of hanging

        threads = [threading.Thread(target=work, args=(r,)) for r in range(self.world)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        for e in err:
            if e is not None:
                raise e
        return out

    def close(self):
        for c in self.contexts:
            c.close()
        self.contexts = []
        if self.h is not None:
            self._lib.sc_local_group_free(self.h)
            self.h = None
