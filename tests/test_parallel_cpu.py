"""Host-side logic of the row-sharded fit, exercised with world_size-2 gloo process groups on CPU."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from seacells_b200.parallel import merge_partials, slab_range, slab_size


def test_slabs_cover_every_cell_once():
    for n in (1, 7, 64, 1000, 1_000_003):
        for world in (1, 2, 3, 4, 8):
            c = slab_size(n, world)
            assert c * world >= n and c * world <= n + world
            seen = []
            for r in range(world):
                b, e = slab_range(r, world, n)
                assert 0 <= b <= e <= n and e - b <= c
                seen.extend(range(b, e)) if n < 5000 else None
                if r:
                    assert b == min(n, r * c)
            if n < 5000:
                assert seen == list(range(n))


def _dense_rule(Gcol, is_cand):
    """np.argmin semantics with the candidate restriction."""
    cand = np.where(is_cand)[0]
    if cand.size and Gcol[cand].min() < 0:
        return int(np.argmin(np.where(is_cand, Gcol, np.inf)))
    return int(np.argmin(Gcol))


def _local_partial(Gcol, is_cand, b, e):
    INF, BIG = 1.0e300, 0x7fffffff
    idx = np.arange(b, e)
    c = idx[is_cand[b:e]]
    out = [INF, BIG, INF, BIG]
    if c.size:
        j = c[np.argmin(Gcol[c])]
        out[0], out[1] = Gcol[j], j
    if not (out[0] < 0):
        nc = idx[~is_cand[b:e]]
        if nc.size:
            z = nc[Gcol[nc] == 0.0]
            if z.size:
                out[2], out[3] = 0.0, z[0]
            else:
                j = nc[np.argmin(Gcol[nc])]
                out[2], out[3] = Gcol[j], j
    return out


def _worker(rank, world, port, n, k, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(123)            # identical data on every rank
    G = rng.normal(size=(n, k))
    is_cand = rng.random((n, k)) < 0.3
    G[~is_cand] = np.abs(G[~is_cand])            # outside the support of t2 the gradient is >= 0
    G[~is_cand & (rng.random((n, k)) < 0.2)] = 0.0
    G[:, 0] = np.abs(G[:, 0])                    # a column with no negative entry: the zero rule decides
    G[:, 1] = np.abs(G[:, 1]) + 0.5              # no negative entry and no zero: first minimum of the whole column
    is_cand[:, 2] = True                         # a "hub" column: every cell is a candidate
    b, e = slab_range(rank, world, n)
    part = torch.tensor([_local_partial(G[:, a], is_cand[:, a], b, e) for a in range(k)], dtype=torch.float64)
    gathered = [torch.zeros_like(part) for _ in range(world)]
    dist.all_gather(gathered, part)
    amin = merge_partials(torch.stack(gathered).numpy())
    expect = np.array([_dense_rule(G[:, a], is_cand[:, a]) for a in range(k)])
    ok = bool(np.array_equal(amin, expect))
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        ret.put(int(flag.item()))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_partial_merge_matches_dense_argmin_gloo(world):
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29500 + os.getpid() % 2000 + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, 997, 12, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ret.get(timeout=5) == 1
