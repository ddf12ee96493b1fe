"""The row-sharded (multi-GPU) code path.  The exchange lives in the library (comm.cuh): here it runs with ranks as
threads of one process on a single GPU (LocalComm) and, when >= 2 GPUs exist, with one process per GPU over NCCL.
The sharded fit must equal the reference's golden vectors and the single-GPU fit bit for bit."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import seacells_oracle as O
from tests.helpers import csr_from, load_golden, same_sparse

pytestmark = pytest.mark.gpu


def _local_fit(M, k, T, arch, A0n, world, n_iter, X=None, devices=None):
    """`world` ranks as threads; every rank holds the same inputs and makes the same calls.  Returns per-rank results."""
    from seacells_b200 import FitEngine, build_kernel
    from seacells_b200.parallel import LocalGroup
    grp = LocalGroup(world, devices)
    n = M.shape[0]

    def work(rank, ctx):
        out = {}
        if X is not None:   # sharded kernel construction: every rank searches its own query slab
            km = build_kernel(X, 15, "union", ctx)
            out["M"], out["sigma"], out["idx"], out["d2"] = km.export()
            km.close()
        eng = FitEngine(n, k, T, ctx)
        out["shard"] = eng.shard()
        eng.set_kernel(M)
        eng.set_archetypes(arch)
        eng.set_A(A0n)
        rss, it, conv, thr = eng.run(max_iter=n_iter, min_iter=n_iter)
        out.update(rss=rss, A=eng.A(), B=eng.B(), labels=eng.hard_assignments(), KB=eng.KB(), hubs=eng.stats()[
            "hub_archetypes_last_update_B"])
        eng.close()
        return out

    try:
        return grp.run(work)
    finally:
        grp.close()


@pytest.mark.parametrize("world", [2, 3])
def test_local_ranks_reproduce_the_reference(world):
    g = load_golden("rna_n600_k8_sparse")
    M = csr_from(g, "M")
    res = _local_fit(M, int(g["k"]), int(g["T"]), g["archetypes"], O.l1_normalise_columns(csr_from(g, "A0")), world,
                     int(g["n_iter"]), X=g["X32"])
    from seacells_b200.parallel import slab_range
    for r, out in enumerate(res):
        assert out["shard"] == slab_range(r, world, M.shape[0])
        np.testing.assert_allclose(out["rss"], g["RSS_iters"], rtol=1e-9)
        assert out["rss"] == res[0]["rss"], "the RSS must agree bit for bit across ranks"
        assert same_sparse(out["A"], csr_from(g, "A_final"))
        assert same_sparse(out["B"], csr_from(g, "B_final"))
        assert np.array_equal(out["labels"], g["labels"])
        # sharded kernel construction: same kNN graph and kernel matrix on every rank
        D = O.knn_distance_graph(out["idx"], out["d2"], M.shape[0])
        assert same_sparse(D, csr_from(g, "D"))
        assert np.array_equal(out["M"].indices, M.indices)
        np.testing.assert_allclose(out["M"].data, M.data, rtol=1e-13, atol=0)
        assert same_sparse(out["KB"], res[0]["KB"])  # the Z_ accessor evaluates every row on every rank


def test_local_ranks_with_hub_archetypes():
    """5000 cells: hub archetypes exist, so the cell-major hub block and the job-wide hub selection are exercised; the
    sharded result must equal the single-engine result bit for bit."""
    from seacells_b200 import FitEngine
    from seacells_b200.synth import synth_archetypes, synth_assignments, synth_embedding
    n, k = 5000, 66
    X = synth_embedding(n, 50, 0)
    M, *_ = O.build_kernel(X.astype(np.float64), 15)
    arch, A0n = synth_archetypes(n, k), O.l1_normalise_columns(synth_assignments(n, k))
    res = _local_fit(M, k, 50, arch, A0n, 3, 2)
    one = FitEngine(n, k)
    one.set_kernel(M)
    one.set_archetypes(arch)
    one.set_A(A0n)
    rss1, *_ = one.run(max_iter=2, min_iter=2)
    assert one.stats()["hub_archetypes_last_update_B"] > 0
    for out in res:
        assert out["hubs"] == one.stats()["hub_archetypes_last_update_B"]
        assert out["rss"] == rss1
        assert same_sparse(out["A"], one.A()) and same_sparse(out["B"], one.B())


def test_local_ranks_on_two_devices():
    """The thread transport also drives several GPUs from one process (peer copies)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    g = load_golden("atac_n900_k12_sparse")
    M = csr_from(g, "M")
    res = _local_fit(M, int(g["k"]), int(g["T"]), g["archetypes"], O.l1_normalise_columns(csr_from(g, "A0")), 2,
                     int(g["n_iter"]), devices=[0, 1])
    for out in res:
        assert same_sparse(out["A"], csr_from(g, "A_final")) and same_sparse(out["B"], csr_from(g, "B_final"))


def _nccl_worker(rank, world, port, name, ret):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from seacells_b200 import FitEngine, _lib, build_kernel
    from seacells_b200.parallel import init_nccl
    g = load_golden(name)
    ctx = init_nccl(_lib.default_context(rank))
    assert (ctx.rank, ctx.world) == (rank, world)
    km = build_kernel(g["X32"], 15, "union", ctx)      # query slabs sharded, all-gathered on the device
    Mb, sigma, idx, d2 = km.export()
    io, do = O.knn_exact(g["X32"], 15)
    ok = np.array_equal(idx, io) and np.array_equal(d2, do)
    M = csr_from(g, "M")
    ok = ok and np.array_equal(Mb.indices, M.indices)
    eng = FitEngine(M.shape[0], int(g["k"]), int(g["T"]), ctx)
    eng.set_kernel(M)
    eng.set_archetypes(g["archetypes"])
    eng.set_A(O.l1_normalise_columns(csr_from(g, "A0")))
    rss, n_iter, conv, _ = eng.run(max_iter=int(g["max_iter"]), min_iter=int(g["min_iter"]), eps=float(g["eps"]))
    ok = ok and n_iter == int(g["n_iter"]) and np.allclose(rss, g["RSS_iters"], rtol=1e-9)
    ok = ok and same_sparse(eng.A(), csr_from(g, "A_final")) and same_sparse(eng.B(), csr_from(g, "B_final"))
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        ret.put(int(flag.item()))
    dist.destroy_process_group()


def test_nccl_sharded_fit_two_gpus():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    world = 2
    mpc = mp.get_context("spawn")
    ret = mpc.Queue()
    port = 29700 + os.getpid() % 1000
    procs = [mpc.Process(target=_nccl_worker, args=(r, world, port, "atac_n900_k12_sparse", ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    assert ret.get(timeout=5) == 1
