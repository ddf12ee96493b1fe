"""The row-sharded (multi-GPU) code path: emulated on one GPU in-process, and for real over NCCL when >= 2 GPUs exist."""
import ctypes as C
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import seacells_oracle as O
from tests.helpers import csr_from, load_golden, same_sparse

pytestmark = pytest.mark.gpu


def _emulated_fit(g, world):
    """`world` engines on one GPU, each owning a slab of cells; the two exchanges are done by hand with torch copies."""
    import torch
    from seacells_b200 import FitEngine
    from seacells_b200.parallel import device_view, slab_range, slab_size
    M = csr_from(g, "M")
    n, k, T = M.shape[0], int(g["k"]), int(g["T"])
    A0 = O.l1_normalise_columns(csr_from(g, "A0"))
    engs = []
    for r in range(world):
        e = FitEngine(n, k, T)
        e.set_kernel(M)
        e.set_archetypes(g["archetypes"])
        e.set_A(A0)
        b, en = slab_range(r, world, n)
        e.ctx.check(e.ctx.lib.sc_fit_set_shard(e.h, b, en), "set_shard")
        engs.append(e)
    lib, ctx = engs[0].ctx.lib, engs[0].ctx
    S, c = engs[0].S, slab_size(n, world)

    def slots(e):
        pi, pv, pc, rows = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_longlong()
        ctx.check(lib.sc_fit_A_slots_device(e.h, C.byref(pi), C.byref(pv), C.byref(pc), C.byref(rows)), "slots")
        R = int(rows.value)
        return (device_view(pi.value, (R, S), torch.int32), device_view(pv.value, (R, S), torch.float64),
                device_view(pc.value, (R,), torch.int32))

    def update_A():
        for e in engs:
            e.update_A(0.0)
        ctx.sync()
        views = [slots(e) for e in engs]
        for r in range(world):          # "all-gather": every engine receives every other engine's slab
            b, en = slab_range(r, world, n)
            for q in range(world):
                if q != r:
                    for dst, src in zip(views[q], views[r]):
                        dst[b:en].copy_(src[b:en])
        torch.cuda.synchronize()

    def update_B():
        parts = [torch.zeros(k * 4, dtype=torch.float64, device="cuda") for _ in engs]
        for e in engs:
            ctx.check(lib.sc_fit_update_B_begin(e.h), "begin")
        for t in range(T):
            for e, p in zip(engs, parts):
                ctx.check(lib.sc_fit_update_B_eval(e.h, t, C.c_void_p(p.data_ptr())), "eval")
            ctx.sync()
            gathered = torch.cat(parts).contiguous()
            torch.cuda.synchronize()
            for e in engs:
                ctx.check(lib.sc_fit_update_B_step(e.h, t, C.c_void_p(gathered.data_ptr()), world), "step")
            ctx.sync()
        for e in engs:
            ctx.check(lib.sc_fit_update_B_end(e.h), "end")

    update_A()
    rss = [engs[0].rss()]
    for it in range(int(g["n_iter"])):
        update_A()
        update_B()
        r = [e.rss() for e in engs]
        assert len(set(r)) == 1, "replicated RSS must agree bit for bit across ranks"
        rss.append(r[0])
    return engs, rss


@pytest.mark.parametrize("world", [2, 3])
def test_emulated_shards_reproduce_the_reference(world):
    g = load_golden("rna_n600_k8_sparse")
    engs, rss = _emulated_fit(g, world)
    np.testing.assert_allclose(rss, g["RSS_iters"], rtol=1e-9)
    for e in engs:
        assert same_sparse(e.A(), csr_from(g, "A_final"))
        assert same_sparse(e.B(), csr_from(g, "B_final"))
        assert np.array_equal(e.hard_assignments(), g["labels"])


def test_emulated_shards_with_hub_archetypes():
    """5000 cells: hub archetypes exist, so the cell-major hub block and the per-rank hub selection are exercised;
    the sharded result must equal the single-engine result bit for bit."""
    from seacells_b200 import FitEngine
    from seacells_b200.synth import synth_archetypes, synth_assignments, synth_embedding
    n, k = 5000, 66
    X = synth_embedding(n, 50, 0)
    M, *_ = O.build_kernel(X.astype(np.float64), 15)
    arch, A0 = synth_archetypes(n, k), synth_assignments(n, k)
    g = {"M_data": M.data, "M_indices": M.indices.astype(np.int32), "M_indptr": M.indptr.astype(np.int64),
         "M_shape": np.array(M.shape), "k": k, "T": 50, "archetypes": arch, "n_iter": 2}
    A0c = sp.csr_matrix(A0)
    g.update({"A0_data": A0c.data, "A0_indices": A0c.indices.astype(np.int32), "A0_indptr": A0c.indptr.astype(np.int64),
              "A0_shape": np.array(A0c.shape)})
    engs, rss = _emulated_fit(g, 3)
    one = FitEngine(n, k)
    one.set_kernel(M)
    one.set_archetypes(arch)
    one.set_A(O.l1_normalise_columns(A0))
    rss1, *_ = one.run(max_iter=2, min_iter=2)
    assert one.stats()["hub_archetypes_last_update_B"] > 0
    assert rss == rss1
    for e in engs:
        assert same_sparse(e.A(), one.A()) and same_sparse(e.B(), one.B())


def _nccl_worker(rank, world, port, name, ret):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from seacells_b200 import FitEngine, _lib
    from seacells_b200.parallel import ShardedFit, sharded_knn
    g = load_golden(name)
    ctx = _lib.default_context(rank)
    idx, d2 = sharded_knn(g["X32"], 15, ctx)
    io, do = O.knn_exact(g["X32"], 15)
    ok = np.array_equal(idx, io) and np.array_equal(d2, do)
    M = csr_from(g, "M")
    eng = FitEngine(M.shape[0], int(g["k"]), int(g["T"]), ctx)
    eng.set_kernel(M)
    eng.set_archetypes(g["archetypes"])
    eng.set_A(O.l1_normalise_columns(csr_from(g, "A0")))
    sf = ShardedFit(eng)
    rss, n_iter, conv, _ = sf.run(max_iter=int(g["max_iter"]), min_iter=int(g["min_iter"]), eps=float(g["eps"]))
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
