"""Per-stage wall times of the row-sharded fit (torchrun --nproc-per-node N scripts/scale_probe_sharded.py n)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from sklearn.preprocessing import normalize
from seacells_b200 import FitEngine, _lib, build_kernel
from seacells_b200.parallel import ShardedFit, sharded_knn
from seacells_b200.synth import synth_archetypes, synth_assignments, synth_embedding
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = _lib.default_context(local)
n = int(sys.argv[1]); k = n // 75
X = synth_embedding(n, 50, 0)
km = build_kernel(X, 15, "union", ctx, knn_graph=sharded_knn(X, 15, ctx))
eng = FitEngine(n, k, 50, ctx); eng.set_kernel(km); eng.set_archetypes(synth_archetypes(n, k))
eng.set_A(normalize(synth_assignments(n, k), axis=0, norm="l1"))
sf = ShardedFit(eng)
sf.update_A(); sf.rss()
for it in range(2):
    ctx.sync(); dist.barrier(); t = time.time(); sf.update_A(); ctx.sync(); ta = time.time() - t
    t = time.time(); sf.update_B(); ctx.sync(); tb = time.time() - t
    t = time.time(); r = sf.rss(); tr = time.time() - t
    if rank == 0: print(f"outer {it+1}: update_A {ta:.3f} update_B {tb:.3f} rss {tr:.3f} -> {r:.6f}", flush=True)
if rank == 0: print({kk: round(v, 4) for kk, v in eng.profile().items()}, flush=True)
dist.destroy_process_group()
