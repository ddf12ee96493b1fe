"""Per-call wall times of the row-sharded engine (launch with torchrun; development aid, not the bench)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from sklearn.preprocessing import normalize
from seacells_b200 import FitEngine, _lib, build_kernel
from seacells_b200.parallel import init_nccl
from seacells_b200.synth import synth_archetypes, synth_assignments, synth_embedding

n = int(sys.argv[1]); k = n // 75; outer = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = init_nccl(_lib.default_context(local))
X = synth_embedding(n, 50, 0)
build_kernel(X[:4096], 15, "union", _lib.Context(local)).close()
t = time.time(); km = build_kernel(X, 15, "union", ctx); ctx.sync(); tb = time.time() - t
eng = FitEngine(n, k, 50, ctx)
t = time.time(); eng.set_kernel(km); ctx.sync(); ts = time.time() - t
eng.set_archetypes(synth_archetypes(n, k)); eng.set_A(normalize(synth_assignments(n, k), axis=0, norm="l1"))
def timed(fn):
    ctx.sync(); dist.barrier(); t = time.time(); r = fn(); ctx.sync(); return time.time() - t, r
ta, _ = timed(eng.update_A); tr, r0 = timed(eng.rss)
if rank == 0: print(f"world {ctx.world} build {tb:.3f}s set_kernel(K) {ts:.3f}s init update_A {ta:.3f}s rss {tr:.3f}s -> {r0!r}", flush=True)
for it in range(outer):
    ta, _ = timed(eng.update_A); tb2, _ = timed(eng.update_B); tr, r = timed(eng.rss)
    if rank == 0:
        print(f"outer {it+1}: update_A {ta:.3f}s update_B {tb2:.3f}s rss {tr:.3f}s -> {r!r}", {k2: round(v, 4) for k2, v in eng.profile().items()}, flush=True)
dist.barrier(); dist.destroy_process_group()
