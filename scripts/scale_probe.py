"""Per-stage wall times of the engine at a given size (development aid; not the bench)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from seacells_b200 import FitEngine, build_kernel
from seacells_b200.synth import synth_archetypes, synth_assignments, synth_embedding
from sklearn.preprocessing import normalize

n = int(sys.argv[1]); k = int(sys.argv[2]) if len(sys.argv) > 2 else n // 75
outer = int(sys.argv[3]) if len(sys.argv) > 3 else 2
def sync(): torch.cuda.synchronize()
t = time.time(); X = synth_embedding(n, 50, 0); print(f"synth {time.time()-t:.2f}s", flush=True)
t = time.time(); km = build_kernel(X, 15); sync(); print(f"kernel build {time.time()-t:.3f}s nnz/row={km.nnz/n:.2f}", flush=True)
eng = FitEngine(n, k)
t = time.time(); eng.set_kernel(km); eng.set_archetypes(synth_archetypes(n, k))
A0 = normalize(synth_assignments(n, k), axis=0, norm="l1"); eng.set_A(A0); print(f"setup {time.time()-t:.3f}s", flush=True)
t = time.time(); eng.update_A(); eng.ctx.sync(); print(f"init update_A {time.time()-t:.3f}s", eng.stats(), flush=True)
t = time.time(); r = eng.rss(); print(f"rss0 {time.time()-t:.3f}s -> {r:.6f}", flush=True)
prev = eng.profile()
for it in range(outer):
    t = time.time(); eng.update_A(); eng.ctx.sync(); ta = time.time()-t
    t = time.time(); eng.update_B(); eng.ctx.sync(); tb = time.time()-t
    t = time.time(); r = eng.rss(); tr = time.time()-t
    cur = eng.profile()
    print(f"outer {it+1}: update_A {ta:.3f}s update_B {tb:.3f}s rss {tr:.3f}s -> {r:.6f}", {k: round(cur[k] - prev[k], 4) for k in cur},
          {k: v for k, v in eng.stats().items() if k in ("t2_nnz_last_update_B", "hub_archetypes_last_update_B", "kb_nnz_last", "gradients_evaluated_last_update_B")}, flush=True)
    prev = cur
free, total = torch.cuda.mem_get_info(); print(f"device mem used {(total-free)/1e9:.1f} GB", flush=True)
print("launches", eng.ctx.launches)
print({k: round(v, 4) for k, v in eng.profile().items()})
