"""construct_kernel_matrix once at n cells x d dims (used under ncu: launch list / full capture of the kNN kernels)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seacells_b200 import _lib, build_kernel
from seacells_b200.synth import synth_embedding
n = int(sys.argv[1]); d = int(sys.argv[2]) if len(sys.argv) > 2 else 50
ctx = _lib.default_context()
X = synth_embedding(n, d, 0)
build_kernel(X[:4096], 15)
t = time.time(); km = build_kernel(X, 15); ctx.sync(); dt = time.time() - t
print(f"construct_kernel_matrix n={n} d={d}: {dt:.3f}s nnz/row={km.nnz / n:.2f} fallback={ctx.lib.sc_knn_fallback_rows(ctx.h)} launches={ctx.launches}")
