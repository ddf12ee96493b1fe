"""Timing of the metacell aggregation kernel (sc_summarize) on synthetic counts: python scripts/summarize_probe.py [n] [genes]."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, scipy.sparse as sp, torch
from seacells_b200 import _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
genes = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
k = n // 75
rng = np.random.default_rng(0)
nnz_row = 250
indptr = np.arange(0, (n + 1) * nnz_row, nnz_row, dtype=np.int64)
indices = np.sort(rng.integers(0, genes, size=(n, nnz_row)), axis=1).astype(np.int32)
indices += (np.arange(nnz_row, dtype=np.int32)[None, :] * 0)
# make columns distinct within a row: spread them
indices = ((np.arange(nnz_row)[None, :] * (genes // nnz_row)) + rng.integers(0, genes // nnz_row, size=(n, nnz_row))).astype(np.int32)
data = np.ceil(rng.random((n, nnz_row)) * 9).astype(np.float32)
grp = rng.integers(0, k, size=n).astype(np.int32)
ctx = _lib.default_context()
dev = torch.device("cuda", 0)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
d_ptr, d_idx, d_val, d_grp = t(indptr), t(indices.ravel()), t(data.ravel()), t(grp)
d_cell = torch.arange(n, dtype=torch.int32, device=dev)
d_w = torch.ones(n, dtype=torch.float64, device=dev)
d_out = torch.empty((k, genes), dtype=torch.float64, device=dev)
torch.cuda.synchronize()
def run():
    ctx.check(ctx.lib.sc_summarize(ctx.h, n, genes, _lib.ptr(d_ptr), _lib.ptr(d_idx), _lib.ptr(d_val), 0, n, _lib.ptr(d_grp),
                                   _lib.ptr(d_cell), _lib.ptr(d_w), k, None, _lib.ptr(d_out)), "sc_summarize")
for _ in range(3): run()
t0 = time.perf_counter()
for _ in range(5): run()
dt = (time.perf_counter() - t0) / 5
alg = n * nnz_row * 8.0 + 8.0 * (n + 1) + 16.0 * n + 8.0 * k * genes
print(f"sc_summarize n={n} genes={genes} k={k} nnz={n*nnz_row}: {dt*1e3:.2f} ms  algorithmic {alg/1e9:.2f} GB -> {alg/dt/1e9:.0f} GB/s "
      f"(device-resident inputs, includes the member sort)")
tot = d_out.sum(0).cpu().numpy()
ref = np.zeros(genes); np.add.at(ref, indices.ravel(), data.ravel().astype(np.float64))
print("column totals conserved:", np.array_equal(tot, ref))
