"""CSR SpMM roofline probe: Y[r x c] = M[r x r] X[r x c] with device-resident operands, CUDA events on the
library stream.  Algorithmic bytes = 12 nnz + 4 (r+1) + 8 c r + 8 c r.   python scripts/spmm_probe.py [n] [c] [reps]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from seacells_b200 import _lib, build_kernel
from seacells_b200.synth import synth_embedding

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
c = int(sys.argv[2]) if len(sys.argv) > 2 else n // 75
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
ctx = _lib.default_context(0)
km = build_kernel(synth_embedding(n, 50, 0), 15, "union", ctx)
M, *_ = km.export(want_knn=False)
dev = torch.device("cuda", 0)
ip = torch.from_numpy(M.indptr.astype(np.int32)).to(dev)
ix = torch.from_numpy(M.indices.astype(np.int32)).to(dev)
dv = torch.from_numpy(M.data).to(dev)
X = torch.rand((n, c), dtype=torch.float64, device=dev)
Y = torch.empty((n, c), dtype=torch.float64, device=dev)
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
torch.cuda.synchronize()
def call():
    ctx.check(ctx.lib.sc_spmm_csr_f64(ctx.h, n, n, c, _lib.ptr(ip), _lib.ptr(ix), _lib.ptr(dv), _lib.ptr(X), _lib.ptr(Y)), "spmm")
for _ in range(3): call()
ts = []
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); call(); e1.record(stream); e1.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = float(np.median(ts))
alg = 12.0 * M.nnz + 4.0 * (n + 1) + 16.0 * c * n
ref = torch.sparse.mm(torch.sparse_csr_tensor(ip.long(), ix.long(), dv, size=(n, n)), X) if n <= 200_000 else None
err = float((ref - Y).abs().max()) if ref is not None else None
print(json.dumps({"n": n, "c": c, "nnz": int(M.nnz), "ms_median": ms, "ms_min": float(min(ts)), "algorithmic_bytes": alg,
                  "GBps": alg / ms / 1e6, "frac_of_6551": alg / ms / 1e6 / 6551.4, "max_abs_err_vs_torch": err}))
