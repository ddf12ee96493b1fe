"""Development check of the tensor-core kNN filter: raw tile values, exactness of the kNN, timings (python scripts/knn_tc_check.py [n_big])."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from seacells_b200 import _lib
from seacells_b200.core import knn
from seacells_b200.synth import synth_embedding

ctx = _lib.default_context()


def probe(n, d, cols):
    X = synth_embedding(n, d, 0)
    s = np.zeros((128, cols), np.float32)
    xc = np.zeros((n, d), np.float32)
    nn = np.zeros(n, np.float32)
    ctx.check(ctx.lib.sc_knn_tc_probe(ctx.h, _lib.ptr(X), 0, n, d, cols, _lib.ptr(s), _lib.ptr(xc), _lib.ptr(nn)), "probe")
    x64 = xc.astype(np.float64)
    ref = x64[:128] @ x64[:cols].T - 0.5 * nn[:cols].astype(np.float64)[None, :]
    err = np.abs(s.astype(np.float64) - ref)
    nr = np.sqrt((x64 ** 2).sum(1))
    scale = (nr[:128, None] + nr[None, :cols]) ** 2
    fac = ctx.lib.sc_knn_tc_error_factor(d)
    print(f"probe n={n} d={d}: max|s-ref|={err.max():.3e}  max 2*err/(|xi|+|xj|)^2={(2 * err / scale).max():.3e}  "
          f"bound factor={fac:.3e}  ratio={(2 * err / scale).max() / fac:.4f}", flush=True)
    bad = np.argwhere(err > 1e-2 * np.abs(ref).max())
    if len(bad):
        print("  LARGE ERRORS at", bad[:10].tolist(), "s", s[tuple(bad[0])], "ref", ref[tuple(bad[0])])
        print("  row0 s  ", s[0, :8]); print("  row0 ref", ref[0, :8])
        print("  row1 s  ", s[1, :8]); print("  row1 ref", ref[1, :8])
        print("  row9 s  ", s[9, :8]); print("  row9 ref", ref[9, :8])
    return err.max()


def exact_check(n, d):
    X = synth_embedding(n, d, 1)
    os.environ.pop("SC_KNN_FFMA", None); os.environ.pop("SC_KNN_EXACT", None)
    t = time.time(); i1, d1 = knn(X, 15, ctx); t1 = time.time() - t
    fb = ctx.lib.sc_knn_fallback_rows(ctx.h)
    os.environ["SC_KNN_EXACT"] = "1"
    t = time.time(); i0, d0 = knn(X, 15, ctx); t0 = time.time() - t
    os.environ.pop("SC_KNN_EXACT")
    print(f"exact check n={n} d={d}: idx equal {np.array_equal(i0, i1)} d2 equal {np.array_equal(d0, d1)} "
          f"fallback rows {fb}  tc {t1:.3f}s  all-fp64 {t0:.3f}s", flush=True)


def timing(n, d=50):
    X = synth_embedding(n, d, 0)
    for name, env in (("tc", None), ("ffma", "SC_KNN_FFMA")):
        if env:
            os.environ[env] = "1"
        knn(X[:4096], 15, ctx)
        t = time.time(); i1, _ = knn(X, 15, ctx); dt = time.time() - t
        fb = ctx.lib.sc_knn_fallback_rows(ctx.h)
        if env:
            os.environ.pop(env)
        print(f"knn n={n} d={d} [{name}]: {dt:.3f}s (incl. H2D/D2H)  fallback rows {fb}  "
              f"{2.0 * n * n * d / dt / 1e12:.1f} algorithmic TFLOP/s", flush=True)


if __name__ == "__main__":
    probe(4096, 50, 4096)
    probe(2048 + 77, 30, 2048)
    probe(3000, 63, 3000)
    exact_check(20000, 50)
    exact_check(5000, 30)
    big = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    timing(big)
