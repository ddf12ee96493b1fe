"""TEST INFRASTRUCTURE ONLY — loads the *unmodified* reference files from /root/reference.

Only `tests/golden/make_golden.py` (and ad-hoc validation in the build container) may import this.
/root/reference does not exist on the GPU box, so nothing under `-m gpu`, `smoke()` or `bench.py`
uses this module; they use `oracle/seacells_oracle.py` (the restatement pinned by the golden vectors).

The reference imports scanpy / palantir / tqdm.notebook at module scope (build_graph.py:8,119, cpu.py:4),
none of which exist here.  We register stub modules and give `scanpy.pp.neighbors` an *exact* brute-force
stand-in: d2(i,j) = sum_c (x_ic - x_jc)^2 accumulated sequentially in fp64, neighbours ordered
by (d2, j), self excluded by index, Euclidean (sqrt) distances stored, explicit zeros eliminated.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np
import scipy.sparse as sp

from oracle.seacells_oracle import knn_exact as exact_knn  # one definition of the kNN stand-in

REF = "/root/reference/SEACells"


def _neighbors(ad, use_rep, n_neighbors, knn=True):
    X = ad.obsm[use_rep]
    n = X.shape[0]
    idx, d2 = exact_knn(X, n_neighbors)
    dist = np.sqrt(d2)
    indptr = np.arange(0, n * (n_neighbors - 1) + 1, n_neighbors - 1)
    D = sp.csr_matrix((dist.ravel(), idx.ravel().astype(np.int64), indptr), shape=(n, n))
    D.eliminate_zeros()
    D.sort_indices()
    ad.obsp["distances"] = D


def install_shims():
    if not os.path.isdir(REF):
        raise RuntimeError("reference tree not present; ref_shim is only usable in the build container")
    pal = types.ModuleType("palantir")
    pal.utils = types.ModuleType("palantir.utils")
    pal.core = types.ModuleType("palantir.core")
    sc = types.ModuleType("scanpy")
    sc.pp = types.ModuleType("scanpy.pp")
    sc.pp.neighbors = _neighbors
    sys.modules.update({"palantir": pal, "palantir.utils": pal.utils, "palantir.core": pal.core,
                        "scanpy": sc, "scanpy.pp": sc.pp})
    import tqdm.notebook
    tqdm.notebook.tqdm = lambda it, *a, **k: it


def load(name: str):
    install_shims()
    pkg = sys.modules.get("SEACells")
    if pkg is None:
        pkg = types.ModuleType("SEACells")
        pkg.__path__ = [REF]
        sys.modules["SEACells"] = pkg
    full = f"SEACells.{name}"
    if full in sys.modules:
        return sys.modules[full]
    spec = importlib.util.spec_from_file_location(full, f"{REF}/{name}.py")
    m = importlib.util.module_from_spec(spec)
    sys.modules[full] = m
    spec.loader.exec_module(m)
    return m
