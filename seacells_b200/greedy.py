"""Greedy adaptive CSSP initialisation (cpu.py:342-423) on the GPU - "next" row of the scope table."""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from . import _lib
from ._lib import ptr


def greedy_centers_from_kernel(M, n_centers: int, ctx=None) -> np.ndarray:
    """`n_centers` cell positions chosen by greedy column-subset selection on K = M M^T (M: scipy CSR kernel matrix)."""
    ctx = ctx or _lib.default_context()
    M = sp.csr_matrix(M)
    if not M.has_sorted_indices:
        M = M.copy()
        M.sort_indices()
    indptr, indices = M.indptr.astype(np.int32), M.indices.astype(np.int32)
    data = np.ascontiguousarray(M.data, np.float64)
    out = np.zeros(int(n_centers), np.int64)
    ctx.check(ctx.lib.sc_greedy_centers(ctx.h, M.shape[0], ptr(indptr), ptr(indices), ptr(data), int(M.nnz),
                                        int(n_centers), ptr(out)), "sc_greedy_centers")
    return out


def greedy_centers(model, n_centers):
    """cpu.py:342-423 for a SEACellsB200 model (uses its kernel matrix)."""
    if model.kernel_matrix is None:
        raise RuntimeError("Must first construct kernel matrix before initializing SEACells.")
    if model.verbose:
        print("Initializing residual matrix using greedy column selection")
    return greedy_centers_from_kernel(model.kernel_matrix, n_centers, model._ctx)
