"""Shared helpers for the tests: golden loading and compressed <-> scipy conversion."""
import os

import numpy as np
import scipy.sparse as sp

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["rna_n600_k8_sparse", "rna_n600_k8_dense_l2", "atac_n900_k12_sparse", "rna_n2000_k27_sparse"]
# pins the oracle away from the default parameters (10 neighbours, 20 Frank-Wolfe steps); oracle tests only
ORACLE_ONLY_CASES = ["rna_n500_k9_nn10_T20"]


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def csr_from(g, prefix):
    return sp.csr_matrix((g[f"{prefix}_data"], g[f"{prefix}_indices"], g[f"{prefix}_indptr"]),
                         shape=tuple(g[f"{prefix}_shape"]))


def same_sparse(a, b):
    a = sp.csr_matrix(a); b = sp.csr_matrix(b)
    a.sum_duplicates(); b.sum_duplicates(); a.sort_indices(); b.sort_indices()
    return (a.shape == b.shape and np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices)
            and np.array_equal(a.data, b.data))


def max_rel_diff(a, b):
    a = sp.csr_matrix(a); b = sp.csr_matrix(b)
    d = abs(a - b)
    if d.nnz == 0:
        return 0.0
    return float(d.max() / max(abs(a).max(), 1e-300))
