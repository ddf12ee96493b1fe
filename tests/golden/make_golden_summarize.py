"""Golden vectors for the metacell summaries, produced by the UNMODIFIED reference core.py (build container only).

    python tests/golden/make_golden_summarize.py

core.summarize_by_SEACell / summarize_by_soft_SEACell (core.py:121-242) are loaded through oracle/ref_shim.py; the
AnnData they build at the end is replaced by a tiny stand-in class (scanpy is not installed), the input AnnData by
seacells_b200.synth.MiniAnnData.  Stored: counts (CSR, integer-valued float32), hard labels, the soft assignment
matrix, cell types, and the reference's outputs.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import pandas as pd
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from seacells_b200.synth import MiniAnnData  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


class _FakeAnnData:
    def __init__(self, X, dtype=None):
        self.X = X
        self.obs = pd.DataFrame(index=pd.RangeIndex(X.shape[0]).astype(str))
        self.layers = {}
        self.obs_names = self.obs.index
        self.var_names = None


def main():
    core = ref_shim.load("core")
    sys.modules["scanpy"].AnnData = _FakeAnnData  # after load(): it (re)installs the stub modules
    rng = np.random.default_rng(7)
    n, genes, k = 800, 300, 11
    counts = sp.random(n, genes, density=0.08, random_state=3, format="csr", dtype=np.float32)
    counts.data = np.ceil(counts.data * 9).astype(np.float32)
    labels_code = rng.integers(0, k, size=n)
    labels = np.array([f"SEACell-{c}" for c in labels_code])
    celltypes = np.array(["T", "B", "NK", "Mono"])[(labels_code % 3 + rng.integers(0, 2, size=n)) % 4]
    # soft assignments: a few weights per cell, columns sum to 1; a few cells with only tiny weights (NaN corner excluded)
    A = np.zeros((k, n))
    for j in range(n):
        idx = rng.choice(k, size=rng.integers(1, 5), replace=False)
        w = rng.random(len(idx)) + 0.1
        A[idx, j] = w / w.sum()
    ad = MiniAnnData(np.zeros((n, 2)), "X_pca", raw=counts, layers={"counts2": (counts * 2).tocsr()})
    ad.obs["SEACell"] = labels
    ad.obs["celltype"] = celltypes
    hard = core.summarize_by_SEACell(ad, "SEACell", celltype_label=None, summarize_layer="raw")
    hard2 = core.summarize_by_SEACell(ad, "SEACell", summarize_layer="counts2")
    soft = core.summarize_by_soft_SEACell(ad, A, celltype_label="celltype", summarize_layer="raw", minimum_weight=0.05)
    sparsified = core.sparsify_assignments(A.T, 0.05)
    out = {
        "counts_indptr": counts.indptr.astype(np.int64), "counts_indices": counts.indices.astype(np.int32),
        "counts_data": counts.data, "counts_shape": np.array(counts.shape, np.int64),
        "labels": labels, "celltypes": celltypes, "A": A,
        "hard_names": np.array(list(hard.obs_names)), "hard_X": np.asarray(hard.X.todense()),
        "hard2_X": np.asarray(hard2.X.todense()),
        "soft_X": np.asarray(soft.X.todense()), "soft_sizes": np.asarray(soft.obs["Pseudo-sizes"], dtype=np.float64),
        "soft_celltype": np.array(list(soft.obs["celltype"])),
        "soft_purity": np.asarray(soft.obs["celltype_purity"], dtype=np.float64),
        "sparsified": np.asarray(sparsified),
    }
    np.savez_compressed(os.path.join(HERE, "summarize_n800_k11.npz"), **out)
    print("wrote summarize_n800_k11.npz", out["hard_X"].shape, out["soft_X"].shape)


if __name__ == "__main__":
    main()
