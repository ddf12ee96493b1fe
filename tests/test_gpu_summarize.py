"""Metacell summaries on the GPU (sc_summarize through seacells_b200.summarize) against the reference's golden
vectors and the oracle.  Tolerances: hard assignment on count data exact; soft assignment <= 1e-12 relative (the
weight sums are formed in a different order than numpy's pairwise reduction)."""
import numpy as np
import pandas as pd
import pytest
import scipy.sparse as sp

from oracle import seacells_oracle as O
from tests.helpers import csr_from, load_golden

pytestmark = pytest.mark.gpu

SOFT_RTOL = 1e-12


def _ad(g, counts):
    from seacells_b200.synth import MiniAnnData
    n = counts.shape[0]
    ad = MiniAnnData(np.zeros((n, 2)), "X_pca", raw=counts, layers={"counts2": (counts * 2).tocsr()}, counts=counts)
    ad.obs["SEACell"] = g["labels"]
    ad.obs["celltype"] = g["celltypes"]
    return ad


def test_summarize_vs_reference_golden():
    from seacells_b200 import sparsify_assignments, summarize_by_SEACell, summarize_by_soft_SEACell
    g = load_golden("summarize_n800_k11")
    counts = csr_from(g, "counts").astype(np.float32)
    ad = _ad(g, counts)
    hard = summarize_by_SEACell(ad, "SEACell", summarize_layer="raw")
    assert list(hard.obs_names) == list(g["hard_names"])
    assert np.array_equal(np.asarray(hard.X.todense()), g["hard_X"])
    assert np.array_equal(np.asarray(hard.layers["raw"].todense()), g["hard_X"])
    assert np.array_equal(np.asarray(summarize_by_SEACell(ad, summarize_layer="counts2").X.todense()), g["hard2_X"])
    assert np.array_equal(np.asarray(summarize_by_SEACell(ad, summarize_layer="X").X.todense()), g["hard_X"])
    with_ct = summarize_by_SEACell(ad, celltype_label="celltype")
    assert {"celltype", "celltype_purity"} <= set(with_ct.obs.columns)
    # soft assignment: dense and sparse A give the same numbers
    assert np.array_equal(sparsify_assignments(g["A"].T, 0.05), g["sparsified"])
    sps = sparsify_assignments(sp.csr_matrix(g["A"].T), 0.05)
    np.testing.assert_allclose(np.asarray(sps.todense()), g["sparsified"], rtol=1e-15, atol=0)
    for A in (g["A"], sp.csr_matrix(g["A"])):
        soft = summarize_by_soft_SEACell(ad, A, celltype_label="celltype", minimum_weight=0.05)
        np.testing.assert_allclose(np.asarray(soft.X.todense()), g["soft_X"], rtol=SOFT_RTOL, atol=0)
        np.testing.assert_allclose(np.asarray(soft.obs["Pseudo-sizes"]), g["soft_sizes"], rtol=SOFT_RTOL)
        assert list(soft.obs["celltype"]) == list(g["soft_celltype"])
        np.testing.assert_allclose(np.asarray(soft.obs["celltype_purity"], dtype=float), g["soft_purity"], rtol=1e-12)
    with pytest.raises(ValueError):
        summarize_by_soft_SEACell(ad, g["A"], celltype_label="nope")


def test_summarize_vs_oracle_ragged():
    """Empty metacells, unassigned cells, float64 non-integer data, a cell below the soft threshold (NaN corner)."""
    from seacells_b200.summarize import aggregate, summarize_by_soft_SEACell
    from seacells_b200.synth import MiniAnnData
    rng = np.random.default_rng(3)
    n, genes, k = 1500, 97, 40
    X = sp.random(n, genes, density=0.2, random_state=5, format="csr", dtype=np.float64)
    X[7] = 0
    X.eliminate_zeros()
    grp = rng.integers(0, k - 3, size=n)            # groups k-3..k-1 stay empty
    w = rng.random(n)
    out = aggregate(X, grp, np.arange(n), w, k)
    ref = np.zeros((k, genes))
    Xd = X.toarray()
    for j in range(n):                              # ascending cell order, product rounded then added
        ref[grp[j]] += w[j] * Xd[j]
    assert np.array_equal(out, ref)
    assert np.all(out[k - 3:] == 0.0)
    # soft with a NaN cell: everything that cell has counts for is NaN in the reference, as are the pseudo-sizes
    A = np.zeros((5, n))
    A[rng.integers(0, 5, size=n), np.arange(n)] = 1.0
    A[:, 11] = 0.01
    ad = MiniAnnData(np.zeros((n, 2)), "X_pca", raw=X)
    soft = summarize_by_soft_SEACell(ad, A, minimum_weight=0.05)
    with np.errstate(all="ignore"):
        ref_expr, ref_sizes = O.summarize_soft(X, A, 0.05)
    got = np.asarray(soft.X.todense())
    assert np.array_equal(np.isnan(got), np.isnan(ref_expr))
    ok = ~np.isnan(ref_expr)
    np.testing.assert_allclose(got[ok], ref_expr[ok], rtol=SOFT_RTOL, atol=0)
    assert np.all(np.isnan(np.asarray(soft.obs["Pseudo-sizes"]))) and np.all(np.isnan(ref_sizes))


def test_summarize_at_scale_properties():
    """100k cells x 2000 genes: column totals are conserved by the hard summary, soft rows are weighted means."""
    from seacells_b200.summarize import aggregate
    rng = np.random.default_rng(0)
    n, genes, k = 100_000, 2000, 1333
    X = sp.random(n, genes, density=0.03, random_state=1, format="csr", dtype=np.float32)
    X.data = np.ceil(X.data * 20).astype(np.float32)
    grp = rng.integers(0, k, size=n)
    out = aggregate(X, grp, np.arange(n), np.ones(n), k)
    assert np.array_equal(out.sum(0), np.asarray(X.sum(0), dtype=np.float64).ravel())   # integer counts: exact
    onehot = sp.csr_matrix((np.ones(n), (grp, np.arange(n))), shape=(k, n))
    assert np.array_equal(out, (onehot @ X.astype(np.float64)).toarray())
