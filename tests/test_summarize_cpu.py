"""Host-side logic of seacells_b200.summarize that needs no GPU: sparsify_assignments (core.py:103-118) on dense and
sparse input against the reference's golden output, and the MiniAnnData row selection the summaries rely on."""
import numpy as np
import scipy.sparse as sp

from tests.helpers import csr_from, load_golden


def test_sparsify_assignments_dense_and_sparse():
    from seacells_b200.summarize import sparsify_assignments
    g = load_golden("summarize_n800_k11")
    W = g["A"].T
    assert np.array_equal(sparsify_assignments(W, 0.05), g["sparsified"])
    Ws = sparsify_assignments(sp.csr_matrix(W), 0.05)
    np.testing.assert_allclose(np.asarray(Ws.todense()), g["sparsified"], rtol=1e-15, atol=0)
    # a row with nothing above the threshold becomes NaN (0/0), as in the reference
    W2 = W.copy()
    W2[5] = 0.01
    out = sparsify_assignments(W2, 0.05)
    assert np.all(np.isnan(out[5])) and np.array_equal(out[6], g["sparsified"][6])


def test_mini_anndata_row_selection():
    from seacells_b200.synth import MiniAnnData
    g = load_golden("summarize_n800_k11")
    counts = csr_from(g, "counts")
    ad = MiniAnnData(np.zeros((counts.shape[0], 2)), "X_pca", raw=counts, layers={"twice": (counts * 2).tocsr()})
    ad.obs["SEACell"] = g["labels"]
    cells = ad.obs_names[ad.obs["SEACell"] == g["labels"][0]]
    sub = ad[cells, :]
    rows = np.nonzero(g["labels"] == g["labels"][0])[0]
    assert sub.shape[0] == len(rows) and list(sub.obs_names) == list(cells)
    assert (sub.raw.X != counts[rows]).nnz == 0 and (sub.layers["twice"] != (counts * 2).tocsr()[rows]).nnz == 0
    assert list(ad.var_names[:2]) == ["g0", "g1"] and ad.shape == counts.shape
