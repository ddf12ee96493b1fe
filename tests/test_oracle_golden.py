"""Pins oracle/seacells_oracle.py against outputs of the unmodified reference (tests/golden/*.npz)."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import seacells_oracle as O
from tests.helpers import CASES, ORACLE_ONLY_CASES, csr_from, load_golden, same_sparse


@pytest.mark.parametrize("name", CASES + ORACLE_ONLY_CASES)
def test_kernel_build_matches_reference(name):
    g = load_golden(name)
    X = g["X32"].astype(np.float64)
    M, sigma, idx, d2 = O.build_kernel(X, int(g["n_neighbors"]), "union")
    D = O.knn_distance_graph(idx, d2, X.shape[0])
    assert same_sparse(D, csr_from(g, "D"))
    assert np.array_equal(sigma, g["sigma"])
    assert same_sparse(M, csr_from(g, "M"))                       # bit-exact for float64 input
    Mi, *_ = O.build_kernel(X, int(g["n_neighbors"]), "intersection", knn=(idx, d2))
    assert same_sparse(Mi, csr_from(g, "Mint"))
    assert abs(M - M.T).max() == 0.0 and np.all(M.diagonal() == 1.0)
    with pytest.raises(ValueError):
        O.symmetrise(D, "bogus")


@pytest.mark.parametrize("name", CASES + ORACLE_ONLY_CASES)
def test_fit_matches_reference(name):
    g = load_golden(name)
    M = csr_from(g, "M")
    dense = bool(g["dense"])
    A0 = csr_from(g, "A0n") if dense else csr_from(g, "A0")
    steps = {}
    r = O.fit(M, g["archetypes"], A0, max_iter=int(g["max_iter"]), min_iter=int(g["min_iter"]),
              convergence_epsilon=float(g["eps"]), max_FW_iter=int(g["T"]), dense=dense, l2_penalty=float(g["l2"]),
              record=lambda s, A, B: steps.__setitem__(s, (sp.csr_matrix(A), sp.csr_matrix(B))) if s <= 2 else None)
    for s in (0, 1, 2):
        assert same_sparse(steps[s][0], csr_from(g, f"A_step{s}")), f"A after step {s}"
        assert same_sparse(steps[s][1], csr_from(g, f"B_step{s}")), f"B after step {s}"
    assert same_sparse(r["A"], csr_from(g, "A_final"))
    assert same_sparse(r["B"], csr_from(g, "B_final"))
    assert r["n_iter"] == int(g["n_iter"]) and r["converged"] == bool(g["converged"])
    np.testing.assert_allclose(r["RSS_iters"], g["RSS_iters"], rtol=1e-12)
    assert np.array_equal(r["labels"], g["labels"])
    si, sw = O.soft_assignments(r["A"], r["B"])
    assert np.array_equal(si, g["soft_idx"]) and np.array_equal(sw, g["soft_w"])
    assert np.array_equal(O.hard_archetypes(r["B"]), g["hard_archetypes"])
    # the n x n-free RSS used for large-n checks agrees with the reference's definition
    assert abs(O.rss_trace_identity(M, r["A"], r["B"]) - g["RSS_iters"][-1]) <= 1e-10 * g["RSS_iters"][-1]


def test_sparse_and_dense_semantics_agree():
    """SURVEY.md headline fact: cpu.py and cpu_dense.py give bit-identical A/B on identical inputs."""
    g = load_golden("rna_n600_k8_sparse")
    M = csr_from(g, "M")
    A0 = O.l1_normalise_columns(csr_from(g, "A0"))
    rs = O.fit(M, g["archetypes"], csr_from(g, "A0"), max_iter=3, min_iter=3)
    rd = O.fit(M, g["archetypes"], A0.toarray(), max_iter=3, min_iter=3, dense=True)
    assert same_sparse(rs["A"], sp.csr_matrix(rd["A"])) and same_sparse(rs["B"], sp.csr_matrix(rd["B"]))


def test_sparse_argmin_rule():
    """Implicit-zero rule of scipy's sparse argmin (scipy/sparse/_data.py), including its `== 0` corner."""
    rng = np.random.default_rng(0)
    for _ in range(50):
        G = sp.random(7, 40, density=0.4, random_state=rng.integers(1 << 30), data_rvs=lambda s: rng.normal(size=s))
        G = sp.csc_matrix(G)
        if rng.random() < 0.5:
            G.data = np.abs(G.data)
        G.data[rng.random(G.nnz) < 0.1] = 0.0
        assert np.array_equal(O.sparse_argmin_cols(G), np.asarray(G.argmin(axis=0)).reshape(-1))


def test_greedy_centers_matches_reference():
    g = load_golden("rna_n600_k8_sparse")
    M = csr_from(g, "M")
    c = O.greedy_centers(M @ M.T, len(g["greedy_centers"]))
    assert np.array_equal(c, g["greedy_centers"])


def test_fit_argument_check():
    g = load_golden("rna_n600_k8_sparse")
    with pytest.raises(ValueError):
        O.fit(csr_from(g, "M"), g["archetypes"], csr_from(g, "A0"), max_iter=3, min_iter=10)


def test_summaries_match_reference():
    """oracle summarize_hard / summarize_soft / sparsify_assignments vs the unmodified core.py (core.py:103-242)."""
    g = load_golden("summarize_n800_k11")
    counts = csr_from(g, "counts")
    names, hard = O.summarize_hard(counts, g["labels"])
    assert list(names) == list(g["hard_names"])                   # order of first appearance
    assert np.array_equal(hard, g["hard_X"])                      # integer counts: exact
    assert np.array_equal(O.summarize_hard(counts * 2, g["labels"])[1], g["hard2_X"])
    assert np.array_equal(O.sparsify_assignments(g["A"].T, 0.05), g["sparsified"])
    soft, sizes = O.summarize_soft(counts, g["A"], 0.05)
    assert np.array_equal(soft, g["soft_X"]) and np.array_equal(sizes, g["soft_sizes"])
