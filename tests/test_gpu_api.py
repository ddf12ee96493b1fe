"""Reference-facing behaviour of the model class and less common parameter values, against the oracle."""
import os
import pickle
import tempfile

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import seacells_oracle as O
from tests.helpers import same_sparse

pytestmark = pytest.mark.gpu


def _data(n, d=50, seed=0, kind="rna"):
    from seacells_b200.synth import synth_embedding
    X = synth_embedding(n, d, seed, kind)
    M, *_ = O.build_kernel(X.astype(np.float64), 15)
    return X, M


@pytest.mark.parametrize("T", [1, 7, 64])
def test_other_max_FW_iter(T):
    from seacells_b200 import FitEngine
    from seacells_b200.synth import synth_archetypes, synth_assignments
    n, k = 900, 12
    X, M = _data(n, seed=5)
    arch, A0 = synth_archetypes(n, k), synth_assignments(n, k)
    ref = O.fit(M, arch, A0, max_iter=3, min_iter=3, max_FW_iter=T)
    eng = FitEngine(n, k, T)
    eng.set_kernel(M)
    eng.set_archetypes(arch)
    eng.set_A(O.l1_normalise_columns(A0))
    rss, n_iter, conv, _ = eng.run(max_iter=3, min_iter=3)
    assert same_sparse(eng.A(), ref["A"]) and same_sparse(eng.B(), ref["B"])
    np.testing.assert_allclose(rss, ref["RSS_iters"], rtol=1e-9)
    with pytest.raises(Exception):
        FitEngine(n, k, 65)          # compile-time limit of the kernels, reported instead of mis-computing


def test_tiny_k_and_duplicate_cells():
    from seacells_b200 import FitEngine, build_kernel
    rng = np.random.default_rng(3)
    n = 403
    X = rng.normal(size=(n, 9)).astype(np.float32)
    # eleven copies of one cell: fewer than n_neighbors//2 non-zero distances remain, so sigma = 0 and the reference
    # produces NaN on those diagonals (0/0) and drops the -inf exponents; the engine must do exactly the same
    Xd = X.copy()
    Xd[50:60] = Xd[49]
    M, sigma, idx, d2 = build_kernel(Xd, 15).export()
    with np.errstate(all="ignore"):
        Mo, so, io, do = O.build_kernel(Xd.astype(np.float64), 15)
    assert np.array_equal(idx, io) and np.array_equal(d2, do) and np.array_equal(sigma, so) and (so == 0).sum() == 11
    assert np.array_equal(M.indptr, Mo.indptr) and np.array_equal(M.indices, Mo.indices)
    assert np.isnan(Mo.data).sum() == 11 and np.allclose(M.data, Mo.data, rtol=1e-13, atol=0, equal_nan=True)
    X[50:54] = X[49]                 # five copies: zero distances are dropped from the graph, sigma stays positive
    km = build_kernel(X, 15)
    M, sigma, idx, d2 = km.export()
    Mo, so, io, do = O.build_kernel(X.astype(np.float64), 15)
    assert np.array_equal(idx, io) and np.array_equal(d2, do) and np.array_equal(sigma, so)
    assert np.array_equal(M.indices, Mo.indices) and np.allclose(M.data, Mo.data, rtol=1e-13, atol=0)
    for k in (2, 3):
        arch = np.arange(k) * 100 + 7
        A0 = sp.csr_matrix(rng.random((k, n)))
        ref = O.fit(Mo, arch, A0, max_iter=2, min_iter=2)
        eng = FitEngine(n, k)
        eng.set_kernel(Mo)
        eng.set_archetypes(arch)
        eng.set_A(O.l1_normalise_columns(A0))
        rss, *_ = eng.run(max_iter=2, min_iter=2)
        assert same_sparse(eng.A(), ref["A"]) and same_sparse(eng.B(), ref["B"])
        np.testing.assert_allclose(rss, ref["RSS_iters"], rtol=1e-9)


def test_default_random_initialisation_follows_the_legacy_rng_stream():
    """cpu.py:257-264: without initial_assignments A0 comes from np.random (randint then random)."""
    from seacells_b200 import SEACells
    from seacells_b200.synth import MiniAnnData, synth_archetypes
    n, k = 700, 9
    X, M = _data(n, seed=8)
    arch = synth_archetypes(n, k)
    np.random.seed(42)
    per = int(k * 0.25)
    r = np.random.randint(0, k, size=(n, per)).reshape(-1)
    A0 = sp.csr_matrix((np.random.random(len(r)), (r, np.repeat(np.arange(n), per))), shape=(k, n))
    ref = O.fit(M, arch, A0, max_iter=2, min_iter=2)
    model = SEACells(MiniAnnData(X), "X_pca", k, verbose=False, use_sparse=True)
    model.add_precomputed_kernel_matrix(M)
    np.random.seed(42)
    try:
        model.fit(max_iter=2, min_iter=2, initial_archetypes=arch)
    except RuntimeWarning:
        pass
    assert same_sparse(model.A_, ref["A"]) and same_sparse(model.B_, ref["B"])
    assert same_sparse(model.A0, O.l1_normalise_columns(A0))


def test_step_explicit_updates_and_persistence():
    from seacells_b200 import SEACells
    from seacells_b200.synth import MiniAnnData, synth_archetypes, synth_assignments
    n, k = 800, 10
    X, M = _data(n, d=30, seed=2, kind="atac")
    arch, A0 = synth_archetypes(n, k), synth_assignments(n, k)
    ad = MiniAnnData(X.astype(np.float64), key="X_svd")          # float64 embedding under the scATAC key
    model = SEACells(ad, "X_svd", k, verbose=False, use_sparse=True, convergence_epsilon=1e-5)
    with pytest.raises(RuntimeError):
        model.step()                                                  # cpu.py:566-579
    model.construct_kernel_matrix(graph_construction="intersection")
    Mi, *_ = O.build_kernel(X.astype(np.float64), 15, "intersection")
    assert np.array_equal(model.kernel_matrix.indices, Mi.indices)
    assert "distances" in ad.obsp and ad.obsp["distances"].shape == (n, n)
    with pytest.raises(ValueError):
        model.construct_kernel_matrix(graph_construction="nonsense")
    model.construct_kernel_matrix()                                   # back to the union graph
    with pytest.raises(RuntimeError):
        model.step()                                                  # not initialised yet
    model.initialize(initial_archetypes=arch, initial_assignments=A0)
    K = model.kernel_matrix @ model.kernel_matrix.T
    A = O.update_A(K, sp.csr_matrix((np.ones(k), (arch, np.arange(k))), shape=(n, k)), O.l1_normalise_columns(A0))
    assert same_sparse(model.A_, A)
    thr = model.convergence_threshold
    model.step()
    model.step()                                                      # "resume" = call step() again
    assert len(model.RSS_iters) == 3 and model.convergence_threshold == thr
    assert list(ad.obs["SEACell"]) == list(model.get_hard_assignments()["SEACell"])
    # explicit-matrix entry points (cpu.py:425, 461) agree with the stateful path
    A_now, B_now = model.A_, model.B_
    B_next = O.update_B(K, A_now, B_now)
    assert same_sparse(model._updateB(A_now, B_now), B_next)
    assert same_sparse(model._updateA(B_next, A_now), O.update_A(K, B_next, A_now))
    np.testing.assert_allclose(model.compute_RSS(A_now, B_now), O.compute_RSS(model.kernel_matrix, A_now, B_now), rtol=1e-9)
    R = model.compute_reconstruction(A_now, B_now)
    assert R.shape == (n, n)
    with tempfile.TemporaryDirectory() as d:
        model.save_assignments(d)                                     # cpu.py:743-764
        assert sorted(os.listdir(d)) == ["A.npz", "B.npz", "SEACells.csv", "kernel_matrix.npz"]
        assert sp.load_npz(os.path.join(d, "A.npz")).shape == (n, k)  # A.npz holds A_.T
        model.save_model(d)                                           # cpu.py:731-741
        m2 = pickle.load(open(os.path.join(d, "model.pkl"), "rb"))
        assert m2.RSS_iters == model.RSS_iters and m2.k == k
        # the pickle keeps the fitted state (cpu.py:731-741 pickles A_, B_, Z_ with the model)
        assert same_sparse(m2.A_, model.A_) and same_sparse(m2.B_, model.B_)
        assert same_sparse(m2.kernel_matrix, model.kernel_matrix)
        assert list(m2.get_hard_assignments()["SEACell"]) == list(model.get_hard_assignments()["SEACell"])
        assert list(m2.get_hard_archetypes()) == list(model.get_hard_archetypes())
        s2, w2 = m2.get_soft_assignments()
        s1, w1 = model.get_soft_assignments()
        assert np.array_equal(w1, w2) and s1.equals(s2)
        np.testing.assert_allclose(m2.compute_RSS(), model.compute_RSS(), rtol=1e-12)
        m2.step()                                                     # a restored model resumes like the original
        model.step()
        assert same_sparse(m2.A_, model.A_) and same_sparse(m2.B_, model.B_)


def test_pickle_keeps_Z_after_fit():
    import pickle
    from seacells_b200 import SEACells
    from seacells_b200.synth import MiniAnnData, synth_archetypes, synth_assignments
    n, k = 700, 9
    X, M = _data(n, seed=4)
    model = SEACells(MiniAnnData(X), "X_pca", k, verbose=False, use_sparse=True)
    model.add_precomputed_kernel_matrix(M)
    try:
        model.fit(max_iter=12, min_iter=3, initial_archetypes=synth_archetypes(n, k),
                  initial_assignments=synth_assignments(n, k))
    except RuntimeWarning:
        pass                                           # raised after the results are stored (cpu.py:647-650)
    m2 = pickle.loads(pickle.dumps(model))
    assert m2._engine is None and m2.Z_ is not None
    assert same_sparse(m2.Z_, model.get_archetype_matrix())
    assert same_sparse(m2.A_, model.A_) and same_sparse(m2.B_, model.B_)


def test_input_validation():
    """Inputs the engine cannot treat like the reference are rejected loudly instead of mis-computed."""
    from seacells_b200 import FitEngine, SEACells, SeacellsB200Error
    from seacells_b200.synth import MiniAnnData, synth_archetypes, synth_assignments
    n, k = 600, 8
    X, M = _data(n, seed=6)
    eng = FitEngine(n, k)
    bad = sp.csr_matrix(M).copy().tolil()
    bad[3, 5] = 0.25                                   # M[3,5] != M[5,3]: K = M M^T is no longer M (M .)
    with pytest.raises(SeacellsB200Error, match="symmetric"):
        eng.set_kernel(bad.tocsr())
    neg = sp.csr_matrix(M).copy()
    neg.data[7] = -neg.data[7]
    with pytest.raises(SeacellsB200Error, match="negative"):
        eng.set_kernel(neg)
    eng.set_kernel(M)
    eng.set_archetypes(synth_archetypes(n, k))
    A0 = O.l1_normalise_columns(synth_assignments(n, k)).tocsc()
    A0.data[0] = -0.5
    with pytest.raises(SeacellsB200Error, match="non-negative"):
        eng.set_A(A0)
    ad = MiniAnnData(X)
    with pytest.raises(ValueError, match="max_franke_wolfe_iters"):
        SEACells(ad, "X_pca", k, verbose=False, max_franke_wolfe_iters=65)
    with pytest.raises(ValueError, match="n_neighbors"):
        SEACells(ad, "X_pca", k, verbose=False, n_neighbors=40)
    with pytest.raises(ValueError):
        eng.run(max_iter=3, min_iter=5)


def test_two_devices_in_one_process():
    """Every extern "C" entry point makes its context's device current (and restores the caller's)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from seacells_b200 import FitEngine, _lib
    from seacells_b200.synth import synth_archetypes, synth_assignments
    n, k = 900, 10
    X, M = _data(n, seed=8)
    arch, A0 = synth_archetypes(n, k), O.l1_normalise_columns(synth_assignments(n, k))
    res = []
    engines = [FitEngine(n, k, 50, _lib.default_context(dv)) for dv in (0, 1)]
    for e in engines:
        e.set_kernel(M); e.set_archetypes(arch); e.set_A(A0)
    for step in range(2):
        for dv, e in enumerate(engines):
            torch.cuda.set_device(1 - dv)               # the caller's current device is the *other* GPU
            e.update_A(); e.update_B()
            assert torch.cuda.current_device() == 1 - dv
    assert same_sparse(engines[0].A(), engines[1].A()) and same_sparse(engines[0].B(), engines[1].B())
    assert engines[0].rss() == engines[1].rss()


def test_evaluate_compactness_and_separation():
    """evaluate.py:6-81 on injected diffusion components (palantir is third party): group statistics as pandas computes
    them, nearest-metacell distances as sklearn's brute-force NearestNeighbors returns them."""
    import pandas as pd
    from sklearn.neighbors import NearestNeighbors
    from seacells_b200 import evaluate
    from seacells_b200.synth import MiniAnnData, synth_embedding
    n, k = 3000, 40
    X = synth_embedding(n, 50, 9)
    rng = np.random.default_rng(3)
    dc = rng.normal(size=(n, 10)) * np.linspace(2.0, 0.2, 10)
    ad = MiniAnnData(X)
    ad.obs["SEACell"] = [f"SEACell-{i}" for i in rng.integers(0, k, size=n)]
    ad.obs["celltype"] = rng.choice(["a", "b", "c"], size=n)
    df = pd.DataFrame(dc, index=ad.obs_names).join(ad.obs["SEACell"])
    comp = evaluate.compactness(ad, diffusion_components=dc)
    np.testing.assert_allclose(comp["compactness"].values, df.groupby("SEACell").var().mean(axis=1).values, rtol=1e-12)
    cent = df.groupby("SEACell").mean()
    for nth in (1, 3):
        sep = evaluate.separation(ad, nth_nbr=nth, diffusion_components=dc)
        ref = NearestNeighbors(n_neighbors=nth, algorithm="brute").fit(cent.values).kneighbors()[0][:, nth - 1]
        assert list(sep.index) == list(cent.index)
        # the reference renames column 1 only (evaluate.py:79-81): for nth_nbr != 1 the column keeps the name nth_nbr
        assert list(sep.columns) == (["separation"] if nth == 1 else [nth])
        np.testing.assert_allclose(sep.iloc[:, 0].values, ref, rtol=1e-8)  # sklearn expands ||x-y||^2 (less exact)
    sep_c = evaluate.separation(ad, nth_nbr=1, cluster="celltype", diffusion_components=dc)
    assert 0 < len(sep_c) <= k
    pur = evaluate.compute_celltype_purity(ad, "celltype")
    assert pur.shape == (k, 2) and (pur["celltype_purity"] <= 1).all()
    with pytest.raises(ImportError):
        evaluate.compactness(ad)           # palantir is not installed here: clear error, no silent fallback
