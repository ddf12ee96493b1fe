"""Parity of the CUDA path (through the C ABI) against the oracle and the reference's golden vectors.

Tolerances (BASELINE.json north_star): kNN index sets bit-exact; A/B entries and per-iteration RSS <= 1e-6 relative
(we assert far tighter: A/B bit-exact where the argmin sequences agree, RSS <= 1e-9); hard assignments identical.
"""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import seacells_oracle as O
from tests.helpers import CASES, csr_from, load_golden, max_rel_diff, same_sparse

pytestmark = pytest.mark.gpu

RSS_RTOL = 1e-9
M_RTOL = 1e-13  # exp() of identical arguments: CUDA and numpy differ by <= 2 ulp


def _engine(n, k, T=50):
    from seacells_b200 import FitEngine
    return FitEngine(n, k, T)


# ------------------------------------------------------------------------------------------------ hot path 1
@pytest.mark.parametrize("name", CASES)
def test_kernel_build_vs_golden(name):
    from seacells_b200 import build_kernel
    g = load_golden(name)
    X32 = g["X32"]
    for X in (X32, X32.astype(np.float64)):  # float32 hand-off and float64 hand-off give the same fp64 arithmetic
        km = build_kernel(X, int(g["n_neighbors"]), "union")
        M, sigma, idx, d2 = km.export()
        n = X.shape[0]
        D = O.knn_distance_graph(idx, d2, n)
        assert same_sparse(D, csr_from(g, "D")), "kNN graph (indices and distances) must be bit-exact"
        assert np.array_equal(sigma, g["sigma"])
        Mg = csr_from(g, "M")
        assert np.array_equal(M.indptr, Mg.indptr) and np.array_equal(M.indices, Mg.indices)
        np.testing.assert_allclose(M.data, Mg.data, rtol=M_RTOL, atol=0)
        assert abs(M - M.T).max() == 0.0, "M must be exactly symmetric"
        assert np.all(M.diagonal() == 1.0)
    Mi, *_ = build_kernel(X32, int(g["n_neighbors"]), "intersection").export()
    Mig = csr_from(g, "Mint")
    assert np.array_equal(Mi.indptr, Mig.indptr) and np.array_equal(Mi.indices, Mig.indices)
    np.testing.assert_allclose(Mi.data, Mig.data, rtol=M_RTOL, atol=0)
    with pytest.raises(ValueError):
        build_kernel(X32, 15, "bogus")


def test_knn_vs_oracle_ragged_and_duplicates():
    """Sizes that are not multiples of the tile, d > one shared-memory chunk, duplicated points (zero distances)."""
    from seacells_b200 import build_kernel, knn
    rng = np.random.default_rng(5)
    for n, d, kn in ((257, 7, 5), (1000, 70, 15), (333, 130, 33)):
        X = rng.normal(size=(n, d)).astype(np.float32)
        X[10] = X[3]
        X[11] = X[3]          # duplicates: distance exactly 0, dropped from the graph like scanpy's eliminate_zeros
        X[n - 1] = X[0]
        idx, d2 = knn(X, kn)
        io, do = O.knn_exact(X, kn)
        assert np.array_equal(idx, io) and np.array_equal(d2, do)
        part_i, part_d = knn(X, kn, row_begin=100, row_end=201)  # a rank's query slab
        assert np.array_equal(part_i, io[100:201]) and np.array_equal(part_d, do[100:201])
        M, sigma, _, _ = build_kernel(X, kn).export()
        Mo, so, _, _ = O.build_kernel(X.astype(np.float64), kn, knn=(io, do))
        assert np.array_equal(sigma, so)
        assert np.array_equal(M.indptr, Mo.indptr) and np.array_equal(M.indices, Mo.indices)
        np.testing.assert_allclose(M.data, Mo.data, rtol=M_RTOL, atol=0)


def test_knn_filter_path_is_exact():
    """n >= 2048 takes the fp32-filter + fp64 re-rank path: it must return exactly what the all-fp64 kernel returns."""
    import subprocess
    import sys
    from seacells_b200 import _lib, knn
    from seacells_b200.synth import synth_embedding
    rng = np.random.default_rng(11)
    ctx = _lib.default_context()
    # (a) against the numpy oracle at a size it can still do
    X = synth_embedding(6000, 50, 3)
    idx, d2 = knn(X, 15)
    io, do = O.knn_exact(X, 15)
    assert np.array_equal(idx, io) and np.array_equal(d2, do)
    assert ctx.lib.sc_knn_fallback_rows(ctx.h) < 60          # the certificate passes almost everywhere
    # (b) adversarial: tight clusters far from the origin (cancellation), duplicates, a far outlier, float64 input
    Xa = (rng.normal(size=(5000, 20)) * 1e-3 + rng.integers(0, 5, size=(5000, 1)) * 50.0 + 1000.0).astype(np.float32)
    Xa[100:110] = Xa[99]
    Xa[4999] = 1e4
    ia, da = knn(Xa, 15)
    ioa, doa = O.knn_exact(Xa, 15)
    assert np.array_equal(ia, ioa) and np.array_equal(da, doa)
    Xd = rng.normal(size=(4000, 33))
    i64, d64 = knn(Xd, 10)
    io64, do64 = O.knn_exact(Xd, 10)
    assert np.array_equal(i64, io64) and np.array_equal(d64, do64)
    # (c) large: filter path vs the all-fp64 kernel (forced through the environment in a child process)
    n = 60000
    Xl = synth_embedding(n, 50, 0)
    il, dl = knn(Xl, 15)
    np.save("/tmp/_knn_X.npy", Xl)
    code = ("import numpy as np, sys; sys.path.insert(0, %r); from seacells_b200 import knn; "
            "X = np.load('/tmp/_knn_X.npy'); i, d = knn(X, 15); np.save('/tmp/_knn_i.npy', i); np.save('/tmp/_knn_d.npy', d)"
            % __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
    subprocess.run([sys.executable, "-c", code], check=True, env=dict(__import__("os").environ, SC_KNN_EXACT="1"))
    assert np.array_equal(il, np.load("/tmp/_knn_i.npy")) and np.array_equal(dl, np.load("/tmp/_knn_d.npy"))


def test_knn_tensor_core_filter():
    """tcgen05 shortlist stage (d <= 127): realised error of the 2xFP16 split product against the bound the certificate
    uses, on benign and on badly scaled data; the FFMA filter (d > 127) and the tensor-core filter both stay exact."""
    from seacells_b200 import _lib, knn
    from seacells_b200.synth import synth_embedding
    ctx = _lib.default_context()
    rng = np.random.default_rng(17)
    cases = [synth_embedding(4096, 50, 0), synth_embedding(2048 + 77, 30, 1),
             (rng.normal(size=(3000, 63)) * np.logspace(-3, 3, 63)).astype(np.float32),
             (rng.normal(size=(2500, 8)) * 1e-2 + rng.integers(0, 3, size=(2500, 1)) * 300.0).astype(np.float32)]
    for X in cases:
        n, d = X.shape
        cols = n
        s = np.zeros((128, cols), np.float32)
        xc = np.zeros((n, d), np.float32)
        nn = np.zeros(n, np.float32)
        ctx.check(ctx.lib.sc_knn_tc_probe(ctx.h, _lib.ptr(X), 0, n, d, cols, _lib.ptr(s), _lib.ptr(xc), _lib.ptr(nn)),
                  "sc_knn_tc_probe")
        x64 = xc.astype(np.float64)
        ref = x64[:128] @ x64.T - 0.5 * nn.astype(np.float64)[None, :]
        nr = np.sqrt((x64 ** 2).sum(1))
        rel = 2.0 * np.abs(s - ref) / np.maximum((nr[:128, None] + nr[None, :]) ** 2, 1e-300)
        bound = ctx.lib.sc_knn_tc_error_factor(d)
        assert rel.max() < 0.25 * bound, (rel.max(), bound)   # measured: ~0.005 of the bound
        idx, d2 = knn(X, 15)
        io, do = O.knn_exact(X, 15)
        assert np.array_equal(idx, io) and np.array_equal(d2, do)
    # 63 < d <= 127: still the tensor-core filter (KP up to 128); d > 127: fp32 FFMA tile filter
    for d in (70, 127, 130):
        X = rng.normal(size=(2500, d)).astype(np.float32)
        idx, d2 = knn(X, 15)
        io, do = O.knn_exact(X, 15)
        assert np.array_equal(idx, io) and np.array_equal(d2, do)
    # values far outside fp16's range (the pre-scale is a power of two taken from the largest row norm)
    for s in (1e-6, 1e6):
        X = (synth_embedding(3000, 20, 2) * s).astype(np.float32)
        idx, d2 = knn(X, 15)
        io, do = O.knn_exact(X, 15)
        assert np.array_equal(idx, io) and np.array_equal(d2, do)
        assert ctx.lib.sc_knn_fallback_rows(ctx.h) == 0
    # query slab (a rank's shard) through the tensor-core filter, not a multiple of the tile
    X = synth_embedding(5000, 50, 4)
    io, do = O.knn_exact(X, 15)
    pi, pd = knn(X, 15, row_begin=1234, row_end=1234 + 1300)
    assert np.array_equal(pi, io[1234:2534]) and np.array_equal(pd, do[1234:2534])


def test_knn_tensor_core_filter_degenerate_inputs():
    """Low-dimensional embeddings, more exact duplicates than the shortlist holds (every tie at the threshold must fall
    back to the exact kernel), a constant column, and n that is not a multiple of the 128-row tiles."""
    from seacells_b200 import _lib, knn
    ctx = _lib.default_context()
    rng = np.random.default_rng(23)
    X = rng.normal(size=(3001, 2)).astype(np.float32)
    X[100:150] = X[99]                      # 51 identical points: their 14 nearest neighbours are all at distance 0
    idx, d2 = knn(X, 15)
    io, do = O.knn_exact(X, 15)
    assert np.array_equal(idx, io) and np.array_equal(d2, do)
    assert ctx.lib.sc_knn_fallback_rows(ctx.h) >= 51
    X = rng.normal(size=(2049, 5)).astype(np.float32)
    X[:, 3] = 7.25                          # constant column (centred to exactly 0)
    idx, d2 = knn(X, 15)
    io, do = O.knn_exact(X, 15)
    assert np.array_equal(idx, io) and np.array_equal(d2, do)
    for first in (1.0, 0.0):                # (almost) all points identical: nothing but ties; largest norm exactly 0
        X = np.zeros((2100, 4), np.float32)
        X[0] = first
        idx, d2 = knn(X, 6)
        io, do = O.knn_exact(X, 6)
        assert np.array_equal(idx, io) and np.array_equal(d2, do)


def test_spmm_csr_dense():
    from seacells_b200 import spmm_csr
    rng = np.random.default_rng(0)
    for r, m, c in ((300, 300, 7), (1000, 800, 256), (513, 700, 333)):
        S = sp.random(r, m, density=0.02, random_state=1, format="csr")
        X = rng.normal(size=(m, c))
        np.testing.assert_allclose(spmm_csr(S, X), S @ X, rtol=1e-12, atol=1e-13)


# ------------------------------------------------------------------------------------------------ hot path 2
def _run_and_compare(g, dense):
    M = csr_from(g, "M")
    n, k, T = M.shape[0], int(g["k"]), int(g["T"])
    l2 = float(g["l2"])
    eng = _engine(n, k, T)
    eng.set_kernel(M)
    eng.set_archetypes(g["archetypes"])
    A0 = csr_from(g, "A0n") if dense else O.l1_normalise_columns(csr_from(g, "A0"))
    eng.set_A(A0)
    eng.update_A(l2)
    rss = [eng.rss()]
    assert same_sparse(eng.A(), csr_from(g, "A_step0")), "A after initialize"
    assert same_sparse(eng.B(), csr_from(g, "B_step0"))
    thr = float(g["eps"]) * rss[0]
    converged, it = False, 0
    while (not converged and it < int(g["max_iter"])) or it < int(g["min_iter"]):
        it += 1
        eng.update_A(l2)
        eng.update_B()
        rss.append(eng.rss())
        if it in (1, 2):
            assert same_sparse(eng.A(), csr_from(g, f"A_step{it}")), f"A after step {it}"
            assert same_sparse(eng.B(), csr_from(g, f"B_step{it}")), f"B after step {it}"
        if abs(rss[-2] - rss[-1]) < thr:
            converged = True
    assert it == int(g["n_iter"]) and converged == bool(g["converged"])
    assert same_sparse(eng.A(), csr_from(g, "A_final"))
    assert same_sparse(eng.B(), csr_from(g, "B_final"))
    np.testing.assert_allclose(rss, g["RSS_iters"], rtol=RSS_RTOL)
    assert np.array_equal(eng.hard_assignments(), g["labels"])
    assert np.array_equal(eng.hard_archetypes(), g["hard_archetypes"])
    si, sw = eng.soft_assignments(5)
    arch = eng.hard_archetypes()
    assert np.array_equal(arch[si], g["soft_idx"]) and np.array_equal(sw, g["soft_w"])
    return eng


@pytest.mark.parametrize("name", CASES)
def test_fit_vs_reference_golden(name):
    g = load_golden(name)
    _run_and_compare(g, bool(g["dense"]))


def test_fit_run_loop_in_library():
    """sc_fit_run (initialize + _fit inside the library) gives the same iterates as stepping by hand."""
    g = load_golden("rna_n600_k8_sparse")
    M = csr_from(g, "M")
    eng = _engine(M.shape[0], int(g["k"]))
    eng.set_kernel(M)
    eng.set_archetypes(g["archetypes"])
    eng.set_A(O.l1_normalise_columns(csr_from(g, "A0")))
    rss, n_iter, conv, thr = eng.run(max_iter=int(g["max_iter"]), min_iter=int(g["min_iter"]), eps=float(g["eps"]))
    assert n_iter == int(g["n_iter"]) and conv == bool(g["converged"])
    np.testing.assert_allclose(rss, g["RSS_iters"], rtol=RSS_RTOL)
    np.testing.assert_allclose(thr, float(g["threshold"]), rtol=RSS_RTOL)
    assert same_sparse(eng.A(), csr_from(g, "A_final")) and same_sparse(eng.B(), csr_from(g, "B_final"))


def _seeded_case(n, k, seed=0):
    from seacells_b200.synth import synth_archetypes, synth_assignments, synth_embedding
    X = synth_embedding(n, 50, seed)
    Mo, *_ = O.build_kernel(X.astype(np.float64), 15)
    return X, Mo, synth_archetypes(n, k), synth_assignments(n, k)


def test_argmin_traces_vs_oracle():
    """Every Frank-Wolfe argmin of _updateA and _updateB equals the oracle's (cpu.py:450, 487)."""
    n, k = 1800, 24
    X, M, arch, A0 = _seeded_case(n, k, seed=3)
    K = M @ M.T
    B = sp.csr_matrix((np.ones(k), (arch, np.arange(k))), shape=(n, k))
    A = O.l1_normalise_columns(A0)
    eng = _engine(n, k)
    eng.set_kernel(M)
    eng.set_archetypes(arch)
    eng.set_A(A)
    for outer in range(3):
        trA, trB = [], []
        A = O.update_A(K, B, A, 50, trace=trA)
        gA = eng.update_A(0.0, trace=True)
        assert np.array_equal(gA, np.array(trA)), f"_updateA argmin sequence, outer {outer}"
        assert same_sparse(eng.A(), A)
        B = O.update_B(K, A, B, 50, trace=trB)
        gB = eng.update_B(trace=True)
        assert np.array_equal(gB, np.array(trB)), f"_updateB argmin sequence, outer {outer}"
        assert same_sparse(eng.B(), B)
        np.testing.assert_allclose(eng.rss(), O.compute_RSS(M, A, B), rtol=RSS_RTOL)


def test_update_A_dense_column_rule_beyond_first_batch():
    """_updateA's implicit-zero rule when the first exact zero of the gradient column lies beyond archetype 31: the
    first 40 archetypes are 40 nearly coincident cells (mutually K-adjacent), so for the cells around them every one
    of G[0..39] is non-zero.  The kernel keeps running products only for archetypes 0..31 and must continue the scan
    with the full evaluation; every argmin has to equal the oracle's."""
    rng = np.random.default_rng(0)
    n, k = 900, 70
    X = rng.normal(size=(n, 3)).astype(np.float32)
    X[:45] = (rng.normal(size=(45, 3)) * 1e-3 + 5.0).astype(np.float32)
    X[45:120] = (rng.normal(size=(75, 3)) * 0.05 + 5.0).astype(np.float32)
    M, *_ = O.build_kernel(X.astype(np.float64), 15, "union")
    K = sp.csr_matrix(M @ M.T)
    arch = np.concatenate([np.arange(40), np.sort(rng.choice(np.arange(120, n), k - 40, replace=False))])
    B = sp.csr_matrix((np.ones(k), (arch, np.arange(k))), shape=(n, k))
    A = O.l1_normalise_columns(sp.csc_matrix((np.ones(n), (rng.integers(0, k, n), np.arange(n))), shape=(k, n)))
    # the situation really occurs: non-negative gradient columns whose first zero is at index >= 32
    t1 = np.asarray(((K @ B).T @ B).todense())
    G = 2 * (t1 @ np.asarray(A.todense()) - np.asarray((K @ B).T.todense()))
    first_zero = np.array([np.argmax(G[:, j] == 0) if (G[:, j] == 0).any() else -1 for j in range(n)])
    assert ((G.min(0) >= 0) & (first_zero >= 32)).sum() >= 10
    eng = _engine(n, k)
    eng.set_kernel(M)
    eng.set_archetypes(arch)
    eng.set_A(A)
    for outer in range(2):
        trA, trB = [], []
        A = O.update_A(K, B, A, 50, trace=trA)
        gA = eng.update_A(0.0, trace=True)
        assert np.array_equal(gA, np.array(trA)), f"_updateA argmin sequence, outer {outer}"
        assert same_sparse(eng.A(), A)
        B = O.update_B(K, A, B, 50, trace=trB)
        gB = eng.update_B(trace=True)
        assert np.array_equal(gB, np.array(trB)), f"_updateB argmin sequence, outer {outer}"
        assert same_sparse(eng.B(), B)


def test_isolated_cells_and_dead_archetypes():
    """Cells with no archetype within two hops (empty t2 column -> implicit-zero rule) and archetypes nobody uses."""
    n, k = 1200, 6
    X, M, _, A0 = _seeded_case(n, k, seed=7)
    arch = np.array([0, 1, 2, 3, 4, 5])  # crowded initialisation: most cells see no archetype at all
    K = M @ M.T
    B = sp.csr_matrix((np.ones(k), (arch, np.arange(k))), shape=(n, k))
    A = O.l1_normalise_columns(A0)
    eng = _engine(n, k)
    eng.set_kernel(M)
    eng.set_archetypes(arch)
    eng.set_A(A)
    for outer in range(2):
        trA, trB = [], []
        A = O.update_A(K, B, A, 50, trace=trA)
        assert np.array_equal(eng.update_A(0.0, trace=True), np.array(trA))
        assert same_sparse(eng.A(), A)
        B = O.update_B(K, A, B, 50, trace=trB)
        assert np.array_equal(eng.update_B(trace=True), np.array(trB))
        assert same_sparse(eng.B(), B)
    # the same inputs through the generic kernel (l2 path with l2 -> tiny would change results; use the dense flavour)
    eng2 = _engine(n, k)
    eng2.set_kernel(M)
    eng2.set_archetypes(arch)
    eng2.set_A(O.l1_normalise_columns(A0).toarray())   # dense A0: every cell starts with k active entries
    A2 = O.update_A_dense(K, sp.csr_matrix((np.ones(k), (arch, np.arange(k))), shape=(n, k)).toarray(),
                          O.l1_normalise_columns(A0).toarray(), 50)
    eng2.update_A(0.0)
    assert same_sparse(eng2.A(), sp.csr_matrix(A2))


@pytest.mark.slow
def test_full_fit_vs_oracle_n5000():
    n, k = 5000, 66
    X, M, arch, A0 = _seeded_case(n, k, seed=0)
    ref = O.fit(M, arch, A0, max_iter=30, min_iter=10)
    eng = _engine(n, k)
    eng.set_kernel(M)
    eng.set_archetypes(arch)
    eng.set_A(O.l1_normalise_columns(A0))
    rss, n_iter, conv, _ = eng.run(max_iter=30, min_iter=10)
    assert n_iter == ref["n_iter"] and conv == ref["converged"]
    np.testing.assert_allclose(rss, ref["RSS_iters"], rtol=RSS_RTOL)
    assert max_rel_diff(eng.A(), ref["A"]) <= 1e-6 and max_rel_diff(eng.B(), ref["B"]) <= 1e-6
    assert same_sparse(eng.A(), ref["A"]) and same_sparse(eng.B(), ref["B"])
    assert np.array_equal(eng.hard_assignments(), ref["labels"])


def test_model_class_end_to_end():
    """The reference-facing class: construct_kernel_matrix -> fit -> accessors, against the oracle."""
    from seacells_b200 import SEACells
    from seacells_b200.synth import MiniAnnData, synth_archetypes, synth_assignments, synth_embedding
    n, k = 1500, 20
    X = synth_embedding(n, 50, 1)
    arch, A0 = synth_archetypes(n, k), synth_assignments(n, k)
    ad = MiniAnnData(X)
    model = SEACells(ad, "X_pca", k, verbose=False, use_sparse=True)
    with pytest.raises(RuntimeError):
        model.initialize(initial_archetypes=arch, initial_assignments=A0)   # no kernel yet (cpu.py:219)
    with pytest.raises(AssertionError):
        model.add_precomputed_kernel_matrix(sp.identity(n - 1, format="csr"))
    model.construct_kernel_matrix()
    with pytest.raises(ValueError):
        model.fit(max_iter=3, min_iter=10)                                   # cpu.py:669
    Mo, *_ = O.build_kernel(X.astype(np.float64), 15)
    assert np.array_equal(model.kernel_matrix.indices, Mo.indices)
    ref = O.fit(model.kernel_matrix, arch, A0, max_iter=4, min_iter=4)
    try:
        model.fit(max_iter=4, min_iter=4, initial_archetypes=arch, initial_assignments=A0)
        raised = False
    except RuntimeWarning:
        raised = True
    assert raised == (not ref["converged"])                                  # cpu.py:647-650
    np.testing.assert_allclose(model.RSS_iters, ref["RSS_iters"], rtol=RSS_RTOL)
    assert same_sparse(model.A_, ref["A"]) and same_sparse(model.B_, ref["B"])
    labels = model.get_hard_assignments()
    assert list(labels["SEACell"]) == [f"SEACell-{i}" for i in ref["labels"]]
    assert list(ad.obs["SEACell"]) == list(labels["SEACell"])
    si, sw = O.soft_assignments(ref["A"], ref["B"])
    soft_labels, soft_w = model.get_soft_assignments()
    assert np.array_equal(soft_w, sw)
    assert list(soft_labels.values[:, 0]) == list(ad.obs_names[si[:, 0]])
    np.testing.assert_allclose(model.compute_RSS(), ref["RSS_iters"][-1], rtol=RSS_RTOL)
    Z = model.get_archetype_matrix()
    K = model.kernel_matrix @ model.kernel_matrix.T
    Zo = (ref["B"].T @ K)                                                    # cpu.py:641
    assert Z.shape == (k, n) and abs(Z - Zo).max() <= 1e-12 * abs(Zo).max()
    # dense flavour (cpu_dense.py): ndarrays out, l2_penalty honoured
    ad2 = MiniAnnData(X)
    md = SEACells(ad2, "X_pca", k, verbose=False, use_sparse=False, l2_penalty=0.02)
    md.add_precomputed_kernel_matrix(model.kernel_matrix)
    A0d = O.l1_normalise_columns(A0).toarray()
    refd = O.fit(model.kernel_matrix, arch, A0d, max_iter=2, min_iter=2, dense=True, l2_penalty=0.02)
    try:
        md.fit(max_iter=2, min_iter=2, initial_archetypes=arch, initial_assignments=A0d)
    except RuntimeWarning:
        pass
    assert isinstance(md.A_, np.ndarray) and np.array_equal(md.A_, refd["A"]) and np.array_equal(md.B_, refd["B"])


def test_properties_at_scale():
    """Size-independent properties at a size the oracle's n x n RSS cannot reach quickly."""
    from seacells_b200 import FitEngine, build_kernel
    from seacells_b200.synth import synth_archetypes, synth_assignments, synth_embedding
    n, k = 40000, 533
    X = synth_embedding(n, 50, 0)
    km = build_kernel(X, 15)
    M, sigma, idx, d2 = km.export()
    assert abs(M - M.T).max() == 0.0 and np.all(M.diagonal() == 1.0) and np.all(np.diff(M.indptr) >= 15)
    assert np.all(np.diff(d2, axis=1) >= 0) and np.array_equal(sigma, np.sqrt(d2[:, 6]))
    eng = FitEngine(n, k)
    eng.set_kernel(km)
    eng.set_archetypes(synth_archetypes(n, k))
    eng.set_A(O.l1_normalise_columns(synth_assignments(n, k)))
    rss, n_iter, conv, _ = eng.run(max_iter=3, min_iter=3)
    A, B = eng.A(), eng.B()
    assert A.min() >= 0 and B.min() >= 0
    np.testing.assert_allclose(np.asarray(A.sum(0)).ravel(), 1.0, atol=4e-15)
    np.testing.assert_allclose(np.asarray(B.sum(0)).ravel(), 1.0, atol=4e-15)
    assert np.abs(A.data * 1275 - np.round(A.data * 1275)).max() < 1e-9     # FW weights are multiples of 1/1275
    assert np.diff(A.tocsc().indptr).max() <= 50 and np.diff(B.tocsc().indptr).max() <= 50
    np.testing.assert_allclose(rss[-1], O.rss_trace_identity(M, A, B), rtol=1e-9)
    assert all(rss[i + 1] <= rss[i] * (1 + 1e-9) for i in range(len(rss) - 1))
    # determinism: a second engine on the same inputs reproduces A and B bit for bit
    eng2 = FitEngine(n, k)
    eng2.set_kernel(M)
    eng2.set_archetypes(synth_archetypes(n, k))
    eng2.set_A(O.l1_normalise_columns(synth_assignments(n, k)))
    rss2, *_ = eng2.run(max_iter=3, min_iter=3)
    assert rss2 == rss and same_sparse(eng2.A(), A) and same_sparse(eng2.B(), B)


def test_greedy_initialisation_vs_reference():
    """cpu.py:342-423 on the GPU: same centres as the unmodified reference (golden) and as the oracle."""
    from seacells_b200 import SEACells
    from seacells_b200.greedy import greedy_centers_from_kernel
    from seacells_b200.synth import MiniAnnData
    g = load_golden("rna_n600_k8_sparse")
    M = csr_from(g, "M")
    c = greedy_centers_from_kernel(M, len(g["greedy_centers"]))
    assert np.array_equal(c, g["greedy_centers"])
    X, Mo, _, _ = _seeded_case(2500, 33, seed=4)
    assert np.array_equal(greedy_centers_from_kernel(Mo, 43), O.greedy_centers(Mo @ Mo.T, 43))
    # through the model class: waypoint_proportion = 0 -> greedy only (cpu.py:176, 188-201)
    ad = MiniAnnData(X)
    model = SEACells(ad, "X_pca", 33, verbose=False, use_sparse=True)
    model.add_precomputed_kernel_matrix(Mo)
    model.waypoint_proportion = 0
    model.initialize_archetypes()
    assert np.array_equal(model.archetypes, O.greedy_centers(Mo @ Mo.T, 43)[:33])


def test_many_candidates_and_wide_rows():
    """Paths that only dense neighbourhoods with many archetypes reach: cells with > 128 candidate archetypes (generic
    _updateA kernel) and K*B rows with > 448 distinct archetypes (dense-accumulator rows of the sparse product)."""
    from seacells_b200.synth import synth_archetypes, synth_assignments
    n, k, d = 1400, 900, 12
    X = np.random.default_rng(1).normal(size=(n, d)).astype(np.float32)     # one blob: K has ~280 entries per row
    M, *_ = O.build_kernel(X.astype(np.float64), 15)
    arch, A0 = synth_archetypes(n, k), synth_assignments(n, k)
    K = M @ M.T
    B = sp.csr_matrix((np.ones(k), (arch, np.arange(k))), shape=(n, k))
    A = O.l1_normalise_columns(A0)
    eng = _engine(n, k)
    eng.set_kernel(M)
    eng.set_archetypes(arch)
    eng.set_A(A)
    A = O.update_A(K, B, A, 50)
    eng.update_A(0.0)
    assert eng.stats()["slow_cells_last_update_A"] > 1000
    assert same_sparse(eng.A(), A)
    B = O.update_B(K, A, B, 50)
    eng.update_B()
    assert same_sparse(eng.B(), B)
    A = O.update_A(K, B, A, 50)
    eng.update_A(0.0)
    assert same_sparse(eng.A(), A)
    np.testing.assert_allclose(eng.rss(), O.compute_RSS(M, A, B), rtol=RSS_RTOL)
    assert eng.stats()["dense_rows_spgemm_total"] > 0


@pytest.mark.slow
def test_config1_10k_cells_two_outer_iterations():
    """BASELINE config 1 size (10k cells, k=133): reaches the hub rows of the Gram kernel (> 8192 cells per row) and
    the hub block of _updateB; A and B must still equal the oracle's bit for bit."""
    n, k = 10000, 133
    X, M, arch, A0 = _seeded_case(n, k, seed=0)
    ref = O.fit(M, arch, A0, max_iter=2, min_iter=2)
    eng = _engine(n, k)
    eng.set_kernel(M)
    eng.set_archetypes(arch)
    eng.set_A(O.l1_normalise_columns(A0))
    rss, n_iter, conv, _ = eng.run(max_iter=2, min_iter=2)
    assert eng.stats()["hub_archetypes_last_update_B"] > 0
    np.testing.assert_allclose(rss, ref["RSS_iters"], rtol=RSS_RTOL)
    assert same_sparse(eng.A(), ref["A"]) and same_sparse(eng.B(), ref["B"])
    assert np.array_equal(eng.hard_assignments(), ref["labels"])
