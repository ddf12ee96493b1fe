"""Parity at BASELINE config sizes against fixtures computed offline by the oracle / the unmodified reference
(tests/golden/make_golden_large.py): config 1 (10k x 50, k=133, full fit), a scATAC-like 20k x 30 case, and config 2
(100k x 50, k=1333, Frank-Wolfe only on an injected kernel matrix, cpu.py:115-128).

Every Frank-Wolfe argmin of every _updateA / _updateB call is compared (cpu.py:450, 487), for the default engine and for
its two switchable summation variants (SC_FIT_FRESH_KB: K B re-evaluated at every step of _updateB; SC_FIT_RESUM_A:
(t1 A) re-summed at every step of _updateA).  A variant that disagrees with a fixture must not be the default.
Tolerances (BASELINE.json north_star): A/B entries and RSS <= 1e-6 relative, hard assignments identical - asserted here as
bit-exact A/B (digests) and RSS <= 1e-9.
"""
import hashlib
import os

import numpy as np
import pytest
import scipy.sparse as sp

from tests.helpers import GOLDEN

pytestmark = pytest.mark.gpu

VARIANTS = {"default": {}, "fresh_kb": {"SC_FIT_FRESH_KB": "1"}, "resum_A": {"SC_FIT_RESUM_A": "1"}}


def _load(name):
    p = os.path.join(GOLDEN, f"large_{name}.npz")
    if not os.path.exists(p):
        pytest.skip(f"{p} not generated")
    return np.load(p)


def _kernel_matrix(g, structure):
    """The fixture's kernel matrix: structure (indptr, indices) from the engine's own build of the same embedding - checked
    against the fixture's digest - and the float32-rounded values the fixture carries for the upper triangle."""
    indptr, indices = structure
    n = int(g["n"])
    h = hashlib.sha256()
    h.update(indptr.astype(np.int64).tobytes())
    h.update(indices.astype(np.int32).tobytes())
    assert h.hexdigest() == str(g["M_structure_sha256"]), "graph structure differs from the oracle's"
    rows = np.repeat(np.arange(n), np.diff(indptr))
    up = indices >= rows
    assert int(up.sum()) == g["M_upper_f32"].shape[0]
    U = sp.csr_matrix((g["M_upper_f32"].astype(np.float64), (rows[up], indices[up])), shape=(n, n))
    M = (U + sp.triu(U, 1).T).tocsr()
    M.sort_indices()
    assert np.array_equal(M.indptr, indptr) and np.array_equal(M.indices, indices)
    return M


def _block_digests(amins, block):
    a = np.ascontiguousarray(amins, dtype=np.int32)
    out = np.empty((a.shape[0] + block - 1) // block, np.uint64)
    for b in range(out.shape[0]):
        out[b] = int.from_bytes(hashlib.blake2b(a[b * block:(b + 1) * block].tobytes(), digest_size=8).digest(), "little")
    return out


def _matrix_digest(A):
    A = sp.csc_matrix(A).copy()
    A.sum_duplicates()
    A.eliminate_zeros()
    A.sort_indices()
    h = hashlib.sha256()
    h.update(A.indptr.astype(np.int64).tobytes())
    h.update(A.indices.astype(np.int32).tobytes())
    h.update(A.data.astype(np.float64).tobytes())
    return h.hexdigest()


_built = {}


def _structure(name, g):
    """Hot path 1 at config size: kNN graph, sigma and graph structure bit-exact against the oracle's digests; values equal
    after rounding to float32 (CUDA exp vs numpy exp: <= 2 ulp in fp64)."""
    if name in _built:
        return _built[name]
    from seacells_b200 import build_kernel
    from seacells_b200.synth import synth_embedding
    X = synth_embedding(int(g["n"]), int(g["d"]), 0, str(g["kind"]))
    km = build_kernel(X, int(g["n_neighbors"]), "union")
    M, sigma, idx, d2 = km.export()
    km.close()
    assert hashlib.sha256(idx.tobytes()).hexdigest() == str(g["knn_idx_sha256"]), "kNN indices differ from the oracle"
    assert hashlib.sha256(d2.tobytes()).hexdigest() == str(g["knn_d2_sha256"]), "kNN distances differ from the oracle"
    assert hashlib.sha256(sigma.tobytes()).hexdigest() == str(g["sigma_sha256"]), "sigma differs from the oracle"
    rows = np.repeat(np.arange(M.shape[0]), np.diff(M.indptr))
    mine = M.data[M.indices >= rows].astype(np.float32)
    assert mine.shape == g["M_upper_f32"].shape
    assert np.mean(mine != g["M_upper_f32"]) < 1e-5  # a 2-ulp fp64 difference can straddle a float32 rounding boundary
    _built[name] = (M.indptr.copy(), M.indices.copy())
    return _built[name]


def _run_case(name, variant, monkeypatch):
    from seacells_b200 import FitEngine
    from seacells_b200.synth import synth_archetypes, synth_assignments
    from oracle.seacells_oracle import l1_normalise_columns
    g = _load(name)
    n, k, T, iters, block = int(g["n"]), int(g["k"]), int(g["T"]), int(g["iters"]), int(g["block"])
    M = _kernel_matrix(g, _structure(name, g))
    for key in ("SC_FIT_FRESH_KB", "SC_FIT_RESUM_A"):
        monkeypatch.delenv(key, raising=False)
    for key, v in VARIANTS[variant].items():
        monkeypatch.setenv(key, v)
    eng = FitEngine(n, k, T)  # the switches are read when the engine is created
    eng.set_kernel(M)
    eng.set_archetypes(synth_archetypes(n, k))
    eng.set_A(l1_normalise_columns(synth_assignments(n, k)))
    digA, trB, rss_ref = g["updA_block_digests"], g["updB_traces"], g["RSS_iters"]
    bad = []  # (call, kind, number of mismatching (step, block) digests or argmins)

    def check_A(call):
        tr = eng.update_A(trace=True)
        mine = np.stack([_block_digests(tr[t], block) for t in range(T)])
        d = int((mine != digA[call]).sum())
        if d:
            first = int(np.argwhere((mine != digA[call]).any(axis=1))[0, 0])
            bad.append((call, "updateA", d, first))

    check_A(0)
    rss = [eng.rss()]
    for it in range(1, iters + 1):
        check_A(it)
        tb = eng.update_B(trace=True)
        d = int((tb != trB[it - 1]).sum())
        if d:
            first = int(np.argwhere((tb != trB[it - 1]).any(axis=1))[0, 0])
            bad.append((it, "updateB", d, first))
        rss.append(eng.rss())
    assert not bad, (f"{name}/{variant}: argmin sequences differ from the oracle: "
                     f"(outer iteration, call, mismatches, first FW step) = {bad}")
    np.testing.assert_allclose(rss, rss_ref, rtol=1e-9, atol=0)
    if "RSS_iters_reference" in g.files:  # the unmodified reference's explicit n x n norm (cpu.py:520-536)
        np.testing.assert_allclose(rss, g["RSS_iters_reference"], rtol=1e-9, atol=0)
    assert np.array_equal(eng.hard_assignments(), g["labels"])
    assert np.array_equal(eng.hard_archetypes(), g["hard_archetypes"])
    assert _matrix_digest(eng.A()) == str(g["A_sha256"]), "A is not bit-identical to the oracle's"
    assert _matrix_digest(eng.B()) == str(g["B_sha256"]), "B is not bit-identical to the oracle's"
    eng.close()


@pytest.mark.parametrize("variant", list(VARIANTS))
def test_config1_10k_full_fit(variant, monkeypatch):
    _run_case("c1", variant, monkeypatch)


@pytest.mark.parametrize("variant", list(VARIANTS))
def test_atac_20k_two_iterations(variant, monkeypatch):
    _run_case("c3s", variant, monkeypatch)


@pytest.mark.parametrize("variant", list(VARIANTS))
def test_config2_100k_fw_only(variant, monkeypatch):
    _run_case("c2", variant, monkeypatch)


def test_config1_through_the_model_class():
    """The same 10k fixture through seacells_b200.SEACells(...).fit(): RSS_iters, labels, convergence bookkeeping."""
    from seacells_b200 import SEACells
    from seacells_b200.synth import MiniAnnData, synth_archetypes, synth_assignments, synth_embedding
    g = _load("c1")
    n, k = int(g["n"]), int(g["k"])
    M = _kernel_matrix(g, _structure("c1", g))
    ad = MiniAnnData(synth_embedding(n, int(g["d"]), 0, "rna"))
    model = SEACells(ad, "X_pca", k, verbose=False, use_sparse=True)
    model.add_precomputed_kernel_matrix(M)
    try:
        model.fit(max_iter=int(g["iters"]), min_iter=int(g["iters"]), initial_archetypes=synth_archetypes(n, k),
                  initial_assignments=synth_assignments(n, k))
    except RuntimeWarning:
        pass
    np.testing.assert_allclose(model.RSS_iters, g["RSS_iters"], rtol=1e-9, atol=0)
    lab = np.array([int(s.split("-")[1]) for s in model.get_hard_assignments()["SEACell"]], dtype=np.int32)
    assert np.array_equal(lab, g["labels"])
    assert _matrix_digest(model.A_) == str(g["A_sha256"]) and _matrix_digest(model.B_) == str(g["B_sha256"])
