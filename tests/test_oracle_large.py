"""The config-size fixtures (tests/golden/large_*.npz) are what the GPU parity tests trust at 10k - 100k cells: check on
the CPU that they are self-consistent and that the oracle still reproduces them (the first _updateA call of config 1)."""
import hashlib
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import seacells_oracle as O
from seacells_b200.synth import synth_archetypes, synth_assignments, synth_embedding
from tests.helpers import GOLDEN

sys.path.insert(0, GOLDEN)


@pytest.mark.parametrize("name,pinned", [("c1", True), ("c3s", True), ("c2", False)])
def test_large_fixture_is_well_formed(name, pinned):
    g = np.load(os.path.join(GOLDEN, f"large_{name}.npz"))
    n, k, T, iters, block = int(g["n"]), int(g["k"]), int(g["T"]), int(g["iters"]), int(g["block"])
    nblocks = (n + block - 1) // block
    assert g["updA_block_digests"].shape == (iters + 1, T, nblocks)
    assert g["updB_traces"].shape == (iters, T, k)
    assert g["RSS_iters"].shape == (iters + 1,) and np.all(np.diff(g["RSS_iters"]) < 0)
    assert g["labels"].shape == (n,) and g["labels"].max() < k
    assert bool(g["pinned_by_reference"]) == pinned
    if pinned:   # the unmodified reference's explicit n x n norm agrees with the trace identity used above 20k cells
        np.testing.assert_allclose(g["RSS_iters_reference"], g["RSS_iters"], rtol=1e-12)
    # upper triangle of a symmetric matrix with unit diagonal: nnz = 2 * upper - n
    assert 2 * g["M_upper_f32"].shape[0] - n == int(g["M_nnz"])
    B = sp.csc_matrix((g["B_data"], g["B_indices"], g["B_indptr"]), shape=(n, k))
    np.testing.assert_allclose(np.asarray(B.sum(0)).ravel(), 1.0, atol=1e-12)
    assert np.array_equal(O.hard_archetypes(B), g["hard_archetypes"])


def test_oracle_reproduces_config1_first_update():
    import make_golden_large as G
    g = np.load(os.path.join(GOLDEN, "large_c1.npz"))
    n, d, k, T = int(g["n"]), int(g["d"]), int(g["k"]), int(g["T"])
    X = synth_embedding(n, d, 0, str(g["kind"])).astype(np.float64)
    knn = G.knn_exact_parallel(X, int(g["n_neighbors"]), procs=min(8, os.cpu_count() or 1))
    assert hashlib.sha256(knn[0].tobytes()).hexdigest() == str(g["knn_idx_sha256"])
    assert hashlib.sha256(knn[1].tobytes()).hexdigest() == str(g["knn_d2_sha256"])
    Mfull, sigma, *_ = O.build_kernel(X, int(g["n_neighbors"]), "union", knn=knn)
    M = G.quantise_f32(Mfull)
    assert G.structure_digest(M) == str(g["M_structure_sha256"])
    assert np.array_equal(G.upper_values(M), g["M_upper_f32"])
    K = M @ M.T
    B0 = sp.csr_matrix((np.ones(k), (synth_archetypes(n, k), np.arange(k))), shape=(n, k))
    tr = []
    O.update_A(K, B0, O.l1_normalise_columns(synth_assignments(n, k)), T, trace=tr)
    mine = np.stack([G.block_digests(a) for a in tr])
    assert np.array_equal(mine, g["updA_block_digests"][0])
