"""Generate golden vectors by running the UNMODIFIED reference (build container only).

    python tests/golden/make_golden.py

Loads /root/reference/SEACells/{build_graph,cpu,cpu_dense}.py through oracle/ref_shim.py (stub scanpy/palantir,
exact brute-force kNN stand-in) and stores, per case, the inputs and the reference's outputs at every stage:
kNN graph, sigma (re-derived from the reference's own helper), kernel matrix M (union and intersection),
A/B after `initialize` and after each outer iteration, RSS_iters, hard labels, soft assignments, greedy centres.
The fixtures pin oracle/seacells_oracle.py (tests/test_oracle_golden.py) and the CUDA engine (tests/test_gpu_*.py).
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from seacells_b200.synth import MiniAnnData, synth_archetypes, synth_assignments, synth_embedding  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def _csr_parts(prefix, m):
    m = sp.csr_matrix(m)
    m.sum_duplicates()
    m.sort_indices()
    return {f"{prefix}_indptr": m.indptr.astype(np.int64), f"{prefix}_indices": m.indices.astype(np.int32),
            f"{prefix}_data": m.data.astype(np.float64), f"{prefix}_shape": np.array(m.shape, dtype=np.int64)}


def _quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def make_case(name, n, d, k, kind="rna", key="X_pca", dense=False, l2=0.0, max_iter=30, min_iter=10,
              eps=1e-3, T=50, seed=0, with_greedy=False, nn=15):
    cpu = ref_shim.load("cpu")
    cpu_dense = ref_shim.load("cpu_dense")
    bg = ref_shim.load("build_graph")
    X32 = synth_embedding(n, d, seed, kind)
    X = X32.astype(np.float64)  # identical values; the reference then computes rbf in fp64
    out = {"X32": X32, "n_neighbors": np.int64(nn), "k": np.int64(k), "max_iter": np.int64(max_iter),
           "min_iter": np.int64(min_iter), "eps": np.float64(eps), "T": np.int64(T), "l2": np.float64(l2),
           "dense": np.bool_(dense)}

    ad = MiniAnnData(X, key)
    cls = cpu_dense.SEACellsCPUDense if dense else cpu.SEACellsCPU
    model = _quiet(cls, ad, key, k, False, 10, nn, eps, l2, T)
    _quiet(model.construct_kernel_matrix)
    D = ad.obsp["distances"]
    out.update(_csr_parts("D", D))
    sig = np.array([bg.kth_neighbor_distance(D, nn // 2, i) for i in range(n)])
    out["sigma"] = sig
    out.update(_csr_parts("M", model.kernel_matrix))

    ad2 = MiniAnnData(X, key)
    g2 = bg.SEACellGraph(ad2, key, verbose=False)
    out.update(_csr_parts("Mint", _quiet(g2.rbf, nn, graph_construction="intersection")))

    arch = synth_archetypes(n, k)
    A0 = synth_assignments(n, k)
    out["archetypes"] = arch
    out.update(_csr_parts("A0", A0))
    A0_in = A0.toarray() / A0.toarray().sum(0, keepdims=True) if dense else A0
    if dense:
        out.update(_csr_parts("A0n", sp.csr_matrix(A0_in)))

    _quiet(model.initialize, initial_archetypes=arch, initial_assignments=A0_in)
    out.update(_csr_parts("A_step0", model.A_))
    out.update(_csr_parts("B_step0", model.B_))
    converged, n_iter = False, 0
    while (not converged and n_iter < max_iter) or n_iter < min_iter:  # the loop of cpu.py:620-639, stepped by hand
        n_iter += 1
        _quiet(model.step)
        if n_iter in (1, 2):
            out.update(_csr_parts(f"A_step{n_iter}", model.A_))
            out.update(_csr_parts(f"B_step{n_iter}", model.B_))
        if abs(model.RSS_iters[-2] - model.RSS_iters[-1]) < model.convergence_threshold:
            converged = True
    out.update(_csr_parts("A_final", model.A_))
    out.update(_csr_parts("B_final", model.B_))
    out["RSS_iters"] = np.array(model.RSS_iters, dtype=np.float64)
    out["n_iter"] = np.int64(n_iter)
    out["converged"] = np.bool_(converged)
    out["threshold"] = np.float64(model.convergence_threshold)
    labels = model.get_hard_assignments()["SEACell"]
    out["labels"] = np.array([int(s.split("-")[1]) for s in labels], dtype=np.int32)

    # soft assignments: the dense class holds the well-formed definition (cpu_dense.py:586-605)
    dm = _quiet(cpu_dense.SEACellsCPUDense, MiniAnnData(X, key), key, k, False)
    dm.A_ = np.asarray(sp.csr_matrix(model.A_).toarray())
    dm.B_ = np.asarray(sp.csr_matrix(model.B_).toarray())
    soft_labels, soft_w = dm.get_soft_assignments()
    names = {nm: i for i, nm in enumerate(dm.ad.obs_names)}
    out["soft_idx"] = np.vectorize(names.get)(soft_labels.values).astype(np.int32)
    out["soft_w"] = np.asarray(soft_w, dtype=np.float64)
    out["hard_archetypes"] = np.array([names[x] for x in dm.get_hard_archetypes()], dtype=np.int32)

    if with_greedy:
        gm = _quiet(cpu.SEACellsCPU, MiniAnnData(X, key), key, k, False)
        _quiet(gm.add_precomputed_kernel_matrix, model.kernel_matrix)
        with contextlib.redirect_stderr(io.StringIO()):
            out["greedy_centers"] = np.asarray(_quiet(gm._get_greedy_centers, k + 10), dtype=np.int64)

    path = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: n={n} k={k} dense={dense} l2={l2} n_iter={n_iter} converged={converged} "
          f"RSS0={model.RSS_iters[0]:.6f} RSS_end={model.RSS_iters[-1]:.6f} -> {os.path.getsize(path)/1e3:.0f} kB")


if __name__ == "__main__":
    make_case("rna_n600_k8_sparse", 600, 50, 8, with_greedy=True)
    make_case("rna_n600_k8_dense_l2", 600, 50, 8, dense=True, l2=0.05)
    make_case("atac_n900_k12_sparse", 900, 30, 12, kind="atac", key="X_svd")
    make_case("rna_n2000_k27_sparse", 2000, 50, 27)
    # CPU-only extra (pins the oracle away from the defaults): 10 neighbours, 20 Frank-Wolfe steps, tighter epsilon
    make_case("rna_n500_k9_nn10_T20", 500, 50, 9, nn=10, T=20, eps=1e-4, seed=5)
