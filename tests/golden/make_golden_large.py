"""Golden vectors at BASELINE config sizes (build container only; hours of CPU, run once, outputs committed).

    python tests/golden/make_golden_large.py c1      # 10 000 x 50 PCs,  k = 133,  full fit (initialize + 10 iterations)
    python tests/golden/make_golden_large.py c3s     # 20 000 x 30 SVD comps (scATAC-like), k = 266, initialize + 2 iterations
    python tests/golden/make_golden_large.py c2      # 100 000 x 50 PCs, k = 1333, initialize + 2 iterations (FW only)

What is pinned (SURVEY.md §8d "gate at n in {600, 3000, 10000} full-fit, plus C2 FW-only via injected M"):
  * the kernel matrix is built by the oracle (exact kNN, fp64) and its values are rounded to float32 so that the
    fixture can carry them compactly and bit-exactly (upper triangle, CSR order); that matrix is injected on both sides
    through add_precomputed_kernel_matrix (cpu.py:115-128), so kernel-build ulps cannot leak into the FW comparison;
  * every Frank-Wolfe step's argmin vector of every _updateA / _updateB call (cpu.py:450, 487): _updateB traces in full,
    _updateA traces as 8-byte blake2b digests per (step, block of 2048 cells) - a mismatch is localised to a block;
  * RSS per outer iteration (trace identity; at 10k also the reference's explicit norm), final hard labels, final B,
    digest of the final A.
For c1 and c3s the UNMODIFIED reference (cpu.SEACellsCPU through oracle/ref_shim.py) is run on the same injected matrix
and must agree with the oracle restatement bit for bit (A, B) - that pins the oracle at config size, not only at n <= 2000.
"""
from __future__ import annotations

import contextlib
import hashlib
import io
import multiprocessing as mp
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import seacells_oracle as O  # noqa: E402
from seacells_b200.synth import MiniAnnData, synth_archetypes, synth_assignments, synth_embedding  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
BLOCK = 2048  # cells per digest block of the _updateA traces

CASES = {
    #        n       d   k     kind    iters  run the unmodified reference too
    "c1": (10_000, 50, 133, "rna", 10, True),
    "c3s": (20_000, 30, 266, "atac", 2, True),
    "c2": (100_000, 50, 1333, "rna", 2, False),
}


# ---------------------------------------------------------------------------------------------- exact kNN, faster
_X = None


def _knn_block(args):
    """Same arithmetic as oracle.knn_exact (sequential fp64 accumulation over c, order (d2, j)); the selection uses a
    partition to the 64 smallest and a stable sort of those, with a full-sort fall-back when a tie crosses the cut."""
    s, e, kk = args
    X = _X
    n, d = X.shape
    acc = np.zeros((e - s, n), dtype=np.float64)
    for c in range(d):
        diff = X[s:e, c][:, None] - X[:, c][None, :]
        acc += diff * diff
    acc[np.arange(e - s), np.arange(s, e)] = np.inf
    cut = 64
    part = np.argpartition(acc, cut, axis=1)[:, :cut + 1]
    pv = np.take_along_axis(acc, part, axis=1)
    idx = np.empty((e - s, kk), np.int32)
    d2 = np.empty((e - s, kk), np.float64)
    for r in range(e - s):
        order = np.lexsort((part[r, :cut], pv[r, :cut]))
        sel = part[r, :cut][order]
        if pv[r, :cut][order][kk - 1] >= pv[r].max():  # a tie could cross the cut: do it the slow way
            sel = np.argsort(acc[r], kind="stable")
        idx[r] = sel[:kk]
        d2[r] = acc[r, sel[:kk]]
    return s, idx, d2


def knn_exact_parallel(X, n_neighbors, procs=8, block=256):
    global _X
    _X = np.asarray(X, dtype=np.float64)
    n = _X.shape[0]
    kk = n_neighbors - 1
    idx = np.empty((n, kk), np.int32)
    d2 = np.empty((n, kk), np.float64)
    tasks = [(s, min(n, s + block), kk) for s in range(0, n, block)]
    with mp.get_context("fork").Pool(procs) as pool:
        for s, i, d in pool.imap_unordered(_knn_block, tasks):
            idx[s:s + i.shape[0]] = i
            d2[s:s + i.shape[0]] = d
    return idx, d2


# ---------------------------------------------------------------------------------------------- fixture helpers
def quantise_f32(M):
    """Round the kernel values to float32 (kept as fp64): symmetric entries round identically, the diagonal stays 1."""
    M = sp.csr_matrix(M).copy()
    M.sort_indices()
    M.data = M.data.astype(np.float32).astype(np.float64)
    return M


def upper_values(M):
    rows = np.repeat(np.arange(M.shape[0]), np.diff(M.indptr))
    return M.data[M.indices >= rows].astype(np.float32)


def structure_digest(M):
    h = hashlib.sha256()
    h.update(M.indptr.astype(np.int64).tobytes())
    h.update(M.indices.astype(np.int32).tobytes())
    return h.hexdigest()


def block_digests(amins, block=BLOCK):
    """[n] int -> uint64 [ceil(n/block)]: 8-byte blake2b of the int32 bytes of every block of cells."""
    a = np.ascontiguousarray(amins, dtype=np.int32)
    out = np.empty((a.shape[0] + block - 1) // block, np.uint64)
    for b in range(out.shape[0]):
        out[b] = int.from_bytes(hashlib.blake2b(a[b * block:(b + 1) * block].tobytes(), digest_size=8).digest(), "little")
    return out


def matrix_digest(A):
    """Canonical digest of a sparse matrix: CSC, sorted, duplicates summed, explicit zeros dropped."""
    A = sp.csc_matrix(A).copy()
    A.sum_duplicates()
    A.eliminate_zeros()
    A.sort_indices()
    h = hashlib.sha256()
    h.update(A.indptr.astype(np.int64).tobytes())
    h.update(A.indices.astype(np.int32).tobytes())
    h.update(A.data.astype(np.float64).tobytes())
    return h.hexdigest()


def make(name):
    n, d, k, kind, iters, with_ref = CASES[name]
    T, nn = 50, 15
    t0 = time.time()
    X32 = synth_embedding(n, d, 0, kind)
    X = X32.astype(np.float64)
    knn = knn_exact_parallel(X, nn)
    print(f"[{name}] kNN {time.time() - t0:.0f}s", flush=True)
    Mfull, sigma, idx, d2 = O.build_kernel(X, nn, "union", knn=knn)
    M = quantise_f32(Mfull)
    assert abs(M - M.T).max() == 0.0 and np.all(M.diagonal() == 1.0)
    print(f"[{name}] kernel matrix nnz={M.nnz} {time.time() - t0:.0f}s", flush=True)
    arch = synth_archetypes(n, k)
    A0 = synth_assignments(n, k)
    out = {"n": np.int64(n), "d": np.int64(d), "k": np.int64(k), "kind": np.array(kind), "T": np.int64(T),
           "n_neighbors": np.int64(nn), "iters": np.int64(iters), "block": np.int64(BLOCK),
           "M_upper_f32": upper_values(M), "M_structure_sha256": np.array(structure_digest(M)),
           "M_nnz": np.int64(M.nnz), "knn_idx_sha256": np.array(hashlib.sha256(knn[0].tobytes()).hexdigest()),
           "knn_d2_sha256": np.array(hashlib.sha256(knn[1].tobytes()).hexdigest()),
           "sigma_sha256": np.array(hashlib.sha256(sigma.astype(np.float64).tobytes()).hexdigest())}

    K = M @ M.T
    B = sp.csr_matrix((np.ones(k), (arch, np.arange(k))), shape=(n, k))
    A = O.l1_normalise_columns(A0)
    rss, A_dig, B_tr = [], [], []

    def upd_A(A, B):
        tr = []
        A = O.update_A(K, B, A, T, trace=tr)
        A_dig.append(np.stack([block_digests(a) for a in tr]))
        return A

    A = upd_A(A, B)
    rss.append(O.rss_trace_identity(M, A, B))
    print(f"[{name}] initialize RSS0={rss[0]:.10f} {time.time() - t0:.0f}s", flush=True)
    snap = {}
    for it in range(1, iters + 1):
        A = upd_A(A, B)
        tr = []
        B = O.update_B(K, A, B, T, trace=tr)
        B_tr.append(np.stack(tr).astype(np.int32))
        rss.append(O.rss_trace_identity(M, A, B))
        print(f"[{name}] iteration {it} RSS={rss[-1]:.10f} {time.time() - t0:.0f}s", flush=True)
        if it <= 2:
            snap[f"labels_iter{it}"] = O.hard_assignments(A).astype(np.int32)
    labels = O.hard_assignments(A).astype(np.int32)
    Bc = sp.csc_matrix(B)
    Bc.sort_indices()
    out.update({"updA_block_digests": np.stack(A_dig), "updB_traces": np.stack(B_tr), "RSS_iters": np.array(rss),
                "labels": labels, "A_sha256": np.array(matrix_digest(A)), "B_sha256": np.array(matrix_digest(B)),
                "A_nnz": np.int64(sp.csc_matrix(A).nnz),
                "B_indptr": Bc.indptr.astype(np.int64), "B_indices": Bc.indices.astype(np.int32), "B_data": Bc.data,
                "hard_archetypes": O.hard_archetypes(B).astype(np.int32), **snap})
    if n <= 20_000:
        out["RSS_explicit_final"] = np.float64(O.compute_RSS(M, A, B))  # cpu.py:520-536 as written (n x n product)

    if with_ref:
        from oracle import ref_shim
        cpu = ref_shim.load("cpu")
        with contextlib.redirect_stdout(io.StringIO()):
            model = cpu.SEACellsCPU(MiniAnnData(X, "X_pca"), "X_pca", k, False)
            model.add_precomputed_kernel_matrix(M)
            model.initialize(initial_archetypes=arch, initial_assignments=A0)
            for _ in range(iters):
                model.step()
        assert matrix_digest(model.A_) == str(out["A_sha256"]), "oracle A differs from the unmodified reference"
        assert matrix_digest(model.B_) == str(out["B_sha256"]), "oracle B differs from the unmodified reference"
        ref_rss = np.array(model.RSS_iters)
        assert np.allclose(ref_rss, out["RSS_iters"], rtol=1e-9, atol=0)
        lab = np.array([int(s.split("-")[1]) for s in model.get_hard_assignments()["SEACell"]], dtype=np.int32)
        assert np.array_equal(lab, labels)
        out["RSS_iters_reference"] = ref_rss
        out["pinned_by_reference"] = np.bool_(True)
        print(f"[{name}] unmodified reference agrees bit for bit (A, B, labels); RSS rel diff "
              f"{np.max(np.abs(ref_rss - out['RSS_iters']) / ref_rss):.2e} {time.time() - t0:.0f}s", flush=True)
    else:
        out["pinned_by_reference"] = np.bool_(False)
    path = os.path.join(HERE, f"large_{name}.npz")
    np.savez_compressed(path, **out)
    print(f"[{name}] -> {path} {os.path.getsize(path) / 1e6:.2f} MB, total {time.time() - t0:.0f}s", flush=True)


if __name__ == "__main__":
    for nm in sys.argv[1:] or ["c1"]:
        make(nm)
