"""TEST INFRASTRUCTURE ONLY — CPU restatement (numpy/scipy) of the SEACells hot path.

This is the parity oracle of the B200 engine.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it; the product package
`seacells_b200` never does (it fails loudly when its CUDA library is missing).

Pinning: `tests/golden/*.npz` hold outputs of the *unmodified* reference files
(/root/reference/SEACells/{build_graph,cpu,cpu_dense}.py, v0.3.3) run in the build container through
`oracle/ref_shim.py`; `tests/test_oracle_golden.py` checks every function below against them
(A, B bit-exact; RSS <= 1e-12 relative; M bit-exact for float64 input).  The reference ships no tests,
golden vectors or fixtures of its own, and its kNN is third-party `scanpy.pp.neighbors`
(build_graph.py:125; scanpy>1.8 per setup.py:20, 1.8.2 per environment.yaml:9 - not consistently pinned,
not installed here).  At that boundary parity is therefore anchored on the exact brute-force definition in
`knn_exact` below, *not* on scanpy's approximate graph for n >= 8192: "kNN parity unpinned against scanpy".

Each function cites the reference lines it restates.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
from scipy.sparse.linalg import norm as _spnorm


# --------------------------------------------------------------------------------------------
# Hot path 1: kNN -> adaptive bandwidth -> Gaussian kernel -> symmetric CSR  (build_graph.py:113-189)
# --------------------------------------------------------------------------------------------
def knn_exact(X: np.ndarray, n_neighbors: int, block: int = 512):
    """Stand-in for scanpy.pp.neighbors (build_graph.py:125-126): exact kNN, self excluded by index.

    d2(i,j) = sum_c (x_ic - x_jc)^2 with the inputs cast to fp64, accumulated sequentially over c with a
    separately rounded multiply and add; neighbours ordered by (d2, j) ascending.
    Returns idx [n, kn-1] int32 and d2 [n, kn-1] float64 (squared distances).
    """
    X = np.asarray(X, dtype=np.float64)
    n, d = X.shape
    kk = n_neighbors - 1
    idx = np.empty((n, kk), dtype=np.int32)
    d2o = np.empty((n, kk), dtype=np.float64)
    XT = np.ascontiguousarray(X.T)
    for s in range(0, n, block):
        e = min(n, s + block)
        acc = np.zeros((e - s, n), dtype=np.float64)
        for c in range(d):
            diff = X[s:e, c][:, None] - XT[c][None, :]
            acc += diff * diff
        acc[np.arange(e - s), np.arange(s, e)] = np.inf
        order = np.argsort(acc, axis=1, kind="stable")[:, :kk]
        idx[s:e] = order
        d2o[s:e] = np.take_along_axis(acc, order, axis=1)
    return idx, d2o


def knn_distance_graph(idx: np.ndarray, d2: np.ndarray, n: int) -> sp.csr_matrix:
    """The `ad.obsp['distances']` matrix scanpy would write: Euclidean distances, zeros eliminated."""
    kk = idx.shape[1]
    indptr = np.arange(0, n * kk + 1, kk)
    D = sp.csr_matrix((np.sqrt(d2).ravel(), idx.ravel().astype(np.int64), indptr), shape=(n, n))
    D.eliminate_zeros()
    D.sort_indices()
    return D


def adaptive_bandwidth(D: sp.csr_matrix, n_neighbors: int) -> np.ndarray:
    """sigma_i (build_graph.py:18-34,139-147): with m = n_neighbors // 2, the rank selection
    `argsort(argsort(-row)) == nnz - m` picks the m-th smallest stored (non-zero) distance of row i;
    rows with fewer than m stored distances select nothing and `np.linalg.norm([])` gives 0.
    """
    m = n_neighbors // 2
    n = D.shape[0]
    sig = np.zeros(n, dtype=D.dtype)
    indptr, data = D.indptr, D.data
    for i in range(n):
        row = data[indptr[i]:indptr[i + 1]]
        if m >= 1 and row.size >= m:
            sig[i] = np.partition(row, m - 1)[m - 1]
    return sig


def symmetrise(D: sp.csr_matrix, graph_construction: str = "union") -> sp.csr_matrix:
    """Binarise + self loop + union/intersection (build_graph.py:129-164)."""
    n = D.shape[0]
    G = sp.csr_matrix((np.ones_like(D.data), D.indices.copy(), D.indptr.copy()), shape=D.shape)
    G = (G + sp.identity(n, format="csr")).tocsr()
    G.data[:] = 1.0
    if graph_construction == "union":
        S = ((G + G.T) > 0).astype(float)
    elif graph_construction in ("intersect", "intersection"):
        S = G.multiply(G.T)
    else:
        raise ValueError(
            f"Parameter graph_construction = {graph_construction} is not valid. "
            "Please select `union` or `intersection`")
    S = sp.csr_matrix(S)
    S.eliminate_zeros()
    S.sort_indices()
    return S


def gaussian_kernel(X: np.ndarray, S: sp.csr_matrix, sigma: np.ndarray) -> sp.csr_matrix:
    """M[i,j] = exp(-||x_i-x_j||^2 / (sigma_i sigma_j)) on the edges of S (build_graph.py:37-61,170-189).

    The reference evaluates all n columns and masks; restricting to the edge columns leaves the per-row
    arithmetic (`np.sum(np.square(x_i - x_j), axis=1)` in the dtype of X, numpy's pairwise reduction over d)
    unchanged.  Entries that evaluate to exactly 0 are dropped, as `lil_matrix(masked_row)` does.
    """
    n = X.shape[0]
    indptr, indices = S.indptr, S.indices
    vals = np.empty(indices.shape[0], dtype=np.float64)
    for i in range(n):
        cols = indices[indptr[i]:indptr[i + 1]]
        num = np.sum(np.square(X[i, :] - X[cols]), axis=1, keepdims=False)
        den = sigma[i] * sigma[cols]
        vals[indptr[i]:indptr[i + 1]] = np.exp(-num / den)
    M = sp.csr_matrix((vals, indices.copy(), indptr.copy()), shape=(n, n))
    M.eliminate_zeros()
    M.sort_indices()
    return M


def build_kernel(X: np.ndarray, n_neighbors: int = 15, graph_construction: str = "union", knn=None):
    """SEACellGraph.rbf (build_graph.py:113-189) with the exact kNN stand-in.  Returns (M, sigma, idx, d2)."""
    n = X.shape[0]
    idx, d2 = knn_exact(X, n_neighbors) if knn is None else knn
    D = knn_distance_graph(idx, d2, n)
    sigma = adaptive_bandwidth(D, n_neighbors)
    S = symmetrise(D, graph_construction)
    M = gaussian_kernel(X, S, sigma)
    return M, sigma, idx, d2


# --------------------------------------------------------------------------------------------
# Hot path 2: Frank-Wolfe updates  (cpu.py:425-536 sparse semantics, cpu_dense.py:342-446 dense semantics)
# --------------------------------------------------------------------------------------------
def sparse_argmin_cols(G) -> np.ndarray:
    """Column argmin with scipy's implicit-zero rule, vectorised.

    Restates scipy/sparse/_data.py `_argminmax_axis` (reached from cpu.py:450,487 because
    `2.0 * np.array(csr)` hands the sparse matrix back): per column, over the *stored* entries in row order,
    take the first minimum; if it is < 0 or the column is full, return its row; otherwise return the first
    structurally missing row, except that a stored minimum of exactly 0 returns
    min(position-of-minimum-in-slice, first-missing-row).  Empty columns give 0.
    """
    G = sp.csc_matrix(G)
    G.sum_duplicates()
    G.sort_indices()
    nrow, ncol = G.shape
    indptr, indices, data = G.indptr, G.indices, G.data
    out = np.zeros(ncol, dtype=np.int64)
    lens = np.diff(indptr)
    nz = np.nonzero(lens)[0]
    if nz.size == 0:
        return out
    starts = indptr[nz]
    col_of = np.repeat(np.arange(nz.size), lens[nz])
    pos = np.arange(data.shape[0]) - np.repeat(starts, lens[nz])
    cmin = np.minimum.reduceat(data, starts)
    big = np.iinfo(np.int64).max
    first_min_pos = np.minimum.reduceat(np.where(data == cmin[col_of], pos, big), starts)
    min_row = indices[starts + first_min_pos]
    first_missing = np.minimum.reduceat(np.where(indices != pos, pos, big), starts)
    first_missing = np.where(first_missing == big, lens[nz], first_missing)
    full = lens[nz] == nrow
    res = np.where((cmin < 0) | full, min_row,
                   np.where(cmin == 0, np.minimum(first_min_pos, first_missing), first_missing))
    out[nz] = res
    return out


def _onehot(rows, ncols_total, shape):
    return sp.csr_matrix((np.ones(len(rows)), (rows, np.arange(ncols_total))), shape=shape)


def update_A(K, B, A_prev, max_FW_iter: int = 50, trace: list | None = None):
    """cpu.py:425-459.  K, B, A_prev scipy sparse; returns sparse A (k x n)."""
    n, k = B.shape
    A = sp.csr_matrix(A_prev)
    t2 = (K @ B).T
    t1 = t2 @ B
    for t in range(max_FW_iter):
        G = 2.0 * (t1 @ A - t2)
        amins = sparse_argmin_cols(G)
        if trace is not None:
            trace.append(amins.astype(np.int32))
        e = _onehot(amins, n, A.shape)
        A = A + 2.0 / (t + 2.0) * (e - A)
    return A


def update_B(K, A, B_prev, max_FW_iter: int = 50, trace: list | None = None):
    """cpu.py:461-498."""
    k, n = A.shape
    B = sp.csr_matrix(B_prev)
    t1 = A @ A.T
    t2 = K @ A.T
    for t in range(max_FW_iter):
        G = 2.0 * (K @ B @ t1 - t2)
        amins = sparse_argmin_cols(G)
        if trace is not None:
            trace.append(amins.astype(np.int32))
        e = _onehot(amins, k, B.shape)
        B = B + 2.0 / (t + 2.0) * (e - B)
    return B


def update_A_dense(K, B, A_prev, max_FW_iter: int = 50, l2_penalty: float = 0.0, trace=None):
    """cpu_dense.py:342-372 (dense ndarray A, B; honours l2_penalty)."""
    n, k = B.shape
    A = np.array(A_prev, dtype=np.float64)
    t2 = np.asarray((K @ B).T)
    t1 = t2 @ B
    cols = np.arange(n)
    for t in range(max_FW_iter):
        G = 2.0 * np.array(t1 @ A - t2) - l2_penalty * A
        amins = np.argmin(G, axis=0)
        if trace is not None:
            trace.append(amins.astype(np.int32))
        e = np.zeros((k, n))
        e[amins, cols] = 1.0
        A += 2.0 / (t + 2.0) * (e - A)
    return A


def update_B_dense(K, A, B_prev, max_FW_iter: int = 50, trace=None):
    """cpu_dense.py:374-407."""
    k, n = A.shape
    B = np.array(B_prev, dtype=np.float64)
    t1 = A @ A.T
    t2 = np.asarray(K @ A.T)
    cols = np.arange(k)
    for t in range(max_FW_iter):
        G = 2.0 * np.array(K @ B @ t1 - t2)
        amins = np.argmin(G, axis=0)
        if trace is not None:
            trace.append(amins.astype(np.int32))
        e = np.zeros((n, k))
        e[amins, cols] = 1.0
        B += 2.0 / (t + 2.0) * (e - B)
    return B


def compute_RSS(M, A, B) -> float:
    """cpu.py:500-536 / cpu_dense.py:409-446: || M - (M B) A ||_F (not squared, on M not K)."""
    if sp.issparse(A) or sp.issparse(B):
        R = M - (M.dot(sp.csr_matrix(B))).dot(sp.csr_matrix(A))
        return float(_spnorm(R))
    return float(np.linalg.norm(M - (M.dot(B)).dot(A)))


def rss_trace_identity(M, A, B) -> float:
    """Same quantity without the n x n product (SURVEY.md §7 hard part 5); used for large-n checks only.

    ||M - M B A||_F^2 = ||M||_F^2 - 2 <K B, A^T> + <(B^T K B) A, A>,  K = M M^T.
    """
    A = sp.csr_matrix(A)
    B = sp.csr_matrix(B)
    MB = M @ B
    KB = M.T @ MB
    t1 = (MB.T @ MB)
    m2 = float((M.multiply(M)).sum())
    cross = float(KB.multiply(A.T).sum())
    quad = float((t1 @ A).multiply(A).sum())
    return float(np.sqrt(max(m2 - 2.0 * cross + quad, 0.0)))


def l1_normalise_columns(A0) -> sp.csr_matrix:
    """sklearn.preprocessing.normalize(A0, axis=0, norm='l1') on a CSR matrix (cpu.py:252-253)."""
    from sklearn.preprocessing import normalize
    return normalize(sp.csr_matrix(A0), axis=0, norm="l1")


def hard_assignments(A) -> np.ndarray:
    """cpu.py:710-722 / cpu_dense.py:607-617: first index of the column maximum."""
    if sp.issparse(A):
        return np.asarray(sp.csc_matrix(A).argmax(0)).reshape(-1)
    return np.asarray(A).argmax(0)


def hard_archetypes(B) -> np.ndarray:
    """cpu.py:724-729: B.argmax(0) (positions; the reference maps them through obs_names)."""
    if sp.issparse(B):
        return np.asarray(sp.csc_matrix(B).argmax(0)).reshape(-1)
    return np.asarray(B).argmax(0)


def soft_assignments(A, B, top: int = 5):
    """cpu_dense.py:586-605: `top` rounds of row-argmax over A^T with masking to -1.
    Returns (archetype cell positions [n, top], weights [n, top])."""
    At = np.array(A.T.todense() if sp.issparse(A) else A.T, dtype=np.float64)
    arch = hard_archetypes(B)
    rows = np.arange(At.shape[0])
    labs, wts = [], []
    for _ in range(top):
        l = At.argmax(1)
        labs.append(arch[l])
        wts.append(At[rows, l].copy())
        At[rows, l] = -1
    return np.vstack(labs).T, np.vstack(wts).T


def fit(M, archetypes, initial_assignments, max_iter: int = 100, min_iter: int = 10,
        convergence_epsilon: float = 1e-3, max_FW_iter: int = 50, dense: bool = False,
        l2_penalty: float = 0.0, record=None):
    """initialize + _fit (cpu.py:204-285, 595-651 / cpu_dense.py:145-202, 527-576) for given archetypes and A0.

    Returns dict(A, B, RSS_iters, n_iter, converged, labels, threshold).  `record(step, A, B)` is called
    after initialize (step 0) and after every outer iteration.
    """
    if max_iter < min_iter:
        raise ValueError("The maximum number of iterations specified is lower than the minimum "
                         "number of iterations specified.")
    n = M.shape[0]
    K = M @ M.T
    k = len(archetypes)
    B0 = sp.csr_matrix((np.ones(k), (np.asarray(archetypes), np.arange(k))), shape=(n, k))
    assert initial_assignments.shape == (k, n)
    if dense:
        B = B0.toarray()
        A0 = np.array(initial_assignments.toarray() if sp.issparse(initial_assignments)
                      else initial_assignments, dtype=np.float64)
        A = update_A_dense(K, B, A0.copy(), max_FW_iter, l2_penalty)
        upA = lambda B_, A_: update_A_dense(K, B_, A_, max_FW_iter, l2_penalty)
        upB = lambda A_, B_: update_B_dense(K, A_, B_, max_FW_iter)
    else:
        B = B0.copy()
        A0 = l1_normalise_columns(initial_assignments)
        A = update_A(K, B, A0.copy(), max_FW_iter)
        upA = lambda B_, A_: update_A(K, B_, A_, max_FW_iter)
        upB = lambda A_, B_: update_B(K, A_, B_, max_FW_iter)
    rss = [compute_RSS(M, A, B)]
    thr = convergence_epsilon * rss[0]
    if record is not None:
        record(0, A, B)
    converged, n_iter = False, 0
    while (not converged and n_iter < max_iter) or n_iter < min_iter:
        n_iter += 1
        A = upA(B, A)
        B = upB(A, B)
        rss.append(compute_RSS(M, A, B))
        if record is not None:
            record(n_iter, A, B)
        if abs(rss[-2] - rss[-1]) < thr:
            converged = True
    return dict(A=A, B=B, RSS_iters=rss, n_iter=n_iter, converged=converged,
                labels=hard_assignments(A), threshold=thr)


# --------------------------------------------------------------------------------------------
# "next" row: greedy adaptive CSSP initialisation  (cpu.py:342-423)
# --------------------------------------------------------------------------------------------
def greedy_centers(K, n_centers: int) -> np.ndarray:
    """cpu.py:342-423: greedy column subset selection on K; returns `n_centers` cell positions."""
    n = K.shape[0]
    K = sp.csr_matrix(K)
    f = np.array((K.multiply(K)).sum(axis=0)).ravel()
    g = np.array(K.diagonal()).ravel()
    omega = np.zeros((n_centers, n))
    centers = np.zeros(n_centers, dtype=int)
    Kc = K.tocsc()
    for j in range(n_centers):
        p = int(np.argmax(f / g))
        delta = Kc[:, p].toarray().squeeze() - np.multiply(omega[:, p].reshape(-1, 1), omega).sum(axis=0).squeeze()
        delta[p] = np.max([0, delta[p]])
        o = delta / np.max([np.sqrt(delta[p]), 1e-6])
        o_sq = np.linalg.norm(o) ** 2
        o_had = np.multiply(o, o)
        pl = np.zeros(n)
        for r in range(j):
            pl += np.dot(omega[r, :], o) * omega[r, :]
        Ko = (K @ o.reshape(-1, 1)).ravel()
        f += -2.0 * np.multiply(o, Ko - pl) + o_sq * o_had
        g += o_had
        omega[j, :] = o
        centers[j] = p
    return centers


# ------------------------------------------------------------------------------------------------ metacell summaries
def sparsify_assignments(A, thresh):
    """core.py:103-118 on a dense n_cells x n_SEACells array: zero below `thresh`, renormalise every row."""
    A = np.array(A, dtype=np.float64, copy=True)
    A[A < thresh] = 0
    with np.errstate(divide="ignore", invalid="ignore"):
        A = A / A.sum(1, keepdims=True)
    return A


def summarize_hard(data, labels):
    """core.py:190-230: per SEACell (order of first appearance, core.py:207) the column sums of its cells' rows.

    data: cells x genes (scipy sparse or ndarray); labels: length-n sequence.  Returns (metacells, dense fp64 matrix)."""
    labels = np.asarray(labels)
    _, first = np.unique(labels, return_index=True)
    metacells = labels[np.sort(first)]
    data = sp.csr_matrix(data)
    out = np.zeros((len(metacells), data.shape[1]), dtype=np.float64)
    for r, m in enumerate(metacells):
        out[r, :] = np.ravel(data[np.nonzero(labels == m)[0]].sum(axis=0))  # core.py:215-225
    return metacells, out


def summarize_soft(data, A, minimum_weight=0.05):
    """core.py:121-181: expression[ix] = sum_cells w[cell, ix] * data[cell, :] / sum_cells w[cell, ix] with
    w = sparsify_assignments(A.T, minimum_weight); pseudo-sizes = w.sum(0).  A: n_SEACells x n_cells dense."""
    W = sparsify_assignments(np.asarray(A).T, minimum_weight)
    data = sp.csr_matrix(data)
    expr = []
    for ix in range(W.shape[1]):
        cw = W[:, ix]
        expr.append(data.multiply(cw[:, np.newaxis]).toarray().sum(0) / cw.sum())  # core.py:158-163
    return np.array(expr), W.sum(0)
