"""Host-side mirror of the reference model class, backed by the sm_100a library.

`SEACells(...)` keeps the signature of SEACells.core.SEACells (core.py:13-25) and returns one model class whose
methods, attributes, argument meaning and error behaviour follow cpu.SEACellsCPU / cpu_dense.SEACellsCPUDense
(cpu.py:115-764, cpu_dense.py:86-644).  All numeric work of the two hot paths - kernel construction and the
Frank-Wolfe fit - happens in libseacells_b200.so through ctypes; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import pandas as pd
import scipy.sparse as sp

from . import _lib
from ._lib import SeacellsB200Error, ptr


# ------------------------------------------------------------------------------------------------ thin handles
class KernelMatrix:
    """Device-resident kernel matrix M produced by `sc_kernel_build` (build_graph.py:188 `self.M`)."""

    def __init__(self, ctx, handle, n, nnz, kk):
        self.ctx, self.h, self.n, self.nnz, self.kk = ctx, handle, n, nnz, kk

    def export(self, want_knn=True):
        n, nnz, kk = self.n, self.nnz, self.kk
        indptr = np.empty(n + 1, np.int32)
        indices = np.empty(nnz, np.int32)
        data = np.empty(nnz, np.float64)
        sigma = np.empty(n, np.float64)
        idx = np.empty((n, kk), np.int32) if want_knn else None
        d2 = np.empty((n, kk), np.float64) if want_knn else None
        self.ctx.check(self.ctx.lib.sc_kernel_export(self.ctx.h, self.h, ptr(indptr), ptr(indices), ptr(data),
                                                     ptr(sigma), ptr(idx), ptr(d2)), "sc_kernel_export")
        M = sp.csr_matrix((data, indices, indptr), shape=(n, n))
        return M, sigma, idx, d2

    def close(self):
        if self.h is not None:
            if getattr(self.ctx, "h", None):  # a handle must not outlive its context's stream
                self.ctx.lib.sc_kernel_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _as_embedding(X):
    """numpy float32/float64 (host) or a torch tensor (host or CUDA) -> (object to take the address of, is_f64)."""
    if hasattr(X, "data_ptr"):
        import torch
        if X.dtype not in (torch.float32, torch.float64):
            X = X.to(torch.float32)
        return X.contiguous(), X.dtype == torch.float64
    X = np.asarray(X)
    if X.dtype != np.float64:
        X = X.astype(np.float32, copy=False)
    return np.ascontiguousarray(X), X.dtype == np.float64


def knn(X, n_neighbors=15, ctx=None, row_begin=0, row_end=None):
    """Exact kNN (replacement for scanpy.pp.neighbors at build_graph.py:125).  Returns (idx, d2) as numpy arrays."""
    ctx = ctx or _lib.default_context()
    Xc, f64 = _as_embedding(X)
    n, d = Xc.shape
    row_end = n if row_end is None else row_end
    kk = n_neighbors - 1
    idx = np.empty((row_end - row_begin, kk), np.int32)
    d2 = np.empty((row_end - row_begin, kk), np.float64)
    ctx.check(ctx.lib.sc_knn(ctx.h, ptr(Xc), int(f64), n, d, n_neighbors, row_begin, row_end, ptr(idx), ptr(d2)),
              "sc_knn")
    return idx, d2


def build_kernel(X, n_neighbors=15, graph_construction="union", ctx=None, knn_graph=None) -> KernelMatrix:
    """SEACellGraph.rbf (build_graph.py:113-189) on the GPU."""
    if graph_construction == "union":
        inter = 0
    elif graph_construction in ("intersect", "intersection"):
        inter = 1
    else:
        raise ValueError(
            f"Parameter graph_construction = {graph_construction} is not valid. "
            "Please select `union` or `intersection`")
    ctx = ctx or _lib.default_context()
    Xc, f64 = _as_embedding(X)
    n, d = Xc.shape
    h = C.c_void_p()
    nnz = C.c_longlong(0)
    if knn_graph is None:
        rc = ctx.lib.sc_kernel_build(ctx.h, ptr(Xc), int(f64), n, d, n_neighbors, inter, C.byref(h), C.byref(nnz))
    else:
        idx = np.ascontiguousarray(knn_graph[0], np.int32)
        d2 = np.ascontiguousarray(knn_graph[1], np.float64)
        rc = ctx.lib.sc_kernel_build_from_knn(ctx.h, ptr(Xc), int(f64), n, d, n_neighbors, ptr(idx), ptr(d2), inter,
                                              C.byref(h), C.byref(nnz))
    ctx.check(rc, "sc_kernel_build")
    return KernelMatrix(ctx, h, n, int(nnz.value), n_neighbors - 1)


def spmm_csr(S: sp.csr_matrix, X: np.ndarray, ctx=None) -> np.ndarray:
    """Y = S @ X with S CSR fp64 and X dense fp64 (host buffers in, host buffer out)."""
    ctx = ctx or _lib.default_context()
    S = sp.csr_matrix(S)
    X = np.ascontiguousarray(X, np.float64)
    r, m = S.shape
    c = X.shape[1]
    Y = np.empty((r, c), np.float64)
    indptr, indices = S.indptr.astype(np.int32), S.indices.astype(np.int32)   # keep the temporaries alive
    data = np.ascontiguousarray(S.data, np.float64)
    ctx.check(ctx.lib.sc_spmm_csr_f64(ctx.h, r, m, c, ptr(indptr), ptr(indices), ptr(data), ptr(X), ptr(Y)),
              "sc_spmm_csr_f64")
    return Y


class FitEngine:
    """Device-resident model state (sc_fit): compressed A (k x n) and B (n x k)."""

    MAX_FW_ITER = 64    # FW_SMAX in csrc/fit_kernels.cuh: slots per column of the compressed A / B
    MAX_NEIGHBORS = 33  # shortlist width of the kNN kernels (csrc/kernel_build.cu) + 1

    def __init__(self, n, k, max_FW_iter=50, ctx=None):
        self.ctx = ctx or _lib.default_context()
        self.n, self.k, self.T = int(n), int(k), int(max_FW_iter)
        if not 1 <= self.T <= self.MAX_FW_ITER:
            raise ValueError(f"max_franke_wolfe_iters must be in [1, {self.MAX_FW_ITER}] for the B200 engine "
                             f"(got {self.T}); see INTEGRATION.md, 'Capacity limits'")
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.sc_fit_create(self.ctx.h, self.n, self.k, self.T, C.byref(h)), "sc_fit_create")
        self.h = h
        self.S = int(self.ctx.lib.sc_fit_slots(self.h))

    # -- inputs
    def set_kernel(self, M):
        lib = self.ctx.lib
        if isinstance(M, KernelMatrix):
            self.ctx.check(lib.sc_fit_set_kernel_from(self.h, M.h), "sc_fit_set_kernel_from")
            return
        M = sp.csr_matrix(M)
        if not M.has_sorted_indices:
            M = M.copy()
            M.sort_indices()
        if M.nnz >= 2 ** 31:
            raise ValueError(f"kernel matrix has {M.nnz} stored entries; the engine indexes them with int32 (< 2^31)")
        indptr = M.indptr.astype(np.int32)
        indices = M.indices.astype(np.int32)
        data = np.ascontiguousarray(M.data, np.float64)
        self.ctx.check(lib.sc_fit_set_kernel(self.h, ptr(indptr), ptr(indices), ptr(data), int(M.nnz)),
                       "sc_fit_set_kernel")

    def set_kernel_device(self, indptr, indices, data, nnz):
        """CSR already resident in HBM (torch CUDA tensors: int32, int32, float64)."""
        self.ctx.check(self.ctx.lib.sc_fit_set_kernel(self.h, ptr(indptr), ptr(indices), ptr(data), int(nnz)),
                       "sc_fit_set_kernel")

    def set_archetypes(self, cells):
        cells = np.ascontiguousarray(np.asarray(cells), np.int64)
        if cells.shape != (self.k,):
            raise ValueError(f"expected {self.k} archetypes, got {cells.shape}")
        self.ctx.check(self.ctx.lib.sc_fit_set_archetypes(self.h, ptr(cells)), "sc_fit_set_archetypes")

    def set_A(self, A0):
        """A0: k x n (scipy sparse or ndarray), used as the `A_prev` of the next update_A."""
        A0 = sp.csc_matrix(A0)
        assert A0.shape == (self.k, self.n)
        A0.sort_indices()
        colptr = A0.indptr.astype(np.int64)
        rows = A0.indices.astype(np.int32)
        vals = np.ascontiguousarray(A0.data, np.float64)
        self.ctx.check(self.ctx.lib.sc_fit_set_A_csc(self.h, ptr(colptr), ptr(rows), ptr(vals)), "sc_fit_set_A_csc")

    def set_A_csc_device(self, colptr, rows, vals):
        self.ctx.check(self.ctx.lib.sc_fit_set_A_csc(self.h, ptr(colptr), ptr(rows), ptr(vals)), "sc_fit_set_A_csc")

    def _slots_from(self, Mx, cols):
        """scipy matrix with `cols` columns -> slot arrays (each column must have <= S entries)."""
        Mx = sp.csc_matrix(Mx)
        cnt = np.diff(Mx.indptr).astype(np.int32)
        if cnt.max(initial=0) > self.S:
            raise ValueError(f"a column has more than S={self.S} entries and cannot be stored in compressed form")
        idx = np.zeros((cols, self.S), np.int32)
        val = np.zeros((cols, self.S), np.float64)
        col_of = np.repeat(np.arange(cols), cnt)
        pos = np.arange(Mx.nnz) - np.repeat(Mx.indptr[:-1], cnt)
        idx[col_of, pos] = Mx.indices
        val[col_of, pos] = Mx.data
        return idx, val, cnt

    def set_B(self, B):
        idx, val, cnt = self._slots_from(B, self.k)
        self.ctx.check(self.ctx.lib.sc_fit_set_B_slots(self.h, ptr(idx), ptr(val), ptr(cnt)), "sc_fit_set_B_slots")

    def set_A_compressed(self, A):
        idx, val, cnt = self._slots_from(A, self.n)
        self.ctx.check(self.ctx.lib.sc_fit_set_A_slots(self.h, ptr(idx), ptr(val), ptr(cnt)), "sc_fit_set_A_slots")

    # -- the hot path
    def update_A(self, l2_penalty=0.0, trace=False):
        tr = np.empty((self.T, self.n), np.int32) if trace else None
        self.ctx.check(self.ctx.lib.sc_fit_update_A(self.h, float(l2_penalty), ptr(tr)), "sc_fit_update_A")
        return tr

    def update_B(self, trace=False):
        tr = np.empty((self.T, self.k), np.int32) if trace else None
        self.ctx.check(self.ctx.lib.sc_fit_update_B(self.h, ptr(tr)), "sc_fit_update_B")
        return tr

    def rss(self) -> float:
        out = C.c_double(0.0)
        self.ctx.check(self.ctx.lib.sc_fit_rss(self.h, C.byref(out)), "sc_fit_rss")
        return float(out.value)

    def run(self, max_iter=100, min_iter=10, eps=1e-3, l2_penalty=0.0, threshold=None):
        """initialize + _fit loop entirely inside the library.  Returns (rss_iters, n_iter, converged, threshold)."""
        thr = C.c_double(threshold if threshold is not None else -1.0)
        n_iter, conv, n_rss = C.c_int(0), C.c_int(0), C.c_int(0)
        if max_iter < min_iter:
            raise ValueError("The maximum number of iterations specified is lower than the minimum number of "
                             "iterations specified.")
        cap = max(max_iter, min_iter) + 2
        rss = np.zeros(cap, np.float64)
        self.ctx.check(self.ctx.lib.sc_fit_run(self.h, int(max_iter), int(min_iter), float(eps), float(l2_penalty),
                                               C.byref(thr), C.byref(n_iter), C.byref(conv), ptr(rss), cap,
                                               C.byref(n_rss)), "sc_fit_run")
        return rss[:n_rss.value].tolist(), int(n_iter.value), bool(conv.value), float(thr.value)

    # -- outputs
    def _get(self, fn, cols):
        idx = np.empty((cols, self.S), np.int32)
        val = np.empty((cols, self.S), np.float64)
        cnt = np.empty(cols, np.int32)
        self.ctx.check(fn(self.h, ptr(idx), ptr(val), ptr(cnt)), "sc_fit_get")
        return idx, val, cnt

    @staticmethod
    def _slots_to_csc(idx, val, cnt, nrows):
        cols, S = idx.shape
        mask = np.arange(S)[None, :] < cnt[:, None]
        indptr = np.zeros(cols + 1, np.int64)
        np.cumsum(cnt, out=indptr[1:])
        m = sp.csc_matrix((val[mask], idx[mask], indptr), shape=(nrows, cols))
        m.sort_indices()
        return m

    def A_slots(self):
        return self._get(self.ctx.lib.sc_fit_get_A, self.n)

    def B_slots(self):
        return self._get(self.ctx.lib.sc_fit_get_B, self.k)

    def A(self) -> sp.csc_matrix:
        """A_ (k x n), CSC."""
        return self._slots_to_csc(*self.A_slots(), self.k)

    def B(self) -> sp.csr_matrix:
        """B_ (n x k), CSR."""
        return self._slots_to_csc(*self.B_slots(), self.n).tocsr()

    def KB(self) -> sp.csr_matrix:
        """K @ B_ (n x k) from the device; Z_ = B_.T @ K is its transpose (cpu.py:641)."""
        nnz = C.c_longlong(0)
        self.ctx.check(self.ctx.lib.sc_fit_get_KB(self.h, C.byref(nnz), None, None, None), "sc_fit_get_KB")
        indptr = np.empty(self.n + 1, np.int64)
        indices = np.empty(nnz.value, np.int32)
        data = np.empty(nnz.value, np.float64)
        self.ctx.check(self.ctx.lib.sc_fit_get_KB(self.h, C.byref(nnz), ptr(indptr), ptr(indices), ptr(data)),
                       "sc_fit_get_KB")
        return sp.csr_matrix((data, indices, indptr), shape=(self.n, self.k))

    def hard_assignments(self) -> np.ndarray:
        lab = np.empty(self.n, np.int32)
        self.ctx.check(self.ctx.lib.sc_fit_hard_assignments(self.h, ptr(lab)), "sc_fit_hard_assignments")
        return lab

    def hard_archetypes(self) -> np.ndarray:
        c = np.empty(self.k, np.int32)
        self.ctx.check(self.ctx.lib.sc_fit_hard_archetypes(self.h, ptr(c)), "sc_fit_hard_archetypes")
        return c

    def soft_assignments(self, top=5):
        idx = np.empty((self.n, top), np.int32)
        w = np.empty((self.n, top), np.float64)
        self.ctx.check(self.ctx.lib.sc_fit_soft_assignments(self.h, top, ptr(idx), ptr(w)), "sc_fit_soft_assignments")
        return idx, w

    def stats(self):
        out = np.zeros(16, np.int64)
        self.ctx.check(self.ctx.lib.sc_fit_stats(self.h, ptr(out)), "sc_fit_stats")
        return {"slow_cells_last_update_A": int(out[0]), "t2_nnz_last_update_B": int(out[1]),
                "hub_archetypes_last_update_B": int(out[2]), "dense_rows_spgemm_total": int(out[3]),
                "gradients_evaluated_last_update_B": int(out[4]), "updB_steps_timed": int(out[5]),
                "updB_eval_ns_total": int(out[6]), "kb_nnz_last": int(out[7]), "hub_steps_timed": int(out[8]),
                "hub_ns_total": int(out[9]), "hub_bytes_per_launch": int(out[10]),
                "chunks_pruned_unread_last_update_B": int(out[11]), "chunk_launches_last_update_B": int(out[12]),
                "updA_long_zero_scans": int(out[13]), "updA_long_zero_scan_batches": int(out[14])}

    def shard(self):
        """Cells [begin, end) this rank owns (the whole range on one GPU)."""
        b, e = C.c_int(0), C.c_int(0)
        self.ctx.check(self.ctx.lib.sc_fit_shard(self.h, C.byref(b), C.byref(e)), "sc_fit_shard")
        return int(b.value), int(e.value)

    STAGES = ["KB_sources", "KE_step", "KB+merge", "t1A", "updA_smem", "updA_generic", "gram", "KAt_rows", "cand_by_arch",
              "updB_step", "rss"]

    def profile(self):
        out = np.zeros(16, np.float64)
        self.ctx.check(self.ctx.lib.sc_fit_profile(self.h, ptr(out)), "sc_fit_profile")
        return dict(zip(self.STAGES, out.tolist()))

    def close(self):
        if getattr(self, "h", None) is not None:
            if getattr(self.ctx, "h", None):  # the context is gone (closed first): its stream no longer exists
                self.ctx.lib.sc_fit_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------ the model class
class SEACellsB200:
    """Drop-in for cpu.SEACellsCPU / cpu_dense.SEACellsCPUDense (same constructor arguments, methods, attributes).

    `use_sparse=True` follows cpu.py (A0 is L1-column-normalised, A_/B_/Z_ are scipy sparse, l2_penalty ignored);
    `use_sparse=False` follows cpu_dense.py (A0 used as given, A_/B_/Z_ are ndarrays, l2_penalty honoured).
    Both run the same compressed engine on the GPU.
    """

    def __init__(self, ad, build_kernel_on: str, n_SEACells: int, verbose: bool = True, n_waypoint_eigs: int = 10,
                 n_neighbors: int = 15, convergence_epsilon: float = 1e-3, l2_penalty: float = 0,
                 max_franke_wolfe_iters: int = 50, use_sparse: bool = False, device: int = 0):
        print("Welcome to SEACells!")
        self.ad = ad
        self.build_kernel_on = build_kernel_on
        self.n_cells = ad.shape[0]
        if not isinstance(n_SEACells, int):
            try:
                n_SEACells = int(n_SEACells)
            except ValueError:
                raise ValueError(
                    f"The number of SEACells specified must be an integer type, not {type(n_SEACells)}")
        self.k = n_SEACells
        # capacity limits of the compiled kernels (the reference has none): fail here, not deep inside fit()
        if not 1 <= int(max_franke_wolfe_iters) <= FitEngine.MAX_FW_ITER:
            raise ValueError(f"max_franke_wolfe_iters must be in [1, {FitEngine.MAX_FW_ITER}] for the B200 engine, "
                             f"not {max_franke_wolfe_iters} (INTEGRATION.md, 'Capacity limits')")
        if not 2 <= int(n_neighbors) <= FitEngine.MAX_NEIGHBORS:
            raise ValueError(f"n_neighbors must be in [2, {FitEngine.MAX_NEIGHBORS}] for the B200 engine, "
                             f"not {n_neighbors} (INTEGRATION.md, 'Capacity limits')")
        self.n_waypoint_eigs = n_waypoint_eigs
        self.waypoint_proportion = 1
        self.n_neighbors = n_neighbors
        self.max_FW_iter = max_franke_wolfe_iters
        self.verbose = verbose
        self.l2_penalty = l2_penalty
        self.use_sparse = use_sparse
        self.RSS_iters = []
        self.convergence_epsilon = convergence_epsilon
        self.convergence_threshold = None
        self.archetypes = None
        self.A0 = None
        self.B0 = None
        self.Z_ = None
        self._device = device
        self._ctx = _lib.default_context(device)  # raises when the library or the GPU is missing
        self._km = None          # KernelMatrix (device) when built by construct_kernel_matrix
        self._M_host = None      # scipy CSR when given by add_precomputed_kernel_matrix / exported
        self._K_host = None
        self._engine = None
        self._A_cache = None
        self._B_cache = None
        self.sigma_ = None

    # ---- kernel matrix -----------------------------------------------------------------------------------------
    @property
    def kernel_matrix(self):
        if self._M_host is None and self._km is not None:
            self._M_host, self.sigma_, _, _ = self._km.export(want_knn=False)
        return self._M_host

    @kernel_matrix.setter
    def kernel_matrix(self, M):
        self._M_host = M
        self._km = None
        self._K_host = None
        self._engine = None
        self._A_cache = self._B_cache = None

    @property
    def K(self):
        """K = M @ M.T (cpu.py:127,157).  Compatibility attribute: the engine applies M twice and never reads it."""
        if self._K_host is None and self.kernel_matrix is not None:
            M = self.kernel_matrix
            self._K_host = M @ M.T
        return self._K_host

    def add_precomputed_kernel_matrix(self, K):
        """cpu.py:115-128."""
        assert K.shape == (self.n_cells, self.n_cells), \
            f"Dimension of kernel matrix must be n_cells = ({self.n_cells},{self.n_cells}), not {K.shape} "
        self.kernel_matrix = sp.csr_matrix(K)

    def construct_kernel_matrix(self, n_neighbors: int = None, graph_construction="union"):
        """cpu.py:130-159 -> build_graph.SEACellGraph.rbf, on the GPU with an exact kNN."""
        if n_neighbors is None:
            n_neighbors = self.n_neighbors
        X = self.ad.obsm[self.build_kernel_on]
        if self.verbose:
            print("Computing kNN graph using the exact B200 brute-force kNN ...")
        print(f"Parameter graph_construction = {graph_construction} being used to build KNN graph...")
        km = build_kernel(X, n_neighbors, graph_construction, self._ctx)
        self._km, self._M_host, self._K_host, self._engine = km, None, None, None
        self._A_cache = self._B_cache = None
        if hasattr(self.ad, "obsp"):
            M, self.sigma_, idx, d2 = km.export(want_knn=True)
            self._M_host = M
            n, kk = idx.shape
            D = sp.csr_matrix((np.sqrt(d2).ravel(), idx.ravel(), np.arange(0, n * kk + 1, kk)), shape=(n, n))
            D.eliminate_zeros()
            self.ad.obsp["distances"] = D   # what scanpy.pp.neighbors leaves behind (build_graph.py:126)
        return

    # ---- initialisation ----------------------------------------------------------------------------------------
    def initialize_archetypes(self):
        """cpu.py:161-202.  The waypoint half needs palantir (third party); the greedy half is in-tree."""
        k = self.k
        if self.waypoint_proportion > 0:
            waypoint_ix = self._get_waypoint_centers(k)
            waypoint_ix = np.random.choice(waypoint_ix, int(len(waypoint_ix) * self.waypoint_proportion),
                                           replace=False)
            from_greedy = self.k - len(waypoint_ix)
            if self.verbose:
                print(f"Selecting {len(waypoint_ix)} cells from waypoint initialization.")
        else:
            from_greedy = self.k
        greedy_ix = self._get_greedy_centers(n_SEACells=from_greedy + 10)
        if self.verbose:
            print(f"Selecting {from_greedy} cells from greedy initialization.")
        if self.waypoint_proportion > 0:
            all_ix = np.hstack([waypoint_ix, greedy_ix])
        else:
            all_ix = np.hstack([greedy_ix])
        unique_ix, ind = np.unique(all_ix, return_index=True)
        all_ix = unique_ix[np.argsort(ind)][:k]
        self.archetypes = all_ix

    def _get_waypoint_centers(self, n_waypoints=None):
        """cpu.py:287-340 (palantir diffusion maps + max-min sampling)."""
        try:
            import palantir
        except ImportError as e:
            raise ImportError(
                "waypoint initialisation needs the third-party `palantir` package (cpu.py:320-334). "
                "Pass `initial_archetypes=` to fit()/initialize(), or set model.waypoint_proportion = 0 "
                "to use the greedy initialisation only.") from e
        k = self.k if n_waypoints is None else n_waypoints
        comps = pd.DataFrame(np.asarray(self.ad.obsm[self.build_kernel_on])).set_index(self.ad.obs_names)
        print(f"Building kernel on {self.build_kernel_on}")
        dm_res = palantir.utils.run_diffusion_maps(comps, n_components=self.n_neighbors)
        dc = palantir.utils.determine_multiscale_space(dm_res, n_eigs=self.n_waypoint_eigs)
        waypoint_init = palantir.core._max_min_sampling(data=dc, num_waypoints=k)
        dc["iix"] = np.arange(len(dc))
        return dc.loc[waypoint_init]["iix"].values

    def _get_greedy_centers(self, n_SEACells=None):
        """cpu.py:342-423 - greedy adaptive CSSP on K.  Next-row component."""
        from .greedy import greedy_centers
        k = self.k if n_SEACells is None else n_SEACells
        return greedy_centers(self, k)

    def _ensure_engine(self, k):
        if self._engine is None or self._engine.k != k or self._engine.T != self.max_FW_iter:
            if self._km is None and self._M_host is None:
                raise RuntimeError("Must first construct kernel matrix before initializing SEACells.")
            if self._ctx is None:  # restored from a pickle
                self._ctx = _lib.default_context(getattr(self, "_device", 0))
            eng = FitEngine(self.n_cells, k, self.max_FW_iter, self._ctx)
            eng.set_kernel(self._km if self._km is not None else self._M_host)
            self._engine = eng
        return self._engine

    def initialize(self, initial_archetypes=None, initial_assignments=None):
        """cpu.py:204-285 / cpu_dense.py:145-202."""
        if self._km is None and self._M_host is None:
            raise RuntimeError("Must first construct kernel matrix before initializing SEACells.")
        n = self.n_cells
        if initial_archetypes is not None:
            if self.verbose:
                print("Using provided list of initial archetypes")
            self.archetypes = initial_archetypes
        if self.archetypes is None:
            self.initialize_archetypes()
        self.k = len(self.archetypes)
        k = self.k
        eng = self._ensure_engine(k)
        rows = np.asarray(self.archetypes)
        B0 = sp.csr_matrix((np.ones(len(rows)), (rows, np.arange(k))), shape=(n, k))
        self.B0 = B0 if self.use_sparse else B0.toarray()
        eng.set_archetypes(rows)

        if initial_assignments is not None:
            A0 = initial_assignments
            assert A0.shape == (k, n), f"Initial assignment matrix should be of shape (k={k} x n={n})"
            if self.use_sparse:
                from sklearn.preprocessing import normalize
                A0 = normalize(sp.csr_matrix(A0), axis=0, norm="l1")
        else:
            if self.use_sparse:
                # roughly 25% of the archetypes per cell, legacy global RNG stream as cpu.py:257-264
                from sklearn.preprocessing import normalize
                per_cell = int(k * 0.25)
                r = np.random.randint(0, k, size=(n, per_cell)).reshape(-1)
                c = np.repeat(np.arange(n), per_cell)
                A0 = sp.csr_matrix((np.random.random(len(r)), (r, c)), shape=(k, n))
                A0 = normalize(A0, axis=0, norm="l1")
            else:
                A0 = np.random.random((k, n))
                A0 /= A0.sum(0)
            if self.verbose:
                print("Randomly initialized A matrix.")
        self.A0 = A0
        eng.set_A(A0)
        eng.update_A(self.l2_penalty if not self.use_sparse else 0.0)
        self._A_cache = self._B_cache = None
        RSS = eng.rss()
        self.RSS_iters.append(RSS)
        if self.convergence_threshold is None:
            self.convergence_threshold = self.convergence_epsilon * RSS
            if self.verbose:
                print(f"Setting convergence threshold at {self.convergence_threshold:.5f}")

    # ---- model state as the reference exposes it ---------------------------------------------------------------
    @property
    def A_(self):
        if self._A_cache is None and self._engine is not None:
            A = self._engine.A()
            self._A_cache = A if self.use_sparse else A.toarray()
        return self._A_cache

    @property
    def B_(self):
        if self._B_cache is None and self._engine is not None:
            B = self._engine.B()
            self._B_cache = B if self.use_sparse else B.toarray()
        return self._B_cache

    def _fitted_engine(self):
        """The engine holding A_ / B_.  A model restored from a pickle (save_model) has host matrices only: the device
        state is rebuilt from them on first use."""
        if self._engine is None:
            if self._A_cache is None or self._B_cache is None:
                raise RuntimeError("Cell to SEACell assignment matrix has not been initialised. "
                                   "Run model.initialize() first.")
            if self._ctx is None:
                self._ctx = _lib.default_context(getattr(self, "_device", 0))
            A, B = sp.csc_matrix(self._A_cache), sp.csr_matrix(self._B_cache)
            eng = FitEngine(self.n_cells, A.shape[0], self.max_FW_iter, self._ctx)
            eng.set_kernel(self._M_host)
            eng.set_A_compressed(A)
            eng.set_B(B)
            self._engine = eng
        return self._engine

    def _updateA(self, B, A_prev):
        """cpu.py:425-459 for explicit matrices (k x n A_prev, n x k B with compressible columns)."""
        eng = self._ensure_engine(B.shape[1])
        eng.set_B(B)
        eng.set_A(A_prev)
        eng.update_A(self.l2_penalty if not self.use_sparse else 0.0)
        A = eng.A()
        self._A_cache = self._B_cache = None
        return A if self.use_sparse else A.toarray()

    def _updateB(self, A, B_prev):
        """cpu.py:461-498 for explicit matrices."""
        eng = self._ensure_engine(A.shape[0])
        eng.set_A_compressed(A)
        eng.set_B(B_prev)
        eng.update_B()
        B = eng.B()
        self._A_cache = self._B_cache = None
        return B if self.use_sparse else B.toarray()

    def compute_reconstruction(self, A=None, B=None):
        """cpu.py:500-518: (M B) A.  Accessor only (n x n); the fit loop never forms it."""
        A = self.A_ if A is None else A
        B = self.B_ if B is None else B
        if A is None or B is None:
            raise RuntimeError("Either assignment matrix A or archetype matrix B is None.")
        return (self.kernel_matrix.dot(B)).dot(A)

    def compute_RSS(self, A=None, B=None):
        """cpu.py:520-536: ||M - (M B) A||_F, evaluated on the GPU without the n x n product."""
        if A is None and B is None:
            if self._engine is None and self._A_cache is None:
                raise RuntimeError("Either assignment matrix A or archetype matrix B is None.")
            return self._fitted_engine().rss()
        A = self.A_ if A is None else A
        B = self.B_ if B is None else B
        if A is None or B is None:
            raise RuntimeError("Either assignment matrix A or archetype matrix B is None.")
        eng = FitEngine(self.n_cells, sp.csr_matrix(A).shape[0], self.max_FW_iter, self._ctx)
        eng.set_kernel(self._km if self._km is not None else self._M_host)
        eng.set_A_compressed(A)
        eng.set_B(B)
        r = eng.rss()
        eng.close()
        return r

    def plot_convergence(self, save_as=None, show=True):
        """cpu.py:538-556."""
        import matplotlib.pyplot as plt
        plt.figure()
        plt.plot(self.RSS_iters)
        plt.title("Reconstruction Error over Iterations")
        plt.xlabel("Iterations")
        plt.ylabel("Squared Error")
        if save_as is not None:
            plt.savefig(save_as, dpi=150)
        if show:
            plt.show()
        plt.close()

    def _step_engine(self):
        eng = self._engine
        eng.update_A(self.l2_penalty if not self.use_sparse else 0.0)
        eng.update_B()
        self._A_cache = self._B_cache = None
        self.RSS_iters.append(eng.rss())

    def step(self):
        """cpu.py:558-593."""
        if self._km is None and self._M_host is None:
            raise RuntimeError("Kernel matrix has not been computed. Run model.construct_kernel_matrix() first.")
        if self._engine is None and self._A_cache is None:
            raise RuntimeError(
                "Cell to SEACell assignment matrix has not been initialised. Run model.initialize() first.")
        self._fitted_engine()
        self._step_engine()
        labels = self.get_hard_assignments()
        self.ad.obs["SEACell"] = labels["SEACell"]
        return

    def _fit(self, max_iter: int = 50, min_iter: int = 10, initial_archetypes=None, initial_assignments=None):
        """cpu.py:595-651.  Labels are written to ad.obs once, after the loop (the reference rewrites them each step)."""
        self.initialize(initial_archetypes=initial_archetypes, initial_assignments=initial_assignments)
        converged = False
        n_iter = 0
        while (not converged and n_iter < max_iter) or n_iter < min_iter:
            n_iter += 1
            if n_iter == 1 or (n_iter) % 10 == 0:
                if self.verbose:
                    print(f"Starting iteration {n_iter}.")
            self._step_engine()
            if n_iter == 1 or (n_iter) % 10 == 0:
                if self.verbose:
                    print(f"Completed iteration {n_iter}.")
            if np.abs(self.RSS_iters[-2] - self.RSS_iters[-1]) < self.convergence_threshold:
                if self.verbose:
                    print(f"Converged after {n_iter} iterations.")
                converged = True
        self.Z_ = None  # computed on access: B_.T @ K  (cpu.py:641)
        self._fitted = True
        labels = self.get_hard_assignments()
        self.ad.obs["SEACell"] = labels["SEACell"]
        if not converged:
            raise RuntimeWarning(
                "Warning: Algorithm has not converged - you may need to increase the maximum number of iterations")
        return

    def fit(self, max_iter: int = 100, min_iter: int = 10, initial_archetypes=None, initial_assignments=None):
        """cpu.py:653-678."""
        if max_iter < min_iter:
            raise ValueError(
                "The maximum number of iterations specified is lower than the minimum number of iterations specified.")
        self._fit(max_iter=max_iter, min_iter=min_iter, initial_archetypes=initial_archetypes,
                  initial_assignments=initial_assignments)

    def get_archetype_matrix(self):
        """cpu.py:680-682: Z_ = B_.T @ K (k x n)."""
        if self.Z_ is None and (self._engine is not None or self._A_cache is not None):
            Z = self._fitted_engine().KB().T.tocsc()   # (K B)^T = B^T K since K is symmetric; computed on the GPU
            self.Z_ = Z if self.use_sparse else Z.toarray()
        return self.Z_

    def get_soft_assignments(self):
        """cpu_dense.py:586-605 (the well-formed definition): top-5 (label, weight) per cell."""
        idx, w = self._fitted_engine().soft_assignments(5)
        arch = self.get_hard_archetypes()
        soft_labels = pd.DataFrame(np.asarray(arch)[idx])
        soft_labels.index = self.ad.obs_names
        return soft_labels, w

    def get_hard_assignments(self):
        """cpu.py:710-722."""
        assmts = self._fitted_engine().hard_assignments()
        df = pd.DataFrame({"SEACell": [f"SEACell-{i}" for i in assmts]})
        df.index = self.ad.obs_names
        df.index.name = "index"
        return df

    def get_hard_archetypes(self):
        """cpu.py:724-729."""
        return self.ad.obs_names[self._fitted_engine().hard_archetypes()]

    def save_model(self, outdir):
        """cpu.py:731-741 (the device handles are dropped from the pickle; host state is kept)."""
        import pickle
        with open(outdir + "/model.pkl", "wb") as f:
            pickle.dump(self, f)
        return None

    def __getstate__(self):
        # the reference pickles A_, B_, Z_, kernel_matrix with the model (cpu.py:731-741): materialise them on the host
        fitted = self._engine is not None or self._A_cache is not None
        st = dict(self.__dict__)
        st["_A_cache"], st["_B_cache"] = self.A_, self.B_
        st["Z_"] = self.get_archetype_matrix() if fitted and getattr(self, "_fitted", False) else self.Z_
        st["_M_host"] = self.kernel_matrix
        st["_K_host"] = None  # K = M @ M.T is recomputed on access
        for key in ("_ctx", "_km", "_engine"):
            st[key] = None
        return st

    def __setstate__(self, st):
        self.__dict__.update(st)  # device handles are re-created lazily (_fitted_engine / _ensure_engine)

    def save_assignments(self, outdir):
        """cpu.py:743-764: SEACells.csv, kernel_matrix.npz, A.npz (= A_.T, n x k), B.npz."""
        import os
        from scipy.sparse import save_npz
        os.makedirs(outdir, exist_ok=True)
        save_npz(outdir + "/kernel_matrix.npz", self.kernel_matrix)
        save_npz(outdir + "/A.npz", sp.csr_matrix(self.A_).T)
        save_npz(outdir + "/B.npz", sp.csr_matrix(self.B_))
        labels = self.get_hard_assignments()
        labels.to_csv(outdir + "/SEACells.csv")
        return None


def SEACells(ad, build_kernel_on: str, n_SEACells: int, use_gpu: bool = False, verbose: bool = True,
             n_waypoint_eigs: int = 10, n_neighbors: int = 15, convergence_epsilon: float = 1e-3,
             l2_penalty: float = 0, max_franke_wolfe_iters: int = 50, use_sparse: bool = False):
    """Same signature as SEACells.core.SEACells (core.py:13-25).

    Every combination runs on the GPU engine; `use_gpu` is accepted for compatibility.  The reference forbids
    use_sparse with use_gpu (core.py:44-46) because its cupy backend is dense-only - that restriction is kept so that
    existing call sites behave identically.
    """
    if use_sparse:
        assert not use_gpu, "Sparse matrix operations are only supported for CPU implementation."
    return SEACellsB200(ad, build_kernel_on, n_SEACells, verbose, n_waypoint_eigs, n_neighbors, convergence_epsilon,
                        l2_penalty, max_franke_wolfe_iters, use_sparse=use_sparse)


# SEACells.core also hosts the metacell summaries (core.py:103-242); keep `from seacells_b200.core import ...` working
from .summarize import sparsify_assignments, summarize_by_SEACell, summarize_by_soft_SEACell  # noqa: E402,F401
