"""Metacell aggregation on the GPU: drop-ins for SEACells.core.summarize_by_SEACell (core.py:190-242),
summarize_by_soft_SEACell (core.py:121-187) and sparsify_assignments (core.py:103-118).

The reference loops over the metacells in Python and sums a dense copy of every slice of the count matrix; here the
(SEACell, cell, weight) triplets and the CSR count matrix go to `sc_summarize` (csrc/summarize.cu), which adds the member
rows of every metacell in ascending cell order.  Results: identical to the reference for the hard assignment on count
data, <= 1e-12 relative for the soft assignment (the weight sums are accumulated in a different order).

The return value is an `anndata.AnnData` when anndata is importable, otherwise a `MetacellSummary` with the same
fields the reference fills in (`X`, `obs`, `obs_names`, `var_names`, `layers['raw']`).
"""
from __future__ import annotations

import numpy as np
import pandas as pd
import scipy.sparse as sp

from . import _lib
from ._lib import ptr


class MetacellSummary:
    """Minimal stand-in for the AnnData the reference returns (used only when anndata is not installed)."""

    def __init__(self, X, obs_names=None, var_names=None):
        self.X = X
        self.shape = X.shape
        self.obs_names = pd.Index([str(i) for i in range(X.shape[0])] if obs_names is None else obs_names)
        self.var_names = pd.Index(var_names) if var_names is not None else pd.Index([str(j) for j in range(X.shape[1])])
        self.obs = pd.DataFrame(index=self.obs_names)
        self.layers = {}


def _make_result(X, obs_names, var_names):
    try:
        import anndata  # noqa: F401
        out = anndata.AnnData(X)
        if obs_names is not None:
            out.obs_names = obs_names
        out.var_names = var_names
        return out
    except ImportError:
        return MetacellSummary(X, obs_names, var_names)


def _select_layer(ad, summarize_layer, hard):
    """core.py:146-149 (soft) and core.py:215-225 (hard)."""
    if hard and summarize_layer == "X":
        return ad.X
    if summarize_layer == "raw" and getattr(ad, "raw", None) is not None:
        return ad.raw.X
    return ad.layers[summarize_layer]


def aggregate(data, group, cell, weight, n_groups, divisor=None, ctx=None):
    """out[g, :] = sum_{(g, cell, w)} w * data[cell, :] (cells ascending) / divisor[g]; dense fp64 [n_groups x n_genes]."""
    ctx = ctx or _lib.default_context()
    X = sp.csr_matrix(data)
    if X.dtype not in (np.float32, np.float64):
        X = X.astype(np.float64)
    X.sort_indices()
    n, n_genes = X.shape
    indptr = np.ascontiguousarray(X.indptr, np.int64)
    indices = np.ascontiguousarray(X.indices, np.int32)
    vals = np.ascontiguousarray(X.data)
    group = np.ascontiguousarray(group, np.int32)
    cell = np.ascontiguousarray(cell, np.int32)
    weight = np.ascontiguousarray(weight, np.float64)
    div = None if divisor is None else np.ascontiguousarray(divisor, np.float64)
    out = np.empty((n_groups, n_genes), np.float64)
    ctx.check(ctx.lib.sc_summarize(ctx.h, n, n_genes, ptr(indptr), ptr(indices), ptr(vals), int(X.dtype == np.float64),
                                   len(group), ptr(group), ptr(cell), ptr(weight), n_groups, ptr(div), ptr(out)),
              "sc_summarize")
    return out


def sparsify_assignments(A, thresh: float):
    """core.py:103-118: zero every weight below `thresh`, renormalise each row (cell) to sum 1.  A: n_cells x n_SEACells."""
    if sp.issparse(A):
        W = sp.csr_matrix(A, dtype=np.float64, copy=True)
        W.data[W.data < thresh] = 0.0
        W.eliminate_zeros()
        rs = np.asarray(W.sum(1)).ravel()
        with np.errstate(divide="ignore", invalid="ignore"):
            W = sp.diags(1.0 / rs) @ W
        return sp.csr_matrix(W)
    W = np.array(A, dtype=np.float64, copy=True)
    W[W < thresh] = 0
    with np.errstate(divide="ignore", invalid="ignore"):
        W = W / W.sum(1, keepdims=True)
    return W


def summarize_by_SEACell(ad, SEACells_label="SEACell", celltype_label=None, summarize_layer="raw", ctx=None):
    """core.py:190-242: raw counts summed over the cells of every SEACell (rows in order of first appearance)."""
    labels = ad.obs[SEACells_label]
    codes, metacells = pd.factorize(labels)           # pd.unique order == order of first appearance (core.py:207)
    data = _select_layer(ad, summarize_layer, hard=True)
    keep = codes >= 0
    cells = np.nonzero(keep)[0]
    summ = aggregate(data, codes[keep], cells, np.ones(len(cells)), len(metacells), None, ctx)
    X = sp.csr_matrix(summ)
    out = _make_result(X, pd.Index(metacells).astype(str), ad.var_names)
    out.layers["raw"] = sp.csr_matrix(summ)
    if celltype_label is not None:
        try:  # core.py:233-240 (evaluate.compute_celltype_purity, evaluate.py:118-139)
            g = ad.obs.groupby(SEACells_label)[celltype_label]
            purity = pd.concat([g.agg(lambda x: x.value_counts().index[0]),
                                g.agg(lambda x: x.value_counts().values[0] / x.value_counts().values.sum())], axis=1)
            purity.columns = [celltype_label, f"{celltype_label}_purity"]
            purity.index = purity.index.astype(str)
            out.obs = out.obs.join(purity)
        except Exception as e:  # noqa: BLE001 - the reference swallows and reports this too
            print(f"Cell type purity failed with Exception {e}")
    return out


def summarize_by_soft_SEACell(ad, A, celltype_label=None, summarize_layer="raw", minimum_weight: float = 0.05, ctx=None):
    """core.py:121-187: expression of every SEACell as the weight-averaged counts of its (soft-assigned) cells.

    A: n_SEACells x n_cells assignment matrix (ndarray or scipy sparse), e.g. `model.A_`."""
    if celltype_label is not None and celltype_label not in ad.obs.columns:
        raise ValueError(f"Celltype label {celltype_label} not present in ad.obs")
    data = _select_layer(ad, summarize_layer, hard=False)
    At = A.T if not sp.issparse(A) else sp.csr_matrix(A).T
    W = sparsify_assignments(sp.csr_matrix(At) if sp.issparse(At) else np.asarray(At), minimum_weight)
    Wc = sp.coo_matrix(W)                               # explicit entries only; NaN rows handled below
    k = Wc.shape[1]
    finite = np.isfinite(Wc.data)
    grp, cell, w = Wc.col[finite], Wc.row[finite], Wc.data[finite]
    sizes = np.zeros(k)
    np.add.at(sizes, grp, w)                            # A.sum(0) (core.py:180) up to summation order
    expr = aggregate(data, grp, cell, w, k, sizes, ctx)
    # a cell whose weights are all below the threshold gets 0/0 = NaN on every SEACell in the reference (core.py:114):
    # every weight sum, hence every expression value and pseudo-size, is then NaN (core.py:161-163, 180)
    if sp.issparse(W):
        bad = ~np.isfinite(np.asarray(W.sum(1)).ravel())
    else:
        bad = ~np.isfinite(W).all(1)
    if bad.any():
        expr[:] = np.nan
        sizes[:] = np.nan
    out = _make_result(sp.csr_matrix(expr), None, ad.var_names)
    out.obs["Pseudo-sizes"] = sizes
    if celltype_label is not None:
        ct = np.asarray(ad.obs[celltype_label])[cell]
        df = pd.DataFrame({"g": grp, "ct": ct, "w": w}).groupby(["g", "ct"], sort=True)["w"].sum().reset_index()
        tot = df.groupby("g")["w"].transform("sum")
        df["p"] = df["w"] / tot
        top = df.sort_values(["g", "w"], ascending=[True, False], kind="stable").groupby("g").head(1).set_index("g")
        out.obs["celltype"] = top["ct"].reindex(range(k)).to_numpy()
        out.obs["celltype_purity"] = top["p"].reindex(range(k)).to_numpy()
    return out
