"""Deterministic synthetic embeddings for the BASELINE.json configs.

Nothing here touches the GPU or the oracle; bench.py, the tests and the golden-vector script all
draw their inputs from these generators so that every arm sees byte-identical data.
"""
from __future__ import annotations

import math

import numpy as np
import pandas as pd
import scipy.sparse as sp


def synth_embedding(n: int, d: int = 50, seed: int = 0, kind: str = "rna") -> np.ndarray:
    """Gaussian-mixture embedding with a PCA-like decaying spectrum, float32, C-contiguous.

    kind="rna":  per-dimension noise sd linspace(3.0, 0.3, d)   (configs 1, 2, 4, 5: 'X_pca')
    kind="atac": per-dimension noise sd linspace(2.0, 0.5, d)   (config 3: 'X_svd', d=30)
    """
    rng = np.random.default_rng(seed)
    ncl = max(20, int(20 * math.sqrt(math.ceil(n / 1e4))))
    centres = rng.normal(0.0, 4.0, size=(ncl, d))
    lab = rng.integers(0, ncl, size=n)
    sd = np.linspace(3.0, 0.3, d) if kind == "rna" else np.linspace(2.0, 0.5, d)
    X = centres[lab] + rng.normal(0.0, 1.0, size=(n, d)) * sd
    return np.ascontiguousarray(X.astype(np.float32))


def synth_archetypes(n: int, k: int, seed: int = 1) -> np.ndarray:
    """k distinct cell positions used as `initial_archetypes` in timing / parity runs."""
    return np.sort(np.random.default_rng(seed).choice(n, size=k, replace=False)).astype(np.int64)


def synth_assignments(n: int, k: int, seed: int = 2, per_cell: int = 4) -> sp.csr_matrix:
    """Sparse k x n `initial_assignments`: `per_cell` uniform draws per cell (duplicates summed).

    Column L1 normalisation is left to the consumer, exactly as the reference does (cpu.py:252-253).
    """
    rng = np.random.default_rng(seed)
    rows = rng.integers(0, k, size=(n, per_cell)).reshape(-1)
    cols = np.repeat(np.arange(n), per_cell)
    vals = rng.random(rows.shape[0]) + 0.05
    A0 = sp.csr_matrix((vals, (rows, cols)), shape=(k, n))
    A0.sum_duplicates()
    return A0


class _Raw:
    def __init__(self, X):
        self.X = X


class MiniAnnData:
    """The slice of the AnnData interface the hot path and the metacell summaries touch.

    `.obsm` (dict of n x d arrays), `.obsp` (dict), `.obs` (DataFrame), `.obs_names` (Index), `.shape`; optionally a
    cells x genes matrix `.X` with `.var_names`, `.raw.X`, `.layers` and `ad[cells, :]` row selection by name.
    A real anndata.AnnData works in its place; anndata itself is not a dependency.
    """

    def __init__(self, X: np.ndarray, key: str = "X_pca", names=None, counts=None, var_names=None, raw=None,
                 layers=None):
        n = X.shape[0]
        self.obsm = {key: X}
        self.obsp = {}
        self.uns = {}
        self.shape = (n, X.shape[1])
        self.obs_names = pd.Index([f"c{i}" for i in range(n)] if names is None else names)
        self.obs = pd.DataFrame(index=self.obs_names)
        self.X = counts
        self.raw = _Raw(raw) if raw is not None else None
        self.layers = dict(layers or {})
        n_var = None
        for m in (counts, raw, *self.layers.values()):
            if m is not None:
                n_var = m.shape[1]
                break
        if var_names is not None:
            self.var_names = pd.Index(var_names)
        elif n_var is not None:
            self.var_names = pd.Index([f"g{j}" for j in range(n_var)])
        if n_var is not None:
            self.shape = (n, n_var)

    def __getitem__(self, key):
        cells, cols = key
        if not (isinstance(cols, slice) and cols == slice(None)):
            raise NotImplementedError("MiniAnnData only supports ad[cells, :]")
        pos = self.obs_names.get_indexer(pd.Index(cells))
        key0 = next(iter(self.obsm))
        sub = MiniAnnData(self.obsm[key0][pos], key0, names=self.obs_names[pos],
                          counts=None if self.X is None else self.X[pos],
                          var_names=getattr(self, "var_names", None),
                          raw=None if self.raw is None else self.raw.X[pos],
                          layers={k: v[pos] for k, v in self.layers.items()})
        sub.obs = self.obs.iloc[pos]
        return sub

    @property
    def n_obs(self):
        return self.shape[0]
