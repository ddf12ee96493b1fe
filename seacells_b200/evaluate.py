"""Metacell quality metrics of SEACells/evaluate.py (compactness :6-26, separation :29-81, celltype purity :118-139).

"Next" row f-4 of the scope table.  The reference computes both metrics on palantir's multiscale
diffusion space (`palantir.utils.run_diffusion_maps` + `determine_multiscale_space`, third party); that step is used
when palantir is importable and can be replaced by passing the components directly (`diffusion_components=`), which is
how the tests pin the arithmetic that follows it.  The group statistics are pandas on the host (k x 10 numbers); the
nearest-metacell search of `separation` runs on the engine's exact kNN (`sc_knn`, the kernel that replaces
scanpy.pp.neighbors) - sklearn's NearestNeighbors in the reference.
"""
from __future__ import annotations

import numpy as np
import pandas as pd


def _multiscale_space(ad, low_dim_embedding, n_eigs, diffusion_components):
    if diffusion_components is not None:
        dc = diffusion_components
        if not isinstance(dc, pd.DataFrame):
            dc = pd.DataFrame(np.asarray(dc))
        dc = dc.copy()
        dc.index = ad.obs_names
        return dc
    try:
        import palantir
    except ImportError as e:
        raise ImportError(
            "compactness / separation are defined on palantir's multiscale diffusion space (evaluate.py:19-21, 44-46); "
            "install palantir or pass the components as `diffusion_components=`") from e
    components = pd.DataFrame(ad.obsm[low_dim_embedding]).set_index(ad.obs_names)
    dm_res = palantir.utils.run_diffusion_maps(components)
    return palantir.utils.determine_multiscale_space(dm_res, n_eigs=n_eigs)


def compactness(ad, low_dim_embedding="X_pca", SEACells_label="SEACell", diffusion_components=None):
    """evaluate.py:6-26: per metacell, the variance (ddof = 1) of every diffusion component over its cells, averaged
    over the components."""
    dc = _multiscale_space(ad, low_dim_embedding, 10, diffusion_components)
    return pd.DataFrame(dc.join(ad.obs[SEACells_label]).groupby(SEACells_label).var().mean(axis=1)).rename(
        columns={0: "compactness"})


def separation(ad, low_dim_embedding="X_pca", nth_nbr=1, cluster=None, SEACells_label="SEACell",
               diffusion_components=None, ctx=None):
    """evaluate.py:29-81: distance, in diffusion space, from every metacell's centroid to its nth nearest other
    metacell (optionally only where that neighbour has the same majority `cluster`)."""
    from .core import knn
    dc = _multiscale_space(ad, low_dim_embedding, 10, diffusion_components)
    metacells_dcs = dc.join(ad.obs[SEACells_label], how="inner").groupby(SEACells_label).mean()
    X = np.ascontiguousarray(metacells_dcs.values, dtype=np.float64)
    # NearestNeighbors(n_neighbors=nth_nbr).kneighbors() on the training set excludes the point itself: same as the
    # engine's kNN with n_neighbors = nth_nbr + 1 (self counted, then dropped)
    idx, d2 = knn(X, nth_nbr + 1, ctx)
    dists = pd.DataFrame(np.sqrt(d2)).set_index(metacells_dcs.index)
    dists.columns += 1
    nbr_cells = np.array(metacells_dcs.index)[idx]
    metacells_nbrs = pd.DataFrame(nbr_cells)
    metacells_nbrs.index = metacells_dcs.index
    metacells_nbrs.columns += 1
    if cluster is not None:
        clusters = ad.obs.groupby(SEACells_label)[cluster].agg(lambda x: x.value_counts().index[0])
        nbr_clusters = pd.DataFrame(np.asarray(clusters.values, dtype=object)[idx]).set_index(clusters.index)  # (Arrow-backed strings
        # of pandas 3 cannot be indexed with a 2-D array, which is what the reference does at evaluate.py:70)
        nbr_clusters.columns = metacells_nbrs.columns
        nbr_clusters = nbr_clusters.join(pd.DataFrame(clusters))
        clusters_match = nbr_clusters.eq(nbr_clusters[cluster], axis=0)
        return pd.DataFrame(dists[nth_nbr][clusters_match[nth_nbr]]).rename(columns={1: "separation"})
    return pd.DataFrame(dists[nth_nbr]).rename(columns={1: "separation"})


def celltype_frac(x, col_name):
    """evaluate.py:118-121."""
    val_counts = x[col_name].value_counts()
    return val_counts.values[0] / val_counts.values.sum()


def compute_celltype_purity(ad, col_name):
    """evaluate.py:124-139: per metacell, the most abundant value of `col_name` and its share."""
    grouped = ad.obs.groupby("SEACell")
    celltype_fraction = grouped.apply(lambda x: celltype_frac(x, col_name))
    celltype = grouped.apply(lambda x: x[col_name].value_counts().index[0])
    return pd.concat([celltype, celltype_fraction], axis=1).rename(columns={0: col_name, 1: f"{col_name}_purity"})
