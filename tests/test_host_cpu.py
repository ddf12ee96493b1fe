"""Host-side pieces that need no GPU: the reference arm of bench.py and the metacell quality metrics (with the engine's
kNN replaced by the oracle's, which it is bit-identical to)."""
import os
import sys

import numpy as np
import pandas as pd
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)   # the pool's workers (spawn) import the module by name: it must be importable
    import bench
    return bench


@pytest.mark.timeout(300)
def test_reference_arm_replicas():
    """`bench.py --impl reference` runs one single-threaded oracle fit per host core; here 2 replicas of a tiny sample."""
    b = _bench()
    per_rep, iters, k = b.cpu_fit_replicas(450, 2, 2)
    assert len(per_rep) == 2 and all(t > 0 for t in per_rep)
    assert iters >= 10 and k == 6
    dt, it1, k1 = b.cpu_fit_sample(450)
    assert it1 == iters and k1 == k
    assert "450 cells" in b.workload_name(450)


def test_evaluate_metrics_with_oracle_knn(monkeypatch):
    """evaluate.py:6-81 / 118-139 on injected diffusion components."""
    from sklearn.neighbors import NearestNeighbors
    from oracle import seacells_oracle as O
    import seacells_b200.core as core
    from seacells_b200 import evaluate
    from seacells_b200.synth import MiniAnnData, synth_embedding
    monkeypatch.setattr(core, "knn", lambda X, kn, ctx=None: O.knn_exact(X, kn))
    n, k = 900, 15
    rng = np.random.default_rng(3)
    dc = rng.normal(size=(n, 10)) * np.linspace(2.0, 0.2, 10)
    ad = MiniAnnData(synth_embedding(n, 50, 9))
    ad.obs["SEACell"] = [f"SEACell-{i}" for i in rng.integers(0, k, size=n)]
    ad.obs["celltype"] = rng.choice(["a", "b", "c"], size=n)
    df = pd.DataFrame(dc, index=ad.obs_names).join(ad.obs["SEACell"])
    comp = evaluate.compactness(ad, diffusion_components=dc)
    np.testing.assert_allclose(comp["compactness"].values, df.groupby("SEACell").var().mean(axis=1).values, rtol=1e-12)
    cent = df.groupby("SEACell").mean()
    sep = evaluate.separation(ad, nth_nbr=1, diffusion_components=dc)
    ref = NearestNeighbors(n_neighbors=1, algorithm="brute").fit(cent.values).kneighbors()[0][:, 0]
    assert list(sep.columns) == ["separation"] and list(sep.index) == list(cent.index)
    np.testing.assert_allclose(sep["separation"].values, ref, rtol=1e-8)
    assert 0 < len(evaluate.separation(ad, cluster="celltype", diffusion_components=dc)) <= k
    pur = evaluate.compute_celltype_purity(ad, "celltype")
    assert pur.shape == (k, 2) and ((pur["celltype_purity"] > 0) & (pur["celltype_purity"] <= 1)).all()
    with pytest.raises(ImportError):
        evaluate.compactness(ad)
