"""How far apart are the two _updateA summation variants at a size no CPU oracle reaches?  Runs the full fit with the
default (running products) and with SC_FIT_RESUM_A=1 (re-summed) and reports the per-iteration RSS, the first outer
iteration whose RSS differs and the number of cells whose hard label differs.   python scripts/variant_diff.py [n]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sklearn.preprocessing import normalize
from seacells_b200 import FitEngine, build_kernel
from seacells_b200.synth import synth_archetypes, synth_assignments, synth_embedding
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
k = n // 75
km = build_kernel(synth_embedding(n, 50, 0), 15)
arch = synth_archetypes(n, k); A0 = normalize(synth_assignments(n, k), axis=0, norm="l1")
res = {}
for name, env in (("default", {}), ("resum_A", {"SC_FIT_RESUM_A": "1"})):
    os.environ.pop("SC_FIT_RESUM_A", None); os.environ.update(env)
    eng = FitEngine(n, k); eng.set_kernel(km); eng.set_archetypes(arch); eng.set_A(A0)
    rss, it, conv, _ = eng.run(max_iter=100, min_iter=10)
    res[name] = (rss, eng.hard_assignments().copy()); eng.close()
ra, la = res["default"]; rb, lb = res["resum_A"]
first = next((i for i, (x, y) in enumerate(zip(ra, rb)) if x != y), None)
print(json.dumps({"n": n, "k": k, "rss_default": ra, "rss_resum_A": rb, "first_outer_iteration_with_different_rss": first,
                  "max_rel_rss_diff": max(abs(x - y) / x for x, y in zip(ra, rb)),
                  "cells_with_different_hard_label": int((la != lb).sum()), "fraction": float((la != lb).mean())}))
