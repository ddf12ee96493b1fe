"""Whole-fit seconds at n cells for a list of environment-switch settings (development aid).
    python scripts/fit_sweep.py 1000000 "SC_R0_SEED=0" "SC_R0_SEED=16" ...   (each engine reads the switches when created)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sklearn.preprocessing import normalize
from seacells_b200 import FitEngine, build_kernel
from seacells_b200.synth import synth_archetypes, synth_assignments, synth_embedding
import hashlib
n = int(sys.argv[1]); k = n // 75
REPS = int(os.environ.get("FIT_SWEEP_REPS", "3"))
X = synth_embedding(n, 50, 0)
km = build_kernel(X, 15)
arch = synth_archetypes(n, k); A0 = normalize(synth_assignments(n, k), axis=0, norm="l1")
for cfg in sys.argv[2:] or [""]:
    sw = dict(kv.split("=") for kv in cfg.split(",") if kv)
    for key in ("SC_R0_SEED", "SC_HUB_MAX", "SC_FIT_FRESH_KB", "SC_FIT_RESUM_A", "SC_NO_OVERLAP", "SC_HUB_U"): os.environ.pop(key, None)
    os.environ.update(sw)
    eng = FitEngine(n, k); eng.set_kernel(km)
    ts = []
    for rep in range(REPS):
        eng.set_archetypes(arch); eng.set_A(A0); eng.ctx.sync()
        s0 = eng.stats(); t = time.time(); rss, it, conv, _ = eng.run(max_iter=100, min_iter=10); eng.ctx.sync(); ts.append(time.time() - t); s1 = eng.stats()
    grp = (s1["updB_eval_ns_total"] - s0["updB_eval_ns_total"]) / max(1, s1["updB_steps_timed"] - s0["updB_steps_timed"]) / 1e6
    hub = (s1["hub_ns_total"] - s0["hub_ns_total"]) / max(1, s1["hub_steps_timed"] - s0["hub_steps_timed"]) / 1e6
    sha = hashlib.sha256(eng.hard_assignments().astype(np.int32).tobytes()).hexdigest()[:16]
    print(f"{cfg or 'default':28s} labels {sha} fit s {['%.3f' % x for x in ts]} iters {it} rss_last {rss[-1]!r} group ms/step {grp:.3f} hub ms/step {hub:.3f} hub bytes/launch {s1['hub_bytes_per_launch']/1e9:.3f} GB evaluated {s1['gradients_evaluated_last_update_B']/1e6:.0f}M t2nnz {s1['t2_nnz_last_update_B']/1e6:.0f}M hubs {s1['hub_archetypes_last_update_B']} chunks pruned/launched {s1['chunks_pruned_unread_last_update_B']}/{s1['chunk_launches_last_update_B']} updA long scans {s1['updA_long_zero_scans']} batches {s1['updA_long_zero_scan_batches']}", flush=True)
    eng.close()
