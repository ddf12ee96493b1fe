"""Small fit used under ncu (launch list / full capture): n cells, k archetypes, 1 outer iteration."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sklearn.preprocessing import normalize
from seacells_b200 import FitEngine, build_kernel
from seacells_b200.synth import synth_archetypes, synth_assignments, synth_embedding
n = int(sys.argv[1]); k = n // 75
X = synth_embedding(n, 50, 0)
km = build_kernel(X, 15)
eng = FitEngine(n, k)
eng.set_kernel(km); eng.set_archetypes(synth_archetypes(n, k))
eng.set_A(normalize(synth_assignments(n, k), axis=0, norm="l1"))
eng.update_A(); print("rss0", eng.rss())
eng.update_A(); eng.update_B(); print("rss1", eng.rss(), eng.stats())
