export SC_KNN_TC_TIME=1
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown --format=csv -lms 100 > gpurun_out/clocks_knn.csv &
SMI=$!
python - <<'PY'
import sys, os, time
sys.path.insert(0, os.getcwd())
from seacells_b200 import _lib
from seacells_b200.core import knn
from seacells_b200.synth import synth_embedding
ctx = _lib.default_context()
X = synth_embedding(1000000, 50, 0)
knn(X[:4096], 15, ctx)
for i in range(4):
    t = time.time(); knn(X, 15, ctx); print("knn", time.time() - t, flush=True)
PY
kill $SMI
