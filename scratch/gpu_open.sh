set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "openness" 2>&1 | tail -2
python - <<'PY'
import sys; sys.path.insert(0, ".")
import torch
from fujishadergpu_b200 import kernels as k
from scratch.bench_algos import timeit, report
S = 32768
d = k.synth_dem((S, S))
b, m = timeit(lambda: k.openness(d, openness_type="positive", num_directions=8, max_distance=256, pixel_size=1.0), n=5)
report("openness positive 8 dir r=256", S, b, m)
b, m = timeit(lambda: k.openness(d, openness_type="negative", num_directions=16, max_distance=50, pixel_size=1.0), n=5)
report("openness negative 16 dir r=50", S, b, m)
PY
