timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu5.log 2>&1; tail -6 gpurun_out/pytest_gpu5.log
python - <<'PY' 2>&1 | tail -6
import sys; sys.path.insert(0, ".")
import torch
from fujishadergpu_b200 import kernels as k
from fujishadergpu_b200.algorithms import _norm_stats as ns
d = k.synth_dem((65536, 65536), seed=20261019)
wins = ns.stratified_windows(65536, 65536, 0, 65536, 0, 65536, grid=3, tile=8256)
views = [d[y:y + th, x:x + tw] for (y, x, tw, th) in wins]
for i in range(4):
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); nv = k.count_samples(views, finite_only=True); b.record(); torch.cuda.synchronize()
    print(f"count_samples run {i}: {a.elapsed_time(b):.3f} ms")
p = {"radii": [2, 8, 32, 128, 512, 2048], "weights": [32/63,16/63,8/63,4/63,2/63,1/63], "pixel_size": 1.0}
for i in range(3):
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); st = ns.compute_norm_stats_device(d, "topousm_fast", p); b.record(); torch.cuda.synchronize()
    print(f"compute_norm_stats_device run {i}: {a.elapsed_time(b):.3f} ms")
PY
