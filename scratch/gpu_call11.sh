timeout 900 python scratch/check_v6.py 16384 > gpurun_out/check_v6_prefetch.log 2>&1
grep -c "^ok" gpurun_out/check_v6_prefetch.log; grep "FAIL" gpurun_out/check_v6_prefetch.log | head; tail -16 gpurun_out/check_v6_prefetch.log
python - <<'PY' 2>&1 | tail -4
import sys; sys.path.insert(0, ".")
import torch
from fujishadergpu_b200 import kernels as k
from fujishadergpu_b200.algorithms import _norm_stats as ns
d = k.synth_dem((65536, 65536), seed=20261019)
wins = ns.stratified_windows(65536, 65536, 0, 65536, 0, 65536, grid=3, tile=8256)
views = [d[y:y + th, x:x + tw] for (y, x, tw, th) in wins]
for i in range(3):
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); nv = k.count_samples(views, finite_only=True); b.record(); torch.cuda.synchronize()
    print(f"count_samples run {i}: {a.elapsed_time(b):.3f} ms", nv[:2])
PY
