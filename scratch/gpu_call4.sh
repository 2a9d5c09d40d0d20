timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu3.log 2>&1; tail -15 gpurun_out/pytest_gpu3.log
python - <<'PY' > gpurun_out/ao_bench.log 2>&1
import sys; sys.path.insert(0, ".")
import torch
from fujishadergpu_b200 import kernels as k
for S in (16384,):
    d = k.synth_dem((S, S))
    for kw in (dict(num_samples=16, radius=10.0), dict(num_samples=16, radius=64.0)):
        for _ in range(2): k.ambient_occlusion(d, **kw)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); k.ambient_occlusion(d, **kw); b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        print(f"ambient_occlusion {S}^2 {kw}: {ms:.3f} ms = {S*S/ms/1e6:.1f} Gpx/s = {8*S*S/ms/1e6:.0f} GB/s algorithmic")
PY
cat gpurun_out/ao_bench.log
