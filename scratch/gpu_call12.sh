timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "curvature or gradient or spatial" > gpurun_out/pytest_curv.log 2>&1; tail -5 gpurun_out/pytest_curv.log
python - <<'PY' 2>&1 | tail -8
import os, sys; sys.path.insert(0, ".")
import torch
from fujishadergpu_b200 import kernels as k
d = k.synth_dem((16384, 16384))
for env in ("1", None, "1", None):
    if env: os.environ["FSG_GRAD_TILED"] = env
    else: os.environ.pop("FSG_GRAD_TILED", None)
    for ct in ("mean", "profile"):
        for _ in range(2): k.curvature(d, curvature_type=ct, pixel_scale_x=1.0, pixel_scale_y=-1.0)
        ts = []
        for _ in range(5):
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record(); k.curvature(d, curvature_type=ct, pixel_scale_x=1.0, pixel_scale_y=-1.0); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        print(f"curvature {ct} 16384^2 {'tiled' if env else 'streaming'}: {ts[2]:.3f} ms = {8*16384**2/ts[2]/1e6:.0f} GB/s")
PY
