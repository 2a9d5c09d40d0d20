timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "openness" > gpurun_out/pytest_open.log 2>&1; tail -3 gpurun_out/pytest_open.log
python - <<'PY' 2>&1 | tail -4
import os, sys; sys.path.insert(0, ".")
import torch
from fujishadergpu_b200 import kernels as k
d = k.synth_dem((32768, 32768))
for neg in ("positive", "negative"):
    for _ in range(2): o = k.openness(d, openness_type=neg, num_directions=8, max_distance=256, pixel_size=1.0)
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); o = k.openness(d, openness_type=neg, num_directions=8, max_distance=256, pixel_size=1.0); b.record(); torch.cuda.synchronize()
    print(f"openness {neg} 8 dir r=256 32768^2: {a.elapsed_time(b):.2f} ms")
os.environ["FSG_OPENNESS_GENERIC"] = "1"
dd = d[:4096, :4096].contiguous()
g = k.openness(dd, openness_type="positive", num_directions=8, max_distance=256, pixel_size=1.0)
os.environ.pop("FSG_OPENNESS_GENERIC")
f = k.openness(dd, openness_type="positive", num_directions=8, max_distance=256, pixel_size=1.0)
print("interior == generic (4096^2):", bool(torch.equal(torch.nan_to_num(g, nan=-7.0), torch.nan_to_num(f, nan=-7.0))))
PY
