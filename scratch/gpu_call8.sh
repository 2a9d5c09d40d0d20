timeout 900 python -m pytest tests/test_cog_io.py -m gpu -x -q > gpurun_out/pytest_cog.log 2>&1; tail -12 gpurun_out/pytest_cog.log
python - <<'PY' 2>&1 | tail -5
import sys, time; sys.path.insert(0, ".")
import torch
from fujishadergpu_b200 import kernels as k
from fujishadergpu_b200.io import cog_writer as cw
S = 16384
d = k.synth_dem((S, S))
out = k.topousm_fast(d, radii=[2, 8, 32, 128, 512, 2048], weights=[32/63,16/63,8/63,4/63,2/63,1/63], norm_scale=14.65, output_dtype="uint8",
                     qp={"a_coef": 107.99319, "b_coef": 128.0, "dn_min": 1, "dn_max": 255})
torch.cuda.synchronize()
t0 = time.perf_counter(); st = cw.write_cog("/tmp/big.tif", out); dt = time.perf_counter() - t0
print(f"write_cog {S}^2 uint8: {dt:.2f} s, {st['bytes']/1e6:.1f} MB, {S*S/dt/1e6:.1f} Mpx/s")
a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
a.record(); lv = out
for _ in range(8): lv = k.overview_average(lv, 0)
b.record(); torch.cuda.synchronize(); print(f"overview cascade (8 levels): {a.elapsed_time(b):.3f} ms")
PY
