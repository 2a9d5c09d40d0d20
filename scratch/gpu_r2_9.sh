set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -k "gradient or curvature or slope or hillshade or spatial" > gpurun_out/pytest_grad.log 2>&1; tail -8 gpurun_out/pytest_grad.log
python scratch/bench_grad.py > gpurun_out/bench_grad.log 2>&1; cat gpurun_out/bench_grad.log
python - <<'PY' > gpurun_out/grad_err.log 2>&1
import sys; sys.path.insert(0, ".")
import numpy as np, torch
from fujishadergpu_b200 import kernels as k
from oracle import terrain_oracle as orc
d = k.synth_dem((2048, 2048), seed=7, nodata=False)
dem = d.cpu().numpy()
NUP = dict(pixel_scale_x=1.0, pixel_scale_y=-1.0)
def rep(name, got, want):
    got = got.cpu().numpy().astype(np.float64); want = want.astype(np.float64)
    err = np.abs(got - want); bar = 1e-6 + 1e-5 * np.abs(want)
    print(f"{name:22s} max abs {err.max():.3e}  max rel {np.max(err / np.maximum(np.abs(want), 1e-30)):.3e}  max err/bar {np.max(err / bar):.3f}")
rep("hillshade", k.hillshade(d, **NUP), orc.hillshade_block(dem, **NUP))
for u in ("degree", "radian", "percent"):
    rep("slope " + u, k.slope(d, unit=u, **NUP), orc.slope_block(dem, unit=u, **NUP))
steep = d * 40.0
rep("slope deg (z x40)", k.slope(steep, unit="degree", **NUP), orc.slope_block(dem * np.float32(40.0), unit="degree", **NUP))
rep("slope pct (z x40)", k.slope(steep, unit="percent", **NUP), orc.slope_block(dem * np.float32(40.0), unit="percent", **NUP))
for t in ("mean", "gaussian", "planform", "profile"):
    g = k.curvature(d, curvature_type=t, **NUP).cpu().numpy().astype(np.float64)
    w = orc.curvature_block(dem, curvature_type=t, **NUP).astype(np.float64)
    ok = w >= 0.05
    err = np.abs(g - w)
    print(f"curvature {t:9s} well-conditioned: max abs {err[ok].max():.3e} max err/bar {np.max(err[ok] / (1e-6 + 1e-5 * w[ok])):.3f}; saturated pre-gamma max {np.abs(g[~ok]**2.2 - w[~ok]**2.2).max() if (~ok).any() else 0:.3e}")
PY
cat gpurun_out/grad_err.log
