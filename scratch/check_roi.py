import sys
sys.path.insert(0, ".")
import torch
from fujishadergpu_b200 import kernels as k
from fujishadergpu_b200.algorithms._norm_stats import compute_norm_stats_device
W6 = [32 / 63, 16 / 63, 8 / 63, 4 / 63, 2 / 63, 1 / 63]
for shape, nod in (((3000, 2500), False), ((3000, 2500), True), ((9000, 8300), False)):
    d = k.synth_dem(shape, seed=3, nodata=nod)
    full = k.topousm_fast(d, radii=[2, 8, 32, 128, 512, 2048], weights=W6)
    m = min(2064, shape[0] // 3, shape[1] // 3)
    out = torch.full(shape, float("nan"), device="cuda")
    k.topousm_fast(d, radii=[2, 8, 32, 128, 512, 2048], weights=W6, out=out, roi=(m, shape[0] - 2 * m, m, shape[1] - 2 * m))
    a, b = full[m:-m, m:-m], out[m:-m, m:-m]
    ok = torch.equal(torch.isnan(a), torch.isnan(b)) and torch.equal(torch.nan_to_num(a), torch.nan_to_num(b))
    print("roi", shape, nod, "ok" if ok else "FAIL", "untouched rows are NaN:", bool(torch.isnan(out[: m - 64]).all()))
d = k.synth_dem((16384, 16384), seed=9)
print("stats", compute_norm_stats_device(d, "topousm_fast", {"radii": [2, 8, 32, 128, 512, 2048], "weights": W6, "pixel_size": 1.0}))
