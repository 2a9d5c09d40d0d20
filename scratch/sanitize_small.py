import os, sys
sys.path.insert(0, ".")
import torch
from fujishadergpu_b200 import kernels as k
W6 = [32 / 63, 16 / 63, 8 / 63, 4 / 63, 2 / 63, 1 / 63]
for shape, nod in (((200, 600), True), ((97, 1100), False)):
    d = k.synth_dem(shape, seed=5, nodata=nod)
    for env in ([], ["FSG_V6_CFGB"], ["FSG_FUSED_V7"], ["FSG_FUSED_V5"], ["FSG_FORCE_GENERIC"]):
        for key in ("FSG_FORCE_GENERIC", "FSG_FUSED_V5", "FSG_FUSED_V7", "FSG_V6_CFGB"):
            os.environ.pop(key, None)
        for key in env:
            os.environ[key] = "1"
        o = k.topousm_fast(d, radii=[2, 8, 32, 128, 512, 2048], weights=W6, norm_scale=12.0)
    for key in ("FSG_FORCE_GENERIC", "FSG_FUSED_V5", "FSG_FUSED_V7", "FSG_V6_CFGB"):
        os.environ.pop(key, None)
    k.hillshade(d, pixel_scale_x=1.0, pixel_scale_y=-1.0); k.slope(d); k.curvature(d)
    k.openness(d, num_directions=8, max_distance=40)
    k.gaussian_nan(d, 3.0); k.decimate(d, 4)
    k.percentile([d], 99.0, take_abs=True)
torch.cuda.synchronize()
print("done")
