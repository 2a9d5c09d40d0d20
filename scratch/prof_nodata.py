"""One main pass on the 65536^2 NoData DEM (for an ncu launch list)."""
import sys
sys.path.insert(0, ".")
import torch
from fujishadergpu_b200 import kernels as k
S = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
W6 = [32 / 63, 16 / 63, 8 / 63, 4 / 63, 2 / 63, 1 / 63]
R = [2, 8, 32, 128, 512, 2048]
d = k.synth_dem((S, S), seed=20261019, nodata=True)
ws = torch.empty(max(256, k.topousm_fast_workspace_bytes((S, S), R, 1.0)), dtype=torch.uint8, device="cuda")
out = torch.empty((S, S), dtype=torch.float32, device="cuda")
for _ in range(2):
    k.topousm_fast(d, radii=R, weights=W6, norm_scale=14.65, workspace=ws, out=out)
torch.cuda.synchronize()
