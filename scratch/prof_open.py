import sys
sys.path.insert(0, ".")
import torch
from fujishadergpu_b200 import kernels as k
S = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
d = k.synth_dem((S, S))
for _ in range(3):
    o = k.openness(d, openness_type="positive", num_directions=8, max_distance=256, pixel_scale_x=1.0, pixel_scale_y=-1.0)
torch.cuda.synchronize()
