import sys
sys.path.insert(0, ".")
import torch
from fujishadergpu_b200 import kernels as k
from fujishadergpu_b200.algorithms import _norm_stats as ns
W6 = [32 / 63, 16 / 63, 8 / 63, 4 / 63, 2 / 63, 1 / 63]
d = k.synth_dem((65536, 65536), seed=20261019)
p = {"radii": [2, 8, 32, 128, 512, 2048], "weights": W6, "pixel_size": 1.0}
ns.compute_norm_stats_device(d, "topousm_fast", p)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
ns.compute_norm_stats_device(d, "topousm_fast", p)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
