import torch,sys
sys.path.insert(0,".")
from fujishadergpu_b200 import kernels as k
d=k.synth_dem((16384,16384))
for i in range(2): o=k.hillshade(d,pixel_scale_x=1.0,pixel_scale_y=-1.0)
for i in range(2): o=k.slope(d,pixel_scale_x=1.0,pixel_scale_y=-1.0)
for i in range(2): o=k.curvature(d,pixel_scale_x=1.0,pixel_scale_y=-1.0)
torch.cuda.synchronize()
