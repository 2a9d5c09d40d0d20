import torch,sys
sys.path.insert(0,".")
from fujishadergpu_b200 import kernels as k
S=int(sys.argv[1]) if len(sys.argv)>1 else 16384
d=k.synth_dem((S,S))
w=[32/63,16/63,8/63,4/63,2/63,1/63]
for i in range(2): o=k.topousm_fast(d,radii=[2,8,32,128,512,2048],weights=w,norm_scale=14.5)
torch.cuda.synchronize()
