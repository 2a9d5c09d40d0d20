import torch,sys
sys.path.insert(0,".")
from fujishadergpu_b200 import kernels as k
S=int(sys.argv[1]) if len(sys.argv)>1 else 65536
d=k.synth_dem((S,S))
w=[32/63,16/63,8/63,4/63,2/63,1/63]
ws=torch.empty(max(256,k.topousm_fast_workspace_bytes((S,S),[2,8,32,128,512,2048],1.0)),dtype=torch.uint8,device="cuda")
out=torch.empty((S,S),dtype=torch.float32,device="cuda")
for i in range(2): k.topousm_fast(d,radii=[2,8,32,128,512,2048],weights=w,norm_scale=14.5,workspace=ws,out=out)
torch.cuda.synchronize()
