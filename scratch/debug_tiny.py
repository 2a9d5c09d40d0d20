import numpy as np, torch, sys
sys.path.insert(0, '.')
from oracle import terrain_oracle as orc
from fujishadergpu_b200 import kernels as k
dem = orc.synth_dem(50,1,seed=51); dem[25,0]=np.nan
d = torch.from_numpy(dem).cuda()
for radii in ([3],[100],[3,100]):
    want = orc.topousm_fast_block(dem, radii=radii)
    got = k.topousm_fast(d, radii=radii).cpu().numpy()
    bad = np.where(~np.isnan(want) & (got!=want))[0]
    print(radii, "mismatch rows", bad.tolist(), [(float(got[i,0]), float(want[i,0])) for i in bad[:4]])
sm = k.decimate(d,4).cpu().numpy(); so = orc.decimate_valid_mean(dem,4)
print("decimate equal", np.array_equal(sm, so, equal_nan=True))
