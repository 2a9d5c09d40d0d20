import os, sys
sys.path.insert(0, ".")
import torch
from fujishadergpu_b200 import kernels as k
from scratch.bench_algos import timeit, report
fails = 0
for shape, nod in (((1500, 1700), True), ((1500, 1700), False), ((700, 3001), True)):
    d = k.synth_dem(shape, seed=5, nodata=nod)
    for typ in ("positive", "negative"):
        for nd, md in ((8, 256), (16, 50), (8, 31)):
            os.environ["FSG_OPENNESS_GENERIC"] = "1"
            ref = k.openness(d, openness_type=typ, num_directions=nd, max_distance=md, pixel_size=1.0)
            os.environ.pop("FSG_OPENNESS_GENERIC")
            got = k.openness(d, openness_type=typ, num_directions=nd, max_distance=md, pixel_size=1.0)
            torch.cuda.synchronize()
            an, bn = torch.isnan(ref), torch.isnan(got)
            ok = torch.equal(an, bn) and torch.equal(torch.nan_to_num(ref), torch.nan_to_num(got))
            fails += 0 if ok else 1
            print("ok  " if ok else "FAIL", shape, nod, typ, nd, md, float((torch.nan_to_num(ref) - torch.nan_to_num(got)).abs().max()), flush=True)
print("FAILS", fails)
S = 16384
d = k.synth_dem((S, S))
b, m = timeit(lambda: k.openness(d, openness_type="positive", num_directions=8, max_distance=256, pixel_size=1.0), n=3, warm=1)
report("openness positive 8 dir r=256 (fast)", S, b, m)
os.environ["FSG_OPENNESS_GENERIC"] = "1"
b, m = timeit(lambda: k.openness(d, openness_type="positive", num_directions=8, max_distance=256, pixel_size=1.0), n=3, warm=1)
report("openness positive 8 dir r=256 (generic)", S, b, m)
