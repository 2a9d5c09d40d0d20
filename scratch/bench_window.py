"""One statistics window (8256^2, region of interest = the trimmed interior) by profiler tag, and alone."""
import sys
sys.path.insert(0, ".")
import torch
from fujishadergpu_b200 import kernels as k
R = [2, 8, 32, 128, 512, 2048]
W6 = [32 / 63, 16 / 63, 8 / 63, 4 / 63, 2 / 63, 1 / 63]
S, m = 8256, 2064
d = k.synth_dem((S, S), seed=3)
names = {1: "fused_main", 2: "pyramid", 3: "coarse", 6: "fused_roi"}
def run():
    return k.topousm_fast(d, radii=R, weights=W6, roi=(m, S - 2 * m, m, S - 2 * m))
for _ in range(3):
    run()
torch.cuda.synchronize()
k.profile_enable(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    run()
e1.record(); torch.cuda.synchronize()
prof = k.profile_read(); k.profile_enable(False)
print(f"window total {e0.elapsed_time(e1) / 10:.3f} ms")
per = {}
for tag, ms in prof:
    per.setdefault(tag, []).append(ms)
for tag, v in per.items():
    n = len(v) // 10
    print(names.get(tag, tag), [round(sum(v[i::n]) / 10, 3) for i in range(n)])
