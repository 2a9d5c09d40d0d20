"""Stage timings of the statistics pre-pass (one window of 8256^2, ROI 4128^2) and of the percentile."""
import sys, time
sys.path.insert(0, ".")
import torch
from fujishadergpu_b200 import kernels as k
from fujishadergpu_b200.algorithms import _norm_stats as ns
W6 = [32 / 63, 16 / 63, 8 / 63, 4 / 63, 2 / 63, 1 / 63]
R = [2, 8, 32, 128, 512, 2048]
d = k.synth_dem((65536, 65536), seed=20261019)
wins = ns.stratified_windows(65536, 65536, 0, 65536, 0, 65536, grid=3, tile=8256)
views = [d[y:y + th, x:x + tw] for (y, x, tw, th) in wins]
m = 2064

def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e

def run_win(v):
    return k.topousm_fast(v, radii=R, weights=W6, roi=(m, v.shape[0] - 2 * m, m, v.shape[1] - 2 * m))

for _ in range(2):
    outs = [run_win(v) for v in views]
torch.cuda.synchronize()
k.profile_enable(True)
a = ev(); outs = [run_win(v) for v in views]; b = ev(); torch.cuda.synchronize()
prof = k.profile_read(); k.profile_enable(False)
print(f"9 windows, topousm_fast(roi): {a.elapsed_time(b):.2f} ms total")
import collections
agg = collections.defaultdict(list)
for tag, ms in prof: agg[tag].append(ms)
for tag, v in agg.items(): print(f"  profile tag {tag}: n={len(v)} sum={sum(v):.2f} ms mean={sum(v)/len(v):.3f}")
pooled = [o[m:-m, m:-m] for o in outs]
a = ev(); nv = k.count_samples(views, finite_only=True); b = ev(); torch.cuda.synchronize()
print(f"count_samples: {a.elapsed_time(b):.2f} ms")
for _ in range(2): s = k.percentile(pooled, 99.0, take_abs=True)
a = ev(); s = k.percentile(pooled, 99.0, take_abs=True); b = ev(); torch.cuda.synchronize()
print(f"percentile (staged): {a.elapsed_time(b):.2f} ms -> {s}")
a = ev(); s2 = k.percentile_two_pass(pooled, 99.0, take_abs=True); b = ev(); torch.cuda.synchronize()
print(f"percentile (two pass): {a.elapsed_time(b):.2f} ms -> {s2}")
t0 = time.perf_counter(); a = ev()
st = ns.compute_norm_stats_device(d, "topousm_fast", {"radii": R, "weights": W6, "pixel_size": 1.0})
b = ev(); torch.cuda.synchronize()
print(f"compute_norm_stats_device: {a.elapsed_time(b):.2f} ms (wall {(time.perf_counter()-t0)*1e3:.2f}) -> {st}")
# one window alone, with empty-cache allocation excluded
ws = torch.empty(max(256, k.topousm_fast_workspace_bytes(views[4].shape, R, 1.0)), dtype=torch.uint8, device="cuda")
out = torch.empty(views[4].shape, dtype=torch.float32, device="cuda")
for _ in range(2): k.topousm_fast(views[4], radii=R, weights=W6, roi=(m, 4128, m, 4128), workspace=ws, out=out)
k.profile_enable(True)
a = ev(); k.topousm_fast(views[4], radii=R, weights=W6, roi=(m, 4128, m, 4128), workspace=ws, out=out); b = ev(); torch.cuda.synchronize()
prof = k.profile_read(); k.profile_enable(False)
print(f"one window (preallocated): {a.elapsed_time(b):.3f} ms; profile {prof}")
