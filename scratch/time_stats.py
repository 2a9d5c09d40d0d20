import sys, time
sys.path.insert(0, ".")
import torch
from fujishadergpu_b200 import kernels as k
from fujishadergpu_b200.algorithms import _norm_stats as ns
W6 = [32 / 63, 16 / 63, 8 / 63, 4 / 63, 2 / 63, 1 / 63]
d = k.synth_dem((65536, 65536), seed=20261019)
p = {"radii": [2, 8, 32, 128, 512, 2048], "weights": W6, "pixel_size": 1.0}
for _ in range(2):
    ns.compute_norm_stats_device(d, "topousm_fast", p)
torch.cuda.synchronize()
# wall-clock pieces
import contextlib
marks = []
orig_order = k.order_stats
orig_topo = k.topousm_fast
def timed(name, fn):
    def w(*a, **kw):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = fn(*a, **kw)
        torch.cuda.synchronize(); marks.append((name, (time.perf_counter() - t0) * 1e3))
        return r
    return w
k.order_stats = timed("order_stats", orig_order)
k.topousm_fast = timed("topousm_fast(roi)", orig_topo)
ns._k = k
torch.cuda.synchronize(); t0 = time.perf_counter()
ns.compute_norm_stats_device(d, "topousm_fast", p)
torch.cuda.synchronize(); tot = (time.perf_counter() - t0) * 1e3
import collections
agg = collections.defaultdict(float); cnt = collections.Counter()
for n, t in marks:
    agg[n] += t; cnt[n] += 1
print("total (with syncs)", round(tot, 2), {n: (cnt[n], round(t, 2)) for n, t in agg.items()})
