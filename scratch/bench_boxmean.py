"""Times the decimated-grid box means: single pass (box_mean2d_kernel) vs the two passes, and the pyramid."""
import os, sys
sys.path.insert(0, ".")
import torch
from fujishadergpu_b200 import kernels as k

def timeit(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

for (h, w, size) in ((16384, 16384, 65), (4096, 4096, 65), (4096, 4096, 257), (2064, 2064, 65), (516, 516, 65), (516, 516, 257),
                     (2048, 16384, 65), (512, 4096, 257)):
    g = torch.rand((h, w), device="cuda") * 900 + 100
    for two in (True, False):
        os.environ.pop("FSG_BOX_TWO_PASS", None)
        if two:
            os.environ["FSG_BOX_TWO_PASS"] = "1"
        k.reload_debug_switches()
        ms = timeit(lambda: k.grid_mean_band(g, 0, h, size, 0, h))
        print(f"{h}x{w} size {size:3d} {'two-pass' if two else 'one-pass'}: {ms:7.3f} ms  {h*w/ms/1e6:7.2f} Gcell/s  ({8*h*w/ms/1e6:.0f} GB/s algorithmic)")
os.environ.pop("FSG_BOX_TWO_PASS", None); k.reload_debug_switches()
