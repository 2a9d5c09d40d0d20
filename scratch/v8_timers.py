"""Debug build (-DFSG_V8_TIMERS): per-warp busy / barrier-wait cycles of one v8 CTA."""
import ctypes as C, os, subprocess, sys
sys.path.insert(0, ".")
import torch
from fujishadergpu_b200 import kernels as k, _lib
S = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
W6 = [32 / 63, 16 / 63, 8 / 63, 4 / 63, 2 / 63, 1 / 63]
d = k.synth_dem((S, S))
R = [2, 8, 32, 128, 512, 2048]
for _ in range(2):
    k.topousm_fast(d, radii=R, weights=W6, norm_scale=14.65)
torch.cuda.synchronize()
lib = _lib.load()
buf = (C.c_ulonglong * 48)()
lib.fsg_debug_v8_timers.restype = C.c_int
print("rc", lib.fsg_debug_v8_timers(buf))
print("warp  busyA  waitA  busyB  waitB   (kcycles)")
for w in range(12):
    v = [buf[4 * w + q] / 1e3 for q in range(4)]
    print(f"{w:4d} " + " ".join(f"{x:7.1f}" for x in v))
