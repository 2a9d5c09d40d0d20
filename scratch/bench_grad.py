import sys, json
sys.path.insert(0, ".")
import torch
from fujishadergpu_b200 import kernels as k
from scratch.bench_algos import timeit, report
S = 16384
d = k.synth_dem((S, S))
b, m = timeit(lambda: k.hillshade(d, pixel_scale_x=1.0, pixel_scale_y=-1.0)); report("hillshade local f32", S, b, m)
b, m = timeit(lambda: k.slope(d, unit="degree", pixel_scale_x=1.0, pixel_scale_y=-1.0)); report("slope degree f32", S, b, m)
b, m = timeit(lambda: k.hillshade(d, pixel_scale_x=30.0, pixel_scale_y=-30.0)); report("hillshade px=30 (IEEE div)", S, b, m)
b, m = timeit(lambda: k.curvature(d, curvature_type="mean", pixel_scale_x=1.0, pixel_scale_y=-1.0)); report("curvature mean f32", S, b, m)
