"""Secondary timings (BASELINE.json configs 0,1,3 and the pre-pass pieces): Mpx/s and achieved GB/s at 8 B/px.
python scratch/bench_algos.py [--quick]   -> one JSON line per case on stdout"""
import json, sys, time
sys.path.insert(0, ".")
import torch
from fujishadergpu_b200 import kernels as k

PEAK = 6547.2
try:
    PEAK = float(json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"])
except Exception:
    pass


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[0], ts[len(ts) // 2]


def report(name, S, best, med, bpp=8.0):
    px = S * S
    print(json.dumps({"case": name, "size": S, "best_ms": round(best, 4), "median_ms": round(med, 4),
                      "mpx_s": round(px / med / 1e3, 1), "achieved_gbs": round(bpp * px / med / 1e6, 1),
                      "frac_of_measured_hbm": round(bpp * px / med / 1e6 / PEAK, 4), "bytes_per_px": bpp}), flush=True)


def main():
    quick = "--quick" in sys.argv
    S1, S2, S4 = (4096, 8192, 8192) if quick else (4096, 16384, 32768)
    d = k.synth_dem((S1, S1))
    b, m = timeit(lambda: k.hillshade(d, pixel_scale_x=1.0, pixel_scale_y=-1.0)); report("hillshade local f32", S1, b, m)
    b, m = timeit(lambda: k.hillshade(d, pixel_scale_x=1.0, pixel_scale_y=-1.0, output_dtype="uint8",
                                      qp={"a": 254.0, "b": 1.0, "lo": 1, "hi": 255})) if False else (0, 0)
    del d
    d = k.synth_dem((S2, S2))
    out = None
    b, m = timeit(lambda: k.hillshade(d, pixel_scale_x=1.0, pixel_scale_y=-1.0)); report("hillshade local f32", S2, b, m)
    b, m = timeit(lambda: k.slope(d, unit="degree", pixel_scale_x=1.0, pixel_scale_y=-1.0)); report("slope degree f32", S2, b, m)
    for ct in ("mean", "profile"):
        b, m = timeit(lambda: k.curvature(d, curvature_type=ct, pixel_scale_x=1.0, pixel_scale_y=-1.0)); report(f"curvature {ct} f32", S2, b, m)
    del d
    torch.cuda.empty_cache()
    d = k.synth_dem((S4, S4))
    b, m = timeit(lambda: k.openness(d, openness_type="positive", num_directions=8, max_distance=256, pixel_size=1.0), n=3, warm=1)
    report("openness positive 8 dir r=256", S4, b, m)
    b, m = timeit(lambda: k.decimate(d, 4), n=3, warm=1); report("decimate f=4 (pyramid)", S4, b, m, bpp=4.25)
    w = [32 / 63, 16 / 63, 8 / 63, 4 / 63, 2 / 63, 1 / 63]
    b, m = timeit(lambda: k.topousm_fast(d, radii=[2, 8, 32, 128, 512, 2048], weights=w, norm_scale=14.65), n=3, warm=1)
    report("topousm_fast 6 radii main pass", S4, b, m)
    b, m = timeit(lambda: k.topousm_fast(d, radii=[2, 8, 32], weights=[4 / 7, 2 / 7, 1 / 7], norm_scale=14.65), n=3, warm=1)
    report("topousm_fast radii 2,8,32 only", S4, b, m)
    b, m = timeit(lambda: k.topousm_fast(d, radii=[128, 512, 2048], weights=[4 / 7, 2 / 7, 1 / 7], norm_scale=14.65), n=3, warm=1)
    report("topousm_fast radii 128,512,2048 only", S4, b, m)
    b, m = timeit(lambda: k.topousm_fast(d, radii=[2], weights=[1.0], norm_scale=14.65), n=3, warm=1)
    report("topousm_fast radius 2 only", S4, b, m)
    b, m = timeit(lambda: k.topousm_fast(d, radii=[32], weights=[1.0], norm_scale=14.65), n=3, warm=1)
    report("topousm_fast radius 32 only", S4, b, m)


if __name__ == "__main__":
    main()
