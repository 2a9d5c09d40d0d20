"""A/B check on the GPU: fused_kernel_v8 (interior fast path + v6 borders / NaN blocks) vs fused_kernel_v6 alone,
bit for bit, on dense / NoData rasters, regions of interest and integer outputs; then timings."""
import os, sys
sys.path.insert(0, ".")
import torch
from fujishadergpu_b200 import kernels as k

W6 = [32 / 63, 16 / 63, 8 / 63, 4 / 63, 2 / 63, 1 / 63]
SW = ("FSG_FORCE_GENERIC", "FSG_FUSED_V5", "FSG_NO_BULK", "FSG_V6_CFGB", "FSG_NO_V8")


def run(d, radii, w, env, **kw):
    for key in SW:
        os.environ.pop(key, None)
    for key in env:
        os.environ[key] = "1"
    k.reload_debug_switches()
    o = k.topousm_fast(d, radii=radii, weights=w, norm_scale=kw.pop("norm_scale", 14.65), **kw)
    torch.cuda.synchronize()
    return o


def same(a, b):
    if a.dtype != torch.float32:
        ok = torch.equal(a, b)
        return ok, "identical" if ok else f"{int((a != b).sum())} px differ"
    an, bn = torch.isnan(a), torch.isnan(b)
    if not torch.equal(an, bn):
        return False, f"nan masks differ ({int((an != bn).sum())} px)"
    ok = torch.equal(torch.where(an, torch.zeros_like(a), a).view(torch.int32), torch.where(bn, torch.zeros_like(b), b).view(torch.int32))
    if ok:
        return True, "bit-identical"
    diff = (torch.where(an, torch.zeros_like(a), a) - torch.where(bn, torch.zeros_like(b), b)).abs()
    idx = torch.nonzero(diff > 0)[:6].tolist()
    return False, f"max diff {float(diff.max()):.3e} at {int((diff > 0).sum())} px, first {idx}"


fails = 0
quick = "--quick" in sys.argv
cases = [((700, 5000), False), ((3000, 2900), False), ((3000, 2900), True), ((2051, 1797), True), ((4096, 4096), False),
         ((4113, 2053), True), ((1024, 8192 + 64), False)]
if not quick:
    cases += [((8192, 8192), True), ((16384, 4096), False)]
for shape, nod in cases:
    d = k.synth_dem(shape, seed=11 + shape[0], nodata=nod)
    if nod:   # a few isolated NaNs and a NaN block in the interior
        d[shape[0] // 2, shape[1] // 3] = float("nan")
        d[shape[0] // 3: shape[0] // 3 + 40, shape[1] // 2: shape[1] // 2 + 70] = float("nan")
    for radii, w in (([2, 8, 32, 128, 512, 2048], W6), ([2, 8, 32], [4 / 7, 2 / 7, 1 / 7]), ([2, 8, 32, 128], [0.4, 0.3, 0.2, 0.1])):
        for ns in (14.65, None):
            ref = run(d, radii, w, ["FSG_NO_V8"], norm_scale=ns)
            got = run(d, radii, w, [], norm_scale=ns)
            ok, msg = same(ref, got)
            fails += 0 if ok else 1
            print(("ok  " if ok else "FAIL"), shape, "nodata" if nod else "dense", radii, "scale", ns, msg, flush=True)
        if len(radii) == 6:
            for od in ("uint8", "int16"):
                qp = {"a_coef": 107.99319, "b_coef": 128.0, "dn_min": 1, "dn_max": 255} if od == "uint8" else {"a_coef": 27863.095, "b_coef": 0.0, "dn_min": -32767, "dn_max": 32767}
                r8 = run(d, radii, w, ["FSG_NO_V8"], output_dtype=od, qp=qp)
                g8 = run(d, radii, w, [], output_dtype=od, qp=qp)
                ok, msg = same(r8, g8)
                fails += 0 if ok else 1
                print(("ok  " if ok else "FAIL"), shape, od, msg, flush=True)
            H, W = shape
            for roi in ((H // 5, H // 2, W // 7, W // 2), (0, H // 3, W // 2, W - W // 2), (H // 2, H - H // 2, 0, W // 3)):
                o1 = torch.zeros(shape, dtype=torch.float32, device="cuda")
                o2 = torch.zeros(shape, dtype=torch.float32, device="cuda")
                run(d, radii, w, ["FSG_NO_V8"], roi=roi, out=o1)
                run(d, radii, w, [], roi=roi, out=o2)
                sl = (slice(roi[0], roi[0] + roi[1]), slice(roi[2], roi[2] + roi[3]))
                ok, msg = same(o1[sl], o2[sl])
                fails += 0 if ok else 1
                print(("ok  " if ok else "FAIL"), shape, "roi", roi, msg, flush=True)
print("FAILS", fails, flush=True)

S = 16384
for a in sys.argv[1:]:
    if a.isdigit():
        S = int(a)
d = k.synth_dem((S, S))
ws = torch.empty(max(256, k.topousm_fast_workspace_bytes((S, S), [2, 8, 32, 128, 512, 2048], 1.0)), dtype=torch.uint8, device="cuda")
out = torch.empty((S, S), dtype=torch.float32, device="cuda")
for name, env in (("v6", ["FSG_NO_V8"]), ("v8", [])):
    for radii, w in (([2, 8, 32, 128, 512, 2048], W6), ([2, 8, 32], [4 / 7, 2 / 7, 1 / 7])):
        run(d, radii, w, env, workspace=ws, out=out)
        k.profile_enable(True)
        for _ in range(3):
            k.topousm_fast(d, radii=radii, weights=w, norm_scale=14.65, workspace=ws, out=out)
        torch.cuda.synchronize()
        prof = k.profile_read()
        k.profile_enable(False)
        fused = sorted(ms for tag, ms in prof if tag == 1)
        print(f"{name:10s} {str(radii):28s} fused pass {fused[len(fused)//2]:8.3f} ms  ({S*S/fused[len(fused)//2]/1e6:7.1f} Gpx/s)", flush=True)
