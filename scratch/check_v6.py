"""A/B check on the GPU: fused_kernel_v6 vs the general fused_kernel (bit-for-bit) on dense / NoData rasters, then timings."""
import os, sys, time
sys.path.insert(0, ".")
import torch
from fujishadergpu_b200 import kernels as k

W6 = [32 / 63, 16 / 63, 8 / 63, 4 / 63, 2 / 63, 1 / 63]


def run(d, radii, w, env, **kw):
    for key in ("FSG_FORCE_GENERIC", "FSG_FUSED_V5", "FSG_FUSED_V7", "FSG_NO_BULK", "FSG_V6_CFGB"):
        os.environ.pop(key, None)
    for key in env:
        os.environ[key] = "1"
    o = k.topousm_fast(d, radii=radii, weights=w, norm_scale=14.65, **kw)
    torch.cuda.synchronize()
    return o


def same(a, b):
    an, bn = torch.isnan(a), torch.isnan(b)
    if not torch.equal(an, bn):
        return False, f"nan masks differ ({int((an != bn).sum())} px)"
    ok = torch.equal(torch.where(an, torch.zeros_like(a), a), torch.where(bn, torch.zeros_like(b), b))
    if ok:
        return True, "bit-identical"
    diff = (torch.where(an, torch.zeros_like(a), a) - torch.where(bn, torch.zeros_like(b), b)).abs()
    return False, f"max diff {float(diff.max()):.3e} at {int((diff > 0).sum())} px"


fails = 0
cases = [((64, 5000), True), ((40, 300), False), ((1000, 1500), True), ((1000, 1500), False), ((3000, 2900), False), ((2051, 1797), True), ((4096, 4096), False),
         ((700, 5000), True), ((5000, 300), False), ((129, 264 * 3), True), ((4096 + 17, 2048 + 5), True)]
for shape, nod in cases:
    d = k.synth_dem(shape, seed=11 + shape[0], nodata=nod)
    if nod:   # a few isolated NaNs and a NaN block in the interior
        d[shape[0] // 2, shape[1] // 3] = float("nan")
        d[shape[0] // 3: shape[0] // 3 + 40, shape[1] // 2: shape[1] // 2 + 70] = float("nan")
    for radii, w in (([2, 8, 32, 128, 512, 2048], W6), ([2, 8, 32], [4 / 7, 2 / 7, 1 / 7]), ([128, 3, 17], [0.2, 0.5, 0.3]),
                     ([32], [1.0]), ([50, 2], [0.5, 0.5])):
        ref = run(d, radii, w, ["FSG_FORCE_GENERIC"])
        for env in ([], ["FSG_V6_CFGB"], ["FSG_V6_CFGB", "FSG_NO_BULK"]):
            got = run(d, radii, w, env)
            ok, msg = same(ref, got)
            if not ok:
                fails += 1
            print(("ok  " if ok else "FAIL"), shape, "nodata" if nod else "dense", radii, env, msg, flush=True)
            if not ok and "max diff" in msg:
                an = torch.isnan(ref)
                dd = (torch.where(an, torch.zeros_like(ref), ref) - torch.where(an, torch.zeros_like(got), got)).abs()
                idx = torch.nonzero(dd > 0)[:6].tolist()
                print("     first diffs (row, col):", idx, flush=True)
        if radii == [2, 8, 32, 128, 512, 2048]:
            for od in ("uint8", "int16"):
                qp = {"a_coef": 107.99319, "b_coef": 128.0, "dn_min": 1, "dn_max": 255} if od == "uint8" else {"a_coef": 27863.095, "b_coef": 0.0, "dn_min": -32767, "dn_max": 32767}
                try:
                    r8 = run(d, radii, w, ["FSG_FORCE_GENERIC"], output_dtype=od, qp=qp)
                    g8 = run(d, radii, w, [], output_dtype=od, qp=qp)
                    ok = torch.equal(r8, g8)
                    fails += 0 if ok else 1
                    print(("ok  " if ok else "FAIL"), shape, od, flush=True)
                except Exception as exc:
                    print("skip", od, repr(exc)[:120])
print("FAILS", fails)

S = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
d = k.synth_dem((S, S))
ws = torch.empty(max(256, k.topousm_fast_workspace_bytes((S, S), [2, 8, 32, 128, 512, 2048], 1.0)), dtype=torch.uint8, device="cuda")
out = torch.empty((S, S), dtype=torch.float32, device="cuda")
for name, env in (("v5", ["FSG_FUSED_V5"]), ("v6", []), ("v6 cfgB", ["FSG_V6_CFGB"]), ("v7", ["FSG_FUSED_V7"])):
    for radii, w in (([2, 8, 32, 128, 512, 2048], W6), ([2, 8, 32], [4 / 7, 2 / 7, 1 / 7]), ([128, 512, 2048], [4 / 7, 2 / 7, 1 / 7]), ([2], [1.0])):
        run(d, radii, w, env, workspace=ws, out=out)
        k.profile_enable(True)
        for _ in range(3):
            k.topousm_fast(d, radii=radii, weights=w, norm_scale=14.65, workspace=ws, out=out)
        torch.cuda.synchronize()
        prof = k.profile_read()
        k.profile_enable(False)
        fused = sorted(ms for tag, ms in prof if tag == 1)
        print(f"{name:10s} {str(radii):28s} fused kernel {fused[len(fused)//2]:8.3f} ms  ({S*S/fused[len(fused)//2]/1e6:7.1f} Gpx/s)", flush=True)
