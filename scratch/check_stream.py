import sys, time
sys.path.insert(0, ".")
import torch
from fujishadergpu_b200 import kernels as k
from fujishadergpu_b200.core.tile_processor import HostTilePipeline, StreamedTopoPipeline
W6 = [32 / 63, 16 / 63, 8 / 63, 4 / 63, 2 / 63, 1 / 63]
params = {"radii": [2, 8, 32, 128, 512, 2048], "weights": W6, "pixel_size": 1.0}
for shape, nod, C in (((9000, 8300), False, 1024), ((9000, 8300), True, 2048), ((5000, 4111), False, 512)):
    d = k.synth_dem(shape, seed=77, nodata=nod)
    hin = torch.empty(shape, dtype=torch.float32).pin_memory(); hin.copy_(d)
    for od in ("uint8", "float32"):
        a = HostTilePipeline(shape, "topousm_fast", params, output_dtype=od)
        b = StreamedTopoPipeline(shape, params, output_dtype=od, chunk_rows=C)
        dt = torch.uint8 if od == "uint8" else torch.float32
        ha = torch.empty(shape, dtype=dt).pin_memory(); hb = torch.empty(shape, dtype=dt).pin_memory()
        a.run(hin, ha); b.run(hin, hb); b.run(hin, hb)
        if od == "uint8":
            ok = torch.equal(ha, hb)
        else:
            ok = torch.equal(torch.isnan(ha), torch.isnan(hb)) and torch.equal(torch.nan_to_num(ha), torch.nan_to_num(hb))
        print("stream", shape, nod, C, od, "ok" if ok else "FAIL", flush=True)
        del a, b
if len(sys.argv) > 1:
    S = int(sys.argv[1])
    d = k.synth_dem((S, S), seed=20261019)
    hin = torch.empty((S, S), dtype=torch.float32, pin_memory=True)
    for r in range(0, S, 4096):
        hin[r:r + 4096].copy_(d[r:r + 4096])
    hout = torch.empty((S, S), dtype=torch.uint8, pin_memory=True)
    del d
    for cls, kw in ((StreamedTopoPipeline, {}), (HostTilePipeline, None)):
        pipe = cls((S, S), params, output_dtype="uint8") if kw is not None else cls((S, S), "topousm_fast", params, output_dtype="uint8")
        pipe.run(hin, hout)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(2):
            pipe.run(hin, hout)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 2
        print(cls.__name__, S, f"{dt*1e3:.1f} ms  {S*S/dt/1e6:.0f} Mpx/s", flush=True)
        del pipe
        torch.cuda.empty_cache()
