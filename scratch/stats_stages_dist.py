"""Stage timings of sharded_topousm_scale under torchrun (wall clock with device syncs, rank 0 prints)."""
import os, sys, time
sys.path.insert(0, ".")
import torch, torch.distributed as dist
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
local = int(os.environ.get("LOCAL_RANK", "0")); torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from fujishadergpu_b200 import kernels as k
from fujishadergpu_b200.core import sharding as sh
world, rank = dist.get_world_size(), dist.get_rank()
H = W = 65536
R = [2, 8, 32, 128, 512, 2048]; Wt = [32/63,16/63,8/63,4/63,2/63,1/63]
r0, r1 = sh.band_bounds(H, world)[rank]
ext, band = sh.haloed_band(H, W, world, rank, R, device=dev)
k.synth_dem((r1 - r0, W), seed=20261019, device=dev, row0=r0, h_global=H, out=band)
marks = []
def mark(name):
    torch.cuda.synchronize(); marks.append((name, time.perf_counter()))
orig_plan, orig_run, orig_conc, orig_pct = sh.plan_exchange, sh.run_exchanges, None, sh.distributed_percentile
from fujishadergpu_b200 import _device as _dev
orig_conc = _dev.run_concurrently
def run_ex(plans, d=None):
    mark("plans built"); o = orig_run(plans, d); mark("exchange done"); return o
def conc(fns, device, n_streams=3):
    o = orig_conc(fns, device, n_streams); mark("windows computed"); return o
def pct(*a, **kw):
    o = orig_pct(*a, **kw); mark("percentile done"); return o
sh.run_exchanges = run_ex; _dev.run_concurrently = conc; sh.distributed_percentile = pct
for it in range(4):
    dist.barrier(); marks.clear(); mark("start")
    s = sh.sharded_topousm_scale(band, H, rank, world, radii=R, weights=Wt, dist=dist)
    mark("end")
    if rank == 0 and it >= 2:
        t0 = marks[0][1]
        print(f"iter {it} scale {s}: " + "  ".join(f"{n} {1e3*(t-t0):.2f}" for n, t in marks[1:]), flush=True)
dist.barrier(); dist.destroy_process_group()
