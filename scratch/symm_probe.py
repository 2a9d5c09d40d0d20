"""Feasibility probe: torch symmetric memory (peer pointers over NVLink) vs NCCL send/recv for the window gather."""
import os, sys, time
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
import torch.distributed._symmetric_memory as symm
rows, W = 8192, 65536
try:
    t = symm.empty((rows, W), dtype=torch.float32, device=dev)
    hdl = symm.rendezvous(t, dist.group.WORLD)
    print(rank, "rendezvous ok", "ptrs", [hex(p) for p in hdl.buffer_ptrs][:2], flush=True)
except Exception as e:
    print(rank, "SYMM FAILED", repr(e), flush=True)
    dist.destroy_process_group(); sys.exit(0)
t.fill_(float(rank + 1))
torch.cuda.synchronize(); hdl.barrier(); torch.cuda.synchronize()
peer = (rank + 1) % world
pv = hdl.get_buffer(peer, (rows, W), torch.float32)
dst = torch.empty((4127, 8256), dtype=torch.float32, device=dev)
def timeit(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ms = timeit(lambda: dst.copy_(pv[100:4227, 6794:6794 + 8256]))
print(rank, f"peer strided copy 136 MB: {ms:.3f} ms = {dst.numel()*4/ms/1e6:.0f} GB/s; value ok {bool((dst == peer + 1).all())}", flush=True)
sys.path.insert(0, ".")
from fujishadergpu_b200 import kernels as k
ms = timeit(lambda: k.copy_rect(dst, pv[100:4227, 6794:6794 + 8256]))
print(rank, f"peer copy-engine 2-D copy 136 MB: {ms:.3f} ms = {dst.numel()*4/ms/1e6:.0f} GB/s; value ok {bool((dst == peer + 1).all())}", flush=True)
flat = torch.empty(4127 * 8256, dtype=torch.float32, device=dev)
ms = timeit(lambda: flat.copy_(pv.view(-1)[: flat.numel()]))
print(rank, f"peer contiguous copy 136 MB: {ms:.3f} ms = {flat.numel()*4/ms/1e6:.0f} GB/s", flush=True)
# NCCL send/recv of the same volume (ring)
src = torch.empty_like(flat)
def nccl():
    ops = [dist.P2POp(dist.isend, src, (rank + 1) % world), dist.P2POp(dist.irecv, flat, (rank - 1) % world)]
    for w in dist.batch_isend_irecv(ops): w.wait()
ms = timeit(nccl)
print(rank, f"NCCL send/recv 136 MB: {ms:.3f} ms = {flat.numel()*4/ms/1e6:.0f} GB/s", flush=True)
ms = timeit(lambda: hdl.barrier())
print(rank, f"symm barrier: {ms*1e3:.1f} us", flush=True)
x = torch.zeros(2049, dtype=torch.int64, device=dev)
ms = timeit(lambda: dist.all_reduce(x))
print(rank, f"NCCL all_reduce 16 KB: {ms*1e3:.1f} us", flush=True)
dist.destroy_process_group()
