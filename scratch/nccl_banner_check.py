import os, sys
sys.path.insert(0, ".")
import torch, torch.distributed as dist
import bench
local = int(os.environ.get("LOCAL_RANK", "0")); torch.cuda.set_device(local)
dev = torch.device("cuda", local)
with bench.stdout_to_stderr():
    dist.init_process_group("nccl", device_id=dev)
    dist.barrier()
t = torch.ones(4, device=dev); dist.all_reduce(t)
ops = [dist.P2POp(dist.isend, t, 1 - dist.get_rank()), dist.P2POp(dist.irecv, torch.empty_like(t), 1 - dist.get_rank())]
for r in dist.batch_isend_irecv(ops): r.wait()
torch.cuda.synchronize()
if dist.get_rank() == 0:
    print('{"ok": true, "sum": %d}' % int(t[0].item()))
dist.barrier(); dist.destroy_process_group()
