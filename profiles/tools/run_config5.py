"""BASELINE config 5: topousm_fast on a 131072^2 f32 DEM (64 GiB), row-band sharded over N GPUs, uint8 COG output.

    python profiles/tools/run_config5.py [--size 131072] [--out /tmp/config5.tif]                      (N = 1)
    python -m torch.distributed.run --nproc-per-node N ... profiles/tools/run_config5.py --gpus N      (N = 2, 4, 8)

Prints one JSON line: Mpx/s of the compute step (statistics pre-pass + main pass, CUDA events, max over ranks), bytes
and device time of the NVLink exchanges per step, the COG writer's phases, checksums of the uint8 result (equal for
every N = bit-identical output) and a read-back check of the file against the device result.
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--size", type=int, default=131072)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--out", default="/tmp/config5.tif")
    ap.add_argument("--no-write", action="store_true")
    a = ap.parse_args()
    from fujishadergpu_b200 import kernels as k
    from fujishadergpu_b200.core import sharding as sh
    from fujishadergpu_b200.io.cog_sharded import tile_aligned_bounds, write_cog_sharded
    from fujishadergpu_b200.io.geotiff_reader import read_geotiff
    from fujishadergpu_b200.io.output_encoding import quantize_params, resolve_output_range
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    d = None
    if world > 1:
        saved = os.dup(1); os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()
        sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)
        d = dist
    S = int(a.size)
    H = W = S
    radii = [2, 8, 32, 128, 512, 2048]
    raw = [2.0 ** (5 - i) for i in range(6)]
    weights = [v / sum(raw) for v in raw]
    own = sh.band_bounds(H, world)
    assert own == tile_aligned_bounds(H, world), "band boundaries must be multiples of the 512-row COG blocks"
    r0, r1 = own[rank]
    ext, band = sh.haloed_band(H, W, world, rank, radii, device=dev,
                               peer_group=dist.group.WORLD if (world > 1 and not os.environ.get("FSG_NO_PEER")) else None)
    k.synth_dem((r1 - r0, W), seed=20261017 + 5, device=dev, row0=r0, h_global=H, out=band)
    qp = quantize_params(*resolve_output_range("topousm_fast"), "uint8")
    out8 = torch.empty((r1 - r0, W), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()

    def step():
        for _ in range(3):
            _res, scale_dev, spec = sh.topousm_fast_sharded_step(band, H, rank, world, radii=radii, weights=weights, dist=d,
                                                                 output_dtype="uint8", qp=qp, out=out8, dem_ext=ext)
            if spec.ok():
                break
        return scale_dev

    step(); step()
    torch.cuda.synchronize()
    if d is not None:
        d.barrier()
    sh.EXCHANGE_LOG = {"sent": 0, "received": 0, "events": []}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        scale_dev = step()
    e1.record()
    torch.cuda.synchronize()
    log, sh.EXCHANGE_LOG = sh.EXCHANGE_LOG, None
    ms = torch.tensor([e0.elapsed_time(e1) / a.steps], dtype=torch.float64, device=dev)
    xms = torch.tensor([sum(x.elapsed_time(y) for x, y in log["events"]) / a.steps], dtype=torch.float64, device=dev)
    xb = torch.tensor([log["sent"] / a.steps, log["received"] / a.steps], dtype=torch.float64, device=dev)
    # checksums of the uint8 band: plain sum and a position-weighted sum (global row / column)
    cs = torch.zeros(2, dtype=torch.int64, device=dev)
    colw = (torch.arange(W, device=dev, dtype=torch.int64) % 251) + 1
    for y in range(0, r1 - r0, 4096):
        blk = out8[y:y + 4096].to(torch.int64)
        roww = ((torch.arange(r0 + y, r0 + y + blk.shape[0], device=dev, dtype=torch.int64) % 241) + 1).unsqueeze(1)
        cs[0] += blk.sum()
        cs[1] += (blk * colw * roww).sum()
        del blk
    if d is not None:
        d.all_reduce(ms, op=d.ReduceOp.MAX); d.all_reduce(xms, op=d.ReduceOp.MAX)
        d.all_reduce(xb, op=d.ReduceOp.SUM); d.all_reduce(cs, op=d.ReduceOp.SUM)
    wstats, check = None, None
    if not a.no_write:
        t0 = time.perf_counter()
        try:
            wstats = write_cog_sharded(a.out, out8, H, rank, world, dist=d)
            wstats["total_s"] = time.perf_counter() - t0
        except OSError as exc:    # e.g. no space left on the scratch file system
            wstats = {"error": repr(exc)[:200]}
        if rank == 0 and "error" not in wstats:   # read back: a strip of full-resolution tiles of rank 0's band and the smallest overview
            t0 = time.perf_counter()
            rows = min(1024, r1 - r0)
            got, meta = read_geotiff(a.out, window=(0, rows, 0, min(W, 8192))) if "window" in read_geotiff.__code__.co_varnames else (None, None)
            if got is not None:
                want = out8[:rows, :min(W, 8192)].cpu().numpy()
                check = {"window_equal": bool(np.array_equal(got, want)), "nodata": meta.get("nodata"),
                         "read_s": time.perf_counter() - t0}
    if rank == 0:
        px = H * W
        line = {"config": 5, "workload": f"topousm_fast, radii 2..2048, {S}x{S} f32 DEM ({px * 4 / 2**30:.0f} GiB), row bands on {world} GPU(s), uint8 COG",
                "n_gpus": world, "steps": a.steps, "ms_per_step": float(ms.item()), "mpx_s": px / (float(ms.item()) * 1e-3) / 1e6,
                "scale_p99": float(scale_dev.item()),
                "exchange": {"bytes_sent_per_step_all_ranks": float(xb[0].item()), "bytes_received_per_step_all_ranks": float(xb[1].item()),
                             "device_ms_per_step_max_rank": float(xms.item()),
                             "gb_s_per_rank": (float(xb[1].item()) / max(1, world)) / max(1e-9, float(xms.item()) * 1e-3) / 1e9},
                "checksum": {"sum": int(cs[0].item()), "weighted": int(cs[1].item())}, "cog": wstats, "readback": check}
        print(json.dumps(line))
    if d is not None:
        d.barrier()
        d.destroy_process_group()


if __name__ == "__main__":
    main()
