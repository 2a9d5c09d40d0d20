#!/usr/bin/env python
"""bench.py -- headline benchmark: Mpx/s of topousm_fast on a synthetic 65536^2 float32 DEM.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size S]

One "step" = one whole job over the raster: the stratified-window statistics pre-pass (9 windows)
followed by the fused main pass (pyramid -> coarse means -> fused full-resolution kernel with
normalisation).  `value` is measured with the DEM resident in HBM (CUDA events, max over ranks);
`e2e` runs the same job through the host-buffer API (pinned host DEM -> device -> uint8 result back
to pinned host memory) with the copies inside the timed region.

`--impl reference` times the CPU arm: the reference has no CPU path and cannot be installed here
(CuPy/Dask/rasterio are not in the image), so -- per the task contract -- it is the oracle port of
the reference's block function (oracle/terrain_oracle.py, NumPy/SciPy) on all host cores, on a
bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RADII = [2, 8, 32, 128, 512, 2048]
METRIC = "Mpx/s topousm_fast on 65536^2 f32 DEM"
UNIT = "Mpx/s"
ALGO_BYTES_PER_PX = 8.0  # read f32 DEM once + write f32 result once


def _weights():
    raw = [2.0 ** (len(RADII) - 1 - i) for i in range(len(RADII))]
    return [v / sum(raw) for v in raw]


def host_cores() -> int:
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except AttributeError:
        pass
    try:
        q, p = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = min(n, max(1, -(-int(q) // int(p))))
    except Exception:
        pass
    return max(1, n)


# ------------------------------------------------------------------------------------------------
# CPU arm (oracle port of the reference block function)
# ------------------------------------------------------------------------------------------------
def _cpu_tile(args):
    seed, side = args
    from oracle import terrain_oracle as orc
    dem = orc.synth_dem(side, side, seed=seed)
    raw = orc.topousm_fast_block(dem, radii=RADII, weights=_weights())
    st = orc.abs_p99_scale(raw)
    out = orc.normalise_by_scale(raw, st)
    return float(out[side // 2, side // 2])


def cpu_throughput(cores: int, side: int, tiles: int):
    """Mpx/s of the oracle on `tiles` independent side x side tiles over `cores` processes."""
    import multiprocessing as mp
    jobs = [(1000 + i, side) for i in range(tiles)]
    t0 = time.perf_counter()
    if cores == 1:
        for j in jobs:
            _cpu_tile(j)
    else:
        with mp.get_context("spawn").Pool(cores) as pool:
            pool.map(_cpu_tile, jobs)
    dt = time.perf_counter() - t0
    return tiles * side * side / dt / 1e6, dt


def run_reference_arm(a) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    side = 2048
    tiles = max(cores, 8)
    vals = []
    for _ in range(max(0, a.warmup > 0)):
        cpu_throughput(cores, side, min(tiles, cores))
    for _ in range(max(1, min(a.steps, 3))):
        v, dt = cpu_throughput(cores, side, tiles)
        vals.append((v, dt))
    v = sum(x for x, _ in vals) / len(vals)
    ms = 1e3 * sum(d for _, d in vals) / len(vals)
    sample = f"{tiles} independent {side}x{side} tiles of the same synthetic DEM family per step, oracle port"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": len(vals),
        "warmup": int(a.warmup > 0), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32 (f64 accumulators)", "data": "synthetic",
        "config": {"workload": "topousm_fast --mode spatial radii 2,8,32,128,512,2048 weights 2^n (CPU sample)",
                   "radii": RADII, "sample": sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


PROF_TAGS = {1: "fused_main", 2: "pyramid", 3: "coarse_means", 6: "fused_stats_windows"}


def kernel_breakdown(prof, steps):
    """Device milliseconds per step of the library's profiler tags (events around the launches, own stream)."""
    out = {}
    for tag, ms in prof:
        name = PROF_TAGS.get(tag)
        if name:
            out[name] = out.get(name, 0.0) + ms / max(1, steps)
    return {k: round(v, 3) for k, v in out.items()}


def reference_tools():
    """What a GPU-vs-CPU comparison outside this repo could use on the box (none ship with the image)."""
    import importlib.util
    import shutil
    return {"gdaldem": shutil.which("gdaldem"), "cupy": importlib.util.find_spec("cupy") is not None,
            "dask": importlib.util.find_spec("dask") is not None, "osgeo": importlib.util.find_spec("osgeo") is not None}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class stdout_to_stderr:
    """Context manager: C-level and Python-level writes to stdout go to stderr inside the block."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


def run_gpu_arm(a) -> None:
    import torch
    import torch.distributed as dist
    from fujishadergpu_b200 import kernels as k
    from fujishadergpu_b200.algorithms._norm_stats import compute_norm_stats_device

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner on stdout while the communicator comes up: send file descriptor 1 to
        # stderr for that moment so that stdout carries the one JSON line only
        with stdout_to_stderr():
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
    if world > 1:
        from fujishadergpu_b200.core import sharding
        return sharding.bench_sharded(a, dist, dev, METRIC, UNIT, RADII, _weights(), clock_sampler=ClockSampler,
                                      peak=measured_peak_gbs())

    S = int(a.size)
    H = W = S
    weights = _weights()
    params = {"radii": RADII, "weights": weights, "pixel_size": 1.0}
    dem = k.synth_dem((H, W), seed=20261017 + 2, device=dev)
    out = torch.empty((H, W), dtype=torch.float32, device=dev)
    ws = torch.empty(max(256, k.topousm_fast_workspace_bytes((H, W), RADII, 1.0)), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()

    from fujishadergpu_b200.core import sharding as sh

    redo = [0]

    def step():
        # one rank of the sharded orchestration (the same code as N > 1): the scale-independent part of the main
        # pass (pyramid, coarse means) runs on a side stream underneath the statistics pre-pass; no host
        # synchronisation inside the step (the p99 scale stays on the device, planning values are speculated and
        # checked once everything is enqueued -- a wrong guess repeats the step, counted in `respeculated_steps`)
        for _attempt in range(3):
            _res, scale_dev, spec = sh.topousm_fast_sharded_step(dem, H, 0, 1, radii=RADII, weights=weights,
                                                                 pixel_size=1.0, out=out, dem_ext=dem)
            if spec.ok():
                break
            redo[0] += 1
        return scale_dev

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    redo[0] = 0
    sampler = ClockSampler(local)
    sampler.start()
    k.reset_launch_count()
    k.profile_enable(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(a.steps):
        scale_dev = step()
    ev1.record()
    torch.cuda.synchronize()
    total_ms = ev0.elapsed_time(ev1)
    st = (float(scale_dev.item()),)
    if not (st[0] == st[0]):   # no valid statistics window (cannot happen on the synthetic DEM)
        raise RuntimeError("the statistics pre-pass found no valid window")
    # bit pattern checksum of the result (sum of the uint32 words): identical for every N proves bit-identity
    checksum = 0
    for r in range(0, H, 8192):
        checksum += int(out[r:r + 8192].view(torch.int32).sum(dtype=torch.int64).item())
    checksum &= (1 << 64) - 1
    launches = k.launch_count()
    prof = k.profile_read(max_records=4096)
    k.profile_enable(False)
    clocks = sampler.stop()
    ms_step = total_ms / a.steps
    value = H * W / (ms_step * 1e-3) / 1e6

    # main-pass-only timing (stats known), for the breakdown
    torch.cuda.synchronize()
    m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    m0.record()
    for _ in range(a.steps):
        k.topousm_fast(dem, radii=RADII, weights=weights, pixel_size=1.0, norm_scale=float(st[0]), workspace=ws, out=out)
    m1.record()
    torch.cuda.synchronize()
    main_ms = m0.elapsed_time(m1) / a.steps

    # roofline of the dominant kernel: the fused full-resolution kernel of the MAIN pass (the largest
    # launch of each step; the nine stats windows launch the same kernel on 8256^2 windows)
    main_fused = [ms for tag, ms in prof if tag == 1]   # tag 6 = the regions of interest of the statistics windows
    if len(main_fused) != a.steps + redo[0]:
        raise RuntimeError(f"profiler recorded {len(main_fused)} main fused passes for {a.steps} steps")
    fused_ms = sum(main_fused) / max(1, len(main_fused))
    peak, peak_src = measured_peak_gbs()
    achieved = ALGO_BYTES_PER_PX * H * W / (fused_ms * 1e-3) / 1e9 if fused_ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "fused_kernel_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic = tj.get("dram_bytes_per_px", None)
            traffic = traffic * H * W if traffic is not None else None
        except Exception:
            traffic = None

    # ---- the same main pass on a DEM with NoData (a wedge along the left edge and three ellipses: 4 % of the pixels;
    # their 256-row blocks leave the interior fast path for the general kernel) -- reported, not the headline
    nodata_ms = None
    try:
        if a.no_nodata_variant:
            raise RuntimeError("skipped (--no-nodata-variant)")
        k.synth_dem((H, W), seed=20261017 + 2, device=dev, nodata=True, out=dem)
        for _ in range(2):
            k.topousm_fast(dem, radii=RADII, weights=weights, pixel_size=1.0, norm_scale=float(st[0]), workspace=ws, out=out)
        torch.cuda.synchronize()
        n0, n1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0.record()
        for _ in range(3):
            k.topousm_fast(dem, radii=RADII, weights=weights, pixel_size=1.0, norm_scale=float(st[0]), workspace=ws, out=out)
        n1.record()
        torch.cuda.synchronize()
        nodata_ms = {"main_pass_ms": n0.elapsed_time(n1) / 3,
                     "nodata_fraction": float(torch.isnan(dem[::16, ::16]).float().mean().item())}
    except Exception as exc:  # noqa: BLE001
        nodata_ms = {"error": repr(exc)}
    k.synth_dem((H, W), seed=20261017 + 2, device=dev, out=dem)   # back to the dense DEM for the host-buffer leg

    # ---- e2e: host buffers through the public host API (uint8 result) ----
    e2e = None
    try:
        if a.no_e2e:
            raise RuntimeError("skipped (--no-e2e)")
        from fujishadergpu_b200.core.tile_processor import StreamedTopoPipeline
        del out
        torch.cuda.empty_cache()
        e2e_side = S
        pipe = StreamedTopoPipeline((e2e_side, e2e_side), params, output_dtype="uint8", device=dev)
        hin = torch.empty((e2e_side, e2e_side), dtype=torch.float32, pin_memory=True)
        for r in range(0, e2e_side, 4096):
            hin[r:r + 4096].copy_(dem[r:r + 4096])
        hout = torch.empty((e2e_side, e2e_side), dtype=torch.uint8, pin_memory=True)
        torch.cuda.synchronize()
        pipe.run(hin, hout)  # warm-up
        n_e2e = max(1, min(a.steps, 3))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            pipe.run(hin, hout)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n_e2e
        e2e = {"value": e2e_side * e2e_side / dt / 1e6, "unit": UNIT, "h2d_bytes_per_step": pipe.h2d_bytes,
               "d2h_bytes_per_step": pipe.d2h_bytes, "ms_per_step": dt * 1e3, "output_dtype": "uint8",
               "pipeline": "StreamedTopoPipeline: chunked H2D (stats-window rows first), per-chunk decimation, global scale while uploading, row bands + D2H overlapped",
               "steps": n_e2e}
        del pipe, hin, hout
    except Exception as exc:  # pinned allocation can fail on small hosts; say so instead of faking
        e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "error": repr(exc)[:200]}

    # ---- CPU baseline: the oracle port, one process, bounded sample ----
    cpu = None
    if not a.no_cpu_baseline:
        v, dt = cpu_throughput(1, 2048, 3)
        cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"3 tiles of 2048x2048, same radii/weights, NumPy/SciPy oracle, {dt:.1f} s"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32 (f64 window sums)", "data": "synthetic",
        "config": {"workload": f"topousm_fast --mode spatial, radii 2,8,32,128,512,2048, 2^n weights, {S}x{S} f32 DEM, "
                               "f32 output, stats pre-pass (9 stratified 8256^2 windows) + main pass per step",
                   "radii": RADII, "size": S, "l2_policy": "inputs (16 GiB) far larger than the 126 MB L2",
                   "main_pass_ms": main_ms, "stats_prepass_ms": ms_step - main_ms,
                   "main_pass_mpx_s": H * W / (main_ms * 1e-3) / 1e6,
                   "fused_kernel_ms": fused_ms, "kernel_ms_per_step": kernel_breakdown(prof, a.steps),
                   "nodata_variant": nodata_ms, "reference_tools_on_this_box": reference_tools(),
                   "scale_p99": float(st[0]), "respeculated_steps": redo[0],
                   "out_checksum": f"{checksum:016x}"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": "fsg::fused_kernel_v8<3> + fused_kernel_v6 on the raster borders (whole fused pass of the main pass)", "peak_source": peak_src,
                     "algorithmic_bytes_per_px": ALGO_BYTES_PER_PX},
        "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=65536)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-nodata-variant", action="store_true", help="skip the NoData main-pass timing (profiling runs)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs)")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup
    if a.impl == "reference":
        return run_reference_arm(a)
    return run_gpu_arm(a)


if __name__ == "__main__":
    main()
