"""One COG from the row bands of several GPUs (BASELINE config 5: topousm_fast on 131072^2, uint8 COG output).

The reference gathers every chunk to ONE writer thread that hands it to GDAL's COG driver
(core/dask_processor.py:486-505, 712 write_cog_da_chunked; creation options :201-228).  Here every rank keeps its
band on its own GPU and does its own share of the file:

    1. AVERAGE overview cascade of the own band on the device (fsg_overview_average).  Bands start at multiples of the
       block size, so for the first levels a band's overview rows are whole tile rows of that level ("distributed"
       levels); the small levels above are gathered to rank 0.
    2. prediction + ZSTD of the own 512 x 512 tiles on the rank's host threads (device -> pinned host, one tile row
       at a time).
    3. the per-tile byte counts are all-gathered; every rank derives the same layout (IFDs first, then tile data
       from the smallest overview to full resolution, row-major tiles, GDAL's block leaders / trailers) and writes
       its tiles at their offsets with pwrite; rank 0 writes the header, the IFDs and the gathered levels.

The file is byte-identical to the one io/cog_writer.write_cog writes from the whole raster on one GPU (tested).
"""
from __future__ import annotations

import os
import struct
import time
from concurrent.futures import ThreadPoolExecutor
from typing import Optional

import numpy as np

from . import cog_writer as cw


def tile_aligned_bounds(H: int, world: int, blocksize: int = 512):
    """Row bands whose boundaries are multiples of the block size (core/sharding.band_bounds with align=blocksize)."""
    from ..core.sharding import band_bounds
    return band_bounds(H, world, align=blocksize)


def write_cog_sharded(path: str, band, H: int, rank: int, world: int, *, dist=None, nodata="auto", transform=None,
                      epsg=None, blocksize: int = 512, level: int = 1, overview_count: int = 8,
                      num_threads: Optional[int] = None, gdal_ghost: bool = True) -> dict:
    """`band` = this rank's rows tile_aligned_bounds(H, world)[rank] of the H x W result (device tensor, uint8 / int16 /
    float32).  Every rank calls this with the same `path` (a file system all ranks see).  Returns timing / size
    statistics (rank 0: of the whole file)."""
    import torch
    from .. import _device as _dev
    from .. import kernels as _k
    t = _dev.as_tensor(band)
    if t.ndim != 2 or t.dtype not in (torch.uint8, torch.int16, torch.float32):
        raise ValueError("write_cog_sharded: 2-D uint8 / int16 / float32 device band expected")
    bs = int(blocksize)
    W = int(t.shape[1])
    own = tile_aligned_bounds(H, world, bs)
    r0, r1 = own[rank]
    if int(t.shape[0]) != r1 - r0:
        raise ValueError("write_cog_sharded: the band does not match tile_aligned_bounds()")
    if nodata == "auto":
        nodata = float("nan") if t.dtype == torch.float32 else 0
    dt = np.dtype(str(t.dtype).replace("torch.", ""))
    pred = cw.predictor_for_dtype(dt)
    comp = cw.COMPRESSION_ZSTD
    fill = np.nan if dt.kind == "f" else (0 if nodata is None else nodata)
    n_ov = cw.overview_levels((H, W), overview_count)
    shapes = [(H, W)]
    for _ in range(n_ov):
        shapes.append(((shapes[-1][0] + 1) // 2, (shapes[-1][1] + 1) // 2))
    multi = dist is not None and world > 1

    # ---- 1. overview cascade of the own band; which levels stay distributed
    t0 = time.perf_counter()
    starts = [[a for (a, _b) in own]]          # band start rows per level
    ends = [[b for (_a, b) in own]]
    distributed = [True]
    for k in range(1, n_ov + 1):
        prev_s, prev_e = starts[-1], ends[-1]
        ok = distributed[-1] and all(s % 2 == 0 for s in prev_s)
        s_k = [s // 2 for s in prev_s]
        e_k = [(e + 1) // 2 for e in prev_e]
        ok = ok and all(s % bs == 0 for s in s_k) and all(e_k[q] == s_k[q + 1] for q in range(world - 1))
        starts.append(s_k); ends.append(e_k); distributed.append(bool(ok))
    # the first level that is no longer distributed is gathered to rank 0, which continues the cascade alone
    first_g = next((k for k in range(1, n_ov + 1) if not distributed[k]), None)
    lv = [t]
    for k in range(1, (first_g if first_g is not None else n_ov) + 1):
        # a band's overview equals the same rows of the raster's overview as long as the band starts on an even row
        if any(s_ % 2 for s_ in starts[k - 1]):
            raise ValueError("write_cog_sharded: band boundaries must stay even down to the first gathered level")
        lv.append(_k.overview_average(lv[-1], None if t.dtype == torch.float32 else nodata))
    gathered = {}
    if first_g is not None:
        src = lv[first_g]
        if src is None:
            raise ValueError("write_cog_sharded: band boundaries are not even rows of the first gathered level")
        if multi:
            parts = [None] * world
            dist.all_gather_object(parts, src.cpu().numpy())
            if rank == 0:
                whole = torch.from_numpy(np.concatenate(parts, axis=0)).to(t.device)
        else:
            whole = src
        if rank == 0:
            gathered[first_g] = whole
            for k in range(first_g + 1, n_ov + 1):
                gathered[k] = _k.overview_average(gathered[k - 1], None if t.dtype == torch.float32 else nodata)
    torch.cuda.synchronize(t.device)
    t_ov = time.perf_counter() - t0

    # ---- 2. compress the own tiles (distributed levels) and, on rank 0, the gathered levels
    t0 = time.perf_counter()
    nthreads = int(num_threads or max(1, min(32, (os.cpu_count() or 1) // max(1, world))))
    pin = torch.empty((bs, W), dtype=t.dtype, pin_memory=True)
    blobs = {}      # level -> list of blobs of this rank's tiles, row major
    with ThreadPoolExecutor(max_workers=nthreads) as pool:
        def level_blobs(a):
            out = []
            h, w = int(a.shape[0]), int(a.shape[1])
            for y0 in range(0, h, bs):
                y1 = min(h, y0 + bs)
                view = pin[: y1 - y0, :w]
                view.copy_(a[y0:y1], non_blocking=False)
                out.extend(cw.encode_tile_rows(view.numpy(), dt, bs, fill, pred, comp, level, pool))
            return out

        for k in range(0, n_ov + 1):
            if distributed[k]:
                blobs[k] = level_blobs(lv[k]) if r1 > r0 else []
            elif rank == 0:
                blobs[k] = level_blobs(gathered[k])
    t_comp = time.perf_counter() - t0

    # ---- 3. layout: every rank derives it from the all-gathered byte counts
    t0 = time.perf_counter()
    counts_mine = {k: [len(b) for b in v] for k, v in blobs.items()}
    if multi:
        all_counts = [None] * world
        dist.all_gather_object(all_counts, counts_mine)
    else:
        all_counts = [counts_mine]
    ifds, grids, ifd_off, data_start = cw._make_ifds(shapes, dt, comp, pred, bs, True, nodata, transform, epsg, gdal_ghost)
    extra = 8 if gdal_ghost else 0       # 4-byte leader + 4-byte trailer around every tile
    cur = data_start
    my_pos = {}
    for k in range(n_ov, -1, -1):          # smallest overview first, full resolution last
        offs, cnts = [], []
        owners = range(world) if distributed[k] else [0]
        for q in owners:
            cq = all_counts[q].get(k, [])
            if q == rank:
                my_pos[k] = cur
            for c in cq:
                offs.append(cur + (4 if gdal_ghost else 0))
                cnts.append(c)
                cur += c + extra
        h, w, ty, tx = grids[k]
        if len(offs) != ty * tx:
            raise RuntimeError(f"write_cog_sharded: level {k} has {len(offs)} tiles, expected {ty * tx}")
        ifds[k].set_payload(cw.TAG_TILEOFFS, cw.T_LONG8, offs)
        ifds[k].set_payload(cw.TAG_TILECOUNTS, cw.T_LONG8, cnts)
    total = cur
    if rank == 0:
        with open(path, "wb") as fh:
            fh.truncate(total)
            cw._write_header_and_ifds(fh, ifds, ifd_off, True, gdal_ghost)
    if multi:
        dist.barrier()
    fd = os.open(path, os.O_WRONLY)
    try:
        for k, v in blobs.items():
            pos = my_pos[k]
            buf = bytearray()
            start = pos
            for b in v:
                if gdal_ghost:
                    buf += struct.pack("<I", len(b))
                buf += b
                if gdal_ghost:
                    buf += b[-4:]
                if len(buf) >= (64 << 20):
                    os.pwrite(fd, bytes(buf), start)
                    start += len(buf)
                    buf = bytearray()
            if buf:
                os.pwrite(fd, bytes(buf), start)
    finally:
        os.close(fd)
    if multi:
        dist.barrier()
    t_write = time.perf_counter() - t0
    return {"bytes": int(total), "levels": n_ov + 1, "distributed_levels": int(sum(distributed)), "overview_s": t_ov,
            "compress_s": t_comp, "write_s": t_write, "threads": nthreads, "predictor": pred, "blocksize": bs}


__all__ = ["write_cog_sharded", "tile_aligned_bounds"]
