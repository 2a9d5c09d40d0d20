"""Global normalisation statistics from stratified full-resolution windows
(reference: algorithms/_norm_stats.py).  The window geometry is the reference's; the windows are
sliced from a DEVICE-RESIDENT raster instead of being read through rasterio, the block function
is the CUDA pipeline, and the pooled percentile is an exact GPU selection."""
from __future__ import annotations

import logging
from typing import Optional

from .. import _device as _dev
from .. import kernels as _k

logger = logging.getLogger(__name__)

# algorithm -> (impl module, raw block function, stat function)   (reference :29-61, hot-path rows)
_NORM_STAT_SPECS = {
    "topousm_fast": ("_impl_topousm_fast", "compute_topousm_fast_efficient_block", "topousm_fast_stat_func"),
    "openness": ("_impl_openness", "compute_openness_vectorized", "robust_unsigned_stretch_stat_func"),
    "ambient_occlusion": ("_impl_ambient_occlusion", "compute_ambient_occlusion_block", "robust_unsigned_stretch_stat_func"),
}


def stratified_windows(width: int, height: int, by0: int, by1: int, bx0: int, bx1: int, *,
                       grid: int = 3, tile: int = 4096) -> list:
    """reference :64-100 -- unique (wy0, wx0, win_w, win_h) windows on a grid x grid layout."""
    step_y = max(1, (int(by1) - int(by0)) // int(grid))
    step_x = max(1, (int(bx1) - int(bx0)) // int(grid))
    picked, seen = [], set()
    for iy in range(int(grid)):
        for ix in range(int(grid)):
            mid_y = int(by0) + iy * step_y + step_y // 2
            mid_x = int(bx0) + ix * step_x + step_x // 2
            top = int(min(max(0, mid_y - tile // 2), max(0, int(height) - tile)))
            left = int(min(max(0, mid_x - tile // 2), max(0, int(width) - tile)))
            win = (top, left, min(tile, int(width) - left), min(tile, int(height) - top))
            if win not in seen:
                seen.add(win)
                picked.append(win)
    return picked


def _norm_stat_max_scale(merged: dict) -> float:
    """reference :103-124 (keys the hot-path algorithms use)."""
    found = []
    for key in ("radii", "scales"):
        v = merged.get(key)
        if isinstance(v, (list, tuple)) and v:
            found.append(max(float(x) for x in v))
    for key in ("kernel_size", "radius", "max_distance"):
        if merged.get(key):
            found.append(float(merged[key]))
    return max(found) if found else 16.0


def _norm_stat_halo_pixels(algorithm: str, merged: dict) -> int:
    """reference :127-153: topousm_fast / openness -> max scale + 16."""
    return int(float(_norm_stat_max_scale(merged)) + 16)


def _norm_stat_window_geometry(algorithm: str, merged: dict, max_tile: int = 4096):
    """reference :150-162 -> (margin, tile)."""
    margin = max(1, int(_norm_stat_halo_pixels(algorithm, merged)))
    return margin, max(min(2048, max(1, int(max_tile))), 4 * margin)


def valid_bbox_host(host_arr, cov: Optional[int] = None):
    """(by0, by1, bx0, bx1) of the valid data from a <= 512 px nearest overview of a HOST raster
    (reference :254-264), or None when nothing is valid."""
    import numpy as np
    a = host_arr.numpy() if hasattr(host_arr, "numpy") else np.asarray(host_arr)
    H, W = int(a.shape[0]), int(a.shape[1])
    cov = max(1, max(W, H) // 512) if cov is None else int(cov)
    ov = a[::cov, ::cov][: max(1, H // cov), : max(1, W // cov)]
    ok = np.isfinite(ov)
    if not ok.any():
        return None
    rows = np.nonzero(ok.any(axis=1))[0]
    cols = np.nonzero(ok.any(axis=0))[0]
    return (int(rows.min()) * cov, min(H, (int(rows.max()) + 1) * cov), int(cols.min()) * cov, min(W, (int(cols.max()) + 1) * cov))


def compute_norm_stats_device(dem, algorithm: str, params: dict, *, grid: int = 3, max_tile: int = 4096,
                              min_valid_frac: float = 0.02, bbox=None) -> Optional[tuple]:
    """Device-resident twin of _compute_norm_stats_tiled (reference :176-298)."""
    import inspect
    import torch
    from importlib import import_module
    from .dask_registry import ALGORITHMS

    spec = _NORM_STAT_SPECS.get(algorithm)
    if spec is None:
        return None
    mod = import_module(f"{__package__}.{spec[0]}")
    block_func, stat_func = getattr(mod, spec[1]), getattr(mod, spec[2])
    merged = {**(ALGORITHMS[algorithm].get_default_params() or {}), **(params or {})}
    accepted = set(inspect.signature(block_func).parameters)
    kw = {k: merged[k] for k in list(merged) if k in accepted and merged[k] is not None}
    margin, tile = _norm_stat_window_geometry(algorithm, merged, max_tile)

    t = _dev.as_f32_2d(dem)
    H, W = int(t.shape[0]), int(t.shape[1])
    if bbox is not None:
        by0, by1, bx0, bx1 = [int(v) for v in bbox]
    else:
        # coarse overview -> bounding box of valid data (reference :254-264; nearest-sampled <=512 px view)
        cov = max(1, max(W, H) // 512)
        ov = t[::cov, ::cov][: max(1, H // cov), : max(1, W // cov)]
        ok = torch.isfinite(ov)
        if not bool(ok.any()):
            return None
        rows = torch.nonzero(ok.any(dim=1)).flatten()
        cols = torch.nonzero(ok.any(dim=0)).flatten()
        by0, by1 = int(rows.min()) * cov, min(H, (int(rows.max()) + 1) * cov)
        bx0, bx1 = int(cols.min()) * cov, min(W, (int(cols.max()) + 1) * cov)

    wins = [t[wy0:wy0 + th, wx0:wx0 + tw]
            for wy0, wx0, tw, th in stratified_windows(W, H, by0, by1, bx0, bx1, grid=grid, tile=min(tile, max(W, H)))]
    # valid fraction of every window in one counting launch (one host sync instead of one per window)
    n_valid = _k.count_samples(wins, finite_only=True) if wins else []
    jobs = []
    for win, nv in zip(wins, n_valid):
        if nv < min_valid_frac * float(win.numel()):
            continue
        m = int(min(margin, win.shape[0] // 3, win.shape[1] // 3))

        def job(win=win, m=m):
            if algorithm == "topousm_fast" and m > 0:
                # same block function, same values: the kernel is only asked for the region the trim keeps
                # (reference :275-277 computes the whole window and throws the margin away)
                raw = _k.topousm_fast(win, radii=kw.get("radii") or [4, 16, 64], weights=kw.get("weights"),
                                      pixel_size=kw.get("pixel_size", 1.0), norm_scale=None,
                                      roi=(m, int(win.shape[0]) - 2 * m, m, int(win.shape[1]) - 2 * m))
            else:
                raw = _dev.as_tensor(block_func(win, **kw))
            return raw[m:-m, m:-m] if m > 0 else raw

        jobs.append(job)
    # the windows are independent: their (small) launches overlap on a few side streams
    pooled = [r for r in _dev.run_concurrently(jobs, t.device) if r.numel()]
    if not pooled:
        return None
    stats = stat_func(pooled)
    if not stats or not (float(stats[-1]) == float(stats[-1])) or float(stats[-1]) <= 1e-9:
        return None
    logger.info("%s global stats from %d full-res windows (tile=%d, margin=%d): %s",
                algorithm, len(pooled), tile, margin, stats)
    return stats


def inject_global_stats(dem, algorithm: str, params: dict) -> dict:
    """reference :301-350 (hot-path step 2 only): adds params['global_stats'] when the algorithm has one."""
    if algorithm in _NORM_STAT_SPECS and "global_stats" not in params:
        st = compute_norm_stats_device(dem, algorithm, params)
        if st is not None:
            params["global_stats"] = st
    return params


__all__ = ["_NORM_STAT_SPECS", "valid_bbox_host", "stratified_windows", "_norm_stat_window_geometry", "compute_norm_stats_device",
           "inject_global_stats"]
