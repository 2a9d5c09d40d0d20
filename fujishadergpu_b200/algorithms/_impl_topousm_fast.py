"""TopoUSM Fast (reference: algorithms/_impl_topousm_fast.py)."""
from __future__ import annotations

import logging
from typing import List, Optional

from .. import kernels as _k
from .. import _device as _dev
from ._base import DaskAlgorithm
from ._global_stats import apply_global_normalization
from ._normalization import topousm_fast_norm_func, topousm_fast_stat_func

logger = logging.getLogger(__name__)


def compute_topousm_fast_efficient_block(block, *, radii: Optional[List[int]] = None,
                                         weights: Optional[List[float]] = None, pixel_size: float = 1.0):
    """reference :49-100 -- sum_i w_i (block - mean_i), raw (un-normalised)."""
    if radii is None:
        radii = [4, 16, 64]
    out = _k.topousm_fast(block, radii=radii, weights=weights, pixel_size=pixel_size, norm_scale=None)
    return _dev.like_input(out, block)


def topousm_fast_default_large_radius_threshold(min_chunk: int) -> int:
    """reference :103-110."""
    return int(max(256, int(min_chunk) // 16))


def split_radii_by_threshold(radii, weights, threshold):
    """reference :113-130 -- weights are NOT renormalised."""
    count = len(radii)
    if weights is None or len(weights) != count:
        weights = [1.0 / count] * count
    near_r, near_w, far_r, far_w = [], [], [], []
    for r, w in zip(radii, weights):
        if int(r) > int(threshold):
            far_r.append(int(r)); far_w.append(float(w))
        else:
            near_r.append(int(r)); near_w.append(float(w))
    return near_r, near_w, far_r, far_w


def compute_topousm_fast_large_coarse_field(coarse_dem, *, large_radii, large_weights, decimation: float):
    """reference :133-155 -- Sum_i w_i * mean_{r_i}(coarse_dem) on the overview grid (NaN-aware box, 'reflect';
    sigma-1 Gaussian when the scaled radius is <= 1), accumulated in list order in f32."""
    import torch
    t = _dev.as_f32_2d(coarse_dem)
    gh = int(t.shape[0])
    field = None
    for r, w in zip(large_radii, large_weights):
        r_coarse = max(1, int(round(float(r) / max(float(decimation), 1.0))))
        size = 0 if r_coarse <= 1 else 2 * r_coarse + 1      # size 0 = sigma-1 Gaussian, 'nearest'
        mean_c = _k.grid_mean_band(t, 0, gh, size, 0, gh)
        if field is None:
            field = torch.empty_like(mean_c)
            _k.combine(field, mean_c, float(w), "first_weighted")
        else:
            _k.combine(field, mean_c, float(w), "add_weighted")
    if field is None:
        raise ValueError("large_radii must not be empty")
    return _dev.like_input(field, coarse_dem)


def _topousm_fast_add_large_block(block, *, coarse_field, w_large, off_r, off_c, full_h, full_w, block_info=None):
    """reference :158-186 -- W_large*block - bilinear(field) at the block's global position."""
    if block_info is not None and block_info.get(0) is not None:
        (r0, _r1), (c0, _c1) = block_info[0]["array-location"][0], block_info[0]["array-location"][1]
    else:
        r0, c0 = 0, 0
    out = _k.topousm_large_part(block, coarse_field, w_large=float(w_large), off_r=int(r0) + int(off_r),
                                off_c=int(c0) + int(off_c), full_h=int(full_h), full_w=int(full_w))
    return _dev.like_input(out, block)


def compute_topousm_fast_input_sample_stats(gpu_arr, *, radii, weights=None, pixel_size=1.0) -> tuple:
    """reference :219-266 -- scale from a bounded central window."""
    t = _dev.as_f32_2d(gpu_arr)
    h, w = int(t.shape[0]), int(t.shape[1])
    max_r = max((int(r) for r in radii), default=1)
    win = max(256, int(min(h, w, max(4096, max_r * 4))))
    y0, x0 = max(0, (h - win) // 2), max(0, (w - win) // 2)
    sample = t[y0:min(h, y0 + win), x0:min(w, x0 + win)]
    if sample.numel() == 0:
        return (1.0,)
    raw = _k.topousm_fast(sample, radii=[max(1, int(r)) for r in radii], weights=weights,
                          pixel_size=float(pixel_size), norm_scale=None)
    return topousm_fast_stat_func(raw)


class TopoUSMFastAlgorithm(DaskAlgorithm):
    """reference :269-357.  Raw block + normalisation are ONE fused launch sequence when
    ``global_stats`` is known (the normal, orchestrated case)."""

    def process(self, gpu_arr, **params):
        pixel_size = params.get("pixel_size", 1.0)
        radii = params.get("radii", None)
        weights = params.get("weights", None)
        if radii is None:
            radii = self._determine_optimal_radii(pixel_size)
        stats = params.get("global_stats", None)
        stats_ok = isinstance(stats, (tuple, list)) and len(stats) >= 1 and float(stats[0]) > 1e-9
        coarse_field = params.get("_topousm_fast_coarse_field", None)
        if coarse_field is not None:
            # overview large-radius split (reference :282-306)
            small_r = params.get("_topousm_fast_small_radii", [])
            small_w = params.get("_topousm_fast_small_weights", None)
            w_large = float(params.get("_topousm_fast_w_large", 0.0))
            off_r, off_c = params.get("_topousm_fast_field_offset", (0, 0))
            t = _dev.as_f32_2d(gpu_arr)
            full_h, full_w = params.get("_topousm_fast_full_shape", tuple(t.shape))
            raw = _k.topousm_large_part(t, coarse_field, w_large=w_large, off_r=off_r, off_c=off_c,
                                        full_h=full_h, full_w=full_w)
            if small_r:
                small = _k.topousm_fast(t, radii=small_r, weights=small_w, pixel_size=pixel_size, norm_scale=None)
                _k.combine(small, raw, 1.0, "add_weighted")     # small + large (reference :299-303)
                raw = small
            if not stats_ok:
                stats = topousm_fast_stat_func(raw)
            return apply_global_normalization(_dev.like_input(raw, gpu_arr), topousm_fast_norm_func, stats)
        if not stats_ok:
            stats = compute_topousm_fast_input_sample_stats(gpu_arr, radii=radii, weights=weights, pixel_size=pixel_size)
        if not (isinstance(stats, (tuple, list)) and len(stats) >= 1 and float(stats[0]) > 1e-9):
            stats = (1.0,)
        out = _k.topousm_fast(gpu_arr, radii=radii, weights=weights, pixel_size=pixel_size, norm_scale=float(stats[0]))
        return _dev.like_input(out, gpu_arr)

    def _determine_optimal_radii(self, pixel_size: float) -> List[int]:
        """reference :336-346."""
        picked = {max(2, min(int(d / pixel_size), 256)) for d in (5, 20, 80, 320)}
        return sorted(picked)

    def get_default_params(self) -> dict:
        return {"mode": "radius", "radii": None, "weights": None, "sigmas": None, "agg": "mean", "auto_sigma": False}


__all__ = ["compute_topousm_fast_efficient_block", "TopoUSMFastAlgorithm",
           "topousm_fast_default_large_radius_threshold", "split_radii_by_threshold",
           "compute_topousm_fast_input_sample_stats", "_topousm_fast_add_large_block"]
