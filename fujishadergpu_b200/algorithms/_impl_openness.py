"""Openness (reference: algorithms/_impl_openness.py)."""
from __future__ import annotations

from .. import kernels as _k
from .. import _device as _dev
from ._base import DaskAlgorithm
from ._global_stats import apply_display_stretch_dask, robust_unsigned_stretch_stat_func  # noqa: F401 (re-export)
from ._nan_utils import (overlap_whole, _combine_multiscale_dask, _downsample_nan_aware, _radius_to_downsample_factor,
                         _resolve_spatial_radii_weights, _upsample_to_shape, large_radius_threshold,
                         multiscale_response_fields)


def compute_openness_vectorized(block, *, openness_type="positive", num_directions=16, max_distance=50,
                                pixel_size=1.0, pixel_scale_x=None, pixel_scale_y=None):
    """reference :31-132."""
    out = _k.openness(block, openness_type=openness_type, num_directions=num_directions, max_distance=max_distance,
                      pixel_size=pixel_size, pixel_scale_x=pixel_scale_x, pixel_scale_y=pixel_scale_y)
    return _dev.like_input(out, block)


def compute_openness_spatial_block(block, *, openness_type="positive", num_directions=16, max_distance=50,
                                   pixel_size=1.0, pixel_scale_x=None, pixel_scale_y=None):
    """reference :135-164 -- decimate, run on the small grid with scaled steps, zoom back."""
    ds = _radius_to_downsample_factor(float(max_distance), block_shape=None, pixel_size=pixel_size,
                                      algorithm_name="openness")
    kw = dict(openness_type=openness_type, num_directions=num_directions)
    if ds <= 1:
        return compute_openness_vectorized(block, max_distance=max_distance, pixel_size=pixel_size,
                                           pixel_scale_x=pixel_scale_x, pixel_scale_y=pixel_scale_y, **kw)
    t = _dev.as_f32_2d(block)
    small = _k.decimate(t, ds)
    psx = float(abs(float(pixel_scale_x)) * ds) if pixel_scale_x is not None else None
    psy = float(abs(float(pixel_scale_y)) * ds) if pixel_scale_y is not None else None
    rs = _k.openness(small, max_distance=max(2, int(round(float(max_distance) / float(ds)))),
                     pixel_size=float(pixel_size) * float(ds), pixel_scale_x=psx, pixel_scale_y=psy, **kw)
    return _dev.like_input(_k.upsample(rs, t.shape), block)


class OpennessAlgorithm(DaskAlgorithm):
    """reference :167-227.  local: the full-resolution block function; spatial: one (decimated) run per radius,
    radii above the large-radius threshold on a coarsened DEM, responses mixed by _combine_multiscale_dask."""

    def process(self, gpu_arr, **params):
        kw = dict(openness_type=params.get("openness_type", "positive"),
                  num_directions=params.get("num_directions", 16), pixel_size=params.get("pixel_size", 1.0),
                  pixel_scale_x=params.get("pixel_scale_x"), pixel_scale_y=params.get("pixel_scale_y"))
        mode = str(params.get("mode", "local")).lower()
        if mode == "spatial":
            if hasattr(gpu_arr, "map_overlap"):
                raise NotImplementedError("openness: spatial mode takes a device block, not a dask array, on the B200 path")
            radii, weights = _resolve_spatial_radii_weights(params.get("radii"), params.get("weights", None),
                                                            kw["pixel_size"])
            thr = large_radius_threshold(gpu_arr, fallback=max(radii) if radii else 64)
            responses = multiscale_response_fields(
                gpu_arr, [float(int(max(2, round(float(r))))) for r in radii],
                block_fn=compute_openness_spatial_block, radius_kw="max_distance",
                depth_for_scale=lambda md: int(md) + 1, is_large=lambda md: int(md) > thr,
                pixel_size=kw["pixel_size"], pixel_scale_x=kw["pixel_scale_x"], pixel_scale_y=kw["pixel_scale_y"],
                coarse_dem=params.get("_overview_coarse_dem"), coarse_decimation=params.get("_overview_decimation"),
                openness_type=kw["openness_type"], num_directions=kw["num_directions"])
            result = _combine_multiscale_dask(responses, weights=weights, agg=params.get("agg", "mean"))
        elif hasattr(gpu_arr, "map_overlap"):
            md = params.get("max_distance", 50)
            result = gpu_arr.map_overlap(compute_openness_vectorized, depth=md + 1, boundary="reflect",
                                         dtype="float32", max_distance=md, **kw)
        else:
            md = params.get("max_distance", 50)
            result = overlap_whole(gpu_arr, compute_openness_vectorized, int(md) + 1, max_distance=md, **kw)
        return apply_display_stretch_dask(result, params.get("global_stats"))

    def get_default_params(self) -> dict:
        return {"openness_type": "positive", "num_directions": 16, "max_distance": 50, "pixel_size": 1.0,
                "mode": "local", "radii": None, "weights": None}


__all__ = ["compute_openness_vectorized", "compute_openness_spatial_block", "OpennessAlgorithm"]
