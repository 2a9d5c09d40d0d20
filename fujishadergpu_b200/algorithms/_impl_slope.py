"""Slope (reference: algorithms/_impl_slope.py)."""
from __future__ import annotations

from .. import kernels as _k
from .. import _device as _dev
from ._base import DaskAlgorithm
from ._impl_hillshade import spatial_responses
from ._nan_utils import overlap_whole, _combine_multiscale_dask, _resolve_spatial_radii_weights, _smooth_for_radius


def compute_slope_block(block, *, unit="degree", pixel_size=1.0, pixel_scale_x=None, pixel_scale_y=None):
    """reference :19-35."""
    out = _k.slope(block, unit=unit, pixel_size=pixel_size, pixel_scale_x=pixel_scale_x, pixel_scale_y=pixel_scale_y)
    return _dev.like_input(out, block)


def compute_slope_spatial_block(block, *, unit="degree", pixel_size=1.0, pixel_scale_x=None, pixel_scale_y=None,
                                radius=4.0):
    """reference :38-45 -- slope of the Gaussian-smoothed block."""
    smoothed = _smooth_for_radius(block, radius, pixel_size=pixel_size, algorithm_name="slope")
    return compute_slope_block(smoothed, unit=unit, pixel_size=pixel_size, pixel_scale_x=pixel_scale_x,
                               pixel_scale_y=pixel_scale_y)


class SlopeAlgorithm(DaskAlgorithm):
    """reference :48-80; whole raster == one block."""

    def process(self, gpu_arr, **params):
        mode = str(params.get("mode", "local")).lower()
        kw = dict(unit=params.get("unit", "degree"), pixel_size=params.get("pixel_size", 1.0),
                  pixel_scale_x=params.get("pixel_scale_x"), pixel_scale_y=params.get("pixel_scale_y"))
        if mode == "spatial":
            if hasattr(gpu_arr, "map_overlap"):
                raise NotImplementedError("slope: spatial mode takes a device block, not a dask array, on the B200 path")
            radii, weights = _resolve_spatial_radii_weights(params.get("radii"), params.get("weights", None), kw["pixel_size"])
            responses = spatial_responses(gpu_arr, radii, params, block_fn=compute_slope_spatial_block,
                                          depth_for_scale=lambda rr: max(2, int(float(rr) * 2 + 1)), unit=kw["unit"])
            return _combine_multiscale_dask(responses, weights=weights, agg=params.get("agg", "mean"))
        if hasattr(gpu_arr, "map_overlap"):
            return gpu_arr.map_overlap(compute_slope_block, depth=1, boundary="reflect", dtype="float32", **kw)
        return overlap_whole(gpu_arr, compute_slope_block, 1, **kw)

    def get_default_params(self) -> dict:
        return {"unit": "degree", "pixel_size": 1.0, "mode": "local", "radii": None, "weights": None}


__all__ = ["compute_slope_block", "compute_slope_spatial_block", "SlopeAlgorithm"]
