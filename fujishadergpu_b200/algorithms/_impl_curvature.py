"""Curvature (reference: algorithms/_impl_curvature.py)."""
from __future__ import annotations

from .. import kernels as _k
from .. import _device as _dev
from ._base import DaskAlgorithm
from ._impl_hillshade import spatial_responses
from ._nan_utils import overlap_whole, _combine_multiscale_dask, _resolve_spatial_radii_weights, _smooth_for_radius


def compute_curvature_block(block, *, curvature_type="mean", pixel_size=1.0, pixel_scale_x=None, pixel_scale_y=None):
    """reference :19-57."""
    out = _k.curvature(block, curvature_type=curvature_type, pixel_size=pixel_size,
                       pixel_scale_x=pixel_scale_x, pixel_scale_y=pixel_scale_y)
    return _dev.like_input(out, block)


def compute_curvature_spatial_block(block, *, radius=4.0, curvature_type="mean", pixel_size=1.0,
                                    pixel_scale_x=None, pixel_scale_y=None):
    """reference :72-78 (the closure _curv_spatial) -- curvature of the Gaussian-smoothed block."""
    smoothed = _smooth_for_radius(block, radius, pixel_size=pixel_size, algorithm_name="curvature")
    return compute_curvature_block(smoothed, curvature_type=curvature_type, pixel_size=pixel_size,
                                   pixel_scale_x=pixel_scale_x, pixel_scale_y=pixel_scale_y)


class CurvatureAlgorithm(DaskAlgorithm):
    """reference :60-99; whole raster == one block."""

    def process(self, gpu_arr, **params):
        mode = str(params.get("mode", "local")).lower()
        kw = dict(curvature_type=params.get("curvature_type", "mean"), pixel_size=params.get("pixel_size", 1.0),
                  pixel_scale_x=params.get("pixel_scale_x"), pixel_scale_y=params.get("pixel_scale_y"))
        if mode == "spatial":
            if hasattr(gpu_arr, "map_overlap"):
                raise NotImplementedError("curvature: spatial mode takes a device block, not a dask array, on the B200 path")
            radii, weights = _resolve_spatial_radii_weights(params.get("radii"), params.get("weights", None), kw["pixel_size"])
            responses = spatial_responses(gpu_arr, radii, params, block_fn=compute_curvature_spatial_block,
                                          depth_for_scale=lambda rr: max(3, int(float(rr) * 2 + 2)),
                                          curvature_type=kw["curvature_type"])
            return _combine_multiscale_dask(responses, weights=weights, agg=params.get("agg", "mean"))
        if hasattr(gpu_arr, "map_overlap"):
            return gpu_arr.map_overlap(compute_curvature_block, depth=2, boundary="reflect", dtype="float32", **kw)
        return overlap_whole(gpu_arr, compute_curvature_block, 2, **kw)

    def get_default_params(self) -> dict:
        return {"curvature_type": "mean", "pixel_size": 1.0, "mode": "local", "radii": None, "weights": None}


__all__ = ["compute_curvature_block", "compute_curvature_spatial_block", "CurvatureAlgorithm"]
