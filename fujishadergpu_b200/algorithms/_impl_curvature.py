"""Curvature (reference: algorithms/_impl_curvature.py)."""
from __future__ import annotations

from .. import kernels as _k
from .. import _device as _dev
from ._base import DaskAlgorithm
from ._impl_hillshade import _reject_spatial


def compute_curvature_block(block, *, curvature_type="mean", pixel_size=1.0, pixel_scale_x=None, pixel_scale_y=None):
    """reference :19-57."""
    out = _k.curvature(block, curvature_type=curvature_type, pixel_size=pixel_size,
                       pixel_scale_x=pixel_scale_x, pixel_scale_y=pixel_scale_y)
    return _dev.like_input(out, block)


class CurvatureAlgorithm(DaskAlgorithm):
    """reference :60-99 (local mode)."""

    def process(self, gpu_arr, **params):
        mode = str(params.get("mode", "local")).lower()
        _reject_spatial("curvature", mode, params.get("radii"))
        kw = dict(curvature_type=params.get("curvature_type", "mean"), pixel_size=params.get("pixel_size", 1.0),
                  pixel_scale_x=params.get("pixel_scale_x"), pixel_scale_y=params.get("pixel_scale_y"))
        if hasattr(gpu_arr, "map_overlap"):
            return gpu_arr.map_overlap(compute_curvature_block, depth=2, boundary="reflect", dtype="float32", **kw)
        return compute_curvature_block(gpu_arr, **kw)

    def get_default_params(self) -> dict:
        return {"curvature_type": "mean", "pixel_size": 1.0, "mode": "local", "radii": None, "weights": None}


__all__ = ["compute_curvature_block", "CurvatureAlgorithm"]
