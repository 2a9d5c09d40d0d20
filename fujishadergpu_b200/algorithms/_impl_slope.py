"""Slope (reference: algorithms/_impl_slope.py)."""
from __future__ import annotations

from .. import kernels as _k
from .. import _device as _dev
from ._base import DaskAlgorithm
from ._impl_hillshade import _reject_spatial


def compute_slope_block(block, *, unit="degree", pixel_size=1.0, pixel_scale_x=None, pixel_scale_y=None):
    """reference :19-35."""
    out = _k.slope(block, unit=unit, pixel_size=pixel_size, pixel_scale_x=pixel_scale_x, pixel_scale_y=pixel_scale_y)
    return _dev.like_input(out, block)


class SlopeAlgorithm(DaskAlgorithm):
    """reference :48-80 (local mode)."""

    def process(self, gpu_arr, **params):
        mode = str(params.get("mode", "local")).lower()
        _reject_spatial("slope", mode, params.get("radii"))
        kw = dict(unit=params.get("unit", "degree"), pixel_size=params.get("pixel_size", 1.0),
                  pixel_scale_x=params.get("pixel_scale_x"), pixel_scale_y=params.get("pixel_scale_y"))
        if hasattr(gpu_arr, "map_overlap"):
            return gpu_arr.map_overlap(compute_slope_block, depth=1, boundary="reflect", dtype="float32", **kw)
        return compute_slope_block(gpu_arr, **kw)

    def get_default_params(self) -> dict:
        return {"unit": "degree", "pixel_size": 1.0, "mode": "local", "radii": None, "weights": None}


__all__ = ["compute_slope_block", "SlopeAlgorithm"]
