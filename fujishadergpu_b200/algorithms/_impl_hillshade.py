"""Hillshade (reference: algorithms/_impl_hillshade.py).  The block function keeps the reference's
name and keywords; the arithmetic runs in libfsg_b200 (csrc/fsg_gradient.cu)."""
from __future__ import annotations

from .. import kernels as _k
from .. import _device as _dev
from ._base import Constants, DaskAlgorithm
from ._nan_utils import _resolve_spatial_radii_weights


def compute_hillshade_block(block, *, azimuth=Constants.DEFAULT_AZIMUTH, altitude=Constants.DEFAULT_ALTITUDE,
                            z_factor=1.0, pixel_size=1.0, pixel_scale_x=None, pixel_scale_y=None):
    """reference :20-54."""
    out = _k.hillshade(block, azimuth=azimuth, altitude=altitude, z_factor=z_factor, pixel_size=pixel_size,
                       pixel_scale_x=pixel_scale_x, pixel_scale_y=pixel_scale_y)
    return _dev.like_input(out, block)


def _reject_spatial(name: str, mode: str, radii) -> None:
    if mode == "spatial" and radii is not None and len(radii) > 1:
        raise NotImplementedError(
            f"{name}: --mode spatial with several radii (Gaussian scale-space smoothing) is not part of the "
            "B200 hot path yet (SURVEY.md section 8f rank 1); use --mode local")


class HillshadeAlgorithm(DaskAlgorithm):
    """reference :70-148 (local mode; whole raster == one block)."""

    def process(self, gpu_arr, **params):
        mode = str(params.get("mode", "local")).lower()
        _reject_spatial("hillshade", mode, params.get("radii"))
        kw = dict(azimuth=params.get("azimuth", Constants.DEFAULT_AZIMUTH),
                  altitude=params.get("altitude", Constants.DEFAULT_ALTITUDE),
                  z_factor=params.get("z_factor", 1.0) if params.get("z_factor", 1.0) is not None else 1.0,
                  pixel_size=params.get("pixel_size", 1.0),
                  pixel_scale_x=params.get("pixel_scale_x"), pixel_scale_y=params.get("pixel_scale_y"))
        if hasattr(gpu_arr, "map_overlap"):  # a dask array: same halo contract as the reference (:133)
            return gpu_arr.map_overlap(compute_hillshade_block, depth=1, boundary="reflect", dtype="float32", **kw)
        return compute_hillshade_block(gpu_arr, **kw)

    def get_default_params(self) -> dict:
        return {"azimuth": Constants.DEFAULT_AZIMUTH, "altitude": Constants.DEFAULT_ALTITUDE, "z_factor": 1.0,
                "pixel_size": 1.0, "multiscale": False, "mode": "local", "radii": None, "weights": None, "agg": "mean"}


__all__ = ["compute_hillshade_block", "HillshadeAlgorithm"]
