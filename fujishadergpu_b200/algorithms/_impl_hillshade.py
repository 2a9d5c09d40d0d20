"""Hillshade (reference: algorithms/_impl_hillshade.py).  The block function keeps the reference's
name and keywords; the arithmetic runs in libfsg_b200 (csrc/fsg_gradient.cu)."""
from __future__ import annotations

from .. import kernels as _k
from .. import _device as _dev
from ._base import Constants, DaskAlgorithm
from ._nan_utils import (overlap_whole, _resolve_spatial_radii_weights, _smooth_for_radius, large_radius_threshold,
                         multiscale_response_fields)


def compute_hillshade_block(block, *, azimuth=Constants.DEFAULT_AZIMUTH, altitude=Constants.DEFAULT_ALTITUDE,
                            z_factor=1.0, pixel_size=1.0, pixel_scale_x=None, pixel_scale_y=None):
    """reference :20-54."""
    out = _k.hillshade(block, azimuth=azimuth, altitude=altitude, z_factor=z_factor, pixel_size=pixel_size,
                       pixel_scale_x=pixel_scale_x, pixel_scale_y=pixel_scale_y)
    return _dev.like_input(out, block)


def compute_hillshade_spatial_block(block, *, azimuth=Constants.DEFAULT_AZIMUTH, altitude=Constants.DEFAULT_ALTITUDE,
                                    z_factor=1.0, pixel_size=1.0, pixel_scale_x=None, pixel_scale_y=None, radius=4.0):
    """reference :57-67 -- hillshade of the Gaussian-smoothed block."""
    smoothed = _smooth_for_radius(block, radius, pixel_size=pixel_size, algorithm_name="hillshade")
    return compute_hillshade_block(smoothed, azimuth=azimuth, altitude=altitude, z_factor=z_factor,
                                   pixel_size=pixel_size, pixel_scale_x=pixel_scale_x, pixel_scale_y=pixel_scale_y)


def spatial_responses(block, radii, params, *, block_fn, depth_for_scale, **block_kwargs) -> list:
    """Per-radius responses of a spatial-mode algorithm: radii above the large-radius threshold
    max(256, min(H, W) // 16) are computed on a coarsened copy and sampled back (reference
    _nan_utils.py:441-524 through HillshadeAlgorithm.process :103-113 and its slope / curvature twins)."""
    thr = large_radius_threshold(block, fallback=int(max(radii)) if radii else 64)
    return multiscale_response_fields(
        block, [float(r) for r in radii], block_fn=block_fn, radius_kw="radius", depth_for_scale=depth_for_scale,
        is_large=lambda rr: int(round(float(rr))) > thr, pixel_size=params.get("pixel_size", 1.0),
        pixel_scale_x=params.get("pixel_scale_x"), pixel_scale_y=params.get("pixel_scale_y"),
        coarse_dem=params.get("_overview_coarse_dem"), coarse_decimation=params.get("_overview_decimation"),
        **block_kwargs)


def _weighted_f32(responses, weights):
    """Weighted sum as HillshadeAlgorithm.process / _combine_direct write it: weights normalised in f32,
    out = r0*w0; out = out + ri*wi (reference :118-128, tile/dask_bridge.py:58-64)."""
    import numpy as np
    from ._nan_utils import _accumulate
    w = np.asarray(weights, dtype=np.float32)
    if not (np.isfinite(w).all() and float(w.sum()) > 0):
        return None
    w = w / float(w.sum())
    return _accumulate(responses, [float(x) for x in w], "first_weighted")


class HillshadeAlgorithm(DaskAlgorithm):
    """reference :70-148; whole raster == one block."""

    def process(self, gpu_arr, **params):
        from ._nan_utils import _accumulate, _combine_multiscale_dask
        mode = str(params.get("mode", "local")).lower()
        z = params.get("z_factor", 1.0)
        kw = dict(azimuth=params.get("azimuth", Constants.DEFAULT_AZIMUTH),
                  altitude=params.get("altitude", Constants.DEFAULT_ALTITUDE),
                  z_factor=1.0 if z is None else z, pixel_size=params.get("pixel_size", 1.0),
                  pixel_scale_x=params.get("pixel_scale_x"), pixel_scale_y=params.get("pixel_scale_y"))
        radii = params.get("radii", [1])
        weights = params.get("weights", None)
        agg = params.get("agg", "mean")
        multiscale = bool(params.get("multiscale", False))
        if mode == "spatial":
            radii, auto_w = _resolve_spatial_radii_weights(radii, weights, kw["pixel_size"])
            if weights is None:
                weights = auto_w
            multiscale = True
        else:
            if not isinstance(radii, (list, tuple)) or len(radii) == 0:
                radii = [1]
            radii = [max(1.0, float(r)) for r in radii]
            multiscale = bool(multiscale or len(radii) > 1)
        if mode == "spatial" or (multiscale and len(radii) > 1):
            if hasattr(gpu_arr, "map_overlap"):
                raise NotImplementedError("hillshade: spatial mode takes a device block, not a dask array, on the B200 path")
            results = spatial_responses(gpu_arr, radii, params, block_fn=compute_hillshade_spatial_block,
                                        depth_for_scale=lambda rr: max(2, int(float(rr) * 2 + 1)),
                                        azimuth=kw["azimuth"], altitude=kw["altitude"], z_factor=kw["z_factor"])
            if agg == "stack":
                return _combine_multiscale_dask(results, agg="stack")
            if agg == "mean":
                if isinstance(weights, (list, tuple)) and len(weights) == len(radii):
                    out = _weighted_f32(results, weights) if len(results) > 1 else results[0]
                    if out is not None:
                        return _dev.like_input(out, gpu_arr)
                return _combine_multiscale_dask(results, weights=None, agg="mean")
            if agg in ("min", "max"):
                return _combine_multiscale_dask(results, agg=agg)
            return _combine_multiscale_dask(results, weights=None, agg="mean")
        if hasattr(gpu_arr, "map_overlap"):  # a dask array: same halo contract as the reference (:133)
            return gpu_arr.map_overlap(compute_hillshade_block, depth=1, boundary="reflect", dtype="float32", **kw)
        return overlap_whole(gpu_arr, compute_hillshade_block, 1, **kw)   # one block == the raster, same halo contract

    def get_default_params(self) -> dict:
        return {"azimuth": Constants.DEFAULT_AZIMUTH, "altitude": Constants.DEFAULT_ALTITUDE, "z_factor": 1.0,
                "pixel_size": 1.0, "multiscale": False, "mode": "local", "radii": None, "weights": None, "agg": "mean"}


__all__ = ["compute_hillshade_block", "compute_hillshade_spatial_block", "HillshadeAlgorithm"]
