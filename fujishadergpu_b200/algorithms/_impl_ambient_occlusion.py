"""Ambient occlusion (reference: algorithms/_impl_ambient_occlusion.py)."""
from __future__ import annotations

from .. import kernels as _k
from .. import _device as _dev
from ._base import DaskAlgorithm
from ._global_stats import apply_display_stretch_dask, robust_unsigned_stretch_stat_func  # noqa: F401 (re-export)
from ._nan_utils import (overlap_whole, _combine_multiscale_dask, _radius_to_downsample_factor, _resolve_spatial_radii_weights,
                         large_radius_threshold, multiscale_response_fields)


def compute_ambient_occlusion_block(block, *, num_samples: int = 16, radius: float = 10.0, intensity: float = 1.0,
                                    pixel_size: float = 1.0, pixel_scale_x: float = None, pixel_scale_y: float = None):
    """reference :33-118 -- 4 rings x num_samples edge-replicated gathers, sigma-1 Gaussian, gamma, NaN restore."""
    out = _k.ambient_occlusion(block, num_samples=num_samples, radius=radius, intensity=intensity,
                               pixel_size=pixel_size, pixel_scale_x=pixel_scale_x, pixel_scale_y=pixel_scale_y)
    return _dev.like_input(out, block)


def compute_ambient_occlusion_spatial_block(block, *, num_samples: int = 16, radius: float = 10.0,
                                            intensity: float = 1.0, pixel_size: float = 1.0,
                                            pixel_scale_x: float = None, pixel_scale_y: float = None):
    """reference :121-158 -- decimate, run on the small grid with scaled radius and steps, zoom back."""
    ds = _radius_to_downsample_factor(float(radius), block_shape=None, pixel_size=pixel_size,
                                      algorithm_name="ambient_occlusion")
    if ds <= 1:
        return compute_ambient_occlusion_block(block, num_samples=num_samples, radius=radius, intensity=intensity,
                                               pixel_size=pixel_size, pixel_scale_x=pixel_scale_x,
                                               pixel_scale_y=pixel_scale_y)
    t = _dev.as_f32_2d(block)
    small = _k.decimate(t, ds)
    psx = float(abs(float(pixel_scale_x)) * ds) if pixel_scale_x is not None else None
    psy = float(abs(float(pixel_scale_y)) * ds) if pixel_scale_y is not None else None
    rs = _k.ambient_occlusion(small, num_samples=num_samples, radius=max(1.0, float(radius) / float(ds)),
                              intensity=intensity, pixel_size=float(pixel_size) * float(ds),
                              pixel_scale_x=psx, pixel_scale_y=psy)
    return _dev.like_input(_k.upsample(rs, t.shape), block)


class AmbientOcclusionAlgorithm(DaskAlgorithm):
    """reference :161-224.  local: the full-resolution block function; spatial: one (decimated) run per radius,
    radii above the large-radius threshold on a coarsened DEM, responses mixed by _combine_multiscale_dask."""

    def process(self, gpu_arr, **params):
        kw = dict(num_samples=params.get("num_samples", 16), intensity=params.get("intensity", 1.0),
                  pixel_size=params.get("pixel_size", 1.0), pixel_scale_x=params.get("pixel_scale_x"),
                  pixel_scale_y=params.get("pixel_scale_y"))
        radius = params.get("radius", 10.0)
        mode = str(params.get("mode", "local")).lower()
        if mode == "spatial":
            if hasattr(gpu_arr, "map_overlap"):
                raise NotImplementedError("ambient_occlusion: spatial mode takes a device block, not a dask array, on the B200 path")
            radii, weights = _resolve_spatial_radii_weights(params.get("radii"), params.get("weights", None),
                                                            kw["pixel_size"])
            thr = large_radius_threshold(gpu_arr, fallback=max(radii) if radii else 64)
            responses = multiscale_response_fields(
                gpu_arr, [float(max(1, int(round(float(r))))) for r in radii],
                block_fn=compute_ambient_occlusion_spatial_block, radius_kw="radius",
                depth_for_scale=lambda rr: int(rr) + 1, is_large=lambda rr: int(rr) > thr,
                pixel_size=kw["pixel_size"], pixel_scale_x=kw["pixel_scale_x"], pixel_scale_y=kw["pixel_scale_y"],
                coarse_dem=params.get("_overview_coarse_dem"), coarse_decimation=params.get("_overview_decimation"),
                num_samples=kw["num_samples"], intensity=kw["intensity"])
            result = _combine_multiscale_dask(responses, weights=weights, agg=params.get("agg", "mean"))
        elif hasattr(gpu_arr, "map_overlap"):
            result = gpu_arr.map_overlap(compute_ambient_occlusion_block, depth=int(radius + 1), boundary="reflect",
                                         dtype="float32", radius=radius, **kw)
        else:
            result = overlap_whole(gpu_arr, compute_ambient_occlusion_block, int(radius + 1), radius=radius, **kw)
        return apply_display_stretch_dask(result, params.get("global_stats"))

    def get_default_params(self) -> dict:
        return {"num_samples": 16, "radius": 10.0, "intensity": 1.0, "pixel_size": 1.0, "mode": "local",
                "radii": None, "weights": None}


__all__ = ["compute_ambient_occlusion_block", "compute_ambient_occlusion_spatial_block", "AmbientOcclusionAlgorithm"]
