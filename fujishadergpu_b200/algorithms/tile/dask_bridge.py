"""Tile adapters: run a registry algorithm on one device-resident tile
(reference: algorithms/tile/dask_bridge.py:72-271, 334-405).  On the B200 path every supported
combination is a direct call of the fused block function -- there is no dask fallback graph."""
from __future__ import annotations

import logging

from ... import kernels as _k
from ... import _device as _dev
from .._base import Constants
from .._global_stats import _apply_display_stretch_block

logger = logging.getLogger(__name__)


def _merged_params(algo, params):
    merged = {}
    defaults = algo.get_default_params()
    if isinstance(defaults, dict):
        merged.update(defaults)
    merged.update(params)
    return merged


def _combine_direct(responses, *, weights=None, agg="mean"):
    """reference :28-69 -- stack / max / min / sum / f32-normalised weighted mean / equal mean."""
    import numpy as np
    import torch
    from .._nan_utils import _accumulate, _combine_multiscale_dask
    if not responses:
        raise ValueError("responses must not be empty")
    a = str(agg or "mean").lower()
    if a == "stack" or len(responses) == 1 or a in ("max", "min"):
        return _combine_multiscale_dask(responses, agg=a)
    if a == "sum":
        return _dev.like_input(_accumulate(responses, [1.0] * len(responses), "first_from_zero"), responses[0])
    if isinstance(weights, (list, tuple)) and len(weights) == len(responses):
        w = np.asarray(weights, dtype=np.float32)
        if np.isfinite(w).all() and float(w.sum()) > 0:
            w = w / float(w.sum())
            return _dev.like_input(_accumulate(responses, [float(x) for x in w], "first_weighted"), responses[0])
    inv = np.float32(1.0 / float(len(responses)))
    return _dev.like_input(_accumulate(responses, [float(inv)] * len(responses), "first_from_zero"), responses[0])


def _direct_hillshade(block, p):
    from .._impl_hillshade import compute_hillshade_spatial_block
    radii = p.get("radii", [1])
    if not isinstance(radii, (list, tuple)) or len(radii) == 0:
        radii = [1]
    radii = [max(1.0, float(r)) for r in radii]
    z = p.get("z_factor", 1.0)
    if bool(p.get("multiscale", False) or len(radii) > 1) and len(radii) > 1:   # reference :94-104
        responses = [compute_hillshade_spatial_block(
            block, azimuth=p.get("azimuth", Constants.DEFAULT_AZIMUTH), altitude=p.get("altitude", Constants.DEFAULT_ALTITUDE),
            z_factor=1.0 if z is None else z, pixel_size=p.get("pixel_size", 1.0), pixel_scale_x=p.get("pixel_scale_x"),
            pixel_scale_y=p.get("pixel_scale_y"), radius=float(r)) for r in radii]
        return _combine_direct(responses, weights=p.get("weights"), agg=p.get("agg", "mean"))
    return _dev.like_input(_k.hillshade(
        block, azimuth=p.get("azimuth", Constants.DEFAULT_AZIMUTH), altitude=p.get("altitude", Constants.DEFAULT_ALTITUDE),
        z_factor=1.0 if z is None else z, pixel_size=p.get("pixel_size", 1.0),
        pixel_scale_x=p.get("pixel_scale_x"), pixel_scale_y=p.get("pixel_scale_y")), block)


def _direct_slope(block, p):
    return _dev.like_input(_k.slope(block, unit=p.get("unit", "degree"), pixel_size=p.get("pixel_size", 1.0),
                                    pixel_scale_x=p.get("pixel_scale_x"), pixel_scale_y=p.get("pixel_scale_y")), block)


def _direct_curvature(block, p):
    return _dev.like_input(_k.curvature(block, curvature_type=p.get("curvature_type", "mean"),
                                        pixel_size=p.get("pixel_size", 1.0), pixel_scale_x=p.get("pixel_scale_x"),
                                        pixel_scale_y=p.get("pixel_scale_y")), block)


def _direct_openness(block, p):
    # display stretch fused into the kernel epilogue (reference :208-222)
    return _dev.like_input(_k.openness(
        block, openness_type=p.get("openness_type", "positive"), num_directions=p.get("num_directions", 16),
        max_distance=p.get("max_distance", 50), pixel_size=p.get("pixel_size", 1.0),
        pixel_scale_x=p.get("pixel_scale_x"), pixel_scale_y=p.get("pixel_scale_y"),
        stretch=p.get("global_stats")), block)


def _direct_ambient_occlusion(block, p):
    # reference :190-206; display stretch fused into the kernel epilogue
    return _dev.like_input(_k.ambient_occlusion(
        block, num_samples=p.get("num_samples", 16), radius=p.get("radius", 10.0), intensity=p.get("intensity", 1.0),
        pixel_size=p.get("pixel_size", 1.0), pixel_scale_x=p.get("pixel_scale_x"), pixel_scale_y=p.get("pixel_scale_y"),
        stretch=p.get("global_stats")), block)


def _direct_topousm_fast(block, p, algo):
    """reference :225-271 -- raw + normalisation fused when global_stats is present."""
    stats = p.get("global_stats")
    if p.get("_topousm_fast_coarse_field") is None and not (
            isinstance(stats, (tuple, list)) and len(stats) >= 1 and float(stats[0]) > 1e-9):
        logger.warning("topousm_fast: global_stats missing on tile direct path; estimating from this tile only.")
        from .._normalization import topousm_fast_stat_func
        radii = p.get("radii") or algo._determine_optimal_radii(p.get("pixel_size", 1.0))
        raw = _k.topousm_fast(block, radii=radii, weights=p.get("weights"), pixel_size=p.get("pixel_size", 1.0))
        return _dev.like_input(_k.scale(raw, topousm_fast_stat_func(raw)[0]), block)
    return algo.process(block, **p)


_DIRECT = {
    "HillshadeAlgorithm": _direct_hillshade, "SlopeAlgorithm": _direct_slope,
    "CurvatureAlgorithm": _direct_curvature, "OpennessAlgorithm": _direct_openness,
    "AmbientOcclusionAlgorithm": _direct_ambient_occlusion,
}


def _process_direct(algo, class_name, dem_gpu, params):
    p = _merged_params(algo, params)
    if str(p.get("mode", "local")).lower() == "spatial" and class_name in _DIRECT:
        return algo.process(dem_gpu, **p)   # the reference falls back to a single-chunk dask graph here (:340-346)
    if class_name in _DIRECT:
        return _DIRECT[class_name](dem_gpu, p)
    if class_name == "TopoUSMFastAlgorithm":
        return _direct_topousm_fast(dem_gpu, p, algo)
    raise NotImplementedError(class_name)


class DaskSharedTileAdapter:
    """reference :374-405."""

    dask_algorithm_cls = None

    def __init__(self):
        if self.dask_algorithm_cls is None:
            raise ValueError("dask_algorithm_cls must be set in subclass")
        self._algo = self.dask_algorithm_cls()

    def get_default_params(self):
        return self._algo.get_default_params()

    def process(self, dem_gpu, **params):
        return _process_direct(self._algo, self.dask_algorithm_cls.__name__, dem_gpu, params)


def tile_adapter_for(name: str, module: str):
    """The tile-backend class of registry algorithm `name` (core/tile_processor.py finds it by class name in
    algorithms/tile/<name>.py): a DaskSharedTileAdapter bound to the class registered in dask_registry.ALGORITHMS."""
    from ..dask_registry import ALGORITHMS
    algo_cls = type(ALGORITHMS[name])
    return type(algo_cls.__name__, (DaskSharedTileAdapter,),
                {"dask_algorithm_cls": algo_cls, "__module__": module,
                 "__doc__": f"Tile adapter of {name} (direct CUDA block function; reference: algorithms/tile/{name}.py)."})

