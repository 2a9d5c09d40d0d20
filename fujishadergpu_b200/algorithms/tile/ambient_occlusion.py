"""Tile-backend class for ambient_occlusion (reference: algorithms/tile/ambient_occlusion.py)."""
from .._impl_ambient_occlusion import AmbientOcclusionAlgorithm as _DaskAmbientOcclusionAlgorithm
from .dask_bridge import DaskSharedTileAdapter


class AmbientOcclusionAlgorithm(DaskSharedTileAdapter):
    dask_algorithm_cls = _DaskAmbientOcclusionAlgorithm


__all__ = ["AmbientOcclusionAlgorithm"]
