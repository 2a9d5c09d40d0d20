"""Tile-backend class for curvature (reference: algorithms/tile/curvature.py)."""
from .._impl_curvature import CurvatureAlgorithm as _DaskCurvatureAlgorithm
from .dask_bridge import DaskSharedTileAdapter


class CurvatureAlgorithm(DaskSharedTileAdapter):
    dask_algorithm_cls = _DaskCurvatureAlgorithm
