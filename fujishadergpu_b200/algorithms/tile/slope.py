"""Tile-backend class for slope (reference: algorithms/tile/slope.py)."""
from .._impl_slope import SlopeAlgorithm as _DaskSlopeAlgorithm
from .dask_bridge import DaskSharedTileAdapter


class SlopeAlgorithm(DaskSharedTileAdapter):
    dask_algorithm_cls = _DaskSlopeAlgorithm
