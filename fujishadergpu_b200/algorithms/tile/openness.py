"""Tile-backend class for openness (reference: algorithms/tile/openness.py)."""
from .._impl_openness import OpennessAlgorithm as _DaskOpennessAlgorithm
from .dask_bridge import DaskSharedTileAdapter


class OpennessAlgorithm(DaskSharedTileAdapter):
    dask_algorithm_cls = _DaskOpennessAlgorithm
