"""Tile-backend class for hillshade (reference: algorithms/tile/hillshade.py)."""
from .._impl_hillshade import HillshadeAlgorithm as _DaskHillshadeAlgorithm
from .dask_bridge import DaskSharedTileAdapter


class HillshadeAlgorithm(DaskSharedTileAdapter):
    dask_algorithm_cls = _DaskHillshadeAlgorithm
