"""Tile-backend class for topousm_fast (reference: algorithms/tile/topousm_fast.py)."""
from .._impl_topousm_fast import TopoUSMFastAlgorithm as _DaskTopoUSMFastAlgorithm
from .dask_bridge import DaskSharedTileAdapter


class TopoUSMFastAlgorithm(DaskSharedTileAdapter):
    dask_algorithm_cls = _DaskTopoUSMFastAlgorithm
