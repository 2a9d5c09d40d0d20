"""Algorithm package: registry + block functions of the accelerated hot path."""
from .dask_registry import ALGORITHMS, DaskAlgorithm  # noqa: F401
from .tile_shared import TileAlgorithm  # noqa: F401

__all__ = ["ALGORITHMS", "DaskAlgorithm", "TileAlgorithm"]
