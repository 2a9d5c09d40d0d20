"""Tile-backend algorithm interface (reference: algorithms/tile_shared.py:19-44)."""
from __future__ import annotations

from abc import ABC, abstractmethod


class TileAlgorithm(ABC):
    """``process(dem_gpu, **params) -> device array`` on one (padded) tile."""

    @abstractmethod
    def process(self, dem_gpu, **params):
        ...

    def get_default_params(self) -> dict:
        return {}


__all__ = ["TileAlgorithm"]
