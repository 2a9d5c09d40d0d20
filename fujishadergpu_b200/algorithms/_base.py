"""Constants and the algorithm ABC of the reference's plugin API
(reference: algorithms/_base.py:12-18 and :42-53)."""
from __future__ import annotations

from abc import ABC, abstractmethod


class Constants:
    DEFAULT_GAMMA = 1 / 2.2
    DEFAULT_AZIMUTH = 315
    DEFAULT_ALTITUDE = 45
    MAX_DEPTH = 150
    NAN_FILL_VALUE_POSITIVE = -1e6
    NAN_FILL_VALUE_NEGATIVE = 1e6


class DaskAlgorithm(ABC):
    """``process(gpu_arr, **params)`` / ``get_default_params()`` -- the interface
    ``core/dask_processor.run_pipeline`` and the tile adapters call."""

    @abstractmethod
    def process(self, gpu_arr, **params):
        ...

    @abstractmethod
    def get_default_params(self) -> dict:
        ...


__all__ = ["Constants", "DaskAlgorithm"]
