"""SlopeAlgorithm of the tile backend, looked up by name in this module (core/tile_processor.py:807-820 of the reference)."""
from .dask_bridge import tile_adapter_for

SlopeAlgorithm = tile_adapter_for("slope", __name__)
