"""OpennessAlgorithm of the tile backend, looked up by name in this module (core/tile_processor.py:807-820 of the reference)."""
from .dask_bridge import tile_adapter_for

OpennessAlgorithm = tile_adapter_for("openness", __name__)
