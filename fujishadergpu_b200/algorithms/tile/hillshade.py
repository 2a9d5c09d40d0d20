"""HillshadeAlgorithm of the tile backend, looked up by name in this module (core/tile_processor.py:807-820 of the reference)."""
from .dask_bridge import tile_adapter_for

HillshadeAlgorithm = tile_adapter_for("hillshade", __name__)
