"""TopoUSMFastAlgorithm of the tile backend, looked up by name in this module (core/tile_processor.py:807-820 of the reference)."""
from .dask_bridge import tile_adapter_for

TopoUSMFastAlgorithm = tile_adapter_for("topousm_fast", __name__)
