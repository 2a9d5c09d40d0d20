"""AmbientOcclusionAlgorithm of the tile backend, looked up by name in this module (core/tile_processor.py:807-820 of the reference)."""
from .dask_bridge import tile_adapter_for

AmbientOcclusionAlgorithm = tile_adapter_for("ambient_occlusion", __name__)
