"""Registry of algorithm instances by canonical name (reference: algorithms/dask_registry.py:27-49).

The five hot-path algorithms and ambient_occlusion are served by the B200 library.  The
reference's other fifteen names are
kept in ``UNACCELERATED`` so a drop-in caller gets a clear error instead of a silent fallback."""
from __future__ import annotations

from ._base import DaskAlgorithm
from ._impl_ambient_occlusion import AmbientOcclusionAlgorithm
from ._impl_curvature import CurvatureAlgorithm
from ._impl_hillshade import HillshadeAlgorithm
from ._impl_openness import OpennessAlgorithm
from ._impl_slope import SlopeAlgorithm
from ._impl_topousm_fast import TopoUSMFastAlgorithm

ALGORITHMS = {
    "topousm_fast": TopoUSMFastAlgorithm(),
    "hillshade": HillshadeAlgorithm(),
    "slope": SlopeAlgorithm(),
    "curvature": CurvatureAlgorithm(),
    "openness": OpennessAlgorithm(),
    "ambient_occlusion": AmbientOcclusionAlgorithm(),
}

UNACCELERATED = (
    "specular", "atmospheric_scattering", "multiscale_terrain", "blur", "visual_saliency", "npr_edges",
    "fractal_anomaly", "scale_space_surprise", "multi_light_uncertainty",
    "structure_tensor", "frangi", "lic", "phase_congruency", "tv_decomposition", "scale_drift",
)


def get_algorithm(name: str) -> DaskAlgorithm:
    key = str(name).lower()
    if key in ALGORITHMS:
        return ALGORITHMS[key]
    if key in UNACCELERATED:
        raise NotImplementedError(
            f"algorithm {key!r} is outside the B200 hot path (topousm_fast, hillshade, slope, curvature, "
            "openness, ambient_occlusion); run it with the reference implementation")
    raise KeyError(f"unknown algorithm {name!r}")


__all__ = ["ALGORITHMS", "UNACCELERATED", "DaskAlgorithm", "get_algorithm"]
