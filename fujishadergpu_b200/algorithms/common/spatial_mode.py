"""Scale construction for --mode local / spatial (reference: algorithms/common/spatial_mode.py).

radii  : the ladder 2, 8, 32, 128, 512, 2048 px cut at min(2048, short_side / 10); never empty.
weights: 2^(n-1-i), L1-normalised, for whatever n results.
Pure host arithmetic; reproduced bit for bit (tests/test_host_logic.py)."""
from __future__ import annotations

from typing import List, Optional, Tuple

AUTO_RADII_SEQUENCE: Tuple[int, ...] = (2, 8, 32, 128, 512, 2048)
AUTO_RADIUS_MAX: int = 2048

RADII_DRIVEN_ALGOS = frozenset((
    "topousm_fast", "hillshade", "slope", "specular", "atmospheric_scattering", "curvature",
    "ambient_occlusion", "openness", "multi_light_uncertainty", "npr_edges", "structure_tensor", "frangi"))
MULTISCALE_REQUIRED_ALGOS = frozenset((
    "fractal_anomaly", "scale_space_surprise", "visual_saliency", "scale_drift", "phase_congruency"))

LOCAL_RADII = [1]
LOCAL_WEIGHTS = [1.0]


def auto_spatial_radii(short_side_px: Optional[float]) -> List[int]:
    ceiling = float(AUTO_RADIUS_MAX)
    if short_side_px is not None:
        ceiling = min(ceiling, float(short_side_px) / 10.0)
    picked = [step for step in AUTO_RADII_SEQUENCE if float(step) <= ceiling]
    return picked or [AUTO_RADII_SEQUENCE[0]]


def auto_spatial_weights(n: int) -> List[float]:
    if n <= 0:
        return []
    powers = [2.0 ** (n - 1 - k) for k in range(n)]
    norm = sum(powers)
    return [p / norm for p in powers]


def auto_spatial_profile(short_side_px: Optional[float], radii: Optional[List[int]] = None):
    chosen = auto_spatial_radii(short_side_px) if radii is None else [int(round(float(v))) for v in radii]
    return chosen, auto_spatial_weights(len(chosen))
