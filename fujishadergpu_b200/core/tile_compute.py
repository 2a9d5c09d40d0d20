"""Tile-pipeline helpers on the hot path (reference: core/tile_compute.py)."""
from __future__ import annotations

import math
from typing import Any, Dict, Iterable, List, Optional, Tuple


def deduplicate_radii_weights(radii: Iterable[int], weights: Optional[Iterable[float]]):
    """reference :9-40."""
    seq = list(radii)
    index: Dict[int, int] = {}
    uniq: List[int] = []
    for r in seq:
        if r not in index:
            index[r] = len(uniq)
            uniq.append(r)
    if weights is None:
        return uniq, None
    wseq = list(weights)
    if len(wseq) != len(seq):
        return uniq, None
    pooled = [0.0] * len(uniq)
    for r, w in zip(seq, wseq):
        try:
            val = float(w)
        except (TypeError, ValueError):
            val = 0.0
        if not math.isfinite(val) or val <= 0:
            val = 0.0
        pooled[index[r]] += val
    tot = sum(pooled)
    if tot <= 0:
        return uniq, None
    return uniq, [v / tot for v in pooled]


def _normalize_topousm_fast_radii_and_weights(target_distances, weights, pixel_size: float,
                                              manual_radii: Optional[Iterable] = None,
                                              manual_weights: Optional[Iterable] = None):
    """reference :43-90."""
    if pixel_size <= 0:
        pixel_size = 1.0
    source = manual_radii if manual_radii is not None else target_distances
    if source is None:
        return None, None
    out: List[int] = []
    for v in source:
        try:
            x = float(v)
        except (TypeError, ValueError):
            continue
        if not math.isfinite(x):
            continue
        r = int(round(x / pixel_size)) if manual_radii is None else int(round(x))
        out.append(max(1, r))
    if not out:
        return None, None
    return deduplicate_radii_weights(out, manual_weights if manual_weights is not None else weights)


def run_tile_algorithm(algo_instance, algorithm: str, dem_gpu, sigma: float, multiscale_mode: bool,
                       pixel_size: float, algo_params: Dict[str, Any]):
    """reference :93-129."""
    if algorithm == "topousm_fast":
        radii, ww = _normalize_topousm_fast_radii_and_weights(
            None, None, pixel_size, manual_radii=algo_params.get("radii"), manual_weights=algo_params.get("weights"))
        params = {"multiscale_mode": multiscale_mode, "radii": radii, "weights": ww,
                  "pixel_size": pixel_size, "sigma": sigma}
        for key in ("global_stats", "_topousm_fast_coarse_field", "_topousm_fast_small_radii",
                    "_topousm_fast_small_weights", "_topousm_fast_w_large", "_topousm_fast_full_shape",
                    "_topousm_fast_field_offset"):
            if algo_params.get(key) is not None:
                params[key] = algo_params[key]
        return algo_instance.process(dem_gpu, **params)
    return algo_instance.process(dem_gpu, **{"sigma": sigma, "pixel_size": pixel_size, **algo_params})
