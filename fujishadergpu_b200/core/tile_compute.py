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


def _nodata_is_nan(nodata) -> bool:
    try:
        return nodata is not None and float(nodata) != float(nodata)
    except (TypeError, ValueError):
        return False


def build_nodata_mask(data, nodata):
    """reference core/tile_processor.py:185-196 (_build_nodata_mask): numeric NoData (|x - nodata| <= 1e-6) and NaN
    cells are both NoData; returns a bool array of the input's kind (NumPy in, NumPy out; device in, device out)."""
    if nodata is None:
        return None
    import numpy as np
    if isinstance(data, np.ndarray):
        if _nodata_is_nan(nodata):
            return np.isnan(data)
        return np.isclose(data, float(nodata), rtol=0.0, atol=1e-6) | np.isnan(data)
    import torch
    from .. import _device as _dev
    t = _dev.as_tensor(data)
    if _nodata_is_nan(nodata):
        return torch.isnan(t)
    return ((t - float(nodata)).abs() <= 1e-6) | torch.isnan(t)


def apply_nodata_mask(result_gpu, mask_nodata, nodata):
    """reference :132-152 -- write the NoData value back over the masked pixels (2-D, HxWxC and CxHxW results)."""
    if mask_nodata is None:
        return result_gpu
    import torch
    from .. import _device as _dev
    res = _dev.as_tensor(result_gpu)
    mask = torch.as_tensor(mask_nodata, device=res.device).to(torch.bool)
    fill = float(nodata) if nodata is not None else 0.0
    if res.ndim == 2:
        res[mask] = fill
    elif res.ndim == 3:
        if tuple(res.shape[:2]) == tuple(mask.shape):
            res[mask, :] = fill
        elif tuple(res.shape[-2:]) == tuple(mask.shape):
            res[:, mask] = fill
        else:
            raise ValueError(f"Unsupported result/mask shapes for nodata masking: result={tuple(res.shape)}, "
                             f"mask={tuple(mask.shape)}")
    else:
        raise ValueError(f"Unsupported result ndim for nodata masking: {res.ndim}")
    return result_gpu
