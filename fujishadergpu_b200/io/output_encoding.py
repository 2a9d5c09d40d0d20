"""Integer output encoding (reference: io/output_encoding.py).  Host-side parameter logic is kept;
the array encode runs on the GPU (fused into the kernel epilogues, or ``quantize_array`` below)."""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np

_NORM_HEADROOM = 1.176
OUTPUT_VALUE_RANGES: Dict[str, Tuple[float, float]] = {
    "topousm_fast": (-_NORM_HEADROOM, _NORM_HEADROOM),
    "hillshade": (0.0, 1.0),
    "curvature": (0.0, 1.0),
    "openness": (0.0, 1.0),
    "ambient_occlusion": (0.0, 1.0),
    "slope": (0.0, 90.0),
}
SUPPORTED_OUTPUT_DTYPES = ("float32", "int16", "uint8")
_INT_MAXPOS: Dict[str, int] = {"int16": 32767, "uint8": 255}


def output_nodata_for_dtype(dtype) -> float:
    return float("nan") if np.dtype(dtype).kind == "f" else 0.0


def resolve_output_range(algorithm: str, *, params: Optional[dict] = None,
                         override: Optional[Tuple[float, float]] = None) -> Optional[Tuple[float, float]]:
    """reference :89-127 (hot-path algorithms)."""
    if override is not None:
        lo, hi = float(override[0]), float(override[1])
        if hi > lo:
            return lo, hi
        raise ValueError(f"output range must satisfy high > low, got ({lo!r}, {hi!r})")
    algo = str(algorithm).lower()
    if algo == "slope" and params is not None:
        unit = str(params.get("unit", "degree")).lower()
        if unit == "radian":
            return (0.0, float(np.pi / 2.0))
        if unit != "degree":
            return None
    return OUTPUT_VALUE_RANGES.get(algo)


def quantize_params(lo: float, hi: float, dtype: str) -> Dict[str, float]:
    """reference :130-175 -- DN = clip(round(a*v + b), dn_min, dn_max), NoData -> 0."""
    lo, hi = float(lo), float(hi)
    dt = str(dtype).lower()
    top = _INT_MAXPOS[dt]
    signed = lo < 0.0 < hi
    if signed:
        half = max(abs(lo), abs(hi))
        half = half if half > 0 else 1.0
        if dt == "int16":
            a_coef, b_coef, dn_min, dn_max = top / half, 0.0, -top, top
        else:
            a_coef, b_coef, dn_min, dn_max = (top - 1) / 2.0 / half, (top + 1) / 2.0, 1, top
    else:
        span = (hi - lo) if (hi - lo) > 0 else 1.0
        a_coef = (top - 1) / span
        b_coef, dn_min, dn_max = 1.0 - a_coef * lo, 1, top
    return {"a_coef": float(a_coef), "b_coef": float(b_coef), "dn_min": int(dn_min), "dn_max": int(dn_max),
            "scale": float(1.0 / a_coef), "offset": float(-b_coef / a_coef), "nodata": 0.0, "signed": bool(signed)}


def quantize_array(arr, qp: Dict[str, float], dtype: str):
    """reference :178-190 / core/dask_processor.py:997-1008 -- on the device."""
    from .. import kernels as _k
    from .. import _device as _dev
    return _dev.like_input(_k.encode(arr, qp, str(dtype)), arr)
