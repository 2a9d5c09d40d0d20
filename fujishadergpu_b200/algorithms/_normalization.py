"""topousm_fast statistics / normalisation (reference: algorithms/_normalization.py)."""
from __future__ import annotations

from typing import Tuple

from .. import kernels as _k
from .. import _device as _dev

NORMAL_PERCENTILE = 99.0


def topousm_fast_stat_func(data) -> Tuple[float]:
    """reference :22-32 -- p99 of |x| over non-NaN samples (exact GPU selection); std fallback."""
    chunks = list(data) if isinstance(data, (list, tuple)) else [data]
    scale = _k.percentile(chunks, NORMAL_PERCENTILE, take_abs=True, finite_only=False)
    if scale == scale:  # non-empty sample
        if scale > 1e-9:
            return (scale,)
        import torch
        flat = torch.cat([_dev.as_tensor(c).reshape(-1) for c in chunks])
        flat = flat[~torch.isnan(flat)]
        sd = float(flat.double().std(unbiased=False).item()) if flat.numel() else 0.0  # rare degenerate-tile fallback
        return (sd if sd > 1e-9 else 1.0,)
    return (1.0,)


def topousm_fast_norm_func(block, stats, nan_mask=None):
    """reference :35-41 -- block / scale (no clip)."""
    s = float(stats[0])
    return _dev.like_input(_k.scale(block, s if s > 0 else 0.0), block)


__all__ = ["topousm_fast_stat_func", "topousm_fast_norm_func", "NORMAL_PERCENTILE"]
