"""Shared normalisation / display-stretch helpers (reference: algorithms/_global_stats.py)."""
from __future__ import annotations

from typing import Any, Tuple

from .. import kernels as _k
from .. import _device as _dev


def apply_global_normalization(block, norm_func, stats: Tuple[Any, ...], nan_mask=None):
    """reference :123-153.  The CUDA normalise kernel keeps NaN where the block is NaN, which is the
    restore_nan step of the reference."""
    return norm_func(block, stats, nan_mask)


def _apply_display_stretch_block(result, stats):
    """reference algorithms/tile/dask_bridge.py:173-187 -- max((x - lo)/scale, 0); no-op without stats."""
    if not (isinstance(stats, (tuple, list)) and len(stats) >= 2):
        return result
    lo, sc = float(stats[0]), float(stats[1])
    if not (sc > 1e-12):
        return result
    return _dev.like_input(_k.stretch(result, lo, sc), result)


def apply_display_stretch_dask(result, stats):
    """reference :156-178 (block-level here: the whole raster is one block)."""
    return _apply_display_stretch_block(result, stats)


def robust_unsigned_stretch_stat_func(values) -> Tuple[float, float]:
    """reference :181-203 -- (p1, p99 - p1) of the finite samples."""
    if values is None:
        return (0.0, 0.0)
    chunks = list(values) if isinstance(values, (list, tuple)) else [values]
    lo = _k.percentile(chunks, 1.0, take_abs=False, finite_only=True)
    if lo != lo:
        return (0.0, 0.0)
    hi = _k.percentile(chunks, 99.0, take_abs=False, finite_only=True)
    sc = hi - lo
    if not (sc > 1e-12):
        return (lo, 0.0)
    return (lo, sc)


__all__ = ["apply_global_normalization", "apply_display_stretch_dask", "_apply_display_stretch_block",
           "robust_unsigned_stretch_stat_func"]
