"""Host-side radius / weight resolution and the device helpers the block functions share
(reference: algorithms/_nan_utils.py).  Only the pieces on the hot path are provided; the image
arithmetic itself lives in the CUDA library."""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np

from .. import kernels as _k
from .. import _device as _dev
from .common.spatial_mode import auto_spatial_profile, auto_spatial_radii, auto_spatial_weights

_ALGO_SCORE = {
    "topousm_fast": 1.15, "hillshade": 1.0, "slope": 1.0, "specular": 1.4, "atmospheric_scattering": 1.05,
    "curvature": 1.1, "ambient_occlusion": 1.5, "openness": 1.4, "multi_light_uncertainty": 1.25,
}


def _radius_to_downsample_factor(radius: float, *, block_shape: Optional[Tuple[int, int]] = None,
                                 pixel_size: float = 1.0, algorithm_name: str = "default",
                                 base_radius: float = 24.0, max_factor: int = 16) -> int:
    """reference :555-601 -- power-of-two decimation from the radius; independent of block_shape."""
    r = max(1.0, float(radius))
    px = max(1e-3, float(pixel_size) if pixel_size else 1.0)
    gain = float(_ALGO_SCORE.get(str(algorithm_name), 1.0))
    score = (r / max(1.0, base_radius)) * gain * 1.0 * (max(1.0, 1.0 / px) ** 0.35)
    if score <= 1.0:
        return 1
    return int(max(1, min(2 ** int(np.floor(np.log2(score))), max_factor)))


def _weight_count_matches(weights, expected: int) -> bool:
    if weights is None or isinstance(weights, (str, bytes)):
        return False
    try:
        return len(weights) == expected
    except TypeError:
        return False


def _clean_normalized_weights(weights) -> Optional[List[float]]:
    """reference :142-154 -- clip to >= 0, L1-normalise; None when nothing positive is left."""
    vals: List[float] = []
    for w in weights:
        try:
            x = float(w)
        except (TypeError, ValueError):
            return None
        vals.append(x if np.isfinite(x) and x > 0 else 0.0)
    tot = float(sum(vals))
    if tot <= 0:
        return None
    return [v / tot for v in vals]


def _normalize_spatial_radii(radii, pixel_size: float) -> List[int]:
    """reference :77-98 -- ints > 0, order kept, duplicates dropped; auto ladder when empty."""
    if radii is None:
        return auto_spatial_radii(None)
    seen, kept = set(), []
    for r in radii:
        try:
            v = int(round(float(r)))
        except (TypeError, ValueError):
            continue
        if v > 0 and v not in seen:
            seen.add(v)
            kept.append(v)
    return kept if kept else auto_spatial_radii(None)


def _resolve_spatial_radii_weights(radii, weights, pixel_size: float, short_side_px: Optional[float] = None):
    """reference :101-130."""
    if radii is None:
        rr, ww = auto_spatial_profile(short_side_px)
        if _weight_count_matches(weights, len(rr)):
            user = _clean_normalized_weights(weights)
            if user is not None:
                return rr, user
        return rr, ww
    rr = _normalize_spatial_radii(radii, pixel_size)
    if _weight_count_matches(weights, len(rr)):
        user = _clean_normalized_weights(weights)
        if user is not None:
            return rr, user
    return rr, auto_spatial_weights(len(rr))


def _downsample_nan_aware(block, factor: int):
    """reference :604-668 -- device array in, device array out."""
    if factor <= 1:
        return block
    return _dev.like_input(_k.decimate(block, int(factor)), block)


def _upsample_to_shape(block, target_shape):
    """reference :671-698."""
    return _dev.like_input(_k.upsample(block, target_shape), block)


def handle_nan_with_uniform(block, size: int, mode: str = "reflect"):
    """reference :34-47 -- (NaN-aware box mean, nan_mask); runs in libfsg_b200 (exact window sums)."""
    import torch
    if str(mode) != "reflect":
        raise NotImplementedError("handle_nan_with_uniform: only mode='reflect' is on the B200 path")
    t = _dev.as_f32_2d(block)
    gh = int(t.shape[0])
    return _dev.like_input(_k.grid_mean_band(t, 0, gh, int(size), 0, gh), block), torch.isnan(t)


def restore_nan(result, nan_mask):
    """reference :701-705 -- in-place result[mask] = NaN."""
    r = _dev.as_tensor(result)
    r[_dev.as_tensor(nan_mask).to(r.device)] = float("nan")
    return result


def handle_nan_with_gaussian(block, sigma: float, mode: str = "nearest"):
    """reference :18-31 -- (smoothed, nan_mask); the NaN-aware Gaussian runs in libfsg_b200."""
    import torch
    if str(mode) != "nearest":
        raise NotImplementedError("handle_nan_with_gaussian: only mode='nearest' is on the B200 path")
    t = _dev.as_f32_2d(block)
    return _dev.like_input(_k.gaussian_nan(t, float(sigma)), block), torch.isnan(t)


def _smooth_for_radius(block, radius: float, *, pixel_size: float = 1.0, algorithm_name: str = "default"):
    """reference :527-552 -- NaN-aware Gaussian smoothing controlled by the spatial radius
    (sigma = max(0.5, r/2); decimate -> smooth -> zoom back for large radii)."""
    r = max(1.0, float(radius))
    if r <= 1.0:
        return block
    factor = _radius_to_downsample_factor(r, block_shape=tuple(block.shape), pixel_size=pixel_size,
                                          algorithm_name=algorithm_name)
    if factor <= 1:
        return handle_nan_with_gaussian(block, sigma=max(0.5, r / 2.0), mode="nearest")[0]
    reduced = _downsample_nan_aware(block, factor)
    small = handle_nan_with_gaussian(reduced, sigma=max(0.5, (r / factor) / 2.0), mode="nearest")[0]
    return _upsample_to_shape(small, tuple(block.shape))


def large_radius_threshold(block, fallback: int) -> int:
    """reference :219-228 with one block == one chunk: max(256, min(H, W) // 16)."""
    try:
        min_chunk = min(int(block.shape[0]), int(block.shape[1]))
    except Exception:
        min_chunk = int(fallback)
    return int(max(256, int(min_chunk) // 16))


def coarsen_factor_for_shape(shape, coarse_max: int = 2048) -> int:
    """reference :231-236 -- power-of-two decimation so the longest side is <= coarse_max."""
    longest = max(int(shape[0]), int(shape[1]))
    if longest <= int(coarse_max):
        return 1
    return 1 << int(np.ceil(np.log2(longest / float(coarse_max))))


def _symmetric_pad(t, depth: int):
    """dask map_overlap(boundary='reflect') on a single chunk: edge-inclusive mirror padding."""
    import torch
    ri = torch.from_numpy(np.pad(np.arange(int(t.shape[0])), depth, mode="symmetric")).to(t.device)
    ci = torch.from_numpy(np.pad(np.arange(int(t.shape[1])), depth, mode="symmetric")).to(t.device)
    return t.index_select(0, ri).index_select(1, ci).contiguous()


def overlap_whole(block, block_fn, depth: int, **kw):
    """`block.map_overlap(block_fn, depth, boundary='reflect')` when the whole raster is ONE block (a device array
    handed to Algorithm.process): mirror-pad (edge inclusive) by the depth, evaluate, crop -- so the ring of `depth`
    pixels along the raster edge sees the mirrored neighbours the reference's Dask path gives it, not the block
    function's own edge rule."""
    t = _dev.as_f32_2d(block)
    H, W = int(t.shape[0]), int(t.shape[1])
    pad = max(1, min(int(depth), max(1, min(H, W) - 1)))
    r = _dev.as_f32_2d(block_fn(_symmetric_pad(t, pad), **kw))
    return _dev.like_input(r[pad:pad + H, pad:pad + W].contiguous(), block)


def coarse_large_radius_response(block, *, block_fn, radius_kw: str, radius: float, factor: int, depth_for_radius,
                                 pixel_size: float = 1.0, pixel_scale_x=None, pixel_scale_y=None,
                                 coarse_cache: Optional[dict] = None, coarse_dem=None, coarse_decimation=None,
                                 **block_kwargs):
    """reference :329-438 with one block == the whole raster: the response of a large radius is computed on a
    coarsened DEM (the injected overview when the host provides one, else the F x F NaN-aware block mean with
    the ragged edge trimmed, as da.coarsen(nanmean, trim_excess=True)), with radius and pixel sizes scaled by
    the factor, and sampled back bilinearly at pixel centres; NoData is restored from the DEM."""
    import torch
    t = _dev.as_f32_2d(block)
    H, W = int(t.shape[0]), int(t.shape[1])
    kw = dict(block_kwargs)
    kw.pop("grad_stats_map", None)
    if coarse_dem is not None and coarse_decimation is not None:
        fac = float(coarse_decimation)
        kw[radius_kw] = max(1, int(round(float(radius) / fac)))
        kw["pixel_size"] = float(pixel_size) * fac
        if pixel_scale_x is not None:
            kw["pixel_scale_x"] = float(pixel_scale_x) * fac
        if pixel_scale_y is not None:
            kw["pixel_scale_y"] = float(pixel_scale_y) * fac
        coarse_resp = _dev.as_f32_2d(block_fn(coarse_dem, **kw))
    else:
        F = int(factor)
        if F not in (2, 4, 8, 16):
            raise NotImplementedError(
                f"large spatial radius {radius}: coarsening factor {F} (raster longer than 32768 px) is not on the "
                "B200 path; pass the host's overview as _overview_coarse_dem / _overview_decimation")
        if coarse_cache is not None and "coarse" in coarse_cache:
            coarse = coarse_cache["coarse"]
        else:
            hc, wc = H // F, W // F
            coarse = _k.pyramid_band(t[: hc * F, : wc * F], [F])[0][0]     # block mean of the valid members, no void fill
            if coarse_cache is not None:
                coarse_cache["coarse"] = coarse
        r_coarse = max(1, int(round(float(radius) / float(F))))
        kw[radius_kw] = r_coarse
        kw["pixel_size"] = float(pixel_size) * float(F)
        if pixel_scale_x is not None:
            kw["pixel_scale_x"] = float(pixel_scale_x) * float(F)
        if pixel_scale_y is not None:
            kw["pixel_scale_y"] = float(pixel_scale_y) * float(F)
        d = int(depth_for_radius(r_coarse))
        padded = _symmetric_pad(coarse, d)
        resp = _dev.as_f32_2d(block_fn(padded, **kw))
        coarse_resp = resp[d:d + coarse.shape[0], d:d + coarse.shape[1]].contiguous()
    up = _dev.as_f32_2d(_bilinear_sample_coarse(coarse_resp, 0, H, 0, W, H, W))
    return _dev.like_input(torch.where(torch.isnan(t), torch.full_like(up, float("nan")), up), block)


def multiscale_response_fields(block, scales, *, block_fn, depth_for_scale, radius_kw: str = "scale", is_large=None,
                               pixel_size: float = 1.0, pixel_scale_x=None, pixel_scale_y=None,
                               coarse_dem=None, coarse_decimation=None, max_depth: int = 150, **block_kwargs) -> list:
    """reference :441-524 -- one response per scale; a large scale goes through the coarse path when the raster
    is longer than 2048 px (coarsening factor > 1) or the host injected an overview."""
    F = coarsen_factor_for_shape(tuple(block.shape))
    cache: dict = {}
    out = []
    for s in scales:
        sv = float(s)
        d = int(depth_for_scale(sv))
        large = is_large(sv) if is_large is not None else (d > int(max_depth))
        if large and (F > 1 or (coarse_dem is not None and coarse_decimation is not None)):
            out.append(coarse_large_radius_response(
                block, block_fn=block_fn, radius_kw=radius_kw, radius=sv, factor=F,
                depth_for_radius=lambda sc: max(1, int(depth_for_scale(sc))), pixel_size=pixel_size,
                pixel_scale_x=pixel_scale_x, pixel_scale_y=pixel_scale_y, coarse_cache=cache, coarse_dem=coarse_dem,
                coarse_decimation=coarse_decimation, **block_kwargs))
        else:
            # reference :516-522: map_overlap(block_fn, depth=max(1, min(d, smallest chunk - 1)), boundary='reflect') --
            # with one block == the whole raster that is: mirror-pad (edge inclusive) by the depth, evaluate, crop
            out.append(overlap_whole(block, block_fn, d, pixel_size=pixel_size, pixel_scale_x=pixel_scale_x,
                                     pixel_scale_y=pixel_scale_y, **{radius_kw: sv}, **block_kwargs))
    return out


def _accumulate(responses, weights32, first_mode: str):
    import torch
    acc = torch.empty_like(_dev.as_f32_2d(responses[0]))
    for i, resp in enumerate(responses):
        _k.combine(acc, _dev.as_f32_2d(resp), float(weights32[i]), first_mode if i == 0 else "add_weighted")
    return acc


def _combine_multiscale_dask(responses, *, weights=None, agg: str = "mean"):
    """reference :182-213 on device blocks: stack / max / min / sum / weighted mean (weights cleaned and
    L1-normalised in Python floats, each used as an f32 scalar) / plain mean."""
    import torch
    if not responses:
        raise ValueError("responses must not be empty")
    a = str(agg or "mean").lower()
    if a == "stack":
        return _dev.like_input(torch.stack([_dev.as_f32_2d(r) for r in responses], dim=0), responses[0])
    if len(responses) == 1:
        return responses[0]
    if a in ("max", "min"):
        acc = _dev.as_f32_2d(responses[0]).clone()
        for r in responses[1:]:
            _k.combine(acc, _dev.as_f32_2d(r), 1.0, a)
        return _dev.like_input(acc, responses[0])
    if a == "sum":   # da.sum over the stacked axis: pairwise only from 8 items on, i.e. in order here
        return _dev.like_input(_accumulate(responses, [1.0] * len(responses), "copy"), responses[0])
    if _weight_count_matches(weights, len(responses)):
        clean = _clean_normalized_weights(weights)
        if clean is not None:
            return _dev.like_input(_accumulate(responses, [np.float32(w) for w in clean], "first_weighted"), responses[0])
    # da.mean = sum / n
    acc = _accumulate(responses, [1.0] * len(responses), "copy")
    return _dev.like_input(_k.scale(acc, float(len(responses))), responses[0])


def _bilinear_sample_coarse(coarse, r0: int, r1: int, c0: int, c1: int, full_h: int, full_w: int):
    """reference :255-281 (through the fused large-part kernel with w_large = 0 on a zero block)."""
    import torch
    t = _dev.as_f32_2d(coarse)
    zero = torch.zeros((int(r1 - r0), int(c1 - c0)), dtype=torch.float32, device=t.device)
    up = _k.topousm_large_part(zero, t, w_large=0.0, off_r=int(r0), off_c=int(c0), full_h=int(full_h), full_w=int(full_w))
    return _dev.like_input(-up, coarse)


__all__ = [
    "_radius_to_downsample_factor", "_resolve_spatial_radii_weights", "_normalize_spatial_radii",
    "_clean_normalized_weights", "_weight_count_matches", "_downsample_nan_aware", "_upsample_to_shape",
    "_bilinear_sample_coarse", "handle_nan_with_uniform", "restore_nan", "handle_nan_with_gaussian", "_smooth_for_radius", "large_radius_threshold",
    "_combine_multiscale_dask", "coarsen_factor_for_shape", "coarse_large_radius_response",
    "multiscale_response_fields",
    "overlap_whole",
]
