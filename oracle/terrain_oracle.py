"""CPU oracle for the terrain-shading hot path  --  TEST INFRASTRUCTURE ONLY.

This module is a NumPy/SciPy restatement of the per-block arithmetic of
geoign/FujiShaderGPU v1.0.1 (the reference).  It exists so that the CUDA path
can be checked on machines where the reference itself cannot run (the reference
needs CuPy; neither this container nor the GPU box has it).  Nothing in the
product package may import it: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline / ``--impl reference`` legs do.

Arithmetic model
----------------
The reference's numeric kernels live in CuPy (``cupy-cuda12x>=13.4.0``, un-pinned,
not vendored).  ``cupyx.scipy.ndimage`` is a documented mirror of
``scipy.ndimage`` and ``cupy.gradient/percentile/pad`` mirror NumPy, so the
restatement uses SciPy/NumPy for those primitives:

* separable filters (``uniform_filter``, ``gaussian_filter``) accumulate in f64
  and round to f32 after EACH axis, axis 0 first;
* elementwise math is f32 with every operation individually rounded (no FMA);
* ``zoom(order=1)`` uses the align-corners mapping ``src = dst*(n_in-1)/(n_out-1)``.

Pinning
-------
The reference ships no golden arrays for this path ("parity unpinned" at the
CuPy boundary).  The oracle is therefore pinned against outputs of the
*unmodified reference modules* executed in the build container through the
NumPy import shim in ``oracle/ref_shim`` (see ``oracle/make_golden.py``); those
outputs are committed under ``tests/golden/`` and ``tests/test_oracle_golden.py``
holds the oracle to them, together with the reference's own known-answer tests
(flat openness == 1, analytic curvature, quantiser vector, auto-radii tables).

Each function cites the reference ``file:line`` it follows (paths relative to
``/root/reference/FujiShaderGPU``).
"""
from __future__ import annotations

import math
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np
from scipy import ndimage as _ndi

F32 = np.float32
GAMMA = 1 / 2.2          # algorithms/_base.py:13
AZIMUTH_DEFAULT = 315    # algorithms/_base.py:14
ALTITUDE_DEFAULT = 45    # algorithms/_base.py:15
NORM_HEADROOM = 1.176    # io/output_encoding.py:39
RADII_LADDER = (2, 8, 32, 128, 512, 2048)   # algorithms/common/spatial_mode.py:21
RADIUS_CAP = 2048                            # algorithms/common/spatial_mode.py:22


# --------------------------------------------------------------------------
# a1 / a2 : scale construction (host side, pure Python)
# --------------------------------------------------------------------------
def ladder_radii(short_side_px: Optional[float]) -> List[int]:
    """algorithms/common/spatial_mode.py:60-75 -- ladder truncated to
    min(2048, short_side/10); never empty."""
    cap = float(RADIUS_CAP) if short_side_px is None else min(float(RADIUS_CAP), float(short_side_px) / 10.0)
    keep = [r for r in RADII_LADDER if float(r) <= cap]
    return keep if keep else [RADII_LADDER[0]]


def pow2_weights(n: int) -> List[float]:
    """algorithms/common/spatial_mode.py:78-84 -- 2^(n-1-i) / sum."""
    if n <= 0:
        return []
    raw = [2.0 ** (n - 1 - i) for i in range(n)]
    tot = sum(raw)
    return [v / tot for v in raw]


def scale_profile(short_side_px, radii=None):
    """algorithms/common/spatial_mode.py:87-101."""
    rr = ladder_radii(short_side_px) if radii is None else [int(round(float(r))) for r in radii]
    return rr, pow2_weights(len(rr))


def merge_duplicate_radii(radii: Iterable[int], weights: Optional[Iterable[float]]):
    """core/tile_compute.py:9-40 -- keep first occurrence order, sum the weights
    of duplicates, clip non-finite / non-positive weights to 0, L1-normalise."""
    rl = list(radii)
    slot = {}
    uniq: List[int] = []
    for r in rl:
        if r not in slot:
            slot[r] = len(uniq)
            uniq.append(r)
    if weights is None:
        return uniq, None
    wl = list(weights)
    if len(wl) != len(rl):
        return uniq, None
    acc = [0.0] * len(uniq)
    for r, w in zip(rl, wl):
        try:
            fw = float(w)
        except (TypeError, ValueError):
            fw = 0.0
        if not math.isfinite(fw) or fw <= 0:
            fw = 0.0
        acc[slot[r]] += fw
    tot = sum(acc)
    if tot <= 0:
        return uniq, None
    return uniq, [v / tot for v in acc]


def tile_radii_weights(radii, weights):
    """core/tile_compute.py:43-90 (manual-radii branch): int(round(x)), floor 1."""
    if radii is None:
        return None, None
    out = []
    for v in radii:
        try:
            x = float(v)
        except (TypeError, ValueError):
            continue
        if not math.isfinite(x):
            continue
        out.append(max(1, int(round(x))))
    if not out:
        return None, None
    return merge_duplicate_radii(out, weights)


# --------------------------------------------------------------------------
# a3 : decimation factor
# --------------------------------------------------------------------------
_ALGO_FACTOR = {  # algorithms/_nan_utils.py:571-581
    "topousm_fast": 1.15, "hillshade": 1.0, "slope": 1.0, "specular": 1.4,
    "atmospheric_scattering": 1.05, "curvature": 1.1, "ambient_occlusion": 1.5,
    "openness": 1.4, "multi_light_uncertainty": 1.25,
}


def decimation_factor(radius: float, pixel_size: float = 1.0, algorithm: str = "default",
                      base_radius: float = 24.0, max_factor: int = 16) -> int:
    """algorithms/_nan_utils.py:555-601 -- power-of-two factor from the radius."""
    r = max(1.0, float(radius))
    px = max(1e-3, float(pixel_size) if pixel_size else 1.0)
    k = float(_ALGO_FACTOR.get(str(algorithm), 1.0))
    score = (r / max(1.0, base_radius)) * k * 1.0 * (max(1.0, 1.0 / px) ** 0.35)
    if score <= 1.0:
        return 1
    f = 2 ** int(np.floor(np.log2(score)))
    return int(max(1, min(f, max_factor)))


# --------------------------------------------------------------------------
# a5 / a6 : NaN-aware box / Gaussian means
# --------------------------------------------------------------------------
def box_mean(a: np.ndarray, size: int, mode: str = "reflect") -> np.ndarray:
    """algorithms/_nan_utils.py:34-47.  Returns the mean only."""
    a = np.asarray(a, dtype=F32)
    hole = np.isnan(a)
    if not hole.any():
        return _ndi.uniform_filter(a, size=size, mode=mode)
    ok = (~hole).astype(F32)
    vals = np.where(hole, F32(0), a) * ok
    num = _ndi.uniform_filter(vals, size=size, mode=mode)
    den = _ndi.uniform_filter(ok, size=size, mode=mode)
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(den > 0, num / den, F32(0)).astype(F32)


def gauss_mean(a: np.ndarray, sigma: float, mode: str = "nearest") -> np.ndarray:
    """algorithms/_nan_utils.py:18-31."""
    a = np.asarray(a, dtype=F32)
    hole = np.isnan(a)
    if not hole.any():
        return _ndi.gaussian_filter(a, sigma=sigma, mode=mode)
    ok = (~hole).astype(F32)
    vals = np.where(hole, F32(0), a) * ok
    num = _ndi.gaussian_filter(vals, sigma=sigma, mode=mode)
    den = _ndi.gaussian_filter(ok, sigma=sigma, mode=mode)
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(den > 0, num / den, F32(0)).astype(F32)


def box_mean_exact(a: np.ndarray, size: int) -> np.ndarray:
    """Explicit statement of what ``uniform_filter(f32, size, mode='reflect')``
    computes (the spec the CUDA kernels implement): per axis, the correctly
    rounded f32 of the exact window mean; axis 0 first, f32 between the axes.
    'reflect' = edge-inclusive mirror (``np.pad(mode='symmetric')``); the window
    for output i covers [i - size//2, i + size - 1 - size//2]."""
    out = np.asarray(a, dtype=F32)
    lo = size // 2
    hi = size - 1 - lo
    for ax in (0, 1):
        pad = [(0, 0), (0, 0)]
        pad[ax] = (lo, hi)
        p = np.pad(out.astype(np.float64), pad, mode="symmetric")
        c = np.cumsum(p, axis=ax)
        z = np.zeros_like(np.take(c, [0], axis=ax))
        c = np.concatenate([z, c], axis=ax)
        n = out.shape[ax]
        s = np.take(c, np.arange(size, size + n), axis=ax) - np.take(c, np.arange(0, n), axis=ax)
        out = (s / float(size)).astype(F32)
    return out


def gauss_taps(sigma: float, truncate: float = 4.0) -> np.ndarray:
    """scipy.ndimage._filters._gaussian_kernel1d: radius int(truncate*sigma+0.5),
    exp(-0.5 x^2 / sigma^2) normalised to sum 1 (f64)."""
    r = int(truncate * float(sigma) + 0.5)
    x = np.arange(-r, r + 1)
    w = np.exp(-0.5 / (float(sigma) * float(sigma)) * x ** 2)
    return w / w.sum()


# --------------------------------------------------------------------------
# a7 / a8 : decimate / upsample
# --------------------------------------------------------------------------
def decimate_cells(a: np.ndarray, f: int) -> np.ndarray:
    """First half of algorithms/_nan_utils.py:604-668 -- f x f mean of finite members anchored at (0,0),
    ragged edge NaN-padded, f32 sums in NumPy's reduction order; all-NaN cells stay NaN."""
    f = int(f)
    h, w = a.shape
    oh, ow = max(1, (h + f - 1) // f), max(1, (w + f - 1) // f)
    work = np.asarray(a, dtype=F32)
    if oh * f - h or ow * f - w:
        work = np.pad(work, ((0, oh * f - h), (0, ow * f - w)), mode="constant", constant_values=np.nan)
    cells = work.reshape(oh, f, ow, f)
    fin = np.isfinite(cells)
    cnt = fin.sum(axis=(1, 3), dtype=F32)
    tot = np.where(fin, cells, F32(0)).sum(axis=(1, 3), dtype=F32)
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(cnt > 0, tot / np.maximum(cnt, F32(1)), F32(np.nan)).astype(F32)


def fill_enclosed_voids(coarse: np.ndarray) -> np.ndarray:
    """Second half of algorithms/_nan_utils.py:604-668 -- enclosed coarse voids (Gaussian validity
    weight > 0.5, sigma = max(1, min(h, w)/64)) are filled, exterior voids stay NaN."""
    void = np.isnan(coarse)
    if void.any():
        sig = max(1.0, float(min(coarse.shape)) / 64.0)
        sv = _ndi.gaussian_filter(np.where(void, F32(0), coarse).astype(F32), sigma=sig, mode="nearest")
        sw = _ndi.gaussian_filter((~void).astype(F32), sigma=sig, mode="nearest")
        enclosed = void & (sw > F32(0.5))
        coarse = np.where(enclosed, sv / np.maximum(sw, F32(1e-6)), coarse).astype(F32)
    return coarse.astype(F32)


def decimate_valid_mean(a: np.ndarray, f: int) -> np.ndarray:
    """algorithms/_nan_utils.py:604-668."""
    if f <= 1:
        return a
    return fill_enclosed_voids(decimate_cells(a, f))


def upsample_align_corners(a: np.ndarray, shape: Tuple[int, int]) -> np.ndarray:
    """algorithms/_nan_utils.py:671-698 -- bilinear zoom to ``shape``."""
    th, tw = int(shape[0]), int(shape[1])
    h, w = a.shape
    if (h, w) == (th, tw):
        return a.astype(F32, copy=False)
    zy, zx = th / max(1, h), tw / max(1, w)
    hole = np.isnan(a)

    def fit(o):
        ph, pw = max(0, th - o.shape[0]), max(0, tw - o.shape[1])
        if ph or pw:
            o = np.pad(o, ((0, ph), (0, pw)), mode="edge")
        return o[:th, :tw]

    if not hole.any():
        return fit(_ndi.zoom(a, zoom=(zy, zx), order=1, mode="nearest").astype(F32))
    ok = (~hole).astype(F32)
    vals = np.where(hole, F32(0), a).astype(F32)
    num = _ndi.zoom(vals, zoom=(zy, zx), order=1, mode="nearest")
    den = _ndi.zoom(ok, zoom=(zy, zx), order=1, mode="nearest")
    with np.errstate(divide="ignore", invalid="ignore"):
        o = np.where(den > F32(1e-3), num / np.maximum(den, F32(1e-6)), F32(np.nan)).astype(F32)
    return fit(o)


def bilinear_align_corners_exact(a: np.ndarray, shape: Tuple[int, int]) -> np.ndarray:
    """Explicit statement of ``zoom(order=1, mode='nearest')`` for the plain
    branch (spec for the CUDA kernel): f64 evaluation of the four-tap sum at
    ``ri = i*(h-1)/(H-1)``, rounded once to f32."""
    H, W = int(shape[0]), int(shape[1])
    h, w = a.shape
    ad = a.astype(np.float64)
    ri = np.arange(H, dtype=np.float64) * ((h - 1) / (H - 1) if H > 1 else 0.0)
    ci = np.arange(W, dtype=np.float64) * ((w - 1) / (W - 1) if W > 1 else 0.0)
    r0 = np.floor(ri).astype(np.int64)
    c0 = np.floor(ci).astype(np.int64)
    r1 = np.minimum(r0 + 1, h - 1)
    c1 = np.minimum(c0 + 1, w - 1)
    fr = (ri - r0)[:, None]
    fc = (ci - c0)[None, :]
    top = ad[r0][:, c0] * (1 - fr) * (1 - fc) + ad[r0][:, c1] * (1 - fr) * fc
    bot = ad[r1][:, c0] * fr * (1 - fc) + ad[r1][:, c1] * fr * fc
    return (top + bot).astype(F32)


def sample_overview_field(coarse: np.ndarray, r0: int, r1: int, c0: int, c1: int,
                          full_h: int, full_w: int) -> np.ndarray:
    """algorithms/_nan_utils.py:255-281 -- pixel-centre bilinear sampling with
    f32 coordinates ``(i+0.5)*f32(ch/H)-0.5`` and edge clamping."""
    ch, cw = coarse.shape
    rr = (np.arange(r0, r1, dtype=F32) + F32(0.5)) * F32(ch / float(full_h)) - F32(0.5)
    cc = (np.arange(c0, c1, dtype=F32) + F32(0.5)) * F32(cw / float(full_w)) - F32(0.5)
    coords = np.empty((2, r1 - r0, c1 - c0), dtype=F32)
    coords[0] = rr[:, None]
    coords[1] = cc[None, :]
    return _ndi.map_coordinates(coarse, coords, order=1, mode="nearest").astype(F32)


# --------------------------------------------------------------------------
# a4 : topousm_fast block
# --------------------------------------------------------------------------
def topousm_plan(radii: Sequence[int], pixel_size: float = 1.0):
    """Per-radius evaluation plan ``(radius, ds, kind, param)`` with kind in
    {'box', 'gauss1'} (algorithms/_impl_topousm_fast.py:68-86)."""
    plan = []
    for r in radii:
        ds = decimation_factor(float(r), pixel_size=pixel_size, algorithm="topousm_fast")
        if ds > 1:
            rs = max(1, int(round(float(r) / ds)))
            plan.append((r, ds, "gauss1", 1.0) if rs <= 1 else (r, ds, "box", 2 * rs + 1))
        elif r <= 1:
            plan.append((r, 1, "gauss1", 1.0))
        else:
            plan.append((r, 1, "box", 2 * r + 1))
    return plan


def topousm_fast_block(dem: np.ndarray, *, radii=None, weights=None, pixel_size: float = 1.0) -> np.ndarray:
    """algorithms/_impl_topousm_fast.py:49-100 -- sum_i w_i*(dem - mean_i), f32,
    accumulated in list order, NaN restored."""
    if radii is None:
        radii = [4, 16, 64]
    dem = np.asarray(dem, dtype=F32)
    hole = np.isnan(dem)
    if weights is None:
        wv = np.array([1.0 / len(radii)] * len(radii), dtype=F32)
    else:
        wv = np.asarray(weights, dtype=F32)
        if len(wv) != len(radii):
            raise ValueError(f"Length of weights ({len(wv)}) must match length of radii ({len(radii)})")
    acc = None
    for (r, ds, kind, param), wt in zip(topousm_plan(radii, pixel_size), wv):
        if ds > 1:
            small = decimate_valid_mean(dem, ds)
            m_small = gauss_mean(small, 1.0, "nearest") if kind == "gauss1" else box_mean(small, param, "reflect")
            mean = upsample_align_corners(m_small, dem.shape)
        elif kind == "gauss1":
            mean = gauss_mean(dem, 1.0, "nearest")
        else:
            mean = box_mean(dem, param, "reflect")
        term = wt * (dem - mean)
        acc = term if acc is None else acc + term
    if hole.any():
        acc[hole] = np.nan
    return acc


def abs_p99_scale(values: np.ndarray) -> Tuple[float]:
    """algorithms/_normalization.py:22-32."""
    v = values[~np.isnan(values)]
    if len(v) > 0:
        s = float(np.percentile(np.abs(v), 99.0))
        if s > 1e-9:
            return (s,)
        sd = float(np.std(v))
        return (sd if sd > 1e-9 else 1.0,)
    return (1.0,)


def normalise_by_scale(block: np.ndarray, stats) -> np.ndarray:
    """algorithms/_normalization.py:35-41 + _global_stats.py:123-153."""
    hole = np.isnan(block)
    s = stats[0]
    out = block / s if s > 0 else np.zeros_like(block)
    if hole.any():
        out[hole] = np.nan
    return out.astype(F32)


def topousm_large_field(coarse_dem, *, large_radii, large_weights, decimation) -> np.ndarray:
    """algorithms/_impl_topousm_fast.py:133-155."""
    field = None
    for r, w in zip(large_radii, large_weights):
        rc = max(1, int(round(float(r) / max(decimation, 1.0))))
        m = gauss_mean(coarse_dem, 1.0, "nearest") if rc <= 1 else box_mean(coarse_dem, 2 * rc + 1, "reflect")
        t = F32(w) * m
        field = t if field is None else field + t
    return field.astype(F32)


def topousm_large_part(block, coarse_field, w_large, off_r, off_c, full_h, full_w) -> np.ndarray:
    """algorithms/_impl_topousm_fast.py:158-186 / tile/dask_bridge.py:237-249."""
    up = sample_overview_field(coarse_field, off_r, off_r + block.shape[0], off_c, off_c + block.shape[1], full_h, full_w)
    return (F32(w_large) * block - up).astype(F32)


def split_radii(radii, weights, threshold):
    """algorithms/_impl_topousm_fast.py:113-130 (weights not renormalised)."""
    n = len(radii)
    if weights is None or len(weights) != n:
        weights = [1.0 / n] * n
    sr, sw, lr, lw = [], [], [], []
    for r, w in zip(radii, weights):
        (lr if int(r) > int(threshold) else sr).append(int(r))
        (lw if int(r) > int(threshold) else sw).append(float(w))
    return sr, sw, lr, lw


# --------------------------------------------------------------------------
# a13 : NaN-aware gradient
# --------------------------------------------------------------------------
def _steps(pixel_size, psx, psy, signed=False):
    sy = float(psy if psy is not None else pixel_size)
    sx = float(psx if psx is not None else pixel_size)
    if not signed:
        sy, sx = abs(sy), abs(sx)
    if abs(sy) < 1e-9:
        sy = float(pixel_size if pixel_size else 1.0)
    if abs(sx) < 1e-9:
        sx = float(pixel_size if pixel_size else 1.0)
    return sy, sx


def gapfilled_gradient(dem, scale=1.0, pixel_size=1.0, psx=None, psy=None):
    """algorithms/_nan_utils.py:50-74 -- returns (dy, dx, nan_mask)."""
    hole = np.isnan(dem)
    if hole.any():
        if (~hole).any():
            filled = np.where(hole, gauss_mean(dem, 1.0, "nearest"), dem)
        else:
            filled = np.zeros_like(dem)
    else:
        filled = dem
    sy, sx = _steps(pixel_size, psx, psy, signed=False)
    dy, dx = np.gradient(filled * scale, sy, sx, edge_order=2)
    return dy, dx, hole


def gradient_f32_exact(f: np.ndarray, sy: float, sx: float):
    """Explicit statement of ``np.gradient(f32, sy, sx, edge_order=2)`` (spec for
    the CUDA kernels).  Every product/sum is an individually rounded f32 op;
    coefficients are Python floats rounded once to f32."""
    f = np.asarray(f, dtype=F32)
    outs = []
    for ax, h in ((0, sy), (1, sx)):
        g = np.empty_like(f)
        fm = np.moveaxis(f, ax, 0)
        gm = np.moveaxis(g, ax, 0)
        gm[1:-1] = (fm[2:] - fm[:-2]) / F32(2.0 * h)
        a, b, c = F32(-1.5 / h), F32(2.0 / h), F32(-0.5 / h)
        gm[0] = a * fm[0] + b * fm[1] + c * fm[2]
        a, b, c = F32(0.5 / h), F32(-2.0 / h), F32(1.5 / h)
        gm[-1] = a * fm[-3] + b * fm[-2] + c * fm[-1]
        outs.append(g)
    return outs[0], outs[1]


# --------------------------------------------------------------------------
# a14-a16 : hillshade / slope / curvature
# --------------------------------------------------------------------------
def hillshade_block(dem, *, azimuth=AZIMUTH_DEFAULT, altitude=ALTITUDE_DEFAULT, z_factor=1.0,
                    pixel_size=1.0, pixel_scale_x=None, pixel_scale_y=None) -> np.ndarray:
    """algorithms/_impl_hillshade.py:20-54.  Light vector is evaluated in f64 and
    demoted to f32 before it meets the f32 gradients -- CuPy's scalar-promotion
    rule (the NumPy-2 shim keeps f64 there; the two differ by <= 2e-7)."""
    dem = np.asarray(dem, dtype=F32)
    dy, dx, hole = gapfilled_gradient(dem, scale=z_factor, pixel_size=pixel_size, psx=pixel_scale_x, psy=pixel_scale_y)
    dy = dy.astype(F32, copy=False)
    dx = dx.astype(F32, copy=False)
    sgx = F32(1.0 if (pixel_scale_x is None or float(pixel_scale_x) >= 0.0) else -1.0)
    sgy = F32(1.0 if (pixel_scale_y is None or float(pixel_scale_y) >= 0.0) else -1.0)
    e = dx * sgx
    n = dy * sgy
    norm = np.sqrt(e * e + n * n + F32(1.0))
    alt = np.radians(altitude)
    az = np.radians(float(azimuth))
    lx = F32(np.sin(az) * np.cos(alt))
    ly = F32(np.cos(az) * np.cos(alt))
    lz = F32(np.sin(alt))
    hs = ((-e * lx) + (-n * ly) + lz) / norm
    hs = np.clip(hs, F32(0.0), F32(1.0)).astype(F32)
    if hole.any():
        hs[hole] = np.nan
    return hs


def slope_block(dem, *, unit="degree", pixel_size=1.0, pixel_scale_x=None, pixel_scale_y=None) -> np.ndarray:
    """algorithms/_impl_slope.py:19-35."""
    dem = np.asarray(dem, dtype=F32)
    dy, dx, hole = gapfilled_gradient(dem, scale=1, pixel_size=pixel_size, psx=pixel_scale_x, psy=pixel_scale_y)
    s = np.arctan(np.sqrt(dx ** 2 + dy ** 2))
    if unit == "degree":
        out = np.degrees(s)
    elif unit == "percent":
        out = np.tan(s) * 100
    else:
        out = s
    if hole.any():
        out[hole] = np.nan
    return out.astype(F32)


def curvature_block(dem, *, curvature_type="mean", pixel_size=1.0, pixel_scale_x=None, pixel_scale_y=None) -> np.ndarray:
    """algorithms/_impl_curvature.py:19-57 -- first derivatives with |step|,
    second derivatives with the SIGNED step (reference quirk kept)."""
    dem = np.asarray(dem, dtype=F32)
    hole = np.isnan(dem)
    sy, sx = _steps(pixel_size, pixel_scale_x, pixel_scale_y, signed=True)
    dy, dx, _ = gapfilled_gradient(dem, psx=sx, psy=sy)
    dyy, dyx = np.gradient(dy, sy, sx, edge_order=2)
    dxy, dxx = np.gradient(dx, sy, sx, edge_order=2)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        if curvature_type == "mean":
            p, q, r = dx, dy, dxx
            s = (dxy + dyx) / 2
            t = dyy
            den = np.power(1 + p ** 2 + q ** 2, 1.5)
            num = (1 + q ** 2) * r - 2 * p * q * s + (1 + p ** 2) * t
            k = -num / (2 * den + 1e-10)
        elif curvature_type == "gaussian":
            k = (dxx * dyy - dxy ** 2) / np.power(1 + dx ** 2 + dy ** 2, 2)
        elif curvature_type == "planform":
            k = (-(dy ** 2 * dxx - 2 * dx * dy * dxy + dx ** 2 * dyy) /
                 (np.power(dx ** 2 + dy ** 2, 1.5) + 1e-10))
        else:
            k = (-(dx ** 2 * dxx + 2 * dx * dy * dxy + dy ** 2 * dyy) /
                 ((dx ** 2 + dy ** 2) * np.power(1 + dx ** 2 + dy ** 2, 1.5) + 1e-10))
        out = (np.tanh(k * 100) + 1) / 2
        out = np.power(out, GAMMA)
    if hole.any():
        out[hole] = np.nan
    return out.astype(F32)


# --------------------------------------------------------------------------
# a17 / a18 : openness
# --------------------------------------------------------------------------
# --------------------------------------------------------------------------
# 8f-1 : spatial mode of the gradient family (Gaussian scale space)
# --------------------------------------------------------------------------
def smooth_for_radius(dem, radius, *, pixel_size: float = 1.0, algorithm: str = "default") -> np.ndarray:
    """algorithms/_nan_utils.py:527-552 -- NaN-aware Gaussian (sigma = max(0.5, r/2), 'nearest'),
    evaluated on the decimated grid and zoomed back for large radii."""
    a = np.asarray(dem, dtype=F32)
    r = max(1.0, float(radius))
    if r <= 1.0:
        return a
    f = decimation_factor(r, pixel_size=pixel_size, algorithm=algorithm)
    if f <= 1:
        return gauss_mean(a, max(0.5, r / 2.0), mode="nearest")
    small = decimate_valid_mean(a, f)
    sm = gauss_mean(small, max(0.5, (r / f) / 2.0), mode="nearest")
    return upsample_align_corners(sm, a.shape)


def hillshade_spatial_block(dem, *, radius=4.0, azimuth=AZIMUTH_DEFAULT, altitude=ALTITUDE_DEFAULT, z_factor=1.0,
                            pixel_size=1.0, pixel_scale_x=None, pixel_scale_y=None) -> np.ndarray:
    """algorithms/_impl_hillshade.py:57-67."""
    sm = smooth_for_radius(dem, radius, pixel_size=pixel_size, algorithm="hillshade")
    return hillshade_block(sm, azimuth=azimuth, altitude=altitude, z_factor=z_factor, pixel_size=pixel_size,
                           pixel_scale_x=pixel_scale_x, pixel_scale_y=pixel_scale_y)


def slope_spatial_block(dem, *, radius=4.0, unit="degree", pixel_size=1.0, pixel_scale_x=None,
                        pixel_scale_y=None) -> np.ndarray:
    """algorithms/_impl_slope.py:38-45."""
    sm = smooth_for_radius(dem, radius, pixel_size=pixel_size, algorithm="slope")
    return slope_block(sm, unit=unit, pixel_size=pixel_size, pixel_scale_x=pixel_scale_x, pixel_scale_y=pixel_scale_y)


def curvature_spatial_block(dem, *, radius=4.0, curvature_type="mean", pixel_size=1.0, pixel_scale_x=None,
                            pixel_scale_y=None) -> np.ndarray:
    """algorithms/_impl_curvature.py:72-78."""
    sm = smooth_for_radius(dem, radius, pixel_size=pixel_size, algorithm="curvature")
    return curvature_block(sm, curvature_type=curvature_type, pixel_size=pixel_size, pixel_scale_x=pixel_scale_x,
                           pixel_scale_y=pixel_scale_y)


def coarsen_factor(shape, coarse_max: int = 2048) -> int:
    """algorithms/_nan_utils.py:231-236."""
    longest = max(int(shape[0]), int(shape[1]))
    if longest <= int(coarse_max):
        return 1
    return 1 << int(np.ceil(np.log2(longest / float(coarse_max))))


def large_radius_response(dem, radius, factor: int, block_fn, depth: int, *, pixel_size=1.0, pixel_scale_x=None,
                          pixel_scale_y=None, **kw) -> np.ndarray:
    """algorithms/_nan_utils.py:329-438 for a single-chunk raster without an injected overview: da.coarsen with
    nanmean over factor x factor blocks (ragged edge trimmed), block function with map_overlap(depth,
    boundary='reflect') on the coarse array, pixel-centre bilinear sampling back, NoData restored."""
    a = np.asarray(dem, dtype=F32)
    H, W = a.shape
    F = int(factor)
    hc, wc = H // F, W // F
    cells = a[: hc * F, : wc * F].reshape(hc, F, wc, F)
    with np.errstate(invalid="ignore", divide="ignore"):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            coarse = np.nanmean(cells, axis=(1, 3)).astype(F32)
    r_c = max(1, int(round(float(radius) / float(F))))
    d = int(depth(r_c))
    padded = np.pad(coarse, d, mode="symmetric")
    resp = block_fn(padded, radius=r_c, pixel_size=float(pixel_size) * F,
                    pixel_scale_x=None if pixel_scale_x is None else float(pixel_scale_x) * F,
                    pixel_scale_y=None if pixel_scale_y is None else float(pixel_scale_y) * F, **kw)
    resp = np.ascontiguousarray(resp[d:d + hc, d:d + wc], dtype=F32)
    up = sample_overview_field(resp, 0, H, 0, W, H, W)
    return np.where(np.isnan(a), F32(np.nan), up).astype(F32)


def combine_responses(responses, *, weights=None, agg: str = "mean") -> np.ndarray:
    """algorithms/tile/dask_bridge.py:28-69 (_combine_direct): f32 weights normalised in f32, the sum
    accumulated in list order, every product / sum an individually rounded f32 op."""
    if not responses:
        raise ValueError("responses must not be empty")
    a = str(agg or "mean").lower()
    if a == "stack":
        return np.stack(responses, axis=0).astype(F32, copy=False)
    if len(responses) == 1:
        return responses[0]
    if a == "max":
        out = responses[0]
        for it in responses[1:]:
            out = np.maximum(out, it)
        return out.astype(F32, copy=False)
    if a == "min":
        out = responses[0]
        for it in responses[1:]:
            out = np.minimum(out, it)
        return out.astype(F32, copy=False)
    if a == "sum":
        out = np.zeros_like(responses[0], dtype=F32)
        for it in responses:
            out = out + it
        return out.astype(F32, copy=False)
    if isinstance(weights, (list, tuple)) and len(weights) == len(responses):
        w = np.asarray(weights, dtype=F32)
        if np.isfinite(w).all() and float(w.sum()) > 0:
            w = w / float(w.sum())
            out = responses[0] * F32(w[0])
            for i in range(1, len(responses)):
                out = out + responses[i] * F32(w[i])
            return out.astype(F32, copy=False)
    out = np.zeros_like(responses[0], dtype=F32)
    inv = F32(1.0 / float(len(responses)))
    for it in responses:
        out = out + it * inv
    return out.astype(F32, copy=False)


def openness_offsets(num_directions: int, max_distance: int):
    """Ray sample offsets ``[(dir, ox, oy), ...]`` and the pad depth D
    (algorithms/_impl_openness.py:58-68, 96-100).  Python banker's ``round``."""
    ang = np.linspace(0, 2 * np.pi, num_directions, endpoint=False)
    dirs = np.stack([np.cos(ang), np.sin(ang)], axis=1)
    dist = np.unique((np.linspace(0.1, 1.0, 10) * max_distance).astype(int))
    dist = dist[dist > 0]
    D = int(dist.max()) if dist.size else 0
    offs = []
    for d in range(num_directions):
        for r in dist:
            ox = int(round(float(r) * float(dirs[d][0])))
            oy = int(round(float(r) * float(dirs[d][1])))
            if ox == 0 and oy == 0:
                continue
            offs.append((d, ox, oy))
    return offs, D


def openness_block(dem, *, openness_type="positive", num_directions=16, max_distance=50,
                   pixel_size=1.0, pixel_scale_x=None, pixel_scale_y=None) -> np.ndarray:
    """algorithms/_impl_openness.py:31-132."""
    dem = np.asarray(dem, dtype=F32)
    h, w = dem.shape
    hole = np.isnan(dem)
    positive = openness_type == "positive"
    init = -np.pi / 2 if positive else np.pi / 2
    sx = abs(float(pixel_scale_x)) if pixel_scale_x is not None else float(pixel_size)
    sy = abs(float(pixel_scale_y)) if pixel_scale_y is not None else float(pixel_size)
    if sx < 1e-9:
        sx = float(pixel_size) if pixel_size else 1.0
    if sy < 1e-9:
        sy = float(pixel_size) if pixel_size else 1.0
    offs, D = openness_offsets(num_directions, max_distance)
    if D > 0:
        pv = np.pad(np.where(hole, F32(0.0), dem), D, mode="edge")
        pk = np.pad(~hole, D, mode="constant", constant_values=False)
    asum = np.zeros((h, w), dtype=F32)
    acnt = np.zeros((h, w), dtype=F32)
    for d in range(num_directions):
        ext = np.full((h, w), init, dtype=F32)
        seen = np.zeros((h, w), dtype=bool)
        for (dd, ox, oy) in offs:
            if dd != d:
                continue
            sh = pv[D + oy:D + oy + h, D + ox:D + ox + w]
            sk = pk[D + oy:D + oy + h, D + ox:D + ox + w]
            pd = max(float(np.hypot(float(ox) * sx, float(oy) * sy)), 1e-9)
            ang = np.arctan((sh - dem) / pd)
            ok = sk & ~hole
            ext = np.where(ok, np.maximum(ext, ang) if positive else np.minimum(ext, ang), ext)
            seen |= ok
        dang = (np.pi / 2 - ext) if positive else (np.pi / 2 + ext)
        asum += np.where(seen, dang, F32(0.0))
        acnt += seen.astype(F32)
    o = asum / np.maximum(acnt, F32(1.0))
    o = np.clip(o / (np.pi / 2), 0, 1)
    out = np.power(o, GAMMA)
    if hole.any():
        out[hole] = np.nan
    return out.astype(F32)


def openness_spatial_block(dem, *, openness_type="positive", num_directions=16, max_distance=50,
                           pixel_size=1.0, pixel_scale_x=None, pixel_scale_y=None) -> np.ndarray:
    """algorithms/_impl_openness.py:135-164."""
    ds = decimation_factor(float(max_distance), pixel_size=pixel_size, algorithm="openness")
    if ds <= 1:
        return openness_block(dem, openness_type=openness_type, num_directions=num_directions,
                              max_distance=max_distance, pixel_size=pixel_size,
                              pixel_scale_x=pixel_scale_x, pixel_scale_y=pixel_scale_y)
    small = decimate_valid_mean(np.asarray(dem, dtype=F32), ds)
    psx = float(abs(float(pixel_scale_x)) * ds) if pixel_scale_x is not None else None
    psy = float(abs(float(pixel_scale_y)) * ds) if pixel_scale_y is not None else None
    rs = openness_block(small, openness_type=openness_type, num_directions=num_directions,
                        max_distance=max(2, int(round(float(max_distance) / float(ds)))),
                        pixel_size=float(pixel_size) * float(ds), pixel_scale_x=psx, pixel_scale_y=psy)
    return upsample_align_corners(rs, dem.shape)


def ambient_occlusion_table(num_samples: int, radius: float):
    """Ring sample offsets ``[(ring factor, dx, dy), ...]`` in the reference's loop order and the pad depth D
    (algorithms/_impl_ambient_occlusion.py:52-80): np.round offsets, the (0, 0) ones skipped."""
    angles = np.linspace(0, 2 * np.pi, num_samples, endpoint=False)
    directions = np.stack([np.cos(angles), np.sin(angles)], axis=1)
    D = max(1, int(round(float(radius))))
    offs = []
    for r_factor in (0.25, 0.5, 0.75, 1.0):
        r = radius * r_factor
        dx_all = np.round(r * directions[:, 0]).astype(int)
        dy_all = np.round(r * directions[:, 1]).astype(int)
        for i in range(num_samples):
            dx, dy = int(dx_all[i]), int(dy_all[i])
            if dx == 0 and dy == 0:
                continue
            offs.append((r_factor, dx, dy))
    return offs, D


def ambient_occlusion_block(dem, *, num_samples=16, radius=10.0, intensity=1.0, pixel_size=1.0,
                            pixel_scale_x=None, pixel_scale_y=None) -> np.ndarray:
    """algorithms/_impl_ambient_occlusion.py:33-118."""
    dem = np.asarray(dem, dtype=F32)
    h, w = dem.shape
    hole = np.isnan(dem)
    sx = abs(float(pixel_scale_x)) if pixel_scale_x is not None else float(pixel_size)
    sy = abs(float(pixel_scale_y)) if pixel_scale_y is not None else float(pixel_size)
    if sx < 1e-9:
        sx = float(pixel_size) if pixel_size else 1.0
    if sy < 1e-9:
        sy = float(pixel_size) if pixel_size else 1.0
    offs, D = ambient_occlusion_table(num_samples, radius)
    padded = np.pad(dem, D, mode="edge")
    total = np.zeros((h, w), dtype=F32)
    count = np.zeros((h, w), dtype=F32)
    with np.errstate(invalid="ignore"):
        for (r_factor, dx, dy) in offs:
            sh = padded[D + dy:D + dy + h, D + dx:D + dx + w]
            dist = max(float(np.hypot(float(dx) * sx, float(dy) * sy)), 1e-9)
            ang = np.arctan((sh - dem) / dist)
            occ = np.maximum(0, ang) / (np.pi / 4)
            occ = np.minimum(occ, 1.0)
            ok = ~(np.isnan(sh) | hole)
            total += np.where(ok, occ * (1.0 - (r_factor * 0.3)), 0)
            count += np.where(ok, 1.0, 0)
    count = np.maximum(count, 1.0)
    ao = np.clip(1.0 - (total / count) * intensity, 0, 1)
    if hole.any():
        ao = np.where(hole, 1.0, ao)
    ao = _ndi.gaussian_filter(ao, sigma=1.0, mode="nearest")
    out = np.power(ao, GAMMA)
    if hole.any():
        out[hole] = np.nan
    return out.astype(F32)


def ambient_occlusion_spatial_block(dem, *, num_samples=16, radius=10.0, intensity=1.0, pixel_size=1.0,
                                    pixel_scale_x=None, pixel_scale_y=None) -> np.ndarray:
    """algorithms/_impl_ambient_occlusion.py:121-158."""
    ds = decimation_factor(float(radius), pixel_size=pixel_size, algorithm="ambient_occlusion")
    if ds <= 1:
        return ambient_occlusion_block(dem, num_samples=num_samples, radius=radius, intensity=intensity,
                                       pixel_size=pixel_size, pixel_scale_x=pixel_scale_x, pixel_scale_y=pixel_scale_y)
    small = decimate_valid_mean(np.asarray(dem, dtype=F32), ds)
    psx = float(abs(float(pixel_scale_x)) * ds) if pixel_scale_x is not None else None
    psy = float(abs(float(pixel_scale_y)) * ds) if pixel_scale_y is not None else None
    rs = ambient_occlusion_block(small, num_samples=num_samples, radius=max(1.0, float(radius) / float(ds)),
                                 intensity=intensity, pixel_size=float(pixel_size) * float(ds),
                                 pixel_scale_x=psx, pixel_scale_y=psy)
    return upsample_align_corners(rs, dem.shape)


def overview_average_2x(a: np.ndarray, nodata=None) -> np.ndarray:
    """One 2 x 2 AVERAGE overview level as the COG writer's GPU kernel defines it (GDAL OVERVIEW_RESAMPLING=AVERAGE,
    core/dask_processor.py:201-228): members equal to `nodata` (NaN for floats) are skipped, an empty cell is NoData,
    integers round half up, a mean that collides with the NoData value moves to the nearest valid DN."""
    h, w = a.shape
    oh, ow = (h + 1) // 2, (w + 1) // 2
    is_f = a.dtype.kind == "f"
    pad = np.zeros((oh * 2, ow * 2), dtype=np.float64)
    ok = np.zeros((oh * 2, ow * 2), dtype=bool)
    pad[:h, :w] = np.where(np.isnan(a), 0.0, a) if is_f else a
    ok[:h, :w] = ~np.isnan(a) if is_f else (np.ones_like(a, bool) if nodata is None else a != nodata)
    v = np.where(ok, pad, 0.0)
    s = v[0::2, 0::2] + v[0::2, 1::2]
    s = s + v[1::2, 0::2]
    s = s + v[1::2, 1::2]
    n = ok[0::2, 0::2].astype(np.int64) + ok[0::2, 1::2] + ok[1::2, 0::2] + ok[1::2, 1::2]
    with np.errstate(divide="ignore", invalid="ignore"):
        mean = s / n
    if is_f:
        return np.where(n > 0, mean, np.nan).astype(a.dtype)
    m = np.floor(mean + 0.5)
    if nodata is not None:
        m = np.where((n > 0) & (m == nodata), np.where(mean >= nodata, nodata + 1.0, nodata - 1.0), m)
        m = np.where(n > 0, m, float(nodata))
    return m.astype(a.dtype)


def display_stretch(block, stats) -> np.ndarray:
    """algorithms/tile/dask_bridge.py:173-187 / _global_stats.py:156-178."""
    if not (isinstance(stats, (tuple, list)) and len(stats) >= 2):
        return block
    lo, sc = float(stats[0]), float(stats[1])
    if not (sc > 1e-12):
        return block
    return np.maximum((block - F32(lo)) / F32(sc), F32(0.0)).astype(F32)


def p1_p99_stretch_stats(values) -> Tuple[float, float]:
    """algorithms/_global_stats.py:181-203."""
    if values is None:
        return (0.0, 0.0)
    v = values[np.isfinite(values)]
    if v.size == 0:
        return (0.0, 0.0)
    lo = float(np.percentile(v, 1.0))
    hi = float(np.percentile(v, 99.0))
    sc = hi - lo
    return (lo, sc) if sc > 1e-12 else (lo, 0.0)


# --------------------------------------------------------------------------
# a19 : integer encoding
# --------------------------------------------------------------------------
VALUE_RANGES = {  # io/output_encoding.py:40-73 (hot-path algorithms only)
    "topousm_fast": (-NORM_HEADROOM, NORM_HEADROOM),
    "hillshade": (0.0, 1.0), "curvature": (0.0, 1.0), "openness": (0.0, 1.0),
    "slope": (0.0, 90.0),
}
_MAXPOS = {"int16": 32767, "uint8": 255}


def value_range(algorithm: str, params: Optional[dict] = None, override=None):
    """io/output_encoding.py:89-127 (hot-path subset)."""
    if override is not None:
        lo, hi = float(override[0]), float(override[1])
        if hi > lo:
            return lo, hi
        raise ValueError(f"output range must satisfy high > low, got ({lo!r}, {hi!r})")
    a = str(algorithm).lower()
    if a == "slope" and params is not None:
        u = str(params.get("unit", "degree")).lower()
        if u == "radian":
            return (0.0, float(np.pi / 2.0))
        if u != "degree":
            return None
    return VALUE_RANGES.get(a)


def encode_params(lo: float, hi: float, dtype: str) -> dict:
    """io/output_encoding.py:130-175."""
    lo, hi = float(lo), float(hi)
    dt = str(dtype).lower()
    top = _MAXPOS[dt]
    if lo < 0.0 < hi:
        half = max(abs(lo), abs(hi)) or 1.0
        if dt == "int16":
            a, b, dmin, dmax = top / half, 0.0, -top, top
        else:
            a, b, dmin, dmax = (top - 1) / 2.0 / half, (top + 1) / 2.0, 1, top
        signed = True
    else:
        span = (hi - lo) if (hi - lo) > 0 else 1.0
        a = (top - 1) / span
        b, dmin, dmax = 1.0 - a * lo, 1, top
        signed = False
    return {"a_coef": float(a), "b_coef": float(b), "dn_min": int(dmin), "dn_max": int(dmax),
            "scale": float(1.0 / a), "offset": float(-b / a), "nodata": 0.0, "signed": signed}


def encode_array(arr, qp: dict, dtype: str) -> np.ndarray:
    """io/output_encoding.py:178-190 == core/dask_processor.py:997-1008."""
    a = np.asarray(arr, dtype=F32)
    with np.errstate(invalid="ignore", over="ignore"):
        dn = np.rint(F32(qp["a_coef"]) * a + F32(qp["b_coef"]))
        dn = np.clip(dn, qp["dn_min"], qp["dn_max"])
    dn = np.where(np.isfinite(a), dn, 0.0)
    return dn.astype(np.dtype(dtype))


# --------------------------------------------------------------------------
# a12 / a21 : stats-window and halo geometry (host ints)
# --------------------------------------------------------------------------
def stats_windows(width, height, by0, by1, bx0, bx1, *, grid=3, tile=4096):
    """algorithms/_norm_stats.py:64-100 -- de-duplicated (wy0, wx0, w, h) list."""
    ch = max(1, (int(by1) - int(by0)) // int(grid))
    cw = max(1, (int(bx1) - int(bx0)) // int(grid))
    out, seen = [], set()
    for gy in range(int(grid)):
        for gx in range(int(grid)):
            cy = int(by0) + gy * ch + ch // 2
            cx = int(bx0) + gx * cw + cw // 2
            wy0 = int(min(max(0, cy - tile // 2), max(0, int(height) - tile)))
            wx0 = int(min(max(0, cx - tile // 2), max(0, int(width) - tile)))
            key = (wy0, wx0, min(tile, int(width) - wx0), min(tile, int(height) - wy0))
            if key not in seen:
                seen.add(key)
                out.append(key)
    return out


def stats_window_geometry(algorithm: str, params: dict, max_tile: int = 4096):
    """algorithms/_norm_stats.py:103-162 (hot-path algorithms): (margin, tile)."""
    vals = []
    v = params.get("radii")
    if isinstance(v, (list, tuple)) and v:
        vals.append(max(float(x) for x in v))
    if params.get("max_distance"):
        vals.append(float(params["max_distance"]))
    top = max(vals) if vals else 16.0
    margin = max(1, int(top + 16))
    return margin, max(min(2048, max(1, int(max_tile))), 4 * margin)


def with_overlap(fn, dem: np.ndarray, depth: int, **kw) -> np.ndarray:
    """`dem.map_overlap(fn, depth, boundary='reflect')` with one chunk == the whole raster (the spatial-mode small
    radii, algorithms/_nan_utils.py:516-522): edge-inclusive mirror padding by the depth, the block function, crop."""
    h, w = dem.shape
    d = max(1, min(int(depth), max(1, min(h, w) - 1)))
    out = fn(np.pad(dem, d, mode="symmetric"), **kw)
    return np.ascontiguousarray(out[..., d:d + h, d:d + w])


def halo_depth(algorithm: str, params: dict) -> int:
    """map_overlap depth per algorithm: _impl_topousm_fast.py:203, _impl_hillshade.py:133,
    _impl_slope.py:71, _impl_curvature.py:92, _impl_openness.py:204-206."""
    a = str(algorithm)
    if a == "topousm_fast":
        return int(max(params["radii"]) + 16)
    if a in ("hillshade", "slope"):
        return 1
    if a == "curvature":
        return 2
    if a == "openness":
        return int(params.get("max_distance", 50)) + 1
    raise KeyError(a)


# --------------------------------------------------------------------------
# synthetic DEMs -- CPU generator for tests / fixtures
# --------------------------------------------------------------------------
def synth_dem(h: int, w: int, seed: int = 20261017, nodata: bool = False, relief: float = 250.0) -> np.ndarray:
    """Eight-octave sinusoid terrain + 0.25 m white noise, ~0-900 m, f32."""
    rng = np.random.default_rng(seed)
    rr = np.arange(h, dtype=np.float64)[:, None]
    cc = np.arange(w, dtype=np.float64)[None, :]
    z = np.full((h, w), 400.0)
    for k in range(8):
        lam = 4096.0 / (2 ** k)
        th = rng.uniform(0, 2 * np.pi)
        ph = rng.uniform(0, 2 * np.pi)
        z += relief * 2 ** (-0.9 * k) * np.sin(2 * np.pi * (rr * np.sin(th) + cc * np.cos(th)) / lam + ph)
    z += 0.25 * rng.standard_normal((h, w))
    z = z.astype(F32)
    if nodata:
        r = np.arange(h)[:, None]
        c = np.arange(w)[None, :]
        z[c < 0.02 * w * (1 + np.sin(r / 977.0))] = np.nan
        for (cy, cx, ay, ax) in ((0.3, 0.6, 0.02, 0.05), (0.7, 0.35, 0.06, 0.03), (0.55, 0.8, 0.01, 0.012)):
            m = ((r - cy * h) / max(2.0, ay * h)) ** 2 + ((c - cx * w) / max(2.0, ax * w)) ** 2 <= 1.0
            z[m] = np.nan
    return z
