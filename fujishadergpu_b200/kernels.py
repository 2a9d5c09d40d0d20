"""Functional Python face of the C ABI (include/fsg_b200.h): one function per entry point, taking
device arrays and returning device arrays.  No arithmetic happens here -- only argument marshalling,
output / workspace allocation (through torch's caching allocator) and error translation."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from . import _device as dev
from . import _lib
from ._lib import Encode, Window, check, make_encode, opt

_SLOPE_UNITS = {"degree": 0, "percent": 1, "radian": 2}
_CURV_TYPES = {"mean": 0, "gaussian": 1, "planform": 2, "profile": 3}


def _ptr(t: torch.Tensor) -> C.c_void_p:
    return C.c_void_p(t.data_ptr())


def _window(t: torch.Tensor, out: torch.Tensor, *, h_global=None, buf_row0=0, out_row0=None, out_rows=None) -> Window:
    H = int(t.shape[0]) if h_global is None else int(h_global)
    o0 = int(buf_row0) if out_row0 is None else int(out_row0)
    n = int(out.shape[0]) if out_rows is None else int(out_rows)
    return Window(H, int(t.shape[1]), int(buf_row0), int(t.shape[0]), o0, n, int(t.stride(0)), int(out.stride(0)))


def _band_args(t, band):
    """band = None (whole raster) or dict(h_global, buf_row0, out_row0, out_rows) for a row-band shard."""
    if band is None:
        return dict(h_global=int(t.shape[0]), buf_row0=0, out_row0=0, out_rows=int(t.shape[0]))
    return dict(h_global=int(band["h_global"]), buf_row0=int(band["buf_row0"]),
                out_row0=int(band["out_row0"]), out_rows=int(band["out_rows"]))


def hillshade(dem, *, azimuth=315, altitude=45, z_factor=1.0, pixel_size=1.0, pixel_scale_x=None,
              pixel_scale_y=None, output_dtype="float32", qp=None, band=None) -> torch.Tensor:
    t = dev.as_f32_2d(dem)
    b = _band_args(t, band)
    out = dev.empty_out(t, (b["out_rows"], t.shape[1]), output_dtype)
    win = _window(t, out, **b)
    enc = make_encode(output_dtype, qp)
    z = 1.0 if z_factor is None else float(z_factor)
    check(_lib.load().fsg_hillshade(_ptr(t), _ptr(out), C.byref(win), float(azimuth), float(altitude), z,
                                    float(pixel_size), opt(pixel_scale_x), opt(pixel_scale_y), C.byref(enc),
                                    C.c_void_p(dev.stream_ptr(t))), "fsg_hillshade")
    return out


def slope(dem, *, unit="degree", pixel_size=1.0, pixel_scale_x=None, pixel_scale_y=None,
          output_dtype="float32", qp=None, band=None) -> torch.Tensor:
    t = dev.as_f32_2d(dem)
    b = _band_args(t, band)
    out = dev.empty_out(t, (b["out_rows"], t.shape[1]), output_dtype)
    win = _window(t, out, **b)
    enc = make_encode(output_dtype, qp)
    # the reference treats any unit other than degree/percent as radian (_impl_slope.py:28-33)
    u = _SLOPE_UNITS.get(str(unit), 2)
    check(_lib.load().fsg_slope(_ptr(t), _ptr(out), C.byref(win), u, float(pixel_size), opt(pixel_scale_x),
                                opt(pixel_scale_y), C.byref(enc), C.c_void_p(dev.stream_ptr(t))), "fsg_slope")
    return out


def curvature(dem, *, curvature_type="mean", pixel_size=1.0, pixel_scale_x=None, pixel_scale_y=None,
              output_dtype="float32", qp=None, band=None) -> torch.Tensor:
    t = dev.as_f32_2d(dem)
    b = _band_args(t, band)
    out = dev.empty_out(t, (b["out_rows"], t.shape[1]), output_dtype)
    win = _window(t, out, **b)
    enc = make_encode(output_dtype, qp)
    # anything but mean/gaussian/planform falls to the profile branch (_impl_curvature.py:48)
    ct = _CURV_TYPES.get(str(curvature_type), 3)
    check(_lib.load().fsg_curvature(_ptr(t), _ptr(out), C.byref(win), ct, float(pixel_size), opt(pixel_scale_x),
                                    opt(pixel_scale_y), C.byref(enc), C.c_void_p(dev.stream_ptr(t))), "fsg_curvature")
    return out


def _radii_weights(radii: Sequence, weights: Optional[Sequence]):
    rr = np.asarray([int(r) for r in radii], dtype=np.int32)
    if weights is None:
        # np.array([1/n]*n, dtype=float32)  (_impl_topousm_fast.py:58-59)
        ww = np.asarray([1.0 / len(rr)] * len(rr), dtype=np.float32)
    else:
        ww = np.asarray(weights.cpu() if isinstance(weights, torch.Tensor) else weights, dtype=np.float32).reshape(-1)
        if len(ww) != len(rr):
            raise ValueError(f"Length of weights ({len(ww)}) must match length of radii ({len(rr)})")
    return np.ascontiguousarray(rr), np.ascontiguousarray(ww)


def topousm_fast_workspace_bytes(shape, radii, pixel_size=1.0) -> int:
    rr = np.ascontiguousarray(np.asarray([int(r) for r in radii], dtype=np.int32))
    return int(_lib.load().fsg_topousm_fast_workspace_bytes(
        int(shape[0]), int(shape[1]), rr.ctypes.data_as(C.POINTER(C.c_int32)), len(rr), float(pixel_size)))


def topousm_fast(dem, *, radii, weights=None, pixel_size=1.0, norm_scale=None, output_dtype="float32",
                 qp=None, workspace: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
                 roi=None) -> torch.Tensor:
    """Whole pipeline for one block == whole raster.  norm_scale=None -> raw block output.
    roi = (row0, rows, col0, cols): only that region of `out` is guaranteed to be computed."""
    t = dev.as_f32_2d(dem)
    if radii is None or len(radii) == 0:
        raise ValueError("At least one radius value is required")
    rr, ww = _radii_weights(radii, weights)
    lib = _lib.load()
    need = topousm_fast_workspace_bytes(t.shape, rr, pixel_size)
    if workspace is None or workspace.numel() * workspace.element_size() < need:
        workspace = torch.empty(max(need, 256), dtype=torch.uint8, device=t.device)
    if out is None:
        out = dev.empty_out(t, t.shape, output_dtype)
    enc = make_encode(output_dtype, qp)
    if roi is not None:
        check(lib.fsg_topousm_fast_roi(_ptr(t), _ptr(out), int(t.shape[0]), int(t.shape[1]), int(t.stride(0)),
                                       int(out.stride(0)), rr.ctypes.data_as(C.POINTER(C.c_int32)),
                                       ww.ctypes.data_as(C.POINTER(C.c_float)), len(rr), float(pixel_size),
                                       opt(norm_scale), C.byref(enc), _ptr(workspace),
                                       workspace.numel() * workspace.element_size(), int(roi[0]), int(roi[1]),
                                       int(roi[2]), int(roi[3]), C.c_void_p(dev.stream_ptr(t))), "fsg_topousm_fast_roi")
        return out
    check(lib.fsg_topousm_fast(_ptr(t), _ptr(out), int(t.shape[0]), int(t.shape[1]), int(t.stride(0)),
                               int(out.stride(0)), rr.ctypes.data_as(C.POINTER(C.c_int32)),
                               ww.ctypes.data_as(C.POINTER(C.c_float)), len(rr), float(pixel_size), opt(norm_scale),
                               C.byref(enc), _ptr(workspace), workspace.numel() * workspace.element_size(),
                               C.c_void_p(dev.stream_ptr(t))), "fsg_topousm_fast")
    return out


def topousm_large_part(block, field, *, w_large, off_r, off_c, full_h, full_w) -> torch.Tensor:
    t = dev.as_f32_2d(block)
    f = dev.as_f32_2d(field)
    out = dev.empty_out(t, t.shape, "float32")
    check(_lib.load().fsg_topousm_large_part(_ptr(t), _ptr(out), int(t.shape[0]), int(t.shape[1]), int(t.stride(0)),
                                             int(out.stride(0)), _ptr(f), int(f.shape[0]), int(f.shape[1]),
                                             int(f.stride(0)), int(off_r), int(off_c), int(full_h), int(full_w),
                                             float(w_large), C.c_void_p(dev.stream_ptr(t))), "fsg_topousm_large_part")
    return out


def openness_table(num_directions: int, max_distance: int, pixel_size=1.0, pixel_scale_x=None, pixel_scale_y=None):
    """Ray-sample table built with NumPy exactly as algorithms/_impl_openness.py:58-68,96-109 does
    (host integers and Python floats only -- no image arithmetic)."""
    angles = np.linspace(0, 2 * np.pi, num_directions, endpoint=False)
    dirs = np.stack([np.cos(angles), np.sin(angles)], axis=1)
    distances = np.unique((np.linspace(0.1, 1.0, 10) * max_distance).astype(int))
    distances = distances[distances > 0]
    sx = abs(float(pixel_scale_x)) if pixel_scale_x is not None else float(pixel_size)
    sy = abs(float(pixel_scale_y)) if pixel_scale_y is not None else float(pixel_size)
    if sx < 1e-9:
        sx = float(pixel_size) if pixel_size else 1.0
    if sy < 1e-9:
        sy = float(pixel_size) if pixel_size else 1.0
    start, ox, oy, dist = [0], [], [], []
    for d in range(num_directions):
        for r in distances:
            ax = int(round(float(r) * float(dirs[d][0])))
            ay = int(round(float(r) * float(dirs[d][1])))
            if ax == 0 and ay == 0:
                continue
            ox.append(ax)
            oy.append(ay)
            dist.append(max(float(np.hypot(float(ax) * sx, float(ay) * sy)), 1e-9))
        start.append(len(ox))
    return (np.asarray(start, np.int32), np.asarray(ox, np.int32), np.asarray(oy, np.int32),
            np.asarray(dist, np.float32))


def openness(dem, *, openness_type="positive", num_directions=16, max_distance=50, pixel_size=1.0,
             pixel_scale_x=None, pixel_scale_y=None, stretch=None, output_dtype="float32", qp=None,
             band=None) -> torch.Tensor:
    t = dev.as_f32_2d(dem)
    b = _band_args(t, band)
    out = dev.empty_out(t, (b["out_rows"], t.shape[1]), output_dtype)
    win = _window(t, out, **b)
    enc = make_encode(output_dtype, qp)
    start, ox, oy, dist = openness_table(int(num_directions), int(max_distance), pixel_size, pixel_scale_x, pixel_scale_y)
    lo, sc = (float("nan"), float("nan"))
    if isinstance(stretch, (tuple, list)) and len(stretch) >= 2 and float(stretch[1]) > 1e-12:
        lo, sc = float(stretch[0]), float(stretch[1])
    if len(ox) == 0:  # keep pointers valid
        ox = np.zeros(1, np.int32); oy = np.zeros(1, np.int32); dist = np.ones(1, np.float32)
    i32 = C.POINTER(C.c_int32)
    check(_lib.load().fsg_openness_samples(
        _ptr(t), _ptr(out), C.byref(win), 0 if openness_type == "positive" else 1, int(num_directions),
        start.ctypes.data_as(i32), ox.ctypes.data_as(i32), oy.ctypes.data_as(i32),
        dist.ctypes.data_as(C.POINTER(C.c_float)), lo, sc, C.byref(enc), C.c_void_p(dev.stream_ptr(t))),
        "fsg_openness")
    return out


def ao_table(num_samples: int, radius: float, pixel_size=1.0, pixel_scale_x=None, pixel_scale_y=None):
    """Ring-sample table of compute_ambient_occlusion_block, built with NumPy exactly as
    algorithms/_impl_ambient_occlusion.py:52-88 does (host integers and Python floats only)."""
    angles = np.linspace(0, 2 * np.pi, num_samples, endpoint=False)
    directions = np.stack([np.cos(angles), np.sin(angles)], axis=1)
    sx = abs(float(pixel_scale_x)) if pixel_scale_x is not None else float(pixel_size)
    sy = abs(float(pixel_scale_y)) if pixel_scale_y is not None else float(pixel_size)
    if sx < 1e-9:
        sx = float(pixel_size) if pixel_size else 1.0
    if sy < 1e-9:
        sy = float(pixel_size) if pixel_size else 1.0
    ox, oy, dist, fac = [], [], [], []
    for r_factor in (0.25, 0.5, 0.75, 1.0):
        r = radius * r_factor
        dx_all = np.round(r * directions[:, 0]).astype(int)
        dy_all = np.round(r * directions[:, 1]).astype(int)
        for i in range(num_samples):
            dx, dy = int(dx_all[i]), int(dy_all[i])
            if dx == 0 and dy == 0:
                continue
            ox.append(dx)
            oy.append(dy)
            dist.append(max(float(np.hypot(float(dx) * sx, float(dy) * sy)), 1e-9))
            fac.append(1.0 - (r_factor * 0.3))
    return (np.asarray(ox, np.int32), np.asarray(oy, np.int32), np.asarray(dist, np.float32), np.asarray(fac, np.float32))


def ambient_occlusion(dem, *, num_samples=16, radius=10.0, intensity=1.0, pixel_size=1.0, pixel_scale_x=None,
                      pixel_scale_y=None, stretch=None, output_dtype="float32", qp=None) -> torch.Tensor:
    """compute_ambient_occlusion_block (algorithms/_impl_ambient_occlusion.py:33-118) on one device block."""
    t = dev.as_f32_2d(dem)
    if int(num_samples) < 1 or int(num_samples) > 64:
        raise ValueError("ambient_occlusion: num_samples must be 1..64")
    ox, oy, dist, fac = ao_table(int(num_samples), float(radius), pixel_size, pixel_scale_x, pixel_scale_y)
    n = len(ox)
    if n == 0:  # keep pointers valid
        ox = np.zeros(1, np.int32); oy = np.zeros(1, np.int32); dist = np.ones(1, np.float32); fac = np.ones(1, np.float32)
    lib = _lib.load()
    H, W = int(t.shape[0]), int(t.shape[1])
    out = dev.empty_out(t, (H, W), output_dtype)
    wsb = int(lib.fsg_ambient_occlusion_workspace_bytes(H, W))
    ws = torch.empty(max(wsb, 256), dtype=torch.uint8, device=t.device)
    lo, sc = (float("nan"), float("nan"))
    if isinstance(stretch, (tuple, list)) and len(stretch) >= 2 and float(stretch[1]) > 1e-12:
        lo, sc = float(stretch[0]), float(stretch[1])
    enc = make_encode(output_dtype, qp)
    i32, f32 = C.POINTER(C.c_int32), C.POINTER(C.c_float)
    check(lib.fsg_ambient_occlusion(_ptr(t), _ptr(out), H, W, int(t.stride(0)), int(out.stride(0)), n,
                                    ox.ctypes.data_as(i32), oy.ctypes.data_as(i32), dist.ctypes.data_as(f32),
                                    fac.ctypes.data_as(f32), float(intensity), lo, sc, C.byref(enc), _ptr(ws),
                                    ws.numel(), C.c_void_p(dev.stream_ptr(t))), "fsg_ambient_occlusion")
    return out


def overview_average(arr: torch.Tensor, nodata=None) -> torch.Tensor:
    """One 2 x 2 AVERAGE overview level (ceil(H/2) x ceil(W/2)) of a float32 / int16 / uint8 device raster;
    members equal to `nodata` (NaN for float32) are skipped (core/dask_processor.py:201-228: OVERVIEW_RESAMPLING=AVERAGE)."""
    t = dev.as_tensor(arr)
    kinds = {torch.float32: _lib.FSG_OUT_F32, torch.int16: _lib.FSG_OUT_I16, torch.uint8: _lib.FSG_OUT_U8}
    if t.ndim != 2 or t.dtype not in kinds:
        raise ValueError("overview_average: 2-D float32 / int16 / uint8 raster expected")
    if t.stride(1) != 1:
        t = t.contiguous()
    H, W = int(t.shape[0]), int(t.shape[1])
    out = torch.empty(((H + 1) // 2, (W + 1) // 2), dtype=t.dtype, device=t.device)
    has = nodata is not None and not (isinstance(nodata, float) and nodata != nodata)
    check(_lib.load().fsg_overview_average(_ptr(t), _ptr(out), H, W, int(t.stride(0)), int(out.stride(0)), kinds[t.dtype],
                                           float(nodata) if has else 0.0, 1 if has else 0,
                                           C.c_void_p(dev.stream_ptr(t))), "fsg_overview_average")
    return out


def decimate(dem, factor: int) -> torch.Tensor:
    t = dev.as_f32_2d(dem)
    f = int(factor)
    if f <= 1:
        return t
    H, W = int(t.shape[0]), int(t.shape[1])
    lib = _lib.load()
    need = int(lib.fsg_decimate_workspace_bytes(H, W, f))
    ws = torch.empty(max(need, 256), dtype=torch.uint8, device=t.device)
    out = dev.empty_out(t, ((H + f - 1) // f, (W + f - 1) // f), "float32")
    check(lib.fsg_decimate(_ptr(t), _ptr(out), H, W, int(t.stride(0)), f, _ptr(ws), need,
                           C.c_void_p(dev.stream_ptr(t))), "fsg_decimate")
    return out


def upsample(coarse, shape) -> torch.Tensor:
    t = dev.as_f32_2d(coarse).contiguous()
    H, W = int(shape[0]), int(shape[1])
    if (int(t.shape[0]), int(t.shape[1])) == (H, W):
        return t
    out = dev.empty_out(t, (H, W), "float32")
    ws = torch.empty(256, dtype=torch.uint8, device=t.device)
    check(_lib.load().fsg_upsample(_ptr(t), _ptr(out), int(t.shape[0]), int(t.shape[1]), H, W, _ptr(ws), 256,
                                   C.c_void_p(dev.stream_ptr(t))), "fsg_upsample")
    return out


def encode(arr, qp: dict, output_dtype: str) -> torch.Tensor:
    t = dev.as_tensor(arr)
    if t.dtype != torch.float32:
        t = t.to(torch.float32)
    t = t.contiguous()
    out = dev.empty_out(t, t.shape, output_dtype)
    enc = make_encode(output_dtype, qp)
    check(_lib.load().fsg_encode_f32(_ptr(t), _ptr(out), t.numel(), C.byref(enc), C.c_void_p(dev.stream_ptr(t))),
          "fsg_encode_f32")
    return out


def scale(arr, scale_value: float) -> torch.Tensor:
    t = dev.as_tensor(arr).to(torch.float32).contiguous()
    out = torch.empty_like(t)
    check(_lib.load().fsg_scale_f32(_ptr(t), _ptr(out), t.numel(), float(scale_value), C.c_void_p(dev.stream_ptr(t))),
          "fsg_scale_f32")
    return out


def stretch(arr, lo: float, scale_value: float) -> torch.Tensor:
    t = dev.as_tensor(arr).to(torch.float32).contiguous()
    out = torch.empty_like(t)
    check(_lib.load().fsg_stretch_f32(_ptr(t), _ptr(out), t.numel(), float(lo), float(scale_value),
                                      C.c_void_p(dev.stream_ptr(t))), "fsg_stretch_f32")
    return out


def gaussian_nan(block, sigma: float) -> torch.Tensor:
    """handle_nan_with_gaussian(block, sigma, mode='nearest')[0] (reference _nan_utils.py:18-31)."""
    t = dev.as_f32_2d(block)
    H, W = int(t.shape[0]), int(t.shape[1])
    out = torch.empty((H, W), dtype=torch.float32, device=t.device)
    lib = _lib.load()
    need = int(lib.fsg_gaussian_nan_workspace_bytes(H, W, float(sigma)))
    ws = torch.empty(max(need, 256), dtype=torch.uint8, device=t.device)
    check(lib.fsg_gaussian_nan(_ptr(t), _ptr(out), H, W, int(t.stride(0)), float(sigma), _ptr(ws), need,
                               C.c_void_p(dev.stream_ptr(t))), "fsg_gaussian_nan")
    return out


_COMBINE_MODES = {"first_weighted": 0, "add_weighted": 1, "first_from_zero": 2, "max": 3, "min": 4, "copy": 5}


def combine(acc, resp, w: float, mode: str) -> None:
    """One step of the reference's response combiner (in place on `acc`)."""
    a = dev.as_tensor(resp)
    if not a.is_contiguous():
        a = a.contiguous()
    assert acc.is_contiguous() and acc.dtype == torch.float32 and a.dtype == torch.float32 and acc.numel() == a.numel()
    check(_lib.load().fsg_combine_f32(_ptr(a), _ptr(acc), int(a.numel()), float(w), _COMBINE_MODES[mode],
                                      C.c_void_p(dev.stream_ptr(a))), "fsg_combine_f32")


def _pooled_views(chunks):
    views = [dev.as_tensor(c) for c in chunks]
    views = [v if v.ndim == 2 else v.reshape(1, -1) for v in views]
    views = [v if (v.dtype == torch.float32 and v.stride(1) == 1) else v.to(torch.float32).contiguous() for v in views]
    if len(views) > 16:  # the C entry point pools up to 16 chunks
        views = [torch.cat([v.reshape(-1) for v in views]).reshape(1, -1)]
    return views


def order_stats(chunks: Sequence, rank: int, *, take_abs: bool, finite_only: bool):
    """(a[rank], a[rank+1], n) over the pooled non-NaN (or finite) samples of the 2-D device views in
    `chunks`; rank < 0 only counts.  One host sync per call (stats are host scalars by contract)."""
    views = _pooled_views(chunks)
    n = len(views)
    if n == 0:
        return float("nan"), float("nan"), 0
    lib = _lib.load()
    ptrs = (C.c_void_p * n)(*[v.data_ptr() for v in views])
    rows = (C.c_int64 * n)(*[int(v.shape[0]) for v in views])
    cols = (C.c_int64 * n)(*[int(v.shape[1]) for v in views])
    lds = (C.c_int64 * n)(*[int(v.stride(0)) if v.shape[0] > 1 else int(v.shape[1]) for v in views])
    d = views[0].device
    result = torch.empty(4, dtype=torch.float64, device=d)
    wsb = int(lib.fsg_order_stats_workspace_bytes())
    ws = torch.empty(wsb, dtype=torch.uint8, device=d)
    check(lib.fsg_order_stats(ptrs, rows, cols, lds, n, int(rank), 1 if take_abs else 0, 1 if finite_only else 0,
                              _ptr(result), _ptr(ws), wsb, C.c_void_p(dev.stream_ptr(views[0]))), "fsg_order_stats")
    res = result.cpu().tolist()
    return float(res[0]), float(res[1]), int(res[3])


def count_samples(chunks: Sequence, *, finite_only: bool = True):
    """Per-chunk counts of valid samples for up to 16 2-D device views (one launch, one host sync)."""
    views = [dev.as_tensor(c) for c in chunks]
    views = [v if (v.dtype == torch.float32 and v.stride(1) == 1) else v.to(torch.float32).contiguous() for v in views]
    out = []
    for i in range(0, len(views), 16):
        part = views[i:i + 16]
        n = len(part)
        ptrs = (C.c_void_p * n)(*[v.data_ptr() for v in part])
        rows = (C.c_int64 * n)(*[int(v.shape[0]) for v in part])
        cols = (C.c_int64 * n)(*[int(v.shape[1]) for v in part])
        lds = (C.c_int64 * n)(*[int(v.stride(0)) if v.shape[0] > 1 else int(v.shape[1]) for v in part])
        counts = torch.empty(n, dtype=torch.int64, device=part[0].device)
        check(_lib.load().fsg_count_samples(ptrs, rows, cols, lds, n, 1 if finite_only else 0, _ptr(counts),
                                            C.c_void_p(dev.stream_ptr(part[0]))), "fsg_count_samples")
        out.extend(int(x) for x in counts.cpu().tolist())
    return out


def percentile(chunks: Sequence, q: float, *, take_abs=False, finite_only=False) -> float:
    """np.percentile(sample, q) for an f32 sample (method 'linear'); NaN when the sample is empty.
    One pass structure: device-staged radix select (2 scans of the sample, one host synchronisation)."""
    views = _pooled_views(chunks)
    if not views:
        return float("nan")
    return staged_percentile(views, q, take_abs=take_abs, finite_only=finite_only, device=views[0].device)


def percentile_two_pass(chunks: Sequence, q: float, *, take_abs=False, finite_only=False) -> float:
    """Same result through fsg_order_stats (count first, then a[k], a[k+1]): the host derives the rank.
    Index and interpolation arithmetic are NumPy's own f32 scalar ops
    (numpy/lib/_function_base_impl.py: percentile -> _quantile -> _lerp)."""
    views = _pooled_views(chunks)
    _, _, n = order_stats(views, -1, take_abs=take_abs, finite_only=finite_only)
    if n == 0:
        return float("nan")
    q32 = np.true_divide(q, np.float32(100))
    vi = (n - 1) * q32                      # f32 scalar, like numpy's get_virtual_index
    prev = int(np.floor(vi))
    prev = min(max(prev, 0), n - 1)
    gamma = np.asanyarray(vi - np.floor(vi), dtype=np.asanyarray(vi).dtype)[()]
    lo, hi, _ = order_stats(views, prev, take_abs=take_abs, finite_only=finite_only)
    a, b = np.float32(lo), np.float32(hi)
    diff = b - a
    out = a + diff * gamma
    if gamma >= 0.5:
        out = b - diff * (1 - gamma)
    return float(out)


def synth_dem(shape, *, seed=20261017, nodata=False, device="cuda", row0=0, h_global=None, out=None) -> torch.Tensor:
    rows, W = int(shape[0]), int(shape[1])
    H = rows if h_global is None else int(h_global)
    if out is None:
        out = torch.empty((rows, W), dtype=torch.float32, device=device)
    check(_lib.load().fsg_synth_dem(_ptr(out), H, W, int(row0), rows, int(out.stride(0)), int(seed),
                                    1 if nodata else 0, C.c_void_p(dev.stream_ptr(out))), "fsg_synth_dem")
    return out


def launch_count() -> int:
    return int(_lib.load().fsg_launch_count())


def reset_launch_count() -> None:
    _lib.load().fsg_reset_launch_count()


def profile_enable(on: bool) -> None:
    _lib.load().fsg_profile_enable(1 if on else 0)


def profile_read(max_records: int = 256):
    """[(tag, milliseconds), ...] of the instrumented kernels launched by this thread since the last read."""
    tags = (C.c_int * max_records)()
    ms = (C.c_float * max_records)()
    n = _lib.load().fsg_profile_read(tags, ms, max_records)
    return [(int(tags[i]), float(ms[i])) for i in range(n)]


# ---- row-band shard stages (multi-GPU) ----------------------------------------------------------
def topousm_plan(radii: Sequence, pixel_size: float = 1.0) -> dict:
    """Per-radius evaluation plan: kind 0 = fused full-res box, 1 = decimated level, 2 = full-res plane."""
    rr = np.ascontiguousarray(np.asarray([int(r) for r in radii], dtype=np.int32))
    n = len(rr)
    kind = np.zeros(n, np.int32); factor = np.zeros(n, np.int32); size = np.zeros(n, np.int32)
    halo = C.c_int32(0)
    i32 = C.POINTER(C.c_int32)
    check(_lib.load().fsg_topousm_plan(rr.ctypes.data_as(i32), n, float(pixel_size), kind.ctypes.data_as(i32),
                                       factor.ctypes.data_as(i32), size.ctypes.data_as(i32), C.byref(halo)),
          "fsg_topousm_plan")
    return {"kind": kind.tolist(), "factor": factor.tolist(), "size": size.tolist(), "fused_halo": int(halo.value)}


def pyramid_band(band, factors: Sequence[int]):
    """f x f valid means of a band's own rows for every factor; returns (grids, has_void flags)."""
    t = dev.as_f32_2d(band)
    rows, W = int(t.shape[0]), int(t.shape[1])
    fl = np.ascontiguousarray(np.asarray(list(factors), dtype=np.int32))
    grids = [torch.empty(((rows + f - 1) // f, (W + f - 1) // f), dtype=torch.float32, device=t.device) for f in factors]
    flags = torch.zeros(16, dtype=torch.int32, device=t.device)
    ptrs = (C.c_void_p * len(grids))(*[g.data_ptr() for g in grids])
    check(_lib.load().fsg_pyramid_band(_ptr(t), rows, W, int(t.stride(0)), fl.ctypes.data_as(C.POINTER(C.c_int32)),
                                       len(grids), ptrs, _ptr(flags), C.c_void_p(dev.stream_ptr(t))), "fsg_pyramid_band")
    return grids, flags[: len(grids)]


def grid_mean_band(src, src_row0: int, gh: int, size: int, out_row0: int, out_rows: int) -> torch.Tensor:
    """NaN-aware box (size > 0, 'reflect') / sigma-1 Gaussian (size == 0, 'nearest') mean of grid rows
    [out_row0, out_row0+out_rows); `src` holds global rows [src_row0, src_row0+len(src))."""
    t = dev.as_f32_2d(src)
    gw = int(t.shape[1])
    out = torch.empty((int(out_rows), gw), dtype=torch.float32, device=t.device)
    lib = _lib.load()
    need = int(lib.fsg_grid_mean_band_workspace_bytes(int(out_rows), gw))
    ws = torch.empty(max(need, 256), dtype=torch.uint8, device=t.device)
    check(lib.fsg_grid_mean_band(_ptr(t), int(src_row0), int(t.shape[0]), int(gh), gw, int(t.stride(0)), int(size),
                                 int(out_row0), int(out_rows), _ptr(out), _ptr(ws), need,
                                 C.c_void_p(dev.stream_ptr(t))), "fsg_grid_mean_band")
    return out


def copy_rect(dst: torch.Tensor, src: torch.Tensor) -> torch.Tensor:
    """dst[:, :] = src[:, :] (2-D f32 views with unit column stride) on the copy engines; `src` may be another
    GPU's band mapped through symmetric memory (core/sharding.PeerBands)."""
    if dst.shape != src.shape or dst.ndim != 2 or dst.dtype != torch.float32 or src.dtype != torch.float32:
        raise ValueError("copy_rect: two f32 2-D tensors of one shape expected")
    if dst.stride(1) != 1 or src.stride(1) != 1:
        raise ValueError("copy_rect: unit column stride expected")
    check(_lib.load().fsg_copy_rect_f32(_ptr(dst), int(dst.stride(0)), _ptr(src), int(src.stride(0)), int(dst.shape[0]),
                                        int(dst.shape[1]), C.c_void_p(dev.stream_ptr(dst))), "fsg_copy_rect_f32")
    return dst


def grid_void_fill(grid, inplace: bool = False) -> torch.Tensor:
    """Enclosed-void fill of a whole decimated grid (returns a filled copy, or fills `grid` itself).  The kernels
    are gated by a device-side "grid has NaN" flag, so the call is cheap on a grid without voids."""
    t = dev.as_f32_2d(grid)
    t = t if (inplace and t.is_contiguous()) else t.contiguous().clone()
    gh, gw = int(t.shape[0]), int(t.shape[1])
    sigma = max(1.0, min(gh, gw) / 64.0)
    need = 2 * ((gh * gw * 4 + 255) // 256 * 256) + (int(4 * sigma + 0.5) + 1) * 8 + 2048
    ws = torch.empty(need, dtype=torch.uint8, device=t.device)
    check(_lib.load().fsg_grid_void_fill(_ptr(t), gh, gw, _ptr(ws), need, C.c_void_p(dev.stream_ptr(t))),
          "fsg_grid_void_fill")
    return t


def topousm_fused_band(dem_ext, dem_row0: int, H: int, out_row0: int, out_rows: int, *, radii, weights=None,
                       pixel_size=1.0, term_grids=None, term_grow0=None, norm_scale=None, output_dtype="float32",
                       qp=None, out: Optional[torch.Tensor] = None, workspace: Optional[torch.Tensor] = None,
                       norm_scale_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Fused pass over one row band.  term_grids[i] is None for fused terms, else a window of the term's
    mean grid whose first row is global grid row term_grow0[i].  norm_scale_dev: one f32 on the device that
    replaces norm_scale at kernel time (the launch does not wait for the statistics pre-pass)."""
    t = dev.as_f32_2d(dem_ext)
    W = int(t.shape[1])
    rr, ww = _radii_weights(radii, weights)
    n = len(rr)
    if out is None:
        out = dev.empty_out(t, (int(out_rows), W), output_dtype)
    enc = make_encode(output_dtype, qp)
    grids = [None if g is None else dev.as_f32_2d(g).contiguous() for g in (term_grids or [None] * n)]
    ptrs = (C.c_void_p * n)(*[(g.data_ptr() if g is not None else None) for g in grids])
    g0 = (C.c_int64 * n)(*[int(v) if v is not None else 0 for v in (term_grow0 or [0] * n)])
    gn = (C.c_int64 * n)(*[int(g.shape[0]) if g is not None else 0 for g in grids])
    lib = _lib.load()
    need = int(lib.fsg_topousm_fused_band_workspace_bytes(int(out_rows), W))
    if workspace is None or workspace.numel() * workspace.element_size() < need:
        workspace = torch.empty(max(need, 256), dtype=torch.uint8, device=t.device)
    if norm_scale_dev is not None and (norm_scale_dev.dtype != torch.float32 or not norm_scale_dev.is_cuda):
        raise TypeError("norm_scale_dev must be a float32 CUDA tensor")
    check(lib.fsg_topousm_fused_band_ws(
        _ptr(t), int(dem_row0), int(t.shape[0]), int(H), W, int(t.stride(0)), _ptr(out), int(out_row0), int(out_rows),
        int(out.stride(0)), rr.ctypes.data_as(C.POINTER(C.c_int32)), ww.ctypes.data_as(C.POINTER(C.c_float)), n,
        float(pixel_size), ptrs, g0, gn, opt(norm_scale),
        C.c_void_p(norm_scale_dev.data_ptr()) if norm_scale_dev is not None else None, C.byref(enc),
        _ptr(workspace), workspace.numel() * workspace.element_size(), C.c_void_p(dev.stream_ptr(t))),
        "fsg_topousm_fused_band_ws")
    return out


def valid_bbox(band, first_row: int, cov: int, n_rows: int, n_cols: int, row_index0: int,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """int32[4] on the device: (max of -row, max row, max of -col, max col) over the finite samples of the overview
    band[first_row + i * cov, j * cov] (row indices offset by row_index0); INT_MIN-like when nothing is finite."""
    t = dev.as_f32_2d(band)
    if out is None:
        out = torch.empty(4, dtype=torch.int32, device=t.device)
    check(_lib.load().fsg_valid_bbox(_ptr(t), int(t.stride(0)), int(first_row), int(cov), int(n_rows), int(n_cols),
                                     int(row_index0), _ptr(out), C.c_void_p(dev.stream_ptr(t))), "fsg_valid_bbox")
    return out


def reload_debug_switches() -> None:
    """Re-read the FSG_* environment switches (read once per process otherwise)."""
    _lib.load().fsg_debug_reload_switches()


def _chunk_args(views):
    n = len(views)
    ptrs = (C.c_void_p * max(n, 1))(*[v.data_ptr() for v in views])
    rows = (C.c_int64 * max(n, 1))(*[int(v.shape[0]) for v in views])
    cols = (C.c_int64 * max(n, 1))(*[int(v.shape[1]) for v in views])
    lds = (C.c_int64 * max(n, 1))(*[int(v.stride(0)) if v.shape[0] > 1 else int(v.shape[1]) for v in views])
    return ptrs, rows, cols, lds


def key_histogram(chunks, level: int, prefix: int, mask: int, *, take_abs: bool, finite_only: bool, device):
    """(hist int64[2048], count int) of this rank's sample keys matching (key & mask) == prefix."""
    views = _pooled_views(chunks) if chunks else []
    hist = torch.zeros(2048, dtype=torch.int32, device=device)
    cnt = torch.zeros(1, dtype=torch.int64, device=device)
    ptrs, rows, cols, lds = _chunk_args(views)
    check(_lib.load().fsg_key_histogram(ptrs, rows, cols, lds, len(views), int(level), int(prefix), int(mask),
                                        1 if take_abs else 0, 1 if finite_only else 0, _ptr(hist), _ptr(cnt),
                                        C.c_void_p(int(torch.cuda.current_stream(device).cuda_stream))), "fsg_key_histogram")
    return hist.to(torch.int64), cnt


def key_rank_info(chunks, key: int, *, take_abs: bool, finite_only: bool, device):
    """int64[2]: (#keys <= key, min key > key or 0xffffffff) over this rank's samples."""
    views = _pooled_views(chunks) if chunks else []
    out = torch.zeros(2, dtype=torch.int64, device=device)
    ptrs, rows, cols, lds = _chunk_args(views)
    check(_lib.load().fsg_key_rank_info(ptrs, rows, cols, lds, len(views), int(key), 1 if take_abs else 0,
                                        1 if finite_only else 0, _ptr(out),
                                        C.c_void_p(int(torch.cuda.current_stream(device).cuda_stream))), "fsg_key_rank_info")
    return out


def key_to_float(key: int, take_abs: bool) -> float:
    return float(_lib.load().fsg_key_to_float(int(key), 1 if take_abs else 0))


def staged_percentile(chunks, q: float, *, take_abs: bool, finite_only: bool, device, all_reduce=None,
                      scale_out: Optional[torch.Tensor] = None, min_valid: float = 1e-9, peer_exchange=None,
                      compact: bool = True):
    """np.percentile(sample, q) (method 'linear', f32 sample) of the union of every rank's chunks with the
    selection state kept on the device: the host only enqueues the radix-select stages and, between them,
    `all_reduce(tensor, op)` (op in {"sum", "min"}) of the exchange area -- no host round trip until the
    4-double result is read.  all_reduce=None: single device.  (reference: _normalization.py:22-32)
    scale_out (one f32 on the device): the value is finished on the device (NumPy's f32 lerp; NaN when the sample
    is empty or the value is NaN / <= min_valid) and NO host synchronisation happens; returns scale_out.
    peer_exchange (core/sharding.PeerExchange, instead of all_reduce): the exchange goes through symmetric memory --
    publish, one 8 us stream barrier, every rank sums the slots of all ranks itself -- instead of five collectives."""
    lib = _lib.load()
    views = _pooled_views(chunks) if chunks else []
    nx = int(lib.fsg_select_exchange_words())
    ws = torch.empty((int(lib.fsg_select_workspace_bytes()) + 7) // 8, dtype=torch.int64, device=device)
    res = torch.empty(4, dtype=torch.float64, device=device)
    stream = C.c_void_p(int(torch.cuda.current_stream(device).cuda_stream))
    ptrs, rows, cols, lds = _chunk_args(views)
    ta, fo = 1 if take_abs else 0, 1 if finite_only else 0
    q32 = np.true_divide(q, np.float32(100))       # NumPy's own f32 quantile (python float / f32 -> f32)
    check(lib.fsg_select_begin(_ptr(ws), ws.numel() * 8, stream), "fsg_select_begin")
    px = peer_exchange

    def exchange(stage):
        if px is not None:
            check(lib.fsg_select_peer_publish(_ptr(ws), C.c_void_p(px.my_slots), stage, stream), "fsg_select_peer_publish")
            px.barrier()
            check(lib.fsg_select_peer_reduce(_ptr(ws), C.c_void_p(px.peer_slots_dev), px.world, stage, stream),
                  "fsg_select_peer_reduce")
        elif all_reduce is not None:
            if stage < 3:
                all_reduce(ws[:2049], "sum")
            else:
                all_reduce(ws[2049:2050], "sum")
                all_reduce(ws[2050:2051], "min")

    # level 0 scans the sample; the keys of the selected bucket are copied out once and levels 1, 2 and the a[k+1] pass
    # read those (two scans instead of four)
    n_local = sum(int(v.shape[0]) * int(v.shape[1]) for v in views)
    keys = torch.empty(max(n_local, 1), dtype=torch.int32, device=device) if (compact and n_local) else None
    kptr = C.c_void_p(keys.data_ptr()) if keys is not None else C.c_void_p(0)
    for level in range(3):
        if level == 0 or not compact:
            check(lib.fsg_select_hist(ptrs, rows, cols, lds, len(views), level, ta, fo, _ptr(ws), stream), "fsg_select_hist")
        else:
            check(lib.fsg_select_hist_keys(kptr, level, _ptr(ws), stream), "fsg_select_hist_keys")
        exchange(level)
        check(lib.fsg_select_pick(level, float(q32), _ptr(ws), stream), "fsg_select_pick")
        if level == 0 and compact:
            check(lib.fsg_select_compact(ptrs, rows, cols, lds, len(views), ta, fo, _ptr(ws), kptr, n_local, stream),
                  "fsg_select_compact")
    if compact:
        check(lib.fsg_select_next_keys(kptr, _ptr(ws), stream), "fsg_select_next_keys")
    else:
        check(lib.fsg_select_next(ptrs, rows, cols, lds, len(views), ta, fo, _ptr(ws), stream), "fsg_select_next")
    exchange(3)
    if px is not None:
        px.barrier()   # the slots may be rewritten by the next call once every rank has read them
    if scale_out is not None:
        if scale_out.dtype != torch.float32 or not scale_out.is_cuda:
            raise TypeError("scale_out must be a float32 CUDA tensor")
        check(lib.fsg_select_finish_scale(_ptr(ws), ta, C.c_float(float(q32)), C.c_float(float(min_valid)),
                                          _ptr(scale_out), stream), "fsg_select_finish_scale")
        return scale_out
    check(lib.fsg_select_finish(_ptr(ws), ta, _ptr(res), stream), "fsg_select_finish")
    a_k, a_k1, rank_dev, n = [float(v) for v in res.cpu().tolist()]   # the only host synchronisation
    n = int(n)
    if n == 0:
        return float("nan")
    vi = (n - 1) * q32
    prev = min(max(int(np.floor(vi)), 0), n - 1)
    if prev != int(rank_dev):
        raise _lib.FsgError(f"staged_percentile: device rank {int(rank_dev)} != NumPy rank {prev} (n={n}, q={q})")
    gamma = np.asanyarray(vi - np.floor(vi), dtype=np.asanyarray(vi).dtype)[()]
    a_k, a_k1 = np.float32(a_k), np.float32(a_k1)
    diff = a_k1 - a_k
    out = a_k + diff * gamma
    if gamma >= 0.5:
        out = a_k1 - diff * (1 - gamma)
    return float(out)
