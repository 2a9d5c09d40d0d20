"""`python -m fujishadergpu_b200.cli INPUT.tif OUTPUT.tif --algorithm ... --mode ... --radii ... --weights ...`

The reference's command line (cli/args.py PIPELINE_ARGS + the knobs of the algorithms on the B200 path, cli/base.py
positional input / output, --algorithm, --pixel-size) over the CUDA library: the whole raster is ONE block on a
180 GB B200 (no tile loop, no Dask graph).  Flow = the reference's run_pipeline (core/dask_processor.py:1146-1490):
read the DEM, NoData -> NaN, per-axis pixel scales (north-up: +dx, -dy), radii / weights resolution
(--mode local -> radii [1], weights [1.0]; spatial without --radii -> auto_spatial_profile(short side)),
global statistics pre-pass, the algorithm, integer encoding, NoData re-mask, COG output.
"""
from __future__ import annotations

import argparse
import logging
import time
from typing import List, Optional

import numpy as np

logger = logging.getLogger("fujishadergpu_b200")

RADII_DRIVEN = {"hillshade", "slope", "curvature", "openness", "ambient_occlusion"}


def build_parser() -> argparse.ArgumentParser:
    from .algorithms.dask_registry import ALGORITHMS
    p = argparse.ArgumentParser(prog="fujishadergpu", description="Terrain shading on one B200 (CUDA library, no CuPy / Dask)")
    p.add_argument("input", help="Input DEM (GeoTIFF / COG, single band)")
    p.add_argument("output", help="Output file (COG)")
    p.add_argument("--algorithm", "--algo", default="topousm_fast", choices=sorted(ALGORITHMS))
    p.add_argument("--pixel-size", "--pixel_size", type=float, help="Pixel size in metres (default: from the GeoTIFF)")
    # cli/args.py PIPELINE_ARGS
    p.add_argument("--nodata", type=str, default=None)
    p.add_argument("--output-dtype", choices=["float32", "int16", "uint8"], default="float32")
    p.add_argument("--output-range", type=str, default=None)
    p.add_argument("--mode", choices=["local", "spatial"], default="spatial")
    p.add_argument("--radii", type=str)
    p.add_argument("--weights", type=str)
    p.add_argument("--agg", choices=["mean", "min", "max", "sum", "stack"], default="mean")
    # cli/args.py ALGORITHM_ARGS (the B200-path algorithms)
    p.add_argument("--azimuth", type=float, default=315.0)
    p.add_argument("--altitude", type=float, default=45.0)
    p.add_argument("--z-factor", type=float, default=1.0)
    p.add_argument("--multiscale", action="store_true")
    p.add_argument("--unit", choices=["degree", "percent", "radian"], default="degree")
    p.add_argument("--curvature-type", choices=["mean", "gaussian", "planform", "profile"], default="mean")
    p.add_argument("--radius", type=int, default=10)
    p.add_argument("--openness-type", choices=["positive", "negative"], default="positive")
    p.add_argument("--num-directions", type=int, default=16)
    p.add_argument("--max-distance", type=int, default=50)
    p.add_argument("--num-samples", type=int, default=16)
    p.add_argument("--intensity", type=float, default=1.0)
    p.add_argument("--device", default="cuda:0")
    p.add_argument("--verbose", action="store_true")
    return p


def _parse_list(text: Optional[str], typ) -> Optional[List]:
    if text is None or str(text).strip() == "":
        return None
    return [typ(float(t)) if typ is int else typ(t) for t in str(text).replace(";", ",").split(",") if t.strip()]


def _parse_nodata(text: Optional[str]) -> Optional[float]:
    """cli/args.py parse_nodata_override: a float, or nan / +nan / -nan."""
    if text is None:
        return None
    if str(text).strip().lower() in {"nan", "+nan", "-nan"}:
        return float("nan")
    try:
        return float(text)
    except ValueError as exc:
        raise ValueError(f"Invalid --nodata value: {text}") from exc


def _parse_output_range(text: Optional[str]):
    """cli/args.py parse_output_range: 'lo,hi' with hi > lo."""
    if text is None:
        return None
    try:
        lo_s, hi_s = str(text).split(",")
        lo, hi = float(lo_s), float(hi_s)
    except (ValueError, TypeError) as exc:
        raise ValueError(f"Invalid --output-range {text!r}; expected 'lo,hi' (e.g. 0,90).") from exc
    if not (hi > lo):
        raise ValueError(f"--output-range requires hi > lo, got {text!r}.")
    return (lo, hi)


def resolve_params(args, shape, pixel_size: float, psx: float, psy: float) -> dict:
    """Algorithm parameters as run_pipeline assembles them (core/dask_processor.py:1188-1295)."""
    from .algorithms.common.spatial_mode import LOCAL_RADII, LOCAL_WEIGHTS, auto_spatial_profile
    radii = _parse_list(args.radii, int)
    weights = _parse_list(args.weights, float)
    algo = args.algorithm
    short = min(int(shape[0]), int(shape[1]))
    params = {"mode": args.mode, "agg": args.agg, "intensity": args.intensity, "pixel_size": float(pixel_size),
              "pixel_scale_x": float(psx), "pixel_scale_y": float(psy), "is_geographic_dem": False}
    per_algo = {
        "hillshade": dict(azimuth=args.azimuth, altitude=args.altitude, z_factor=args.z_factor, multiscale=args.multiscale),
        "slope": dict(unit=args.unit),
        "curvature": dict(curvature_type=args.curvature_type),
        "openness": dict(radius=args.radius, openness_type=args.openness_type, num_directions=args.num_directions,
                         max_distance=args.max_distance),
        "ambient_occlusion": dict(num_samples=args.num_samples, radius=args.radius),
        "topousm_fast": {},
    }
    params.update(per_algo[algo])
    local = args.mode == "local"
    if algo == "topousm_fast":
        if local:
            if radii is not None:
                logger.warning("--mode local ignores explicit radii; forcing radii=%s.", LOCAL_RADII)
            radii, weights = list(LOCAL_RADII), list(LOCAL_WEIGHTS)
        elif radii is None:
            radii, weights = auto_spatial_profile(short)
        params.update(mode="radius", radii=radii, weights=weights)
    else:
        if local:
            if radii is not None:
                logger.warning("--mode local ignores explicit radii/scales; forcing radii=%s.", LOCAL_RADII)
            params.update(radii=list(LOCAL_RADII), weights=list(LOCAL_WEIGHTS))
        elif radii is not None:
            params.update(radii=radii, weights=weights)
        elif algo in RADII_DRIVEN:
            ar, aw = auto_spatial_profile(short)
            params.update(radii=ar, weights=weights if weights is not None else aw)
    return params


def resolve_pixel_scales(pixel_size_arg, meta: dict, shape):
    """(pixel_size, pixel_scale_x, pixel_scale_y, is_geographic) as run_pipeline injects them
    (core/dask_processor.py:1188-1221): signed metres per pixel from the raster's metadata (degrees converted at the
    centre latitude for a geographic CRS); an explicit --pixel-size makes the pixels isotropic, keeps the axis signs
    and drops the geographic approximation."""
    from .io.raster_info import metric_pixel_scales
    sx, sy, mean_m, is_geo, _lat = metric_pixel_scales(meta.get("transform"), meta.get("epsg"), shape)
    if pixel_size_arg is not None:
        p = float(pixel_size_arg)
        return p, (p if sx >= 0 else -p), (p if sy >= 0 else -p), False
    return float(mean_m), float(sx), float(sy), bool(is_geo)


def run(args) -> dict:
    import torch
    from . import kernels as _k
    from .algorithms._norm_stats import inject_global_stats
    from .algorithms.dask_registry import ALGORITHMS
    from .io.cog_writer import write_cog
    from .io.geotiff_reader import read_geotiff
    from .io.output_encoding import quantize_array, quantize_params, resolve_output_range

    t0 = time.perf_counter()
    nodata_override = _parse_nodata(args.nodata)
    dem, meta = read_geotiff(args.input)
    nod = nodata_override if nodata_override is not None else meta.get("nodata")
    dem = dem.astype(np.float32, copy=False)
    if nod is not None and nod == nod:
        dem = np.where(np.isclose(dem, np.float32(nod), rtol=0.0, atol=1e-6), np.float32(np.nan), dem)
    pixel_size, psx, psy, is_geo = resolve_pixel_scales(args.pixel_size, meta, dem.shape)
    params = resolve_params(args, dem.shape, pixel_size, psx, psy)
    params["is_geographic_dem"] = bool(is_geo)
    dev = torch.device(args.device)
    d = torch.from_numpy(np.ascontiguousarray(dem)).pin_memory().to(dev, non_blocking=True)
    t_read = time.perf_counter() - t0

    algo = ALGORITHMS[args.algorithm]
    inject_global_stats(d, args.algorithm, params)
    if args.algorithm == "topousm_fast" and not params.get("global_stats"):
        logger.warning("topousm_fast: no valid statistics window; output is not normalised")
    result = algo.process(d, **params)
    if hasattr(result, "ndim") and result.ndim == 3:
        raise NotImplementedError("--agg stack produces a multi-band raster; the COG writer is single-band")
    # final NoData re-mask for every output dtype, before the quantisation (core/dask_processor.py:1479-1547):
    # the spatial-mode smoothing fills voids, and a filled void must not become a valid DN
    nanmask = torch.isnan(d)
    result = torch.where(nanmask if result.ndim == 2 else nanmask.unsqueeze(0), torch.full_like(result, float("nan")), result)
    if args.output_dtype != "float32":
        override = _parse_output_range(args.output_range)
        rng = resolve_output_range(args.algorithm, params=params, override=override)
        if rng is None:
            raise ValueError(f"{args.algorithm}: --output-dtype {args.output_dtype} needs --output-range lo,hi")
        result = quantize_array(result, quantize_params(rng[0], rng[1], args.output_dtype), args.output_dtype)
    torch.cuda.synchronize(dev)
    t_gpu = time.perf_counter() - t0 - t_read
    stats = write_cog(args.output, result, transform=meta.get("transform"), epsg=meta.get("epsg"))
    total = time.perf_counter() - t0
    info = {"algorithm": args.algorithm, "shape": tuple(dem.shape), "params": {k: v for k, v in params.items() if not k.startswith("_")},
            "read_s": t_read, "gpu_s": t_gpu, "write_s": total - t_read - t_gpu, "total_s": total, "bytes": stats["bytes"],
            "launches": _k.launch_count()}
    logger.info("%s", info)
    return info


def main(argv=None) -> int:
    args = build_parser().parse_args(argv)
    logging.basicConfig(level=logging.INFO if args.verbose else logging.WARNING, format="%(levelname)s %(message)s")
    run(args)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
