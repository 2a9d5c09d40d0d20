"""Host-buffer entry point: one tile (= whole raster on a B200) from host memory to host memory
(reference: core/tile_processor.py::process_single_tile :822-1024 -- read window, NoData->NaN,
cp.asarray, run_tile_algorithm, crop, cp.asnumpy, quantize, write).  Here the integer encoding is
fused into the kernel epilogue, so the device->host copy moves 1 or 2 bytes per pixel instead of 4."""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .. import kernels as _k
from ..algorithms._norm_stats import inject_global_stats
from ..io.output_encoding import quantize_params, resolve_output_range

_TORCH_DT = {"float32": torch.float32, "int16": torch.int16, "uint8": torch.uint8}

# reference core/tile_processor.py:64-87 -- canonical name -> tile class (the five on the B200 path)
DEFAULT_ALGORITHMS = {
    "topousm_fast": "TopoUSMFastAlgorithm", "hillshade": "HillshadeAlgorithm", "slope": "SlopeAlgorithm",
    "curvature": "CurvatureAlgorithm", "openness": "OpennessAlgorithm",
    "ambient_occlusion": "AmbientOcclusionAlgorithm",
}
SPATIAL_TILE_ALGORITHMS = {"hillshade", "slope", "curvature", "openness", "ambient_occlusion"}


def _required_padding_for_algorithm(algorithm: str, algo_params: dict, sigma: float, pixel_size: float,
                                    target_distances=None, tile_size: int = 1024) -> int:
    """reference core/tile_processor.py:207-383 for the hot-path algorithms: halo (pixels, rounded up to 32) a tile
    needs so that its core equals the whole-raster result.  topousm_fast: max radius + 16; spatial mode of the
    Gaussian-smoothing algorithms: 2 R + 2 over the radii at or below the large-radius threshold
    max(256, tile_size // 16) (larger ones come from the overview); openness spatial: R + 16."""
    import math
    from .tile_compute import _normalize_topousm_fast_radii_and_weights
    try:
        base = int(math.ceil(max(float(sigma), 0.0) * 5.0))
    except Exception:
        base = 32
    required = max(32, base)
    mode = str(algo_params.get("mode", "spatial")).lower()
    overview_active = mode == "spatial" and bool(algo_params.get("radii")) and algorithm != "topousm_fast"
    thr_switch = max(256, int(tile_size) // 16)
    if algorithm == "topousm_fast":
        radii, _ = _normalize_topousm_fast_radii_and_weights(
            target_distances=target_distances, weights=algo_params.get("weights"), pixel_size=pixel_size,
            manual_radii=algo_params.get("radii"), manual_weights=algo_params.get("weights"))
        if radii:
            required = max(required, int(max(radii) + 16))
    if mode == "spatial" and algorithm in SPATIAL_TILE_ALGORITHMS:
        radii = algo_params.get("radii")
        if not radii:
            from ..algorithms.common.spatial_mode import auto_spatial_radii
            radii = auto_spatial_radii(None)
        try:
            radii_f = [float(r) for r in radii if float(r) > 0]
            if overview_active:
                small = [r for r in radii_f if int(round(r)) <= thr_switch] or [min(radii_f)]
                max_radius = int(round(max(small)))
            else:
                max_radius = max(int(round(r)) for r in radii_f)
        except Exception:
            max_radius = 32
        if algorithm in {"ambient_occlusion", "openness"}:
            required = max(required, int(max_radius + 16))
        else:
            required = max(required, int(max_radius * 2 + 2))
    return max(32, ((required + 31) // 32) * 32)


def _sanitize_spatial_radii_weights_for_tile(algorithm: str, radii, weights, tile_size: int):
    """reference core/tile_processor.py:102-172 -- integer radii for a tile run of a spatial-mode algorithm, duplicates
    merged (weights of duplicates summed).  Gaussian-smoothing algorithms: round, at least 1, non-positive entries
    dropped; ambient_occlusion / openness: at least 2, weights follow their radius, a warning text when the list
    changed.  -> (radii | None, weights | None, warning | None)"""
    from .tile_compute import deduplicate_radii_weights
    if not isinstance(radii, (list, tuple)) or len(radii) == 0:
        return None, weights, None

    def as_float(v):
        try:
            return float(v)
        except (TypeError, ValueError):
            return None

    if algorithm not in {"ambient_occlusion", "openness"}:
        kept = [max(1, int(round(f))) for f in (as_float(v) for v in radii) if f is not None and f > 0]
        if not kept:
            return None, weights, None
        rr, ww = deduplicate_radii_weights(kept, weights)
        return rr, ww, None
    kept_r, kept_w = [], []
    for i, v in enumerate(radii):
        f = as_float(v)
        if f is None or int(round(f)) <= 0:
            continue
        kept_r.append(max(2, int(round(f))))
        if isinstance(weights, (list, tuple)) and i < len(weights):
            w = as_float(weights[i])
            kept_w.append(0.0 if w is None else w)
    if not kept_r:
        return None, None, None
    rr, ww = deduplicate_radii_weights(kept_r, kept_w if len(kept_w) == len(kept_r) else None)
    warn = None
    try:
        if [int(round(float(v))) for v in radii] != rr:
            warn = f"Spatial radii de-duplicated for {algorithm}: {list(radii)} -> {rr}"
    except Exception:
        pass
    return rr, ww, warn


def _load_algorithm(name: str):
    """reference core/tile_processor.py:807-820 -- the tile backend's plug-in lookup: the class DEFAULT_ALGORITHMS
    names, found in algorithms/tile/<name>.py, instantiated."""
    from importlib import import_module
    cls_name = DEFAULT_ALGORITHMS.get(name)
    if cls_name is not None:
        try:
            cls = getattr(import_module(f"..algorithms.tile.{name}", package=__package__), cls_name, None)
            if cls is not None:
                return cls()
        except ImportError as exc:
            import logging
            logging.getLogger(__name__).warning("Failed to load algorithm %s: %s", name, exc)
    raise ValueError(f"Algorithm {name} not found or not available on this platform")


def _replace_nodata_with_nan(data: np.ndarray, nodata):
    """reference :199-204 -- NoData cells (numeric match within 1e-6, or NaN) become NaN; no NoData value: unchanged."""
    from .tile_compute import build_nodata_mask
    mask = build_nodata_mask(data, nodata)
    return data if mask is None else np.where(mask, np.nan, data)


def _format_algorithm_output(result_core: np.ndarray, algorithm: str):
    """reference core/tile_processor.py:606-624 -- host-side format of a finished tile: float32, NaN as the NoData
    of every float output, hillshade clipped to [0, 1] (NaN survives the clip) with an RGB(A)-last result reduced to
    its first band.  -> (array, nodata)"""
    out = np.asarray(result_core).astype(np.float32, copy=False)
    if algorithm == "hillshade":
        if out.ndim == 3 and out.shape[-1] in (3, 4) and out.shape[-2:] != out.shape[:2]:
            out = out[:, :, 0]
        out = np.clip(out, 0.0, 1.0)
    return out, float("nan")


class HostTilePipeline:
    """Re-usable device/host staging for repeated tiles of one shape (no allocation per call)."""

    def __init__(self, shape, algorithm: str, params: dict, output_dtype: str = "float32", device="cuda:0",
                 chunk_rows: int = 4096):
        self.shape = (int(shape[0]), int(shape[1]))
        self.algorithm = str(algorithm)
        self.params = dict(params)
        self.output_dtype = str(output_dtype)
        self.device = torch.device(device)
        self.chunk_rows = int(chunk_rows)
        self.dev_in = torch.empty(self.shape, dtype=torch.float32, device=self.device)
        self.dev_out = torch.empty(self.shape, dtype=_TORCH_DT[self.output_dtype], device=self.device)
        self.qp = None
        if self.output_dtype != "float32":
            rng = resolve_output_range(self.algorithm, params=self.params)
            if rng is None:
                raise ValueError(f"{self.algorithm}: integer output needs an explicit value range")
            self.qp = quantize_params(rng[0], rng[1], self.output_dtype)
        self.workspace = None
        if self.algorithm != "topousm_fast":
            # the other algorithms run their LOCAL block kernel here (one call, no multiscale combine, no display
            # stretch): say so instead of silently ignoring the request -- cli.run / Algorithm.process do the rest
            if str(self.params.get("mode", "local")).lower() == "spatial":
                raise ValueError(f"HostTilePipeline: {self.algorithm} runs in local mode only; use "
                                 "algorithms.ALGORITHMS[...].process (cli.run) for --mode spatial")
            for key in ("radii", "weights", "agg"):
                if self.params.get(key) not in (None, [], ()):
                    raise ValueError(f"HostTilePipeline: '{key}' is a spatial-mode parameter and is not applied here")
        if self.algorithm == "topousm_fast":
            need = _k.topousm_fast_workspace_bytes(self.shape, self.params["radii"], self.params.get("pixel_size", 1.0))
            self.workspace = torch.empty(max(need, 256), dtype=torch.uint8, device=self.device)
        self.h2d_bytes = self.shape[0] * self.shape[1] * 4
        self.d2h_bytes = self.shape[0] * self.shape[1] * self.dev_out.element_size()

    def run(self, host_in: torch.Tensor, host_out: torch.Tensor) -> torch.Tensor:
        """host_in: pinned f32 [H,W] (NaN = NoData); host_out: pinned buffer of the output dtype."""
        H = self.shape[0]
        for r in range(0, H, self.chunk_rows):
            self.dev_in[r:r + self.chunk_rows].copy_(host_in[r:r + self.chunk_rows], non_blocking=True)
        p = dict(self.params)
        algo = self.algorithm
        if algo == "topousm_fast":
            inject_global_stats(self.dev_in, algo, p)
            scale = float(p["global_stats"][0]) if p.get("global_stats") else 1.0
            _k.topousm_fast(self.dev_in, radii=p["radii"], weights=p.get("weights"), pixel_size=p.get("pixel_size", 1.0),
                            norm_scale=scale, output_dtype=self.output_dtype, qp=self.qp, workspace=self.workspace,
                            out=self.dev_out)
        else:
            fn = {"hillshade": _k.hillshade, "slope": _k.slope, "curvature": _k.curvature, "openness": _k.openness,
                  "ambient_occlusion": _k.ambient_occlusion}[algo]
            keys = {"hillshade": ("azimuth", "altitude", "z_factor"), "slope": ("unit",), "curvature": ("curvature_type",),
                    "openness": ("openness_type", "num_directions", "max_distance"),
                    "ambient_occlusion": ("num_samples", "radius", "intensity")}[algo]
            kw = {k: p[k] for k in keys + ("pixel_size", "pixel_scale_x", "pixel_scale_y") if k in p}
            if algo in ("openness", "ambient_occlusion"):   # the reference's [p1, p99] display stretch (global stats)
                inject_global_stats(self.dev_in, algo, p)
                if p.get("global_stats") is not None:
                    kw["stretch"] = p["global_stats"]
            self.dev_out = fn(self.dev_in, output_dtype=self.output_dtype, qp=self.qp, **kw)
        for r in range(0, H, self.chunk_rows):
            host_out[r:r + self.chunk_rows].copy_(self.dev_out[r:r + self.chunk_rows], non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return host_out


class StreamedTopoPipeline(HostTilePipeline):
    """topousm_fast from host memory to host memory with the three engines overlapped: the raster is
    uploaded in row chunks (the chunks the statistics windows need go first), every chunk is decimated as
    soon as it lands, the global scale is computed while the rest is still in flight, and row bands run
    through the band entry points (global coordinates -> bit-identical to the whole-raster call) as soon
    as the rows and decimated rows they read are resident; each finished band goes back to the host on
    a third stream while the next one computes.  (reference: the tile loop of
    core/tile_processor.py:822-1024, where read, compute and write are strictly sequential per tile.)"""

    def __init__(self, shape, params: dict, output_dtype: str = "uint8", device="cuda:0", chunk_rows: int = 4096):
        super().__init__(shape, "topousm_fast", params, output_dtype, device, chunk_rows)
        assert self.chunk_rows % 16 == 0, "chunks must hold whole decimation cells"
        self.s_in = torch.cuda.Stream(device=self.device)
        self.s_out = torch.cuda.Stream(device=self.device)
        H, W = self.shape
        plan = _k.topousm_plan(self.params["radii"], self.params.get("pixel_size", 1.0))
        self.kinds, self.factors, self.sizes, self.R = plan["kind"], plan["factor"], plan["size"], plan["fused_halo"]
        self.levels = sorted({self.factors[i] for i in range(len(self.kinds)) if self.kinds[i] == 1})
        self.full = {f: torch.empty(((H + f - 1) // f, (W + f - 1) // f), dtype=torch.float32, device=self.device)
                     for f in self.levels}
        self.flags = torch.zeros(16, dtype=torch.int32, device=self.device)

    def _band_need(self, r0: int, r1: int):
        """DEM rows [lo, hi) a band reads directly or through its decimated means."""
        from . import sharding as sh
        H = self.shape[0]
        n = len(self.kinds)
        halo = max([self.R] + [4 if self.sizes[i] == 0 else self.sizes[i] // 2 for i in range(n) if self.kinds[i] == 2])
        lo, hi = sh.mirror_need(r0 - halo, r1 - 1 + halo, H, True)
        for i in range(n):
            if self.kinds[i] != 1:
                continue
            f = self.factors[i]
            gh = (H + f - 1) // f
            rs = (gh - 1) / (H - 1) if H > 1 else 1.0
            mlo = int(np.floor(r0 * rs)); mhi = min(gh - 1, int(np.floor((r1 - 1) * rs)) + 1)
            reach = 4 if self.sizes[i] == 0 else self.sizes[i] // 2
            glo, ghi = sh.mirror_need(mlo - reach, mhi + reach, gh, True)
            lo, hi = min(lo, glo * f), max(hi, min(H, ghi * f))
        return lo, hi

    def _whole(self, host_out, radii, weights, px, scale):
        H, C = self.shape[0], self.chunk_rows
        cur = torch.cuda.current_stream(self.device)
        _k.topousm_fast(self.dev_in, radii=radii, weights=weights, pixel_size=px, norm_scale=scale,
                        output_dtype=self.output_dtype, qp=self.qp, workspace=self.workspace, out=self.dev_out)
        for r in range(0, H, C):
            host_out[r:r + C].copy_(self.dev_out[r:r + C], non_blocking=True)
        cur.synchronize()
        return host_out

    def _band(self, r0: int, r1: int, radii, weights, px, scale):
        from . import sharding as sh
        H = self.shape[0]
        n = len(self.kinds)
        halo = max([self.R] + [4 if self.sizes[i] == 0 else self.sizes[i] // 2 for i in range(n) if self.kinds[i] == 2])
        lo, hi = sh.mirror_need(r0 - halo, r1 - 1 + halo, H, True)
        dem_ext = self.dev_in[lo:hi]
        grids, grow0 = [None] * n, [0] * n
        for i in range(n):
            if self.kinds[i] == 1:
                f = self.factors[i]
                gh = (H + f - 1) // f
                rs = (gh - 1) / (H - 1) if H > 1 else 1.0
                mlo = int(np.floor(r0 * rs)); mhi = min(gh - 1, int(np.floor((r1 - 1) * rs)) + 1)
                reach = 4 if self.sizes[i] == 0 else self.sizes[i] // 2
                glo, ghi = sh.mirror_need(mlo - reach, mhi + reach, gh, True)
                grids[i] = _k.grid_mean_band(self.full[f][glo:ghi], glo, gh, self.sizes[i], mlo, mhi - mlo + 1)
                grow0[i] = mlo
            elif self.kinds[i] == 2:
                grids[i] = _k.grid_mean_band(dem_ext, lo, H, self.sizes[i], r0, r1 - r0)
                grow0[i] = r0
        _k.topousm_fused_band(dem_ext, lo, H, r0, r1 - r0, radii=radii, weights=weights, pixel_size=px,
                              term_grids=grids, term_grow0=grow0, norm_scale=scale, output_dtype=self.output_dtype,
                              qp=self.qp, out=self.dev_out[r0:r1])

    def run(self, host_in: torch.Tensor, host_out: torch.Tensor) -> torch.Tensor:
        from ..algorithms._norm_stats import (_norm_stat_window_geometry, compute_norm_stats_device, stratified_windows,
                                              valid_bbox_host)
        H, W = self.shape
        C = self.chunk_rows
        p = dict(self.params)
        radii, weights, px = p["radii"], p.get("weights"), p.get("pixel_size", 1.0)
        cur = torch.cuda.current_stream(self.device)
        nchunks = (H + C - 1) // C
        # valid-data bounding box from a host-side <= 512 px overview -> the windows of the statistics pre-pass;
        # their chunks are uploaded first
        scale = float(p["global_stats"][0]) if p.get("global_stats") else None
        bbox = valid_bbox_host(host_in) if scale is None else None
        first = []
        if scale is None and bbox is not None:
            _margin, tile = _norm_stat_window_geometry("topousm_fast", p)
            wins = stratified_windows(W, H, bbox[0], bbox[1], bbox[2], bbox[3], grid=3, tile=min(tile, max(W, H)))
            first = sorted({c for (wy0, _wx0, _tw, th) in wins
                            for c in range(wy0 // C, min(nchunks, (wy0 + th - 1) // C + 1))})
        order = first + [c for c in range(nchunks) if c not in set(first)]
        self.s_in.wait_stream(cur)
        ev = [None] * nchunks
        with torch.cuda.stream(self.s_in):
            for c in order:
                a, b = c * C, min(H, (c + 1) * C)
                self.dev_in[a:b].copy_(host_in[a:b], non_blocking=True)
                ev[c] = torch.cuda.Event()
                ev[c].record(self.s_in)
        bands = [(r0, min(H, r0 + C)) for r0 in range(0, H, C)]
        needs = []
        for (r0, r1) in bands:
            lo, hi = self._band_need(r0, r1)
            needs.append(set(range(lo // C, min(nchunks, (hi - 1) // C + 1))))
        pending = list(range(len(bands)))
        landed = set()
        self.flags.zero_()

        def launch_ready():
            for bi in list(pending):
                if needs[bi] <= landed:
                    r0, r1 = bands[bi]
                    self._band(r0, r1, radii, weights, px, scale)
                    e = torch.cuda.Event()
                    e.record(cur)
                    self.s_out.wait_event(e)
                    with torch.cuda.stream(self.s_out):
                        host_out[r0:r1].copy_(self.dev_out[r0:r1], non_blocking=True)
                    pending.remove(bi)

        for k_, c in enumerate(order):
            a, b = c * C, min(H, (c + 1) * C)
            cur.wait_event(ev[c])
            if self.levels:   # decimate the chunk as it lands (cells never straddle chunks: C % 16 == 0)
                grids, fl = _k.pyramid_band(self.dev_in[a:b], self.levels)
                for f, g in zip(self.levels, grids):
                    self.full[f][a // f:a // f + g.shape[0]].copy_(g)
                self.flags[: len(self.levels)] |= fl
            landed.add(c)
            if scale is None and k_ + 1 >= len(first):
                # every row the windows read is on the device: global scale now, while the uploads keep running
                st = compute_norm_stats_device(self.dev_in, "topousm_fast", p, bbox=bbox) if bbox is not None else None
                scale = float(st[0]) if st else 1.0
            if scale is not None:
                launch_ready()
        if scale is None:
            scale = 1.0
        launch_ready()
        assert not pending
        if self.levels and bool(self.flags[: len(self.levels)].any()):
            # a decimated cell without any valid pixel: the enclosed-void fill spans 1/16 of the raster side, so
            # this (rare, NoData-heavy) case is redone with the whole-raster call
            self.s_out.synchronize()
            return self._whole(host_out, radii, weights, px, scale)
        self.s_out.synchronize()
        cur.synchronize()
        return host_out


def process_host_raster(dem_host: np.ndarray, algorithm: str, params: dict, output_dtype: str = "float32",
                        nodata: Optional[float] = None, device="cuda:0") -> np.ndarray:
    """Convenience wrapper: NumPy in, NumPy out (pins temporary buffers)."""
    a = np.ascontiguousarray(dem_host, dtype=np.float32)
    if nodata is not None and not np.isnan(float(nodata)):
        a = np.where(np.isclose(a, float(nodata), rtol=0.0, atol=1e-6), np.float32(np.nan), a)
    pipe = HostTilePipeline(a.shape, algorithm, params, output_dtype, device)
    hin = torch.from_numpy(a).pin_memory()
    hout = torch.empty(a.shape, dtype=_TORCH_DT[str(output_dtype)]).pin_memory()
    return pipe.run(hin, hout).numpy()
