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
            fn = {"hillshade": _k.hillshade, "slope": _k.slope, "curvature": _k.curvature, "openness": _k.openness}[algo]
            keys = {"hillshade": ("azimuth", "altitude", "z_factor"), "slope": ("unit",), "curvature": ("curvature_type",),
                    "openness": ("openness_type", "num_directions", "max_distance")}[algo]
            kw = {k: p[k] for k in keys + ("pixel_size", "pixel_scale_x", "pixel_scale_y") if k in p}
            self.dev_out = fn(self.dev_in, output_dtype=self.output_dtype, qp=self.qp, **kw)
        for r in range(0, H, self.chunk_rows):
            host_out[r:r + self.chunk_rows].copy_(self.dev_out[r:r + self.chunk_rows], non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
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
