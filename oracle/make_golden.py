"""Generate tests/golden/*.npz from the UNMODIFIED reference  --  TEST INFRASTRUCTURE ONLY.

Runs in the build container only (needs /root/reference).  The reference's
CuPy-based block functions are imported through the NumPy import shim in
oracle/ref_shim (cupy -> numpy, cupyx.scipy.ndimage -> scipy.ndimage,
dask.array -> stub) and executed on small synthetic DEMs; inputs, parameters and
outputs are stored so that tests on any machine (including the GPU box, where
/root/reference does not exist) can hold both the oracle restatement and the
CUDA path to what the reference itself produced.

    python oracle/make_golden.py            # rewrites tests/golden/
"""
from __future__ import annotations

import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("FSG_REFERENCE_ROOT", "/root/reference")
sys.path[:0] = [os.path.join(HERE, "ref_shim"), REF, ROOT]

import numpy as np  # noqa: E402

from oracle.terrain_oracle import synth_dem  # noqa: E402


def _ref():
    from FujiShaderGPU.algorithms import _impl_topousm_fast as tf
    from FujiShaderGPU.algorithms import _impl_hillshade as hs
    from FujiShaderGPU.algorithms import _impl_slope as sl
    from FujiShaderGPU.algorithms import _impl_curvature as cv
    from FujiShaderGPU.algorithms import _impl_openness as op
    from FujiShaderGPU.algorithms import _nan_utils as nu
    from FujiShaderGPU.algorithms import _normalization as nm
    from FujiShaderGPU.algorithms import _global_stats as gs
    from FujiShaderGPU.algorithms.common import spatial_mode as sm
    from FujiShaderGPU.io import output_encoding as oe
    return tf, hs, sl, cv, op, nu, nm, gs, sm, oe


def ambient_occlusion_section(save) -> None:
    """compute_ambient_occlusion_block / _spatial_block of the reference, through the shim."""
    from FujiShaderGPU.algorithms import _impl_ambient_occlusion as ao
    adense = synth_dem(160, 208, seed=20261031)
    aholes = synth_dem(160, 208, seed=20261032, nodata=True)
    aholes[50:54, 80:91] = np.nan
    aholes[0, 100:103] = np.nan
    arrays = {"adense": adense, "aholes": aholes}
    params = {}
    src = {"adense": adense, "aholes": aholes}
    for cname, (key, kw) in {
        "s16_r10": ("adense", dict(num_samples=16, radius=10.0, intensity=1.0)),
        "s8_r3p5_i2": ("adense", dict(num_samples=8, radius=3.5, intensity=2.0)),
        "s16_r24_northup": ("adense", dict(num_samples=16, radius=24.0, intensity=1.0, pixel_scale_x=1.0, pixel_scale_y=-1.0)),
        "s16_r10_holes": ("aholes", dict(num_samples=16, radius=10.0, intensity=1.0)),
        "s12_r6_holes_geo": ("aholes", dict(num_samples=12, radius=6.0, intensity=0.7, pixel_scale_x=23.7,
                                            pixel_scale_y=-30.9, pixel_size=30.0)),
    }.items():
        arrays[f"local__{cname}"] = ao.compute_ambient_occlusion_block(src[key].copy(), **kw)
        params[f"local__{cname}"] = dict(input=key, kw=kw)
    for cname, (key, kw) in {
        "s16_r64": ("adense", dict(num_samples=16, radius=64.0, intensity=1.0)),
        "s8_r40_holes": ("aholes", dict(num_samples=8, radius=40.0, intensity=1.0)),
    }.items():
        arrays[f"spatial__{cname}"] = ao.compute_ambient_occlusion_spatial_block(src[key].copy(), **kw)
        params[f"spatial__{cname}"] = dict(input=key, kw=kw)
    save("ambient_occlusion", params, **arrays)


def main(out_dir: str, only: str = "") -> None:
    os.makedirs(out_dir, exist_ok=True)
    manifest = {}

    def save(name, params, **arrays):
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **arrays)
        manifest[name] = params

    if only == "ambient_occlusion":   # add one section to the committed fixtures without rewriting the others
        with open(os.path.join(out_dir, "manifest.json")) as fh:
            manifest.update(json.load(fh))
        ambient_occlusion_section(save)
        with open(os.path.join(out_dir, "manifest.json"), "w") as fh:
            json.dump(manifest, fh, indent=1, sort_keys=True, default=lambda o: o.tolist() if hasattr(o, "tolist") else str(o))
        return
    tf, hs, sl, cv, op, nu, nm, gs, sm, oe = _ref()
    ambient_occlusion_section(save)

    # ---------------- gradient family ----------------
    dense = synth_dem(128, 176, seed=20261017)
    holes = synth_dem(128, 176, seed=20261018, nodata=True)
    holes[40:43, 100:104] = np.nan          # small enclosed gap (gauss fill reach)
    holes[0, 150] = np.nan
    holes[127, 60:62] = np.nan                  # NaN on the raster edge
    grad_cases = {
        "dense_none": (dense, dict()),
        "dense_northup": (dense, dict(pixel_scale_x=1.0, pixel_scale_y=-1.0)),
        "dense_px2p5": (dense, dict(pixel_size=2.5)),
        "holes_northup": (holes, dict(pixel_scale_x=1.0, pixel_scale_y=-1.0)),
        "dense_geo": (dense, dict(pixel_scale_x=23.7, pixel_scale_y=-30.9, pixel_size=30.0)),
    }
    arrays = {"dense": dense, "holes": holes}
    params = {}
    for cname, (dem, kw) in grad_cases.items():
        key = "dense" if dem is dense else "holes"
        arrays[f"hillshade__{cname}"] = hs.compute_hillshade_block(dem.copy(), **kw)
        params[f"hillshade__{cname}"] = dict(input=key, kw=kw)
        for unit in ("degree", "percent", "radian"):
            arrays[f"slope_{unit}__{cname}"] = sl.compute_slope_block(dem.copy(), unit=unit, **kw)
            params[f"slope_{unit}__{cname}"] = dict(input=key, kw=dict(unit=unit, **kw))
        for ct in ("mean", "gaussian", "planform", "profile"):
            arrays[f"curvature_{ct}__{cname}"] = cv.compute_curvature_block(dem.copy(), curvature_type=ct, **kw)
            params[f"curvature_{ct}__{cname}"] = dict(input=key, kw=dict(curvature_type=ct, **kw))
    arrays["hillshade__dense_az45_alt30_z2"] = hs.compute_hillshade_block(dense.copy(), azimuth=45, altitude=30, z_factor=2.0)
    params["hillshade__dense_az45_alt30_z2"] = dict(input="dense", kw=dict(azimuth=45, altitude=30, z_factor=2.0))
    save("gradient_family", params, **arrays)

    # ---------------- gradient family, spatial mode (Gaussian scale space) ----------------
    from FujiShaderGPU.algorithms.tile import dask_bridge as db
    sdense = synth_dem(200, 240, seed=20261031)
    sholes = synth_dem(200, 240, seed=20261032, nodata=True)
    sholes[90:93, 100:140] = np.nan
    arrays = {"sdense": sdense, "sholes": sholes}
    params = {}
    nup = dict(pixel_scale_x=1.0, pixel_scale_y=-1.0)
    for key, dem in (("sdense", sdense), ("sholes", sholes)):
        for r in (2.0, 8.0, 32.0, 64.0, 300.0):
            nm_ = f"hillshade_r{int(r)}__{key}"
            arrays[nm_] = hs.compute_hillshade_spatial_block(dem.copy(), radius=r, **nup)
            params[nm_] = dict(input=key, algo="hillshade", kw=dict(radius=r, **nup))
        for r in (3.0, 40.0):
            nm_ = f"slope_r{int(r)}__{key}"
            arrays[nm_] = sl.compute_slope_spatial_block(dem.copy(), radius=r, unit="degree", **nup)
            params[nm_] = dict(input=key, algo="slope", kw=dict(radius=r, unit="degree", **nup))
        for r in (4.0, 30.0):
            nm_ = f"curvature_r{int(r)}__{key}"
            sm_ = nu._smooth_for_radius(dem.copy(), r, pixel_size=1.0, algorithm_name="curvature")
            arrays[nm_] = cv.compute_curvature_block(sm_, curvature_type="mean", **nup)
            params[nm_] = dict(input=key, algo="curvature", kw=dict(radius=r, curvature_type="mean", **nup))
        # the tile adapter's multi-radius hillshade (tile/dask_bridge.py:72-110) and its combiner (:28-69)
        radii = [2, 8, 32, 64]
        wts = sm.auto_spatial_weights(4)
        resp = [hs.compute_hillshade_spatial_block(dem.copy(), radius=float(r), **nup) for r in radii]
        arrays[f"hillshade_multi_weighted__{key}"] = db._combine_direct(resp, weights=wts, agg="mean")
        params[f"hillshade_multi_weighted__{key}"] = dict(input=key, algo="hillshade_multi",
                                                          kw=dict(radii=radii, weights=list(map(float, wts)), agg="mean", **nup))
        arrays[f"hillshade_multi_max__{key}"] = db._combine_direct(resp, weights=None, agg="max")
        params[f"hillshade_multi_max__{key}"] = dict(input=key, algo="hillshade_multi",
                                                     kw=dict(radii=radii, weights=None, agg="max", **nup))
        arrays[f"hillshade_multi_equal__{key}"] = db._combine_direct(resp, weights=None, agg="mean")
        params[f"hillshade_multi_equal__{key}"] = dict(input=key, algo="hillshade_multi",
                                                       kw=dict(radii=radii, weights=None, agg="mean", **nup))
    save("spatial_gradient", params, **arrays)

    # ---------------- topousm_fast ----------------
    tdense = synth_dem(288, 240, seed=20261019)
    tholes = synth_dem(288, 240, seed=20261020, nodata=True)
    tholes[200:203, 150:170] = np.nan
    tvoid = synth_dem(288, 240, seed=20261021)
    tvoid[96:176, 64:160] = np.nan          # fully-NaN coarse cells -> enclosed-void fill
    tvoid[:, :20] = np.nan                  # border-connected exterior stays NaN
    w6 = sm.auto_spatial_weights(6)
    topo_cases = {
        "dense_ladder6": ("tdense", dict(radii=[2, 8, 32, 128, 512, 2048], weights=w6)),
        "dense_default": ("tdense", dict()),
        "dense_local": ("tdense", dict(radii=[1], weights=[1.0])),
        "dense_equal_unsorted": ("tdense", dict(radii=[64, 3, 20, 41, 42, 100, 200])),
        "dense_px0p5": ("tdense", dict(radii=[4, 16, 30, 64], weights=[0.4, 0.3, 0.2, 0.1], pixel_size=0.5)),
        "dense_ds_gauss": ("tdense", dict(radii=[45, 2], weights=[0.5, 0.5], pixel_size=0.03)),
        "holes_ladder6": ("tholes", dict(radii=[2, 8, 32, 128, 512, 2048], weights=w6)),
        "holes_local": ("tholes", dict(radii=[1], weights=[1.0])),
        "void_ladder5": ("tvoid", dict(radii=[2, 8, 32, 128, 512], weights=sm.auto_spatial_weights(5))),
    }
    arrays = {"tdense": tdense, "tholes": tholes, "tvoid": tvoid}
    params = {}
    src = {"tdense": tdense, "tholes": tholes, "tvoid": tvoid}
    for cname, (key, kw) in topo_cases.items():
        raw = tf.compute_topousm_fast_efficient_block(src[key].copy(), **kw)
        arrays[f"raw__{cname}"] = raw
        st = nm.topousm_fast_stat_func(raw)
        if cname in ("dense_ladder6", "holes_ladder6", "dense_local"):
            arrays[f"norm__{cname}"] = gs.apply_global_normalization(raw.copy(), nm.topousm_fast_norm_func, st)
        params[cname] = dict(input=key, kw=kw, scale=float(st[0]))
    # helper-level vectors
    for f in (2, 4, 16):
        arrays[f"decimate{f}__tdense"] = nu._downsample_nan_aware(tdense, f)
        arrays[f"decimate{f}__tvoid"] = nu._downsample_nan_aware(tvoid, f)
        arrays[f"decimate{f}__tholes"] = nu._downsample_nan_aware(tholes, f)
    small = nu._downsample_nan_aware(tdense, 4)
    arrays["upsample4__tdense"] = nu._upsample_to_shape(small, tdense.shape)
    arrays["box17__tdense"] = nu.handle_nan_with_uniform(tdense, size=17, mode="reflect")[0]
    arrays["box17__tholes"] = nu.handle_nan_with_uniform(tholes, size=17, mode="reflect")[0]
    arrays["gauss1__tholes"] = nu.handle_nan_with_gaussian(tholes, sigma=1.0, mode="nearest")[0]
    # overview large-radius split (a10)
    coarse = nu._downsample_nan_aware(tdense, 4)
    field = tf.compute_topousm_fast_large_coarse_field(coarse, large_radii=[128, 512], large_weights=[0.2, 0.1], decimation=4.0)
    arrays["large_field"] = field
    arrays["large_part"] = tf._topousm_fast_add_large_block(
        tdense[100:260, 50:210].copy(), coarse_field=field, w_large=0.3, off_r=100, off_c=50,
        full_h=tdense.shape[0], full_w=tdense.shape[1])
    params["large_part"] = dict(input="tdense", window=[100, 260, 50, 210], w_large=0.3,
                                large_radii=[128, 512], large_weights=[0.2, 0.1], decimation=4.0)
    save("topousm_fast", params, **arrays)

    # ---------------- openness ----------------
    odense = synth_dem(176, 224, seed=20261022)
    oholes = synth_dem(176, 224, seed=20261023, nodata=True)
    oholes[60:64, 90:99] = np.nan
    arrays = {"odense": odense, "oholes": oholes}
    params = {}
    open_cases = {
        "pos8_r64": ("odense", dict(openness_type="positive", num_directions=8, max_distance=64)),
        "neg8_r64": ("odense", dict(openness_type="negative", num_directions=8, max_distance=64)),
        "pos16_r50_northup": ("odense", dict(openness_type="positive", num_directions=16, max_distance=50,
                                             pixel_scale_x=1.0, pixel_scale_y=-1.0)),
        "pos8_r5": ("odense", dict(openness_type="positive", num_directions=8, max_distance=5)),
        "pos8_r64_holes": ("oholes", dict(openness_type="positive", num_directions=8, max_distance=64)),
        "neg16_r20_holes_geo": ("oholes", dict(openness_type="negative", num_directions=16, max_distance=20,
                                               pixel_scale_x=23.7, pixel_scale_y=-30.9, pixel_size=30.0)),
    }
    src = {"odense": odense, "oholes": oholes}
    for cname, (key, kw) in open_cases.items():
        arrays[f"local__{cname}"] = op.compute_openness_vectorized(src[key].copy(), **kw)
        params[f"local__{cname}"] = dict(input=key, kw=kw)
    for cname, (key, kw) in {
        "pos8_r256": ("odense", dict(openness_type="positive", num_directions=8, max_distance=256)),
        "pos8_r100_holes": ("oholes", dict(openness_type="positive", num_directions=8, max_distance=100)),
    }.items():
        arrays[f"spatial__{cname}"] = op.compute_openness_spatial_block(src[key].copy(), **kw)
        params[f"spatial__{cname}"] = dict(input=key, kw=kw)
    loc = arrays["local__pos8_r64"]
    st = gs.robust_unsigned_stretch_stat_func(loc)
    params["stretch"] = dict(of="local__pos8_r64", stats=[float(st[0]), float(st[1])])
    from FujiShaderGPU.algorithms.tile.dask_bridge import _apply_display_stretch_block
    arrays["stretch__pos8_r64"] = _apply_display_stretch_block(loc.copy(), st)
    save("openness", params, **arrays)

    # ---------------- host-side tables ----------------
    tables = {
        "auto_radii": {str(s): sm.auto_spatial_radii(s) for s in (30000, 20480, 10000, 5120, 5119, 300, 15)},
        "auto_radii_none": sm.auto_spatial_radii(None),
        "auto_weights": {str(n): sm.auto_spatial_weights(n) for n in (1, 2, 3, 4, 5, 6, 8)},
        "ds_topousm_px1": {str(r): nu._radius_to_downsample_factor(float(r), pixel_size=1.0, algorithm_name="topousm_fast")
                           for r in (1, 2, 20, 21, 41, 42, 83, 84, 166, 167, 333, 334, 2048, 10000)},
        "ds_openness_px1": {str(r): nu._radius_to_downsample_factor(float(r), pixel_size=1.0, algorithm_name="openness")
                            for r in (5, 34, 35, 68, 69, 137, 138, 256, 274, 275)},
        "ds_topousm_px0p5": {str(r): nu._radius_to_downsample_factor(float(r), pixel_size=0.5, algorithm_name="topousm_fast")
                             for r in (4, 16, 30, 33, 64, 128)},
        "quant": {f"{algo}:{dt}": oe.quantize_params(*oe.resolve_output_range(algo), dt)
                  for algo in ("topousm_fast", "hillshade", "slope", "curvature", "openness")
                  for dt in ("int16", "uint8")},
        "quant_vector": {
            "in": ["inf", "-inf", "nan", "1.0"],
            "uint8_hillshade": oe.quantize_array(np.array([np.inf, -np.inf, np.nan, 1.0], np.float32),
                                                 oe.quantize_params(0.0, 1.0, "uint8"), "uint8").tolist(),
        },
    }
    # tile-adapter level (what core/tile_compute.run_tile_algorithm calls)
    from FujiShaderGPU.algorithms.tile.topousm_fast import TopoUSMFastAlgorithm as TileTopo
    from FujiShaderGPU.core.tile_compute import _normalize_topousm_fast_radii_and_weights
    from FujiShaderGPU.algorithms._norm_stats import stratified_windows, _norm_stat_window_geometry
    tables["tile_radii"] = {
        "dup": list(_normalize_topousm_fast_radii_and_weights(None, None, 1.0, manual_radii=[2, 8.4, 8, 0.2, 32],
                                                             manual_weights=[1, 2, 3, 4, 5])),
        "noweights": list(_normalize_topousm_fast_radii_and_weights(None, None, 1.0, manual_radii=[2, 8, 32])),
    }
    tables["stats_windows"] = {
        "32768": stratified_windows(32768, 32768, 0, 32768, 0, 32768, grid=3, tile=8256),
        "small": stratified_windows(3000, 2000, 100, 1900, 50, 2950, grid=3, tile=2048),
    }
    tables["stats_geometry"] = {
        "ladder6": list(_norm_stat_window_geometry("topousm_fast", {"radii": [2, 8, 32, 128, 512, 2048]})),
        "open256": list(_norm_stat_window_geometry("openness", {"max_distance": 256, "radii": None})),
    }
    adapter = TileTopo()
    tdense = np.load(os.path.join(out_dir, "topousm_fast.npz"))["tdense"]
    out = adapter.process(tdense.copy(), radii=[2, 8, 32, 128], weights=None, pixel_size=1.0, global_stats=(7.25,))
    np.savez_compressed(os.path.join(out_dir, "tile_adapter.npz"), topousm_fast_gs7p25=out)
    manifest["tile_adapter"] = {"topousm_fast_gs7p25": dict(input="topousm_fast.npz:tdense",
                                kw=dict(radii=[2, 8, 32, 128], weights=None, pixel_size=1.0, global_stats=[7.25]))}
    manifest["tables"] = tables
    manifest["_generator"] = {"reference": "geoign/FujiShaderGPU v1.0.1", "numpy": np.__version__,
                              "scipy": __import__("scipy").__version__}
    with open(os.path.join(out_dir, "manifest.json"), "w") as fh:
        json.dump(manifest, fh, indent=1, sort_keys=True, default=lambda o: o.tolist() if hasattr(o, "tolist") else str(o))
    total = sum(os.path.getsize(os.path.join(out_dir, f)) for f in os.listdir(out_dir))
    print(f"wrote {len(os.listdir(out_dir))} files, {total/1e6:.2f} MB -> {out_dir}")


if __name__ == "__main__":
    main(os.path.join(ROOT, "tests", "golden"), only=(sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else ""))
