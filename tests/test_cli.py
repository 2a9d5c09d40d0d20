"""The command line over the B200 path (fujishadergpu_b200/cli.py): GeoTIFF in -> algorithm -> COG out, held to the
oracle's composition of the same steps (reference flow: core/dask_processor.py:1146-1490)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import terrain_oracle as orc  # noqa: E402


def test_parser_matches_reference_flags():
    from fujishadergpu_b200.cli import build_parser, resolve_params
    p = build_parser()
    a = p.parse_args(["in.tif", "out.tif", "--algorithm", "topousm_fast", "--mode", "spatial", "--radii", "2,8,32",
                      "--weights", "0.5,0.3,0.2", "--output-dtype", "uint8"])
    prm = resolve_params(a, (4000, 3000), 1.0, 1.0, -1.0)
    assert prm["radii"] == [2, 8, 32] and prm["weights"] == [0.5, 0.3, 0.2] and prm["mode"] == "radius"
    a = p.parse_args(["in.tif", "out.tif", "--algo", "topousm_fast", "--mode", "local", "--radii", "4,16"])
    prm = resolve_params(a, (4000, 3000), 1.0, 1.0, -1.0)
    assert prm["radii"] == [1] and prm["weights"] == [1.0]          # --mode local ignores explicit radii
    a = p.parse_args(["in.tif", "out.tif", "--algorithm", "topousm_fast"])
    prm = resolve_params(a, (30000, 40000), 1.0, 1.0, -1.0)
    assert prm["radii"] == orc.ladder_radii(30000) and prm["weights"] == orc.pow2_weights(len(prm["radii"]))
    a = p.parse_args(["in.tif", "out.tif", "--algorithm", "hillshade", "--mode", "local", "--azimuth", "300"])
    prm = resolve_params(a, (500, 500), 2.0, 2.0, -2.0)
    assert prm["radii"] == [1] and prm["azimuth"] == 300.0 and prm["pixel_scale_y"] == -2.0
    a = p.parse_args(["in.tif", "out.tif", "--algorithm", "openness", "--max-distance", "64", "--num-directions", "8"])
    prm = resolve_params(a, (5000, 5000), 1.0, 1.0, -1.0)
    assert prm["max_distance"] == 64 and prm["radii"] == orc.ladder_radii(5000)
    # cli/args.py build_algo_params: universal controls and the per-algorithm knobs
    a = p.parse_args(["in.tif", "out.tif", "--algorithm", "ambient_occlusion", "--radius", "12", "--num-samples", "8",
                      "--intensity", "1.5", "--mode", "local"])
    prm = resolve_params(a, (500, 500), 1.0, 1.0, -1.0)
    assert (prm["radius"], prm["num_samples"], prm["intensity"], prm["mode"], prm["agg"]) == (12, 8, 1.5, "local", "mean")
    from fujishadergpu_b200.cli import _parse_nodata, _parse_output_range
    assert _parse_nodata(None) is None and _parse_nodata("-9999") == -9999.0 and _parse_nodata("+NaN") != _parse_nodata("+NaN")
    assert _parse_output_range("0,90") == (0.0, 90.0) and _parse_output_range(None) is None
    for bad in ("90,0", "1", "a,b"):
        with pytest.raises(ValueError):
            _parse_output_range(bad)
    with pytest.raises(ValueError):
        _parse_nodata("none")
    from fujishadergpu_b200.algorithms.dask_registry import ALGORITHMS
    from fujishadergpu_b200.core.tile_processor import DEFAULT_ALGORITHMS
    choices = next(act.choices for act in p._actions if "--algorithm" in act.option_strings)
    assert sorted(choices) == sorted(ALGORITHMS) == sorted(DEFAULT_ALGORITHMS)      # tests/test_registry_cli_sync.py


def test_pixel_scales_from_metadata_match_reference_known_answers():
    """io/raster_info.py of the reference (values produced by its meters_per_degree /
    metric_pixel_scales_from_metadata in the build container) and the --pixel-size override rule."""
    from fujishadergpu_b200.cli import resolve_pixel_scales
    from fujishadergpu_b200.io.raster_info import meters_per_degree, metric_pixel_scales
    known = {0.0: (111319.458, 110574.2727), 35.0: (91288.13767578507, 110940.55217300118),
             60.0: (55799.979000000014, 111412.24020000001), -45.5: (78158.03450200213, 111141.51580157726),
             89.9: (194.94258232671098, 111693.91386062495)}
    for lat, want in known.items():
        assert meters_per_degree(lat) == want
    geo = (138.0, 1 / 3600, 0.0, 36.0, 0.0, -1 / 3600)
    assert metric_pixel_scales(geo, 4326, (3600, 3600)) == (25.202599870885926, -30.8193712366885, 28.010985553787215, True, 35.5)
    assert metric_pixel_scales((0.0, 2.0, 0.0, 10.0, 0.0, -2.0), 6677, (5, 5)) == (2.0, -2.0, 2.0, False, None)
    with pytest.raises(ValueError):
        metric_pixel_scales(geo, None, (3600, 3600))          # degree-sized pixels without a CRS
    assert resolve_pixel_scales(None, {"transform": geo, "epsg": 4326}, (3600, 3600)) == (
        28.010985553787215, 25.202599870885926, -30.8193712366885, True)
    assert resolve_pixel_scales(10.0, {"transform": geo, "epsg": 4326}, (3600, 3600)) == (10.0, 10.0, -10.0, False)
    assert resolve_pixel_scales(None, {}, (10, 10)) == (1.0, 1.0, -1.0, False)


def _write_input(path, dem, nodata=-9999.0, px=2.0):
    from fujishadergpu_b200.io.cog_writer import write_tiff_pyramid
    a = np.where(np.isnan(dem), np.float32(nodata), dem).astype(np.float32)
    write_tiff_pyramid(path, [a], nodata=nodata, transform=(500000.0, px, 0.0, 4100000.0, 0.0, -px), epsg=6677,
                       compress="deflate", bigtiff=False)


@pytest.mark.gpu
def test_cli_end_to_end_against_oracle(tmp_path):
    pytest.importorskip("torch")
    from fujishadergpu_b200.cli import main
    from fujishadergpu_b200.io.geotiff_reader import read_geotiff
    dem = orc.synth_dem(900, 700, seed=61, nodata=True)
    src = str(tmp_path / "dem.tif")
    _write_input(src, dem)
    # hillshade, local, uint8
    out = str(tmp_path / "hs.tif")
    assert main([src, out, "--algorithm", "hillshade", "--mode", "local", "--output-dtype", "uint8"]) == 0
    got, meta = read_geotiff(out)
    # (Algorithm.process on the whole raster = map_overlap(depth=1, boundary='reflect') with one block)
    want = orc.encode_array(orc.with_overlap(orc.hillshade_block, dem, 1, pixel_size=2.0, pixel_scale_x=2.0, pixel_scale_y=-2.0),
                            orc.encode_params(0.0, 1.0, "uint8"), "uint8")
    assert np.array_equal(got == 0, want == 0) and np.abs(got.astype(int) - want.astype(int)).max() <= 1
    assert meta["epsg"] == 6677 and meta["transform"] == (500000.0, 2.0, 0.0, 4100000.0, 0.0, -2.0) and meta["nodata"] == 0.0
    # topousm_fast, explicit radii, float32: raw block / p99 of the stratified windows (here: the whole raster, trimmed)
    out = str(tmp_path / "topo.tif")
    radii, w = [2, 8, 32], [0.5, 0.3, 0.2]
    assert main([src, out, "--algorithm", "topousm_fast", "--radii", "2,8,32", "--weights", "0.5,0.3,0.2"]) == 0
    got, meta = read_geotiff(out)
    raw = orc.topousm_fast_block(dem, radii=radii, weights=w, pixel_size=2.0)
    margin, tile = orc.stats_window_geometry("topousm_fast", {"radii": radii})
    ok = np.isfinite(dem)
    cov = max(1, max(dem.shape) // 512)
    ov = ok[::cov, ::cov][: max(1, dem.shape[0] // cov), : max(1, dem.shape[1] // cov)]
    rows, cols = np.nonzero(ov.any(axis=1))[0], np.nonzero(ov.any(axis=0))[0]
    by0, by1 = int(rows.min()) * cov, min(dem.shape[0], (int(rows.max()) + 1) * cov)
    bx0, bx1 = int(cols.min()) * cov, min(dem.shape[1], (int(cols.max()) + 1) * cov)
    pooled = []
    for (wy0, wx0, tw, th) in orc.stats_windows(dem.shape[1], dem.shape[0], by0, by1, bx0, bx1, grid=3,
                                                tile=min(tile, max(dem.shape))):
        win = dem[wy0:wy0 + th, wx0:wx0 + tw]
        if np.isfinite(win).sum() < 0.02 * win.size:
            continue
        r = orc.topousm_fast_block(win, radii=radii, weights=w, pixel_size=2.0)
        m = int(min(margin, r.shape[0] // 3, r.shape[1] // 3))
        if m > 0:
            r = r[m:-m, m:-m]
        pooled.append(r[~np.isnan(r)])
    scale = orc.abs_p99_scale(np.concatenate(pooled))[0]
    want = orc.normalise_by_scale(raw.copy(), (scale,))
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert np.allclose(got[~np.isnan(want)], want[~np.isnan(want)], rtol=1e-5, atol=1e-6)
    assert meta["nodata"] != meta["nodata"] and meta["predictor"] == 3
    # ambient occlusion, local, int16
    out = str(tmp_path / "ao.tif")
    assert main([src, out, "--algorithm", "ambient_occlusion", "--mode", "local", "--radius", "8", "--output-dtype", "int16"]) == 0
    got, meta = read_geotiff(out)
    assert got.dtype == np.int16 and meta["predictor"] == 2 and np.array_equal(got == 0, np.isnan(dem))


@pytest.mark.gpu
def test_cli_spatial_integer_output_keeps_nodata(tmp_path):
    """--mode spatial smooths with a void-filling Gaussian: the final NoData re-mask must run for integer outputs
    too (core/dask_processor.py:1479-1547), so DN 0 <=> input NoData."""
    pytest.importorskip("torch")
    from fujishadergpu_b200.cli import main
    from fujishadergpu_b200.io.geotiff_reader import read_geotiff
    dem = orc.synth_dem(700, 900, seed=62, nodata=True)
    dem[300:340, 400:470] = np.nan
    src = str(tmp_path / "dem.tif")
    _write_input(src, dem)
    for algo, dtype in (("hillshade", "uint8"), ("slope", "int16"), ("curvature", "uint8")):
        out = str(tmp_path / f"{algo}.tif")
        assert main([src, out, "--algorithm", algo, "--mode", "spatial", "--radii", "2,8", "--output-dtype", dtype]) == 0
        got, meta = read_geotiff(out)
        assert meta["nodata"] == 0.0
        assert np.array_equal(got == 0, np.isnan(dem)), algo
