"""Pin the CPU oracle (oracle/terrain_oracle.py) to outputs of the UNMODIFIED
reference (tests/golden/, produced by oracle/make_golden.py through the NumPy
import shim) and to the reference's own known-answer tests."""
import numpy as np
import pytest

from oracle import terrain_oracle as orc
from conftest import assert_close_f32


def _kw(d):
    return {k: v for k, v in d.items()}


# ---- gradient family -------------------------------------------------------
def _spatial_oracle(dem, meta):
    kw = dict(meta["kw"])
    algo = meta["algo"]
    if algo == "hillshade":
        return orc.hillshade_spatial_block(dem, **kw)
    if algo == "slope":
        return orc.slope_spatial_block(dem, **kw)
    if algo == "curvature":
        return orc.curvature_spatial_block(dem, **kw)
    radii, weights, agg = kw.pop("radii"), kw.pop("weights"), kw.pop("agg")
    resp = [orc.hillshade_spatial_block(dem, radius=float(r), **kw) for r in radii]
    return orc.combine_responses(resp, weights=weights, agg=agg)


def test_spatial_gradient_matches_reference(golden, manifest):
    """Gaussian scale-space ('--mode spatial') block functions and the tile combiner vs the reference."""
    g = golden("spatial_gradient")
    for name, meta in manifest["spatial_gradient"].items():
        got = _spatial_oracle(g[meta["input"]], meta)
        if meta["algo"].startswith("hillshade"):
            assert_close_f32(got, g[name], rtol=0, atol=4e-7, what=name)   # f64-vs-f32 light vector of the shim
        else:
            assert np.array_equal(got, g[name], equal_nan=True), name


def test_gradient_family_matches_reference(golden, manifest):
    g = golden("gradient_family")
    for name, meta in manifest["gradient_family"].items():
        dem = g[meta["input"]]
        kw = _kw(meta["kw"])
        algo = name.split("__")[0]
        if algo == "hillshade":
            got = orc.hillshade_block(dem, **kw)
            # The shim evaluates the light vector in f64 (NumPy-2 promotion), CuPy
            # (and the oracle) in f32: documented <= 2e-7 difference.
            assert_close_f32(got, g[name], rtol=0, atol=4e-7, what=name)
        elif algo.startswith("slope"):
            got = orc.slope_block(dem, **kw)
            assert np.array_equal(got, g[name], equal_nan=True), name
        else:
            got = orc.curvature_block(dem, **kw)
            assert np.array_equal(got, g[name], equal_nan=True), name


def test_explicit_gradient_statement_is_bit_exact():
    rng = np.random.default_rng(5)
    f = (rng.standard_normal((37, 53)) * 300).astype(np.float32)
    for sy, sx in ((1.0, 1.0), (30.9, 23.7), (-1.0, 2.5)):
        wy, wx = np.gradient(f, sy, sx, edge_order=2)
        gy, gx = orc.gradient_f32_exact(f, sy, sx)
        assert np.array_equal(wy, gy) and np.array_equal(wx, gx)


# ---- topousm_fast ----------------------------------------------------------
def test_topousm_fast_matches_reference(golden, manifest):
    g = golden("topousm_fast")
    for cname, meta in manifest["topousm_fast"].items():
        if cname == "large_part":
            continue
        dem = g[meta["input"]]
        raw = orc.topousm_fast_block(dem, **_kw(meta["kw"]))
        assert np.array_equal(raw, g[f"raw__{cname}"], equal_nan=True), cname
        st = orc.abs_p99_scale(raw)
        assert st[0] == pytest.approx(meta["scale"], rel=0, abs=0)
        if f"norm__{cname}" in g:
            assert np.array_equal(orc.normalise_by_scale(raw.copy(), st), g[f"norm__{cname}"], equal_nan=True)


def test_topousm_helpers_match_reference(golden):
    g = golden("topousm_fast")
    for key in ("tdense", "tholes", "tvoid"):
        for f in (2, 4, 16):
            assert np.array_equal(orc.decimate_valid_mean(g[key], f), g[f"decimate{f}__{key}"], equal_nan=True)
    small = orc.decimate_valid_mean(g["tdense"], 4)
    assert np.array_equal(orc.upsample_align_corners(small, g["tdense"].shape), g["upsample4__tdense"])
    assert np.array_equal(orc.box_mean(g["tdense"], 17), g["box17__tdense"])
    assert np.array_equal(orc.box_mean(g["tholes"], 17), g["box17__tholes"])
    assert np.array_equal(orc.gauss_mean(g["tholes"], 1.0), g["gauss1__tholes"])


def test_explicit_box_and_bilinear_statements(golden):
    """The closed-form arithmetic specs the CUDA kernels implement agree with
    scipy.ndimage bit for bit (box) / up to f64 summation order (bilinear)."""
    g = golden("topousm_fast")
    dem = g["tdense"]
    for size in (5, 17, 65, 257, 601):      # 601 > array: re-reflection
        assert np.array_equal(orc.box_mean_exact(dem, size), orc.box_mean(dem, size)), size
    small = orc.decimate_valid_mean(dem, 4)
    a = orc.bilinear_align_corners_exact(small, dem.shape)
    b = orc.upsample_align_corners(small, dem.shape)
    assert (a != b).mean() < 1e-4
    assert np.max(np.abs(a - b)) <= np.spacing(np.float32(np.abs(b).max()))


def test_overview_large_part_matches_reference(golden, manifest):
    g = golden("topousm_fast")
    m = manifest["topousm_fast"]["large_part"]
    dem = g["tdense"]
    coarse = orc.decimate_valid_mean(dem, 4)
    field = orc.topousm_large_field(coarse, large_radii=m["large_radii"], large_weights=m["large_weights"],
                                    decimation=m["decimation"])
    assert np.array_equal(field, g["large_field"])
    r0, r1, c0, c1 = m["window"]
    part = orc.topousm_large_part(dem[r0:r1, c0:c1], field, m["w_large"], r0, c0, dem.shape[0], dem.shape[1])
    assert np.array_equal(part, g["large_part"])


# ---- openness --------------------------------------------------------------
def test_openness_matches_reference(golden, manifest):
    g = golden("openness")
    for name, meta in manifest["openness"].items():
        if name == "stretch":
            continue
        dem = g[meta["input"]]
        fn = orc.openness_block if name.startswith("local__") else orc.openness_spatial_block
        assert np.array_equal(fn(dem, **_kw(meta["kw"])), g[name], equal_nan=True), name
    st = manifest["openness"]["stretch"]
    loc = g[st["of"]]
    got_stats = orc.p1_p99_stretch_stats(loc)
    assert list(got_stats) == st["stats"]
    assert np.array_equal(orc.display_stretch(loc, tuple(st["stats"])), g["stretch__pos8_r64"], equal_nan=True)


# ---- ambient occlusion -----------------------------------
def test_ambient_occlusion_matches_reference(golden, manifest):
    g = golden("ambient_occlusion")
    for name, meta in manifest["ambient_occlusion"].items():
        dem = g[meta["input"]]
        fn = orc.ambient_occlusion_block if name.startswith("local__") else orc.ambient_occlusion_spatial_block
        assert np.array_equal(fn(dem, **_kw(meta["kw"])), g[name], equal_nan=True), name


def _radial(n=101):
    yy, xx = np.mgrid[0:n, 0:n]
    return np.sqrt((xx - n // 2) ** 2 + (yy - n // 2) ** 2)


def test_reference_known_answers_openness():
    """tests/test_openness_yokoyama.py:21-47 of the reference, run on the oracle."""
    r = _radial()
    kw = dict(num_directions=16, max_distance=40)
    peak, pit = (50 - r).astype(np.float32), (r - 50).astype(np.float32)
    assert orc.openness_block(peak, openness_type="positive", **kw)[50, 50] > orc.openness_block(pit, openness_type="positive", **kw)[50, 50]
    assert orc.openness_block(pit, openness_type="negative", **kw)[50, 50] > orc.openness_block(peak, openness_type="negative", **kw)[50, 50]
    flat = np.zeros((101, 101), np.float32)
    for t in ("positive", "negative"):
        assert orc.openness_block(flat, openness_type=t, **kw)[50, 50] == pytest.approx(1.0, abs=1e-3)


def test_reference_known_answers_curvature():
    """tests/test_curvature_analytic.py:26-55 of the reference, run on the oracle."""
    zero = 0.5 ** (1.0 / 2.2)
    x = np.arange(64, dtype=np.float32)[None, :].repeat(64, axis=0)
    y = np.arange(64, dtype=np.float32)[:, None].repeat(64, axis=1)
    cyl = (0.01 * (x - 32.0) ** 2).astype(np.float32)
    assert np.abs(orc.curvature_block(cyl, curvature_type="planform")[20:44, 20:44] - zero).max() < 0.02
    assert np.abs(orc.curvature_block(cyl, curvature_type="profile")[20:44, 20:44] - zero).max() > 0.05
    dome = (-0.01 * ((x - 32.0) ** 2 + (y - 32.0) ** 2)).astype(np.float32)
    for ct in ("planform", "profile"):
        assert np.abs(orc.curvature_block(dome, curvature_type=ct)[20:44, 20:44] - zero).max() > 0.05


def test_reference_known_answers_normalisation():
    """tests/test_topousm_fast_normalization.py:6-44 of the reference."""
    rng = np.random.default_rng(0)
    base = rng.normal(0.0, 0.01, size=(2048,)).astype(np.float32)
    base[:3] = [10.0, -8.0, 12.0]
    assert 0.0 < orc.abs_p99_scale(base)[0] < 0.2
    data = np.concatenate([np.full(800, 0.01, np.float32), np.full(200, 0.5, np.float32)])
    assert orc.abs_p99_scale(data)[0] == pytest.approx(0.5, rel=1e-2)
    v = np.asarray([-4.0, -1.0, 0.0, 1.0, 4.0], dtype=np.float32)
    assert orc.normalise_by_scale(v, (2.0,)).tolist() == pytest.approx([-2.0, -0.5, 0.0, 0.5, 2.0])


# ---- host-side tables ------------------------------------------------------
def test_scale_construction_tables(manifest):
    t = manifest["tables"]
    for s, want in t["auto_radii"].items():
        assert orc.ladder_radii(int(s)) == want
    assert orc.ladder_radii(None) == t["auto_radii_none"]
    for n, want in t["auto_weights"].items():
        assert orc.pow2_weights(int(n)) == want
    # tests/test_spatial_auto_radii.py:31-44 of the reference
    assert orc.ladder_radii(5120) == [2, 8, 32, 128, 512] and orc.ladder_radii(5119) == [2, 8, 32, 128]
    assert orc.ladder_radii(15) == [2]
    r, w = orc.scale_profile(10000, radii=[4, 16, 64, 256])
    assert r == [4, 16, 64, 256] and w == pytest.approx([8 / 15, 4 / 15, 2 / 15, 1 / 15])


def test_decimation_factor_tables(manifest):
    t = manifest["tables"]
    for r, want in t["ds_topousm_px1"].items():
        assert orc.decimation_factor(float(r), 1.0, "topousm_fast") == want
    for r, want in t["ds_openness_px1"].items():
        assert orc.decimation_factor(float(r), 1.0, "openness") == want
    for r, want in t["ds_topousm_px0p5"].items():
        assert orc.decimation_factor(float(r), 0.5, "topousm_fast") == want


def test_quantiser_tables_and_vector(manifest):
    t = manifest["tables"]
    for key, want in t["quant"].items():
        algo, dt = key.split(":")
        got = orc.encode_params(*orc.value_range(algo), dt)
        assert got == want, key
    # tests/test_audit_p1_regressions.py:45-62 of the reference
    v = np.array([np.inf, -np.inf, np.nan, 1.0], np.float32)
    assert orc.encode_array(v, orc.encode_params(0.0, 1.0, "uint8"), "uint8").tolist() == [0, 0, 0, 255] == t["quant_vector"]["uint8_hillshade"]
    assert orc.encode_array(v, orc.encode_params(0.0, 100.0, "int16"), "int16").tolist()[:3] == [0, 0, 0]


def test_tile_radii_and_stats_geometry(manifest):
    t = manifest["tables"]
    got = orc.tile_radii_weights([2, 8.4, 8, 0.2, 32], [1, 2, 3, 4, 5])
    assert [got[0], got[1]] == t["tile_radii"]["dup"]
    got = orc.tile_radii_weights([2, 8, 32], None)
    assert [got[0], got[1]] == t["tile_radii"]["noweights"]
    assert [list(x) for x in orc.stats_windows(32768, 32768, 0, 32768, 0, 32768, grid=3, tile=8256)] == t["stats_windows"]["32768"]
    assert [list(x) for x in orc.stats_windows(3000, 2000, 100, 1900, 50, 2950, grid=3, tile=2048)] == t["stats_windows"]["small"]
    assert list(orc.stats_window_geometry("topousm_fast", {"radii": [2, 8, 32, 128, 512, 2048]})) == t["stats_geometry"]["ladder6"]
    assert list(orc.stats_window_geometry("openness", {"max_distance": 256, "radii": None})) == t["stats_geometry"]["open256"]
