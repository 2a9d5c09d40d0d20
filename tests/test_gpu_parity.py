"""Parity of the CUDA path (through the C ABI / reference-shaped Python API) against
(1) outputs of the unmodified reference (tests/golden) and (2) the CPU oracle on seeded inputs.

Bar: float32 within 1e-5 relative / 1e-6 absolute (on the normalised scale for topousm_fast),
identical NoData masks; uint8/int16 within +-1 DN."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from conftest import assert_close_f32  # noqa: E402
from oracle import terrain_oracle as orc  # noqa: E402


def _cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _np(t):
    return t.detach().cpu().numpy()


def _curvature_close(got, want, name):
    """Curvature display value ((tanh(100k)+1)/2)^(1/2.2).  Where tanh saturates at -1 (output < 0.05)
    the f32 spacing of tanh (2^-24) alone moves the output by (2^-25)^(1/2.2) = 3.8e-4 per ulp, and the
    reference itself evaluates tanhf with the CUDA math library (<= 2 ulp) while the NumPy oracle uses
    its own routine (<= 1 ulp).  There the comparison is made before the gamma: |(t+1)/2| within 4 ulp
    of tanh.  Everywhere else the standard bar applies."""
    assert np.array_equal(np.isnan(got), np.isnan(want)), name
    ok = ~np.isnan(want)
    g, w = got[ok].astype(np.float64), want[ok].astype(np.float64)
    well = w >= 0.05
    err = np.abs(g - w)
    assert np.all(err[well] <= 1e-6 + 1e-5 * np.abs(w[well])), (name, err[well].max())
    if (~well).any():
        pre = np.abs(g[~well] ** 2.2 - w[~well] ** 2.2)
        assert np.all(pre <= 4 * 2.0 ** -25 + 1e-5 * w[~well] ** 2.2), (name, pre.max())


def test_gradient_family_vs_reference_golden(golden, manifest):
    from fujishadergpu_b200.algorithms._impl_hillshade import compute_hillshade_block
    from fujishadergpu_b200.algorithms._impl_slope import compute_slope_block
    from fujishadergpu_b200.algorithms._impl_curvature import compute_curvature_block
    g = golden("gradient_family")
    for name, meta in manifest["gradient_family"].items():
        dem = _cuda(g[meta["input"]])
        algo = name.split("__")[0]
        if algo == "hillshade":
            assert_close_f32(_np(compute_hillshade_block(dem, **meta["kw"])), g[name], what=name)
        elif algo.startswith("slope"):
            assert_close_f32(_np(compute_slope_block(dem, **meta["kw"])), g[name], what=name)
        else:
            _curvature_close(_np(compute_curvature_block(dem, **meta["kw"])), g[name], name)


def test_spatial_gradient_vs_reference_golden(golden, manifest):
    """'--mode spatial' of hillshade / slope / curvature (Gaussian scale space)."""
    from fujishadergpu_b200.algorithms._impl_hillshade import compute_hillshade_spatial_block
    from fujishadergpu_b200.algorithms._impl_slope import compute_slope_spatial_block
    from fujishadergpu_b200.algorithms._impl_curvature import compute_curvature_spatial_block
    from fujishadergpu_b200.algorithms.tile.hillshade import HillshadeAlgorithm as TileHillshade
    g = golden("spatial_gradient")
    for name, meta in manifest["spatial_gradient"].items():
        dem = _cuda(g[meta["input"]])
        kw = dict(meta["kw"])
        algo = meta["algo"]
        if algo == "hillshade":
            assert_close_f32(_np(compute_hillshade_spatial_block(dem, **kw)), g[name], what=name)
        elif algo == "slope":
            assert_close_f32(_np(compute_slope_spatial_block(dem, **kw)), g[name], what=name)
        elif algo == "curvature":
            _curvature_close(_np(compute_curvature_spatial_block(dem, **kw)), g[name], name)
        else:   # the tile adapter's multi-radius path + combiner
            got = TileHillshade().process(dem, multiscale=True, **kw)
            assert_close_f32(_np(got), g[name], what=name)


def test_spatial_mode_algorithm_classes_vs_oracle():
    from fujishadergpu_b200.algorithms.dask_registry import ALGORITHMS
    dem = orc.synth_dem(700, 900, seed=41, nodata=True)
    d = _cuda(dem)
    kw = dict(pixel_scale_x=1.0, pixel_scale_y=-1.0, pixel_size=1.0)
    radii = [2, 8, 32]
    w = orc.pow2_weights(3)
    # slope / curvature: _combine_multiscale_dask (weights cleaned in Python floats)
    clean = [float(x) / float(sum(w)) for x in w]
    depth = {"slope": lambda r: max(2, int(r * 2 + 1)), "curvature": lambda r: max(3, int(r * 2 + 2))}
    for name, fn, extra in (("slope", orc.slope_spatial_block, dict(unit="degree")),
                            ("curvature", orc.curvature_spatial_block, dict(curvature_type="mean"))):
        # small radii: map_overlap(depth, boundary='reflect') on the whole raster (reference _nan_utils.py:516-522)
        resp = [orc.with_overlap(fn, dem, depth[name](float(r)), radius=float(r), **extra, **kw) for r in radii]
        want = resp[0] * np.float32(clean[0])
        for i in range(1, 3):
            want = want + resp[i] * np.float32(clean[i])
        got = _np(ALGORITHMS[name].process(d, mode="spatial", radii=radii, weights=w, **extra, **kw))
        if name == "curvature":
            # the per-radius responses are held to the reference in the golden test (tanh-saturation aware);
            # here the COMBINER is checked exactly, on the device responses themselves
            from fujishadergpu_b200.algorithms._impl_curvature import compute_curvature_spatial_block
            from fujishadergpu_b200.algorithms._nan_utils import _symmetric_pad
            dresp = []
            for r in radii:
                dp = depth[name](float(r))
                full = _np(compute_curvature_spatial_block(_symmetric_pad(d, dp), radius=float(r), **extra, **kw))
                dresp.append(full[dp:dp + dem.shape[0], dp:dp + dem.shape[1]])
            exact = dresp[0] * np.float32(clean[0])
            for i in range(1, 3):
                exact = exact + dresp[i] * np.float32(clean[i])
            assert np.array_equal(got, exact.astype(np.float32), equal_nan=True), name
        else:
            assert_close_f32(got, want, what=name)
    # hillshade: f32-normalised weights, auto radii/weights when none are given
    resp = [orc.with_overlap(orc.hillshade_spatial_block, dem, max(2, int(r * 2 + 1)), radius=float(r), **kw) for r in radii]
    want = orc.combine_responses(resp, weights=w, agg="mean")
    got = _np(ALGORITHMS["hillshade"].process(d, mode="spatial", radii=radii, weights=None, **kw))
    assert_close_f32(got, want, what="hillshade spatial")


def test_spatial_mode_large_radius_coarse_path_vs_oracle():
    """Radii above max(256, min(H,W)//16) on a raster longer than 2048 px: coarsened DEM, block function on the
    reflect-padded coarse array, pixel-centre bilinear sampling back (reference _nan_utils.py:329-438)."""
    from fujishadergpu_b200.algorithms.dask_registry import ALGORITHMS
    dem = orc.synth_dem(2600, 3000, seed=43, nodata=True)
    d = _cuda(dem)
    kw = dict(pixel_scale_x=1.0, pixel_scale_y=-1.0, pixel_size=1.0)
    radii, w = [8, 300], [0.5, 0.5]
    F = orc.coarsen_factor(dem.shape)
    assert F == 2
    # hillshade: f32-normalised weights
    big = orc.large_radius_response(dem, 300.0, F, orc.hillshade_spatial_block, lambda rc: max(2, int(rc * 2 + 1)), **kw)
    small = orc.with_overlap(orc.hillshade_spatial_block, dem, 17, radius=8.0, **kw)
    want = orc.combine_responses([small, big], weights=w, agg="mean")
    got = _np(ALGORITHMS["hillshade"].process(d, mode="spatial", radii=radii, weights=w, **kw))
    assert_close_f32(got, want, rtol=1e-5, atol=2e-6, what="hillshade spatial, large radius")
    # slope: python-float weights
    big = orc.large_radius_response(dem, 300.0, F, orc.slope_spatial_block, lambda rc: max(2, int(rc * 2 + 1)),
                                    unit="degree", **kw)
    small = orc.with_overlap(orc.slope_spatial_block, dem, 17, radius=8.0, unit="degree", **kw)
    want = small * np.float32(0.5) + big * np.float32(0.5)
    got = _np(ALGORITHMS["slope"].process(d, mode="spatial", radii=radii, weights=w, unit="degree", **kw))
    assert_close_f32(got, want, rtol=1e-5, atol=2e-5, what="slope spatial, large radius")
    # an injected overview (what the reference's orchestration passes) is used as given
    ov = _cuda(orc.synth_dem(650, 750, seed=44))
    got = ALGORITHMS["hillshade"].process(d, mode="spatial", radii=[300], weights=[1.0], _overview_coarse_dem=ov,
                                          _overview_decimation=4.0, **kw)
    resp = orc.hillshade_spatial_block(_np(ov), radius=75, pixel_size=4.0, pixel_scale_x=4.0, pixel_scale_y=-4.0)
    want = np.where(np.isnan(dem), np.float32(np.nan), orc.sample_overview_field(resp, 0, 2600, 0, 3000, 2600, 3000))
    assert_close_f32(_np(got), want, rtol=1e-5, atol=2e-6, what="hillshade spatial, injected overview")


def test_gradient_family_large_vs_oracle():
    from fujishadergpu_b200 import kernels as k
    dem = orc.synth_dem(1500, 1111, seed=11, nodata=True)
    d = _cuda(dem)
    kw = dict(pixel_scale_x=1.0, pixel_scale_y=-1.0)
    assert_close_f32(_np(k.hillshade(d, **kw)), orc.hillshade_block(dem, **kw), what="hillshade")
    for unit in ("degree", "percent", "radian"):
        assert_close_f32(_np(k.slope(d, unit=unit, **kw)), orc.slope_block(dem, unit=unit, **kw), what=unit)
    for ct in ("mean", "gaussian", "planform", "profile"):
        _curvature_close(_np(k.curvature(d, curvature_type=ct, **kw)), orc.curvature_block(dem, curvature_type=ct, **kw), ct)


def test_gradient_small_raster_raises_like_numpy():
    from fujishadergpu_b200 import kernels as k
    with pytest.raises(ValueError):
        k.hillshade(torch.zeros((2, 5), device="cuda"))


def test_gradient_encoded_outputs():
    from fujishadergpu_b200 import kernels as k
    from fujishadergpu_b200.io.output_encoding import quantize_params, resolve_output_range
    dem = orc.synth_dem(300, 260, seed=3, nodata=True)
    d = _cuda(dem)
    for algo, fn, ofn in (("hillshade", k.hillshade, orc.hillshade_block), ("slope", k.slope, orc.slope_block)):
        for dt in ("uint8", "int16"):
            qp = quantize_params(*resolve_output_range(algo), dt)
            got = _np(fn(d, output_dtype=dt, qp=qp)).astype(np.int64)
            want = orc.encode_array(ofn(dem), qp, dt).astype(np.int64)
            assert np.array_equal(got == 0, want == 0), (algo, dt)      # identical NoData mask
            assert np.abs(got - want).max() <= 1, (algo, dt)


def _topo_close(got, want, scale, what):
    assert np.array_equal(np.isnan(got), np.isnan(want)), what
    ok = ~np.isnan(want)
    err = np.abs(got[ok].astype(np.float64) - want[ok].astype(np.float64))
    lim = 1e-6 * scale + 1e-5 * np.abs(want[ok].astype(np.float64))
    assert np.all(err <= lim), (what, float(err.max()), float(scale), int((err > lim).sum()))
    return float(np.mean(got[ok] == want[ok]))


def test_topousm_fast_vs_reference_golden(golden, manifest):
    from fujishadergpu_b200.algorithms._impl_topousm_fast import compute_topousm_fast_efficient_block
    from fujishadergpu_b200 import kernels as k
    g = golden("topousm_fast")
    for cname, meta in manifest["topousm_fast"].items():
        if cname == "large_part":
            continue
        dem = _cuda(g[meta["input"]])
        raw = _np(compute_topousm_fast_efficient_block(dem, **meta["kw"]))
        exact = _topo_close(raw, g[f"raw__{cname}"], meta["scale"], cname)
        print(f"{cname}: bit-exact fraction {exact:.5f}")
        if f"norm__{cname}" in g:
            kw = dict(meta["kw"])
            norm = _np(k.topousm_fast(dem, radii=kw.get("radii", [4, 16, 64]), weights=kw.get("weights"),
                                      pixel_size=kw.get("pixel_size", 1.0), norm_scale=meta["scale"]))
            assert_close_f32(norm, g[f"norm__{cname}"], what="norm " + cname)


def test_topousm_helpers_vs_reference_golden(golden, manifest):
    from fujishadergpu_b200 import kernels as k
    g = golden("topousm_fast")
    for key in ("tdense", "tholes", "tvoid"):
        for f in (2, 4, 16):
            got = _np(k.decimate(_cuda(g[key]), f))
            assert_close_f32(got, g[f"decimate{f}__{key}"], rtol=1e-6, atol=0, what=f"decimate{f} {key}")
    small = _cuda(g["decimate4__tdense"])
    assert np.array_equal(_np(k.upsample(small, g["tdense"].shape)), g["upsample4__tdense"])
    m = manifest["topousm_fast"]["large_part"]
    r0, r1, c0, c1 = m["window"]
    part = _np(k.topousm_large_part(_cuda(g["tdense"][r0:r1, c0:c1]), _cuda(g["large_field"]), w_large=m["w_large"],
                                    off_r=r0, off_c=c0, full_h=g["tdense"].shape[0], full_w=g["tdense"].shape[1]))
    assert_close_f32(part, g["large_part"], what="large_part")


def test_topousm_fast_large_vs_oracle():
    """2048 x 1792 with the full ladder (every pyramid level, several strips/bands, ragged edges)."""
    from fujishadergpu_b200 import kernels as k
    for nodata in (False, True):
        dem = orc.synth_dem(2051, 1797, seed=21 + nodata, nodata=nodata)
        radii, w = [2, 8, 32, 128, 512, 2048], orc.pow2_weights(6)
        want = orc.topousm_fast_block(dem, radii=radii, weights=w)
        scale = orc.abs_p99_scale(want)[0]
        got = _np(k.topousm_fast(_cuda(dem), radii=radii, weights=w))
        exact = _topo_close(got, want, scale, f"ladder6 nodata={nodata}")
        print(f"nodata={nodata}: bit-exact fraction {exact:.6f}, scale {scale:.4f}")
        # fused normalisation == IEEE f32 division of the raw output (block / scale)
        for sc in (scale, 3.0, 0.731):
            normed = _np(k.topousm_fast(_cuda(dem), radii=radii, weights=w, norm_scale=sc))
            assert np.array_equal(normed, got / np.float32(sc), equal_nan=True), sc
        qp = orc.encode_params(*orc.value_range("topousm_fast"), "uint8")
        got8 = _np(k.topousm_fast(_cuda(dem), radii=radii, weights=w, norm_scale=scale, output_dtype="uint8", qp=qp))
        want8 = orc.encode_array(orc.normalise_by_scale(want.copy(), (scale,)), qp, "uint8")
        assert np.array_equal(got8 == 0, want8 == 0)
        assert np.abs(got8.astype(int) - want8.astype(int)).max() <= 1


def test_topousm_generic_and_fast_kernels_agree(monkeypatch):
    """The streaming kernel has a general variant (any raster size, multiple mirror reflections) and a
    fast variant; both must give the same bits, on dense and NoData rasters, f32 and uint8."""
    from fujishadergpu_b200 import kernels as k
    radii, w = [2, 8, 32, 128, 512], orc.pow2_weights(5)
    qp = orc.encode_params(*orc.value_range("topousm_fast"), "uint8")
    for nodata in (False, True):
        dem = _cuda(orc.synth_dem(777, 1300, seed=31 + nodata, nodata=nodata))
        monkeypatch.delenv("FSG_FORCE_GENERIC", raising=False)
        fast = _np(k.topousm_fast(dem, radii=radii, weights=w))
        fast8 = _np(k.topousm_fast(dem, radii=radii, weights=w, norm_scale=11.0, output_dtype="uint8", qp=qp))
        monkeypatch.setenv("FSG_FORCE_GENERIC", "1")
        gen = _np(k.topousm_fast(dem, radii=radii, weights=w))
        gen8 = _np(k.topousm_fast(dem, radii=radii, weights=w, norm_scale=11.0, output_dtype="uint8", qp=qp))
        monkeypatch.delenv("FSG_FORCE_GENERIC", raising=False)
        assert np.array_equal(np.isnan(fast), np.isnan(gen))
        scale = float(np.nanpercentile(np.abs(gen), 99))
        _topo_close(fast, gen, scale, f"fast vs generic nodata={nodata}")
        assert np.abs(fast8.astype(int) - gen8.astype(int)).max() <= 1


def test_topousm_tiny_and_ragged_rasters_vs_oracle():
    """Rasters smaller than the window (re-reflection), single rows/columns, ragged strip edges."""
    from fujishadergpu_b200 import kernels as k
    for shape, radii in (((20, 300), [2, 8, 32]), ((300, 19), [2, 8, 32, 128]), ((1, 50), [3]), ((50, 1), [3, 100]),
                         ((33, 257), [41, 42]), ((97, 513), [2, 8, 32, 128, 512, 2048])):
        dem = orc.synth_dem(shape[0], shape[1], seed=shape[0] + shape[1])
        dem[shape[0] // 2, shape[1] // 3] = np.nan
        want = orc.topousm_fast_block(dem, radii=radii)
        got = _np(k.topousm_fast(_cuda(dem), radii=radii))
        scale = max(1e-3, float(np.nanpercentile(np.abs(want), 99)))
        _topo_close(got, want, scale, f"{shape} {radii}")


def test_topousm_overview_large_field_and_nodata_mask():
    """compute_topousm_fast_large_coarse_field (reference _impl_topousm_fast.py:133-155) and the NoData re-mask
    helpers (core/tile_compute.py:132-152, core/tile_processor.py:185-196)."""
    from fujishadergpu_b200.algorithms._impl_topousm_fast import compute_topousm_fast_large_coarse_field
    from fujishadergpu_b200.core.tile_compute import apply_nodata_mask, build_nodata_mask
    coarse = orc.synth_dem(300, 260, seed=51, nodata=True)
    kw = dict(large_radii=[512, 2048, 3], large_weights=[0.2, 0.1, 0.05], decimation=16.0)
    want = orc.topousm_large_field(coarse, **kw)
    got = _np(compute_topousm_fast_large_coarse_field(_cuda(coarse), **kw))
    assert np.array_equal(got, want), np.abs(got - want).max()
    dem = orc.synth_dem(64, 80, seed=52)
    dem[3, 4] = -9999.0
    dem[10, 11] = np.nan
    mask = build_nodata_mask(dem, -9999.0)
    assert mask[3, 4] and mask[10, 11] and mask.sum() == 2
    dmask = build_nodata_mask(_cuda(dem), -9999.0)
    assert np.array_equal(_np(dmask), mask)
    res = torch.ones((64, 80), device="cuda")
    apply_nodata_mask(res, mask, float("nan"))
    assert np.array_equal(np.isnan(_np(res)), mask)
    stack = torch.ones((3, 64, 80), device="cuda")
    apply_nodata_mask(stack, dmask, 0.0)
    assert float(stack[:, 3, 4].abs().sum()) == 0.0 and float(stack.sum()) == 3 * (64 * 80 - 2)


def test_topousm_weights_length_mismatch_raises():
    from fujishadergpu_b200.algorithms._impl_topousm_fast import compute_topousm_fast_efficient_block
    with pytest.raises(ValueError):
        compute_topousm_fast_efficient_block(torch.zeros((64, 64), device="cuda"), radii=[2, 8], weights=[1.0])


def test_stats_percentile_matches_numpy():
    from fujishadergpu_b200.algorithms._normalization import topousm_fast_stat_func
    from fujishadergpu_b200.algorithms._global_stats import robust_unsigned_stretch_stat_func
    rng = np.random.default_rng(4)
    a = (rng.standard_normal((700, 900)) * 3).astype(np.float32)
    a[rng.random(a.shape) < 0.05] = np.nan
    assert topousm_fast_stat_func(_cuda(a))[0] == orc.abs_p99_scale(a)[0]
    b = rng.random((500, 333)).astype(np.float32)
    b[::7, ::5] = np.nan
    assert robust_unsigned_stretch_stat_func(_cuda(b)) == orc.p1_p99_stretch_stats(b)
    # pooled chunks (the stratified-window pre-pass)
    parts = [a[:300, :400], a[350:, 100:], a[100:200, :]]
    pooled = np.concatenate([p[~np.isnan(p)] for p in parts])
    assert topousm_fast_stat_func([_cuda(p) for p in parts])[0] == orc.abs_p99_scale(pooled)[0]
    # reference known answers (tests/test_topousm_fast_normalization.py:21-33)
    data = np.concatenate([np.full(800, 0.01, np.float32), np.full(200, 0.5, np.float32)])
    assert topousm_fast_stat_func(_cuda(data))[0] == pytest.approx(0.5, rel=1e-2)


def test_staged_percentile_matches_numpy_and_two_pass():
    """Device-staged radix select (rank derived on the device with NumPy's f32 virtual-index arithmetic) ==
    np.percentile on the f32 sample == the host-ranked two-pass selection; empty, single, tied and signed samples."""
    from fujishadergpu_b200 import kernels as k
    rng = np.random.default_rng(12)
    cases = []
    for n in (1, 2, 3, 101, 4097, 250_001):
        a = (rng.standard_normal(n) * 50).astype(np.float32)
        cases.append(a)
    tied = np.repeat(np.float32([0.5, -0.5, 0.25, 7.0]), 1000)
    cases.append(tied)
    withnan = (rng.standard_normal(20_000) * 3).astype(np.float32)
    withnan[rng.random(withnan.shape) < 0.3] = np.nan
    withnan[5] = np.inf
    cases.append(withnan)
    for a in cases:
        t = _cuda(a.reshape(1, -1))
        for q in (0.0, 1.0, 50.0, 99.0, 100.0):
            for take_abs, finite_only in ((True, False), (False, True)):
                sample = a[np.isfinite(a)] if finite_only else a[~np.isnan(a)]
                sample = np.abs(sample) if take_abs else sample
                want = float(np.percentile(sample, q)) if sample.size else float("nan")
                got = k.percentile([t], q, take_abs=take_abs, finite_only=finite_only)
                old = k.percentile_two_pass([t], q, take_abs=take_abs, finite_only=finite_only)
                assert (got == want) or (got != got and want != want), (a.size, q, take_abs, got, want)
                assert (got == old) or (got != got and old != old), (a.size, q, take_abs, got, old)
                # four scans of the sample (no compaction of the level-0 bucket): same value
                full = k.staged_percentile([t], q, take_abs=take_abs, finite_only=finite_only, device=t.device, compact=False)
                assert (got == full) or (got != got and full != full), (a.size, q, take_abs, got, full)
    allnan = _cuda(np.full((4, 9), np.nan, np.float32))
    assert k.percentile([allnan], 99.0, take_abs=True) != k.percentile([allnan], 99.0, take_abs=True)   # NaN


def test_openness_vs_reference_golden(golden, manifest):
    from fujishadergpu_b200.algorithms._impl_openness import compute_openness_vectorized, compute_openness_spatial_block
    from fujishadergpu_b200.algorithms._global_stats import _apply_display_stretch_block
    g = golden("openness")
    for name, meta in manifest["openness"].items():
        if name == "stretch":
            continue
        fn = compute_openness_vectorized if name.startswith("local__") else compute_openness_spatial_block
        got = _np(fn(_cuda(g[meta["input"]]), **meta["kw"]))
        assert_close_f32(got, g[name], what=name)
    st = manifest["openness"]["stretch"]
    got = _np(_apply_display_stretch_block(_cuda(g[st["of"]]), tuple(st["stats"])))
    assert_close_f32(got, g["stretch__pos8_r64"], what="stretch")


def test_curvature_streaming_kernel_equals_tile_kernel_bitwise(monkeypatch):
    """curv_stream_kernel (registers, central forms, frame redone) against the tile kernel it replaces for aligned
    rasters: every curvature type, dense / NoData, f32 and uint8, an odd-sized and a band-limited call."""
    from fujishadergpu_b200 import kernels as k
    for shape, nod in (((1500, 2900), False), ((777, 1164), True), ((64, 2048), False), ((9, 12), False)):
        dem = orc.synth_dem(shape[0], shape[1], seed=70 + shape[0], nodata=nod)
        d = _cuda(dem)
        for ct in ("mean", "gaussian", "planform", "profile"):
            kw = dict(curvature_type=ct, pixel_size=1.0, pixel_scale_x=2.0, pixel_scale_y=-2.0)
            monkeypatch.setenv("FSG_GRAD_TILED", "1")
            want = k.curvature(d, **kw)
            want8 = k.curvature(d, output_dtype="uint8", qp={"a_coef": 254.0, "b_coef": 1.0, "dn_min": 1, "dn_max": 255}, **kw)
            monkeypatch.delenv("FSG_GRAD_TILED")
            got = k.curvature(d, **kw)
            got8 = k.curvature(d, output_dtype="uint8", qp={"a_coef": 254.0, "b_coef": 1.0, "dn_min": 1, "dn_max": 255}, **kw)
            assert torch.equal(torch.nan_to_num(got, nan=-7.0), torch.nan_to_num(want, nan=-7.0)), (shape, ct)
            assert torch.equal(got8, want8), (shape, ct)
        if shape[0] >= 777:   # rows [200, 500) of the raster from a buffer holding rows [190, 520)
            monkeypatch.setenv("FSG_GRAD_TILED", "1")
            wb = k.curvature(d[190:520], curvature_type="mean", band=dict(h_global=shape[0], buf_row0=190, out_row0=200, out_rows=300))
            monkeypatch.delenv("FSG_GRAD_TILED")
            gb = k.curvature(d[190:520], curvature_type="mean", band=dict(h_global=shape[0], buf_row0=190, out_row0=200, out_rows=300))
            whole = k.curvature(d, curvature_type="mean")[200:500]
            assert torch.equal(torch.nan_to_num(gb, nan=-7.0), torch.nan_to_num(wb, nan=-7.0))
            assert torch.equal(torch.nan_to_num(gb, nan=-7.0), torch.nan_to_num(whole, nan=-7.0))


def test_ambient_occlusion_vs_reference_golden(golden, manifest):
    """compute_ambient_occlusion_block / _spatial_block against reference outputs."""
    from fujishadergpu_b200.algorithms._impl_ambient_occlusion import (compute_ambient_occlusion_block,
                                                                       compute_ambient_occlusion_spatial_block)
    g = golden("ambient_occlusion")
    for name, meta in manifest["ambient_occlusion"].items():
        fn = compute_ambient_occlusion_block if name.startswith("local__") else compute_ambient_occlusion_spatial_block
        got = _np(fn(_cuda(g[meta["input"]]), **meta["kw"]))
        assert_close_f32(got, g[name], what=name)


def test_ambient_occlusion_classes_stretch_and_encoding():
    """Registry / tile adapter / fused stretch + uint8 encoding against the oracle composition."""
    from fujishadergpu_b200 import kernels as k
    from fujishadergpu_b200.algorithms.dask_registry import ALGORITHMS
    from fujishadergpu_b200.algorithms.tile.ambient_occlusion import AmbientOcclusionAlgorithm as TileAO
    from fujishadergpu_b200.io.output_encoding import quantize_params, resolve_output_range
    dem = orc.synth_dem(300, 420, seed=41, nodata=True)
    d = _cuda(dem)
    kw = dict(num_samples=16, radius=12.0, intensity=1.0, pixel_size=1.0, pixel_scale_x=1.0, pixel_scale_y=-1.0)
    want = orc.ambient_occlusion_block(dem, **kw)
    # Algorithm.process on a device array = map_overlap(depth=radius + 1, boundary='reflect') with one block
    want_ov = orc.with_overlap(orc.ambient_occlusion_block, dem, int(kw.get("radius", 10.0) + 1), **kw)
    assert_close_f32(_np(ALGORITHMS["ambient_occlusion"].process(d, **kw)), want_ov, what="ao local")
    stats = orc.p1_p99_stretch_stats(want)
    want_st = orc.display_stretch(want, stats)
    assert_close_f32(_np(TileAO().process(d, global_stats=stats, **kw)), want_st, rtol=2e-5, atol=2e-6, what="ao tile stretch")
    assert_close_f32(_np(ALGORITHMS["ambient_occlusion"].process(d, global_stats=stats, **kw)),
                     orc.display_stretch(want_ov, stats), rtol=2e-5, atol=2e-6, what="ao dask-class stretch")
    qp = quantize_params(*resolve_output_range("ambient_occlusion"), "uint8")
    got8 = _np(k.ambient_occlusion(d, output_dtype="uint8", qp=qp, **kw)).astype(np.int32)
    want8 = orc.encode_array(want, qp, "uint8").astype(np.int32)
    assert np.array_equal(got8 == 0, want8 == 0)
    assert np.abs(got8 - want8).max() <= 1
    # spatial mode: two radii mixed with the automatic 2^n weights
    radii = [6, 40]
    resp = [orc.with_overlap(orc.ambient_occlusion_spatial_block, dem, int(r) + 1, **{**kw, "radius": float(r)}) for r in radii]
    want_sp = orc.combine_responses(resp, weights=orc.pow2_weights(2), agg="mean")
    got_sp = _np(ALGORITHMS["ambient_occlusion"].process(d, mode="spatial", radii=radii, weights=None, **kw))
    assert_close_f32(got_sp, want_sp, what="ao spatial")


def test_openness_spatial_multi_radius_vs_oracle():
    from fujishadergpu_b200.algorithms.dask_registry import ALGORITHMS
    dem = orc.synth_dem(300, 420, seed=43, nodata=True)
    kw = dict(openness_type="positive", num_directions=8, pixel_size=1.0)
    radii, w = [8, 40, 120], [0.5, 0.3, 0.2]
    resp = [orc.with_overlap(orc.openness_spatial_block, dem, int(max(2, round(float(r)))) + 1,
                             max_distance=float(int(max(2, round(float(r))))), **kw) for r in radii]
    want = orc.combine_responses(resp, weights=w, agg="mean")
    got = _np(ALGORITHMS["openness"].process(_cuda(dem), mode="spatial", radii=radii, weights=w, **kw))
    assert_close_f32(got, want, what="openness spatial 3 radii")


def test_openness_known_answers():
    """tests/test_openness_yokoyama.py:21-47 of the reference, on the CUDA path."""
    from fujishadergpu_b200.algorithms._impl_openness import compute_openness_vectorized as op
    yy, xx = np.mgrid[0:101, 0:101]
    r = np.sqrt((xx - 50) ** 2 + (yy - 50) ** 2)
    peak, pit = _cuda((50 - r).astype(np.float32)), _cuda((r - 50).astype(np.float32))
    kw = dict(num_directions=16, max_distance=40)
    assert op(peak, openness_type="positive", **kw)[50, 50] > op(pit, openness_type="positive", **kw)[50, 50]
    assert op(pit, openness_type="negative", **kw)[50, 50] > op(peak, openness_type="negative", **kw)[50, 50]
    flat = torch.zeros((101, 101), device="cuda")
    for t in ("positive", "negative"):
        assert float(op(flat, openness_type=t, **kw)[50, 50]) == pytest.approx(1.0, abs=1e-3)


def test_tile_adapter_matches_reference(golden, manifest):
    from fujishadergpu_b200.algorithms.tile.topousm_fast import TopoUSMFastAlgorithm
    from fujishadergpu_b200.core.tile_compute import run_tile_algorithm
    meta = manifest["tile_adapter"]["topousm_fast_gs7p25"]
    dem = _cuda(golden("topousm_fast")["tdense"])
    kw = dict(meta["kw"])
    kw["global_stats"] = tuple(kw["global_stats"])
    got = _np(TopoUSMFastAlgorithm().process(dem, **kw))
    assert_close_f32(got, golden("tile_adapter")["topousm_fast_gs7p25"], what="tile adapter")
    got2 = _np(run_tile_algorithm(TopoUSMFastAlgorithm(), "topousm_fast", dem, 1.0, False, 1.0,
                                  {"radii": [2, 8, 32, 128], "global_stats": (7.25,)}))
    assert np.array_equal(got, got2)


def test_encode_known_vector():
    """tests/test_audit_p1_regressions.py:45-62 of the reference."""
    from fujishadergpu_b200.io.output_encoding import quantize_array, quantize_params
    v = _cuda(np.array([np.inf, -np.inf, np.nan, 1.0], np.float32))
    assert _np(quantize_array(v, quantize_params(0.0, 1.0, "uint8"), "uint8")).tolist() == [0, 0, 0, 255]
    assert _np(quantize_array(v, quantize_params(0.0, 100.0, "int16"), "int16")).tolist()[:3] == [0, 0, 0]


def test_library_is_the_path_that_ran():
    from fujishadergpu_b200 import kernels as k
    k.reset_launch_count()
    k.hillshade(torch.zeros((64, 64), device="cuda"))
    assert k.launch_count() >= 1


def test_streamed_host_pipeline_equals_whole_raster_call():
    """Chunked upload + per-band compute + overlapped download == one whole-raster call, byte for byte."""
    from fujishadergpu_b200.core.tile_processor import HostTilePipeline, StreamedTopoPipeline
    shape = (3100, 2700)
    params = {"radii": [2, 8, 32, 128, 512], "weights": orc.pow2_weights(5), "pixel_size": 1.0}
    for nod in (False, True):
        dem = orc.synth_dem(*shape, seed=91, nodata=nod)
        hin = torch.from_numpy(dem).pin_memory()
        for od, dt in (("uint8", torch.uint8), ("float32", torch.float32)):
            whole = torch.empty(shape, dtype=dt).pin_memory()
            banded = torch.empty(shape, dtype=dt).pin_memory()
            HostTilePipeline(shape, "topousm_fast", params, output_dtype=od).run(hin, whole)
            StreamedTopoPipeline(shape, params, output_dtype=od, chunk_rows=512).run(hin, banded)
            assert np.array_equal(whole.numpy(), banded.numpy(), equal_nan=True), (nod, od)


def test_gradient_fast_math_error_budget():
    """hillshade / slope / curvature use approximate SFU forms (rsqrt, sqrt, rcp, ex2, lg2) and a polynomial
    arctangent: the contract is 1e-5 relative / 1e-6 absolute against the reference, and the measured distance to the
    oracle stays far inside it (B200, 2048^2 synthetic DEM: <= 0.06 of the bar for hillshade / slope, <= 0.3 for
    the curvature display value)."""
    from fujishadergpu_b200 import kernels as k
    d = k.synth_dem((2048, 2048), seed=7, nodata=False)
    dem = _np(d)
    nup = dict(pixel_scale_x=1.0, pixel_scale_y=-1.0)

    def frac_of_bar(got, want):
        got, want = _np(got).astype(np.float64), want.astype(np.float64)
        return float(np.max(np.abs(got - want) / (1e-6 + 1e-5 * np.abs(want))))

    assert frac_of_bar(k.hillshade(d, **nup), orc.hillshade_block(dem, **nup)) <= 0.1
    assert frac_of_bar(k.hillshade(d, pixel_scale_x=30.0, pixel_scale_y=-30.0),
                       orc.hillshade_block(dem, pixel_scale_x=30.0, pixel_scale_y=-30.0)) <= 0.1
    for unit in ("degree", "radian", "percent"):
        assert frac_of_bar(k.slope(d, unit=unit, **nup), orc.slope_block(dem, unit=unit, **nup)) <= 0.1, unit
    for ctype in ("mean", "gaussian", "planform", "profile"):
        g = _np(k.curvature(d, curvature_type=ctype, **nup)).astype(np.float64)
        w = orc.curvature_block(dem, curvature_type=ctype, **nup).astype(np.float64)
        ok = w >= 0.05
        assert np.max(np.abs(g[ok] - w[ok]) / (1e-6 + 1e-5 * w[ok])) <= 0.5, ctype
