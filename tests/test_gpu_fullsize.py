"""Parity at BASELINE.json's full sizes.  The NumPy/SciPy oracle cannot evaluate whole rasters of this size in
test time, so the checks are (1) the oracle on crops whose interior depends only on pixels inside the crop,
(2) size-independent properties of the algorithms, and (3) agreement of independent implementations
(general vs streaming kernel, banded vs whole-raster call), always through the C ABI.

    config 0  hillshade local            4096^2   whole-raster oracle
    config 1  slope + curvature          16384^2  crop oracle (centre / edges / corners), NaN-mask identity
    config 2  topousm_fast, 6 radii      32768^2  antisymmetry, general-vs-streaming kernel on bands, crop oracle
                                                  of the full-resolution radii, banded == whole
    config 3  openness 8 dir, r = 256    32768^2  crop oracle with a 256-px halo, flat raster -> 1.0
    config 4  (131072^2 on 8 GPUs) is covered at 16384^2 / 32768^2 by tests/test_gpu_sharding.py
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from conftest import assert_close_f32  # noqa: E402
from oracle import terrain_oracle as orc  # noqa: E402

NUP = dict(pixel_scale_x=1.0, pixel_scale_y=-1.0)


def _np(t):
    return t.detach().cpu().numpy()


def _crops(S, side):
    """(row0, col0) of crops: four corners, four edge centres, the centre."""
    pts = [0, (S - side) // 2, S - side]
    return [(r, c) for r in pts for c in pts]


def _same_bits(a, b):
    an, bn = torch.isnan(a), torch.isnan(b)
    return bool(torch.equal(an, bn)) and bool(torch.equal(torch.nan_to_num(a), torch.nan_to_num(b)))


def test_config0_hillshade_4096_whole_oracle():
    from fujishadergpu_b200 import kernels as k
    d = k.synth_dem((4096, 4096), seed=20261017, nodata=True)
    got = _np(k.hillshade(d, **NUP))
    want = orc.hillshade_block(_np(d), **NUP)
    assert_close_f32(got, want, what="hillshade 4096^2")
    qp = orc.encode_params(*orc.value_range("hillshade"), "uint8")
    got8 = _np(k.hillshade(d, output_dtype="uint8", qp=qp, **NUP))
    want8 = orc.encode_array(want, qp, "uint8")
    assert np.array_equal(got8 == 0, want8 == 0)
    assert np.abs(got8.astype(int) - want8.astype(int)).max() <= 1


def test_config1_slope_curvature_16384_crop_oracle():
    """3x3 / 5x5 stencils + a 4-px gap-fill reach: a crop with a 12-px margin reproduces its interior; at raster
    edges the crop edge IS the raster edge, so the one-sided forms are covered too."""
    from fujishadergpu_b200 import kernels as k
    S, side, m = 16384, 1024, 12
    d = k.synth_dem((S, S), seed=20261018, nodata=True)
    slope = k.slope(d, unit="degree", **NUP)
    curv = k.curvature(d, curvature_type="mean", **NUP)
    assert bool(torch.equal(torch.isnan(slope), torch.isnan(d))) and bool(torch.equal(torch.isnan(curv), torch.isnan(d)))
    for (r0, c0) in _crops(S, side):
        crop = _np(d[r0:r0 + side, c0:c0 + side])
        lo_r, hi_r = (0 if r0 == 0 else m), (side if r0 + side == S else side - m)
        lo_c, hi_c = (0 if c0 == 0 else m), (side if c0 + side == S else side - m)
        ws = orc.slope_block(crop, unit="degree", **NUP)[lo_r:hi_r, lo_c:hi_c]
        gs = _np(slope[r0 + lo_r:r0 + hi_r, c0 + lo_c:c0 + hi_c])
        assert_close_f32(gs, ws, what=f"slope crop {r0},{c0}")
        wc = orc.curvature_block(crop, curvature_type="mean", **NUP)[lo_r:hi_r, lo_c:hi_c]
        gc = _np(curv[r0 + lo_r:r0 + hi_r, c0 + lo_c:c0 + hi_c])
        assert np.array_equal(np.isnan(gc), np.isnan(wc))
        ok = ~np.isnan(wc) & (wc >= 0.05)     # away from the tanh saturation (see test_gpu_parity._curvature_close)
        assert np.all(np.abs(gc[ok] - wc[ok]) <= 1e-6 + 1e-5 * np.abs(wc[ok])), f"curvature crop {r0},{c0}"


def test_config2_topousm_32768_properties(monkeypatch):
    from fujishadergpu_b200 import kernels as k
    S = 32768
    radii, w = [2, 8, 32, 128, 512, 2048], orc.pow2_weights(6)
    d = k.synth_dem((S, S), seed=20261019, nodata=False)
    ws = torch.empty(max(256, k.topousm_fast_workspace_bytes((S, S), radii, 1.0)), dtype=torch.uint8, device="cuda")
    out = k.topousm_fast(d, radii=radii, weights=w, workspace=ws)
    assert bool(torch.isfinite(out).all())
    # (a) antisymmetry: every sum, rounding and division is odd -> topousm(-dem) == -topousm(dem), bit for bit
    neg = k.topousm_fast(-d, radii=radii, weights=w, workspace=ws)
    assert bool(torch.equal(neg, -out)), "antisymmetry"
    del neg
    # (b) p99 scale of the production pre-pass == exact selection over the same windows done by torch
    from fujishadergpu_b200.algorithms._norm_stats import compute_norm_stats_device
    from fujishadergpu_b200.algorithms._norm_stats import _norm_stat_window_geometry, stratified_windows
    st = compute_norm_stats_device(d, "topousm_fast", {"radii": radii, "weights": w, "pixel_size": 1.0})
    margin, tile = _norm_stat_window_geometry("topousm_fast", {"radii": radii})
    wins = stratified_windows(S, S, 0, S, 0, S, grid=3, tile=min(tile, S))   # dense raster: bounding box = raster
    assert len(set(wins)) == 9
    pooled = []
    for (wy0, wx0, tw, th) in wins:      # every window evaluated whole (no region of interest), trimmed, pooled
        r = k.topousm_fast(d[wy0:wy0 + th, wx0:wx0 + tw], radii=radii, weights=w)
        m = int(min(margin, th // 3, tw // 3))
        pooled.append(_np(r[m:-m, m:-m].abs()).ravel())
        del r
    want = np.percentile(np.concatenate(pooled), 99.0)
    del pooled
    assert st is not None and np.float32(st[0]) == np.float32(want), (st, want)
    # (c) the full-resolution radii depend on +-32 px only: oracle on crops (corners / edges / centre)
    near = k.topousm_fast(d, radii=[2, 8, 32], weights=[4 / 7, 2 / 7, 1 / 7], workspace=ws)
    side, m = 768, 40
    for (r0, c0) in _crops(S, side):
        crop = _np(d[r0:r0 + side, c0:c0 + side])
        lo_r, hi_r = (0 if r0 == 0 else m), (side if r0 + side == S else side - m)
        lo_c, hi_c = (0 if c0 == 0 else m), (side if c0 + side == S else side - m)
        want = orc.topousm_fast_block(crop, radii=[2, 8, 32], weights=[4 / 7, 2 / 7, 1 / 7])[lo_r:hi_r, lo_c:hi_c]
        got = _np(near[r0 + lo_r:r0 + hi_r, c0 + lo_c:c0 + hi_c])
        assert np.array_equal(got, want), f"full-resolution radii, crop {r0},{c0}: max diff {np.abs(got - want).max():.3e}"
    del near
    # (d) two independent kernels (general kernel with re-reflection vs the streaming kernel), on three row
    #     bands of the full-size raster through the band entry point
    from fujishadergpu_b200.core.tile_processor import StreamedTopoPipeline
    pipe = StreamedTopoPipeline((S, S), {"radii": radii, "weights": w, "pixel_size": 1.0, "global_stats": (st[0],)},
                                output_dtype="float32", chunk_rows=4096)
    hin = torch.empty((S, S), dtype=torch.float32, pin_memory=True)
    hin.copy_(d)
    hout = torch.empty((S, S), dtype=torch.float32, pin_memory=True)
    pipe.run(hin, hout)
    whole = k.topousm_fast(d, radii=radii, weights=w, norm_scale=st[0], workspace=ws)
    for r0 in (0, 12288, 28672):     # (e) banded + streamed == whole raster, byte for byte
        assert _same_bits(hout[r0:r0 + 4096].cuda(), whole[r0:r0 + 4096]), f"banded vs whole at rows {r0}"
    del pipe, hin, hout, whole
    monkeypatch.setenv("FSG_FORCE_GENERIC", "1")
    sub = d[:, :4096].contiguous()   # the general kernel is slow: a full-height, 4096-wide slab
    gen = k.topousm_fast(sub, radii=radii, weights=w)
    monkeypatch.delenv("FSG_FORCE_GENERIC", raising=False)
    fast = k.topousm_fast(sub, radii=radii, weights=w)
    assert _same_bits(gen, fast), "general vs streaming kernel"


def test_config3_openness_32768_crop_oracle_and_flat():
    from fujishadergpu_b200 import kernels as k
    S, side, halo = 32768, 640, 256
    d = k.synth_dem((S, S), seed=20261020, nodata=True)
    out = k.openness(d, openness_type="positive", num_directions=8, max_distance=256, pixel_size=1.0)
    assert bool(torch.equal(torch.isnan(out), torch.isnan(d)))
    for (r0, c0) in [(0, 0), (S - side, S - side), ((S - side) // 2, (S - side) // 2), (0, (S - side) // 2)]:
        crop = _np(d[r0:r0 + side, c0:c0 + side])
        lo_r, hi_r = (0 if r0 == 0 else halo), (side if r0 + side == S else side - halo)
        lo_c, hi_c = (0 if c0 == 0 else halo), (side if c0 + side == S else side - halo)
        want = orc.openness_block(crop, openness_type="positive", num_directions=8, max_distance=256,
                                  pixel_size=1.0)[lo_r:hi_r, lo_c:hi_c]
        got = _np(out[r0 + lo_r:r0 + hi_r, c0 + lo_c:c0 + hi_c])
        assert_close_f32(got, want, what=f"openness crop {r0},{c0}")
    del out
    flat = torch.full((8192, 8192), 123.5, dtype=torch.float32, device="cuda")   # reference known answer: flat -> 1.0
    o = k.openness(flat, openness_type="positive", num_directions=8, max_distance=256, pixel_size=1.0)
    assert float((o - 1.0).abs().max()) <= 1e-3
