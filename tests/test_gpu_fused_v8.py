"""fused_kernel_v8 (interior fast path of topousm_fast) and the synchronisation-free step around it.

v8 only ever computes pixels whose whole neighbourhood is NaN-free and inside the raster; borders and NaN blocks go
to fused_kernel_v6.  The bar: the combination is bit-identical to v6 alone (which the golden / oracle tests of
tests/test_gpu_parity.py pin to the reference) and to the oracle on a raster small enough for SciPy.
Reference: compute_topousm_fast_efficient_block (algorithms/_impl_topousm_fast.py:49-100), the statistics pre-pass
(algorithms/_norm_stats.py:176-298), topousm_fast_stat_func (algorithms/_normalization.py:22-32).
"""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from conftest import assert_close_f32  # noqa: E402
from oracle import terrain_oracle as orc  # noqa: E402

W6 = [32 / 63, 16 / 63, 8 / 63, 4 / 63, 2 / 63, 1 / 63]
R6 = [2, 8, 32, 128, 512, 2048]
SWITCHES = ("FSG_FORCE_GENERIC", "FSG_FUSED_V5", "FSG_NO_BULK", "FSG_V6_CFGB", "FSG_NO_V8")


def _run(k, d, radii, w, no_v8, **kw):
    for key in SWITCHES:
        os.environ.pop(key, None)
    if no_v8:
        os.environ["FSG_NO_V8"] = "1"
    k.reload_debug_switches()
    try:
        o = k.topousm_fast(d, radii=radii, weights=w, **kw)
        torch.cuda.synchronize()
    finally:
        os.environ.pop("FSG_NO_V8", None)
        k.reload_debug_switches()
    return o


def _same_bits(a, b):
    if a.dtype != torch.float32:
        return bool(torch.equal(a, b))
    an, bn = torch.isnan(a), torch.isnan(b)
    za, zb = torch.where(an, torch.zeros_like(a), a), torch.where(bn, torch.zeros_like(b), b)
    return bool(torch.equal(an, bn)) and bool(torch.equal(za.view(torch.int32), zb.view(torch.int32)))


@pytest.mark.parametrize("shape,nodata", [((700, 5000), False), ((3000, 2900), True), ((4096, 4096), False),
                                          ((1024, 8256), False), ((5000, 1156), True)])
def test_v8_equals_v6_bit_for_bit(shape, nodata):
    from fujishadergpu_b200 import kernels as k
    d = k.synth_dem(shape, seed=31 + shape[0], nodata=nodata)
    if nodata:   # isolated NaNs and a NaN block in the interior: their 256-row blocks go back to v6
        d[shape[0] // 2, shape[1] // 3] = float("nan")
        d[shape[0] // 3: shape[0] // 3 + 40, shape[1] // 2: shape[1] // 2 + 70] = float("nan")
    for radii, w in ((R6, W6), ([2, 8, 32], [4 / 7, 2 / 7, 1 / 7]), ([2, 8, 32, 128], [0.4, 0.3, 0.2, 0.1])):
        for ns in (14.65, None):
            ref = _run(k, d, radii, w, True, norm_scale=ns)
            got = _run(k, d, radii, w, False, norm_scale=ns)
            assert _same_bits(ref, got), (shape, nodata, radii, ns)
    qp8 = {"a_coef": 107.99319, "b_coef": 128.0, "dn_min": 1, "dn_max": 255}
    qp16 = {"a_coef": 27863.095, "b_coef": 0.0, "dn_min": -32767, "dn_max": 32767}
    for od, qp in (("uint8", qp8), ("int16", qp16)):
        ref = _run(k, d, R6, W6, True, norm_scale=14.65, output_dtype=od, qp=qp)
        got = _run(k, d, R6, W6, False, norm_scale=14.65, output_dtype=od, qp=qp)
        assert _same_bits(ref, got), (shape, od)


def test_v8_region_of_interest_equals_full_call():
    """fsg_topousm_fast_roi (the call of the statistics pre-pass) == the full call on the kept region, dense and
    NoData, with and without the fast path."""
    from fujishadergpu_b200 import kernels as k
    for nodata in (False, True):
        d = k.synth_dem((3000, 4100), seed=77, nodata=nodata)
        full = _run(k, d, R6, W6, True, norm_scale=None)
        H, W = d.shape
        for roi in ((600, 1500, 414, 2050), (0, 1000, 2048, 2052), (1500, 1500, 0, 1366), (37, 2900, 33, 4000)):
            for no_v8 in (True, False):
                o = torch.zeros_like(full)
                _run(k, d, R6, W6, no_v8, norm_scale=None, roi=roi, out=o)
                sl = (slice(roi[0], roi[0] + roi[1]), slice(roi[2], roi[2] + roi[3]))
                assert _same_bits(full[sl], o[sl]), (nodata, roi, no_v8)


def test_v8_vs_oracle_whole_raster():
    """1536 x 1728: large enough for v8 strips and NaN-free interior blocks, small enough for the SciPy oracle."""
    from fujishadergpu_b200 import kernels as k
    d = k.synth_dem((1536, 1728), seed=5, nodata=False)
    radii, w = [2, 8, 32, 128], [0.4, 0.3, 0.2, 0.1]
    got = _run(k, d, radii, w, False, norm_scale=None).cpu().numpy()
    want = orc.topousm_fast_block(d.cpu().numpy(), radii=radii, weights=w)
    assert_close_f32(got / 14.0, want / 14.0, what="v8 vs oracle")   # bar on the normalised scale (p99 ~ 14 m)
    assert np.array_equal(got, want)                               # and in fact identical


def test_device_scale_equals_numpy_percentile():
    """fsg_select_finish_scale: the p99 stays on the device and equals np.percentile (f32, method 'linear')."""
    from fujishadergpu_b200 import kernels as k
    g = torch.Generator(device="cuda").manual_seed(3)
    for n in (1, 2, 101, 4097, 1_000_003):
        x = torch.randn(n, generator=g, device="cuda", dtype=torch.float32) * 7.0
        out = torch.empty(1, dtype=torch.float32, device="cuda")
        k.staged_percentile([x.view(1, -1)], 99.0, take_abs=True, finite_only=False, device=x.device, scale_out=out)
        want = np.percentile(np.abs(x.cpu().numpy()), 99.0)
        assert np.float32(out.item()) == np.float32(want), (n, out.item(), want)
    z = torch.zeros(5000, device="cuda")
    out = torch.empty(1, dtype=torch.float32, device="cuda")
    k.staged_percentile([z.view(1, -1)], 99.0, take_abs=True, finite_only=False, device=z.device, scale_out=out)
    assert np.isnan(out.item())   # <= 1e-9: "no global scale" (algorithms/_normalization.py:22-32)


def test_valid_bbox_kernel():
    from fujishadergpu_b200 import kernels as k
    d = torch.full((1000, 1200), float("nan"), device="cuda")
    d[130:777, 250:1111] = 1.0
    cov = 7
    n_rows, n_cols = 1000 // cov, 1200 // cov
    b = k.valid_bbox(d, 0, cov, n_rows, n_cols, 0).cpu().tolist()
    ov = np.isfinite(d.cpu().numpy()[::cov, ::cov][:n_rows, :n_cols])
    rows, cols = np.nonzero(ov.any(axis=1))[0], np.nonzero(ov.any(axis=0))[0]
    assert [-b[0], b[1], -b[2], b[3]] == [rows.min(), rows.max(), cols.min(), cols.max()]
    e = k.valid_bbox(torch.full((64, 64), float("nan"), device="cuda"), 0, 1, 64, 64, 0).cpu().tolist()
    assert e[1] < 0 and e[3] < 0


def test_sync_free_step_equals_sequential_calls():
    """topousm_fast_sharded_step (device-side scale, speculated planning values) == statistics pre-pass read back
    to the host + fsg_topousm_fast with that scale; the NoData raster forces one repeated step (bounding box)."""
    from fujishadergpu_b200 import kernels as k
    from fujishadergpu_b200.algorithms._norm_stats import compute_norm_stats_device
    from fujishadergpu_b200.core import sharding as sh
    radii, w = [2, 8, 32, 128], [0.4, 0.3, 0.2, 0.1]
    for nodata in (False, True):
        d = k.synth_dem((4096, 3072), seed=9, nodata=nodata)
        if nodata:
            d[:300] = float("nan")   # the valid-data bounding box is not the whole raster
        sh.Speculation._memory.clear()
        tries = 0
        for _ in range(3):
            res, scale_dev, spec = sh.topousm_fast_sharded_step(d, 4096, 0, 1, radii=radii, weights=w)
            tries += 1
            if spec.ok():
                break
        assert tries == (2 if nodata else 1)
        st = compute_norm_stats_device(d, "topousm_fast", {"radii": radii, "weights": w, "pixel_size": 1.0})
        assert np.float32(scale_dev.item()) == np.float32(st[0])
        want = k.topousm_fast(d, radii=radii, weights=w, norm_scale=float(st[0]))
        assert _same_bits(res, want)


def test_stats_prepass_nine_windows_vs_oracle():
    """6144^2 with radii <= 128: margin 144, tile 2048 -> the 3 x 3 grid yields nine distinct windows; the scale of
    the device pre-pass (regions of interest, v8 + v6) equals the oracle's window-by-window p99 (reference:
    _compute_norm_stats_tiled, algorithms/_norm_stats.py:176-298), and so does the synchronisation-free step."""
    from fujishadergpu_b200 import kernels as k
    from fujishadergpu_b200.algorithms._norm_stats import compute_norm_stats_device
    from fujishadergpu_b200.core import sharding as sh
    S = 6144
    radii, w = [2, 8, 32, 128], [0.4, 0.3, 0.2, 0.1]
    d = k.synth_dem((S, S), seed=12, nodata=False)
    dem = d.cpu().numpy()
    margin, tile = orc.stats_window_geometry("topousm_fast", {"radii": radii})
    wins = orc.stats_windows(S, S, 0, S, 0, S, grid=3, tile=min(tile, S))
    assert len(set(wins)) == 9 and tile == 2048
    pooled = []
    for (wy0, wx0, tw, th) in wins:
        r = orc.topousm_fast_block(dem[wy0:wy0 + th, wx0:wx0 + tw], radii=radii, weights=w)
        m = int(min(margin, th // 3, tw // 3))
        r = r[m:-m, m:-m]
        pooled.append(r[~np.isnan(r)])
    want = orc.abs_p99_scale(np.concatenate(pooled))[0]
    st = compute_norm_stats_device(d, "topousm_fast", {"radii": radii, "weights": w, "pixel_size": 1.0})
    assert np.float32(st[0]) == np.float32(want), (st, want)
    sh.Speculation._memory.clear()
    _res, scale_dev, spec = sh.topousm_fast_sharded_step(d, S, 0, 1, radii=radii, weights=w)
    assert spec.ok() and np.float32(scale_dev.item()) == np.float32(want)


def _grid_mean(k, g, size, two_pass, row_band=None):
    os.environ.pop("FSG_BOX_TWO_PASS", None)
    if two_pass:
        os.environ["FSG_BOX_TWO_PASS"] = "1"
    k.reload_debug_switches()
    try:
        gh = int(g.shape[0])
        if row_band is None:
            o = k.grid_mean_band(g, 0, gh, size, 0, gh)
        else:   # a row band whose source rows carry the halo the pass needs (mirrored rows included)
            a, b = row_band
            reach = size // 2
            lo, hi = max(0, a - reach), min(gh, b + reach)
            if a - reach < 0:
                hi = max(hi, min(gh, reach - a))
            if b + reach > gh:
                lo = min(lo, max(0, 2 * gh - (b + reach)))
            o = k.grid_mean_band(g[lo:hi], lo, gh, size, a, b - a)
        torch.cuda.synchronize()
    finally:
        os.environ.pop("FSG_BOX_TWO_PASS", None)
        k.reload_debug_switches()
    return o


@pytest.mark.parametrize("shape", [(2064, 2064), (517, 1031), (96, 70), (1200, 193), (33, 4100)])
def test_single_pass_box_mean_equals_two_passes(shape):
    """box_mean2d_kernel (decimated-grid means of the large radii, one pass) == box_axis0 + box_axis1 bit for bit:
    dense, isolated NaNs, a NaN block larger than the window, sizes of both kernel geometries, row bands."""
    from fujishadergpu_b200 import kernels as k
    gen = torch.Generator(device="cuda").manual_seed(shape[0] * 7 + shape[1])
    base = torch.rand(shape, generator=gen, device="cuda") * 900.0 + 100.0
    for variant in ("dense", "nan"):
        g = base.clone()
        if variant == "nan":
            g[shape[0] // 3, shape[1] // 2] = float("nan")
            g[shape[0] // 2: shape[0] // 2 + 90, shape[1] // 4: shape[1] // 4 + 300] = float("nan")
            g[::37, ::53] = float("nan")
        for size in (3, 33, 65, 129, 257):
            if size // 2 >= min(shape):   # mirrored index would leave the grid twice: not a case the plan produces
                continue
            want = _grid_mean(k, g, size, True)
            got = _grid_mean(k, g, size, False)
            assert _same_bits(want, got), (shape, variant, size)
            if shape[0] > 300:
                for band in ((0, 160), (shape[0] // 2 - 7, shape[0] // 2 + 150), (shape[0] - 131, shape[0])):
                    gb = _grid_mean(k, g, size, False, row_band=band)
                    assert _same_bits(want[band[0]:band[1]], gb), (shape, variant, size, band)


def test_single_pass_box_mean_vs_scipy():
    """... and against the reference's own arithmetic: scipy.ndimage.uniform_filter(mode='reflect') on f32
    (algorithms/_nan_utils.py:18-47, handle_nan_with_uniform: U(filled) / U(valid))."""
    from scipy import ndimage
    from fujishadergpu_b200 import kernels as k
    rng = np.random.default_rng(4)
    g = (rng.random((700, 900), dtype=np.float32) * 900 + 100).astype(np.float32)
    for size in (65, 257):
        got = _grid_mean(k, torch.from_numpy(g).cuda(), size, False).cpu().numpy()
        want = ndimage.uniform_filter(g, size=size, mode="reflect")
        assert np.array_equal(got, want), size
    gn = g.copy()
    gn[300:420, 200:520] = np.nan
    gn[::41, ::29] = np.nan
    valid = (~np.isnan(gn)).astype(np.float32)
    filled = np.where(np.isnan(gn), np.float32(0), gn)
    for size in (65, 257):
        got = _grid_mean(k, torch.from_numpy(gn).cuda(), size, False).cpu().numpy()
        num = ndimage.uniform_filter(filled, size=size, mode="reflect")
        den = ndimage.uniform_filter(valid, size=size, mode="reflect")
        want = np.where(den > 0, num / np.where(den > 0, den, 1), np.float32(0)).astype(np.float32)
        assert np.array_equal(got, want), size
