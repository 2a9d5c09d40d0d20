"""the COG container (io/cog_writer.py), the GeoTIFF reader (io/geotiff_reader.py) and the GPU
overview cascade.  GDAL is not in the image: files are read back with the package's reader and with Pillow's
libtiff as an independent decoder; the overview arithmetic is held to the oracle's NumPy restatement."""
import os
import struct
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import terrain_oracle as orc  # noqa: E402
from fujishadergpu_b200.io import cog_writer as cw  # noqa: E402
from fujishadergpu_b200.io.geotiff_reader import read_geotiff  # noqa: E402


def _levels(a, nodata, n):
    out = [a]
    for _ in range(n):
        out.append(orc.overview_average_2x(out[-1], nodata))
    return out


def _sample(dt, shape=(1300, 1100), seed=0):
    rng = np.random.default_rng(seed)
    base = rng.standard_normal(shape).cumsum(axis=1) * 3
    if dt == "uint8":
        a = (np.abs(base) % 254 + 1).astype(np.uint8)
        a[100:140, 200:320] = 0
    elif dt == "int16":
        a = (base * 100).astype(np.int16)
        a[a == 0] = 1
        a[100:140, 200:320] = 0
    else:
        a = base.astype(np.float32)
        a[100:140, 200:320] = np.nan
    return a


@pytest.mark.parametrize("dt", ["uint8", "int16", "float32"])
def test_pyramid_round_trip_and_layout(tmp_path, dt):
    a = _sample(dt)
    nod = float("nan") if dt == "float32" else 0
    lv = _levels(a, None if dt == "float32" else 0, 3)
    p = str(tmp_path / f"{dt}.tif")
    gt = (500000.0, 0.5, 0.0, 4100000.0, 0.0, -0.5)
    st = cw.write_tiff_pyramid(p, lv, nodata=nod, transform=gt, epsg=6677)
    assert st["predictor"] == {"uint8": 1, "int16": 2, "float32": 3}[dt] and st["blocksize"] == 512 and st["bigtiff"]
    for li, want in enumerate(lv):
        got, meta = read_geotiff(p, li)
        assert got.dtype == want.dtype and np.array_equal(got, want, equal_nan=True), (dt, li)
    assert meta["levels"] == 4 and meta["newsubfiletype"] == [0, 1, 1, 1]
    assert meta["compression"] == 50000 and meta["tile"] == (512, 512) and meta["epsg"] == 6677
    assert meta["transform"] == gt
    assert (meta["nodata"] != meta["nodata"]) if dt == "float32" else meta["nodata"] == 0.0
    # window read == slice
    win, _ = read_geotiff(p, 0, window=(400, 300, 500, 333))
    assert np.array_equal(win, a[400:700, 500:833], equal_nan=True)
    # COG layout: every IFD and tile index before the first tile; overview data before full-resolution data
    with open(p, "rb") as fh:
        from fujishadergpu_b200.io.geotiff_reader import _read_ifds
        ifds, big = _read_ifds(fh)
    assert big
    first_data = min(min(i[324]) for i in ifds)
    with open(p, "rb") as fh:
        head = fh.read(16)
    ifd0 = struct.unpack("<Q", head[8:16])[0]
    assert ifd0 == (16 + len(cw.GDAL_GHOST) + 7) // 8 * 8 and first_data > ifd0
    assert max(ifds[3][324]) < min(ifds[2][324]) and max(ifds[1][324]) < min(ifds[0][324])
    # GDAL's structural metadata right after the header, block leaders / trailers, and the validator's verdict
    from fujishadergpu_b200.io.cog_validator import inspect_cog, validate_cog
    with open(p, "rb") as fh:
        fh.seek(16)
        assert fh.read(43) == b"GDAL_STRUCTURAL_METADATA_SIZE=000140 bytes\n"
    info = inspect_cog(p)
    assert info["tiled"] and info["ghost"] and info["leaders_ok"] and info["ifds_before_data"] and info["overview_data_first"]
    assert info["overviews"] == 3 and info["block"] == (512, 512) and info["score"] == 90 and validate_cog(p)
    plain = str(tmp_path / f"{dt}_plain.tif")
    cw.write_tiff_pyramid(plain, lv[:1], nodata=nod, gdal_ghost=False)
    assert not validate_cog(plain) and inspect_cog(plain)["score"] == 40


def test_pillow_decodes_the_same_pixels(tmp_path):
    Image = pytest.importorskip("PIL.Image")
    Image.MAX_IMAGE_PIXELS = None
    for dt in ("uint8", "float32"):
        a = _sample(dt, (900, 700), seed=3)
        lv = _levels(a, None if dt == "float32" else 0, 2)
        p = str(tmp_path / f"pil_{dt}.tif")
        cw.write_tiff_pyramid(p, lv, nodata=float("nan") if dt == "float32" else 0)
        try:
            im = Image.open(p)
            frames = []
            for i in range(3):
                im.seek(i)
                frames.append(np.array(im))
        except Exception as exc:   # a Pillow build without BigTIFF / ZSTD
            pytest.skip(f"Pillow cannot decode this file: {exc!r}")
        for got, want in zip(frames, lv):
            assert np.array_equal(got.astype(want.dtype), want, equal_nan=True), dt


def test_reader_handles_strips_deflate_and_nodata_mask(tmp_path):
    """A classic stripped DEFLATE GeoTIFF with a numeric NoData value (the input side: _build_nodata_mask)."""
    Image = pytest.importorskip("PIL.Image")
    a = _sample("float32", (300, 257), seed=5)
    a[np.isnan(a)] = -9999.0
    p = str(tmp_path / "in.tif")
    try:
        Image.fromarray(a).save(p, compression="tiff_adobe_deflate", tiffinfo={42113: "-9999"})
    except Exception as exc:
        pytest.skip(f"Pillow cannot write the fixture: {exc!r}")
    got, meta = read_geotiff(p, nodata_to_nan=True)
    assert meta["nodata"] == -9999.0 and meta["tile"] is None
    want = a.copy()
    want[a == -9999.0] = np.nan
    assert np.array_equal(got, want, equal_nan=True)


def test_small_and_ragged_rasters_round_trip(tmp_path):
    for shape in ((1, 1), (3, 700), (513, 5), (512, 512)):
        a = _sample("int16", shape, seed=shape[0])
        a[0, 0] = 7
        n = cw.overview_levels(shape)
        lv = _levels(a, 0, n)
        p = str(tmp_path / f"s_{shape[0]}_{shape[1]}.tif")
        cw.write_tiff_pyramid(p, lv, nodata=0)
        for li, want in enumerate(lv):
            got, meta = read_geotiff(p, li)
            assert got.shape == want.shape and np.array_equal(got, want), (shape, li)
        assert meta["levels"] == n + 1


def test_overview_oracle_nodata_rules():
    a = np.array([[1, -1, 5, 0], [1, -1, 0, 0], [-1, -1, 0, 0], [1, 0, 0, 0]], np.int16)
    o = orc.overview_average_2x(a, 0)
    assert o.tolist() == [[1, 5], [-1, 0]]       # mean 0 collides with NoData -> +1; (-1,-1,1)/3 rounds to 0 -> -1; empty -> 0
    f = np.array([[1.0, np.nan], [np.nan, np.nan], [2.0, 4.0]], np.float32)
    assert np.array_equal(orc.overview_average_2x(f), np.array([[1.0], [3.0]], np.float32))
    u = np.array([[1, 2, 255]], np.uint8)
    assert orc.overview_average_2x(u, 0).tolist() == [[2, 255]]       # (1 + 2) / 2 = 1.5 -> 2 (half up)


def test_overview_level_count():
    assert cw.overview_levels((65536, 65536)) == 8
    assert cw.overview_levels((100, 3)) == 7
    assert cw.overview_levels((1, 1)) == 0
    assert cw.predictor_for_dtype("float32") == 3 and cw.predictor_for_dtype("int16") == 2 and cw.predictor_for_dtype("uint8") == 1


@pytest.mark.gpu
@pytest.mark.parametrize("dt", ["uint8", "int16", "float32"])
def test_write_cog_from_device_matches_oracle_pyramid(tmp_path, dt):
    torch = pytest.importorskip("torch")
    a = _sample(dt, (2051, 1797), seed=9)
    p = str(tmp_path / f"cog_{dt}.tif")
    st = cw.write_cog(p, torch.from_numpy(a).cuda(), transform=(0.0, 1.0, 0.0, 0.0, 0.0, -1.0), epsg=6677)
    n = cw.overview_levels(a.shape)
    assert n == 8 and len(st["levels"]) == 9
    lv = _levels(a, None if dt == "float32" else 0, n)
    for li, want in enumerate(lv):
        got, meta = read_geotiff(p, li)
        assert got.shape == want.shape and np.array_equal(got, want, equal_nan=True), (dt, li)
    assert meta["compression"] == 50000 and meta["bigtiff"]


@pytest.mark.gpu
def test_overview_kernel_nodata_collisions():
    torch = pytest.importorskip("torch")
    from fujishadergpu_b200 import kernels as k
    rng = np.random.default_rng(2)
    a = rng.integers(-2, 3, size=(301, 257)).astype(np.int16)       # many zeros (NoData) and means that round to 0
    got = k.overview_average(torch.from_numpy(a).cuda(), 0).cpu().numpy()
    assert np.array_equal(got, orc.overview_average_2x(a, 0))
    f = rng.standard_normal((301, 257)).astype(np.float32)
    f[rng.random(f.shape) < 0.4] = np.nan
    gotf = k.overview_average(torch.from_numpy(f).cuda()).cpu().numpy()
    assert np.array_equal(gotf, orc.overview_average_2x(f), equal_nan=True)


@pytest.mark.gpu
def test_topousm_uint8_to_cog_end_to_end(tmp_path):
    """BASELINE config 5's output side at a small size: fused uint8 encoding on the GPU -> COG -> decoded DN
    identical to the oracle's encoding of the oracle's normalised block (+-1 DN, NoData = 0 identical)."""
    torch = pytest.importorskip("torch")
    from fujishadergpu_b200 import kernels as k
    from fujishadergpu_b200.io.output_encoding import quantize_params, resolve_output_range
    dem = orc.synth_dem(1500, 1300, seed=21, nodata=True)
    radii, w = [2, 8, 32, 128], orc.pow2_weights(4)
    raw = orc.topousm_fast_block(dem, radii=radii, weights=w)
    scale = orc.abs_p99_scale(raw)[0]
    qp = quantize_params(*resolve_output_range("topousm_fast"), "uint8")
    want = orc.encode_array(orc.normalise_by_scale(raw.copy(), (scale,)), qp, "uint8")
    got_dev = k.topousm_fast(torch.from_numpy(dem).cuda(), radii=radii, weights=w, norm_scale=scale, output_dtype="uint8", qp=qp)
    p = str(tmp_path / "topo.tif")
    cw.write_cog(p, got_dev)
    got, meta = read_geotiff(p)
    assert np.array_equal(got == 0, want == 0)
    assert np.abs(got.astype(np.int32) - want.astype(np.int32)).max() <= 1
    assert meta["nodata"] == 0.0 and meta["levels"] == 9


@pytest.mark.gpu
@pytest.mark.parametrize("world", [1, 2, 4])
def test_sharded_cog_writer_equals_single_writer(tmp_path, world):
    """io/cog_sharded.write_cog_sharded (config 5: every rank compresses the tiles of its own band, distributed
    overview levels, gathered small levels) writes byte for byte the file of io/cog_writer.write_cog.  The ranks are
    played one after the other by a stand-in `dist` (real NCCL: tests/test_gpu_sharding.py, profiles/tools/run_config5.py)."""
    torch = pytest.importorskip("torch")
    import threading
    from fujishadergpu_b200.io.cog_sharded import tile_aligned_bounds, write_cog_sharded
    from fujishadergpu_b200.io.cog_writer import write_cog
    H, W = 5 * 1024 + 300, 2100
    g = torch.Generator(device="cuda").manual_seed(5)
    full = (torch.rand((H, W), generator=g, device="cuda") * 255).to(torch.uint8)
    full[1000:1300, 500:900] = 0
    ref = str(tmp_path / "ref.tif")
    write_cog(ref, full)
    out = str(tmp_path / f"sharded{world}.tif")
    own = tile_aligned_bounds(H, world)

    class FakeDist:   # threads play the ranks; all_gather_object / barrier through a shared table
        def __init__(self):
            self.lock = threading.Condition()
            self.round = {}
            self.tl = threading.local()

        def _collect(self, key, value):
            with self.lock:
                slot = self.round.setdefault(key, {})
                slot[self.tl.rank] = value
                self.lock.notify_all()
                self.lock.wait_for(lambda: len(self.round[key]) == world)
                return [self.round[key][q] for q in range(world)]

        def all_gather_object(self, out_list, obj):
            self.tl.n = getattr(self.tl, "n", 0) + 1
            out_list[:] = self._collect(("g", self.tl.n), obj)

        def barrier(self):
            self.tl.n = getattr(self.tl, "n", 0) + 1
            self._collect(("b", self.tl.n), None)

    fd = FakeDist()
    errs = []

    def run(rank):
        try:
            fd.tl.rank = rank
            torch.cuda.set_device(0)
            a, b = own[rank]
            write_cog_sharded(out, full[a:b].contiguous(), H, rank, world, dist=fd if world > 1 else None)
        except Exception as exc:   # pragma: no cover
            errs.append(exc)
            with fd.lock:
                fd.lock.notify_all()

    threads = [threading.Thread(target=run, args=(q,)) for q in range(world)]
    for th in threads:
        th.start()
    for th in threads:
        th.join(timeout=300)
    assert not errs, errs
    assert open(out, "rb").read() == open(ref, "rb").read()


@pytest.mark.parametrize("epsg,geographic", [(4326, True), (6668, True), (6318, True), (7844, True), (6677, False),
                                             (32654, False), (4087, False), (4978, False)])
def test_geokeys_follow_the_shared_geographic_rule(tmp_path, epsg, geographic):
    """The writer tags GTModelType / Geographic- vs ProjectedCSType with the same rule that picked the degree ->
    metre scaling on the way in (io/raster_info.is_geographic_epsg): JGD2011 (6668) is geographic."""
    from fujishadergpu_b200.io.cog_writer import write_tiff_pyramid
    from fujishadergpu_b200.io.raster_info import is_geographic_epsg
    from fujishadergpu_b200.io import geotiff_reader as gr
    assert is_geographic_epsg(epsg) is geographic
    p = str(tmp_path / "g.tif")
    write_tiff_pyramid(p, [np.zeros((64, 64), dtype=np.uint8)], nodata=0, transform=(139.0, 1e-4, 0.0, 36.0, 0.0, -1e-4),
                       epsg=epsg)
    with open(p, "rb") as fh:
        ifds, _big = gr._read_ifds(fh)
    keys = list(ifds[0][34735])
    entries = {keys[i]: keys[i + 3] for i in range(4, len(keys), 4)}
    assert entries[1024] == (2 if geographic else 1)
    assert entries[2048 if geographic else 3072] == epsg and (3072 if geographic else 2048) not in entries
    _a, meta = gr.read_geotiff(p)
    assert meta["epsg"] == epsg
