"""Tiled GeoTIFF / COG container for the encoded output raster.

Reference: get_cog_options (core/dask_processor.py:201-228) and _create_cog_ultra_fast_impl
(io/cog_builder.py:295-310) hand the raster to GDAL's COG driver with COMPRESS=ZSTD LEVEL=1 BLOCKSIZE=512
BIGTIFF=YES, PREDICTOR 3 for float32 / 2 for int16 / none for uint8, OVERVIEW_RESAMPLING=AVERAGE
OVERVIEW_COUNT=8, NoData 0 (integers) or NaN.  GDAL is not in this image, so this module writes that file
itself: the overview levels come from the GPU (kernels.overview_average, a 2 x 2 valid-mean cascade), the tiles
are predicted with NumPy and compressed with libzstd (ctypes, one thread per tile row) and the container is
laid out the way a COG reader expects -- header, every IFD with its tile index first, then the tile data from
the smallest overview to the full-resolution level, with GDAL's structural-metadata "ghost area" after the header
and its block leaders / trailers, so that GDAL reports LAYOUT=COG.  Parity with GDAL's own AVERAGE kernel and its
reading of the ghost area are unpinned (no GDAL here); the container is checked by reading it back with this module's
reader, with io/cog_validator.py and with Pillow's libtiff (tests/test_cog_io.py).
"""
from __future__ import annotations

import ctypes as C
import ctypes.util
import os
import struct
import zlib
from concurrent.futures import ThreadPoolExecutor
from typing import List, Optional, Sequence

import numpy as np

TAG_NEWSUBFILE, TAG_WIDTH, TAG_LENGTH, TAG_BITS, TAG_COMPRESSION, TAG_PHOTOMETRIC = 254, 256, 257, 258, 259, 262
TAG_SPP, TAG_PLANAR, TAG_PREDICTOR, TAG_TILEW, TAG_TILEL, TAG_TILEOFFS, TAG_TILECOUNTS = 277, 284, 317, 322, 323, 324, 325
TAG_SAMPLEFORMAT, TAG_PIXELSCALE, TAG_TIEPOINT, TAG_GEOKEYS, TAG_GDAL_NODATA = 339, 33550, 33922, 34735, 42113
COMPRESSION_NONE, COMPRESSION_DEFLATE, COMPRESSION_ZSTD = 1, 8, 50000
T_BYTE, T_ASCII, T_SHORT, T_LONG, T_DOUBLE, T_LONG8 = 1, 2, 3, 4, 12, 16
_TYPE_FMT = {T_BYTE: "B", T_ASCII: "s", T_SHORT: "H", T_LONG: "I", T_DOUBLE: "d", T_LONG8: "Q"}
_TYPE_SIZE = {T_BYTE: 1, T_ASCII: 1, T_SHORT: 2, T_LONG: 4, T_DOUBLE: 8, T_LONG8: 8}

# GDAL's "ghost area" right after the TIFF header: how its COG driver declares the layout (IFDs before the tile
# data, row-major blocks, every block preceded by its size as uint32 and followed by a repeat of its last 4 bytes).
# A reader that finds it reports LAYOUT=COG (what the reference's io/cog_validator.py:75-82 checks).
GDAL_GHOST_BODY = (b"LAYOUT=IFDS_BEFORE_DATA\nBLOCK_ORDER=ROW_MAJOR\nBLOCK_LEADER=SIZE_AS_UINT4\n"
                   b"BLOCK_TRAILER=LAST_4_BYTES_REPEATED\nKNOWN_INCOMPATIBLE_EDITION=NO\n ")
GDAL_GHOST = b"GDAL_STRUCTURAL_METADATA_SIZE=%06d bytes\n" % len(GDAL_GHOST_BODY) + GDAL_GHOST_BODY

_zstd = None


def _load_zstd():
    global _zstd
    if _zstd is None:
        name = ctypes.util.find_library("zstd") or "libzstd.so.1"
        lib = C.CDLL(name)
        lib.ZSTD_compressBound.restype = C.c_size_t
        lib.ZSTD_compressBound.argtypes = [C.c_size_t]
        lib.ZSTD_compress.restype = C.c_size_t
        lib.ZSTD_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
        lib.ZSTD_decompress.restype = C.c_size_t
        lib.ZSTD_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        lib.ZSTD_isError.restype = C.c_uint
        lib.ZSTD_isError.argtypes = [C.c_size_t]
        _zstd = lib
    return _zstd


def zstd_compress(data: bytes, level: int = 1) -> bytes:
    lib = _load_zstd()
    cap = lib.ZSTD_compressBound(len(data))
    dst = C.create_string_buffer(cap)
    n = lib.ZSTD_compress(dst, cap, data, len(data), int(level))
    if lib.ZSTD_isError(n):
        raise RuntimeError("ZSTD_compress failed")
    return dst.raw[:n]


def zstd_decompress(data: bytes, size: int) -> bytes:
    lib = _load_zstd()
    dst = C.create_string_buffer(size)
    n = lib.ZSTD_decompress(dst, size, data, len(data))
    if lib.ZSTD_isError(n) or n != size:
        raise RuntimeError("ZSTD_decompress failed")
    return dst.raw


def predictor_for_dtype(dtype) -> int:
    """get_cog_options (:221-226): 3 for float32, 2 for 16/32-bit integers, none for uint8."""
    dt = np.dtype(dtype)
    if dt.kind == "f":
        return 3
    if dt.kind in "iu" and dt.itemsize >= 2:
        return 2
    return 1


def apply_predictor(tile: np.ndarray, predictor: int) -> bytes:
    """TIFF predictor on one tile (rows independent): 2 = horizontal differencing of the samples,
    3 = floating-point predictor (bytes of a row regrouped most-significant plane first, then byte differences)."""
    if predictor == 1:
        return np.ascontiguousarray(tile).tobytes()
    if predictor == 2:
        d = tile.copy()
        d[:, 1:] = tile[:, 1:] - tile[:, :-1]          # wraps modulo 2^bits
        return d.tobytes()
    if predictor == 3:
        h, w = tile.shape
        n = tile.dtype.itemsize
        b = np.ascontiguousarray(tile).view(np.uint8).reshape(h, w, n)
        planes = np.ascontiguousarray(b[:, :, ::-1].transpose(0, 2, 1)).reshape(h, w * n)   # little-endian host
        d = planes.copy()
        d[:, 1:] = planes[:, 1:] - planes[:, :-1]
        return d.tobytes()
    raise ValueError(f"unsupported predictor {predictor}")


def undo_predictor(raw: bytes, shape, dtype, predictor: int) -> np.ndarray:
    h, w = shape
    dt = np.dtype(dtype)
    if predictor == 1:
        return np.frombuffer(raw, dtype=dt).reshape(h, w).copy()
    if predictor == 2:
        d = np.frombuffer(raw, dtype=dt).reshape(h, w)
        return np.cumsum(d, axis=1, dtype=dt)
    if predictor == 3:
        n = dt.itemsize
        d = np.frombuffer(raw, dtype=np.uint8).reshape(h, w * n)
        planes = np.cumsum(d, axis=1, dtype=np.uint8).reshape(h, n, w)
        b = np.ascontiguousarray(planes.transpose(0, 2, 1)[:, :, ::-1])
        return b.view(dt).reshape(h, w).copy()
    raise ValueError(f"unsupported predictor {predictor}")


class _Ifd:
    """One image file directory: fixed-size entries + out-of-line values, placed before any tile data."""

    def __init__(self, big: bool):
        self.big = big
        self.entries = []   # (tag, type, count, payload bytes)

    def add(self, tag: int, typ: int, values) -> None:
        if typ == T_ASCII:
            payload = values.encode("ascii") + b"\0"
            count = len(payload)
        else:
            vals = list(values) if isinstance(values, (list, tuple, np.ndarray)) else [values]
            count = len(vals)
            payload = struct.pack("<" + _TYPE_FMT[typ] * count, *vals)
        self.entries.append((tag, typ, count, payload))

    def reserve(self, tag: int, typ: int, count: int) -> None:
        self.entries.append((tag, typ, count, b"\0" * (count * _TYPE_SIZE[typ])))

    def set_payload(self, tag: int, typ: int, values) -> None:
        for i, e in enumerate(self.entries):
            if e[0] == tag:
                self.entries[i] = (tag, typ, len(values), struct.pack("<" + _TYPE_FMT[typ] * len(values), *values))
                return
        raise KeyError(tag)

    def size(self) -> int:
        inline = 8 if self.big else 4
        head = (8 + 20 * len(self.entries) + 8) if self.big else (2 + 12 * len(self.entries) + 4)
        extra = sum((len(p) + 1) // 2 * 2 for (_t, _ty, _c, p) in self.entries if len(p) > inline)
        return head + extra

    def serialise(self, offset: int, next_offset: int) -> bytes:
        inline = 8 if self.big else 4
        ents = sorted(self.entries, key=lambda e: e[0])
        head = (8 + 20 * len(ents) + 8) if self.big else (2 + 12 * len(ents) + 4)
        out = bytearray()
        extra = bytearray()
        out += struct.pack("<Q", len(ents)) if self.big else struct.pack("<H", len(ents))
        for tag, typ, count, payload in ents:
            out += struct.pack("<HH", tag, typ) + (struct.pack("<Q", count) if self.big else struct.pack("<I", count))
            if len(payload) <= inline:
                out += payload + b"\0" * (inline - len(payload))
            else:
                pos = offset + head + len(extra)
                out += struct.pack("<Q", pos) if self.big else struct.pack("<I", pos)
                extra += payload + (b"\0" if len(payload) % 2 else b"")
        out += struct.pack("<Q", next_offset) if self.big else struct.pack("<I", next_offset)
        return bytes(out + extra)


def _geo_entries(ifd: _Ifd, transform, epsg) -> None:
    """GeoTIFF georeferencing of a north-up raster: GDAL geotransform (x0, dx, 0, y0, 0, dy) + EPSG code."""
    if transform is not None:
        x0, dx, rx, y0, ry, dy = [float(v) for v in transform]
        if rx != 0.0 or ry != 0.0:
            raise ValueError("rotated geotransforms are not supported")
        ifd.add(TAG_PIXELSCALE, T_DOUBLE, [abs(dx), abs(dy), 0.0])
        ifd.add(TAG_TIEPOINT, T_DOUBLE, [0.0, 0.0, 0.0, x0, y0, 0.0])
    if epsg is not None:
        from .raster_info import is_geographic_epsg
        code = int(epsg)
        geographic = is_geographic_epsg(code)      # the SAME rule that chose degree -> metre scaling on the way in
        keys = [1, 1, 0, 3,
                1024, 0, 1, 2 if geographic else 1,       # GTModelTypeGeoKey
                1025, 0, 1, 1,                            # GTRasterTypeGeoKey: PixelIsArea
                2048 if geographic else 3072, 0, 1, code]  # GeographicTypeGeoKey / ProjectedCSTypeGeoKey
        ifd.add(TAG_GEOKEYS, T_SHORT, keys)


def _nodata_text(nodata) -> str:
    v = float(nodata)
    if v != v:
        return "nan"
    return str(int(v)) if v.is_integer() else repr(v)


def _sample_format(dt: np.dtype) -> int:
    return {"u": 1, "i": 2, "f": 3}[dt.kind]


def _make_ifds(shapes, dt, comp, pred, bs, bigtiff, nodata, transform, epsg, gdal_ghost):
    """IFDs of a pyramid (level 0 first) with reserved tile tables -> (ifds, [(h, w, tiles_y, tiles_x)], IFD offsets,
    offset of the first tile byte)."""
    ifds: List[_Ifd] = []
    grids = []
    for li, (h, w) in enumerate(shapes):
        ty, tx = (h + bs - 1) // bs, (w + bs - 1) // bs
        grids.append((h, w, ty, tx))
        ifd = _Ifd(bigtiff)
        if li > 0:
            ifd.add(TAG_NEWSUBFILE, T_LONG, 1)
        ifd.add(TAG_WIDTH, T_LONG, w)
        ifd.add(TAG_LENGTH, T_LONG, h)
        ifd.add(TAG_BITS, T_SHORT, dt.itemsize * 8)
        ifd.add(TAG_COMPRESSION, T_SHORT, comp)
        ifd.add(TAG_PHOTOMETRIC, T_SHORT, 1)
        ifd.add(TAG_SPP, T_SHORT, 1)
        ifd.add(TAG_PLANAR, T_SHORT, 1)
        if pred != 1:
            ifd.add(TAG_PREDICTOR, T_SHORT, pred)
        ifd.add(TAG_TILEW, T_SHORT, bs)
        ifd.add(TAG_TILEL, T_SHORT, bs)
        off_t = T_LONG8 if bigtiff else T_LONG
        ifd.reserve(TAG_TILEOFFS, off_t, ty * tx)
        ifd.reserve(TAG_TILECOUNTS, off_t, ty * tx)
        ifd.add(TAG_SAMPLEFORMAT, T_SHORT, _sample_format(dt))
        if li == 0:
            _geo_entries(ifd, transform, epsg)
        if nodata is not None:
            ifd.add(TAG_GDAL_NODATA, T_ASCII, _nodata_text(nodata))
        ifds.append(ifd)
    header = (16 if bigtiff else 8) + (len(GDAL_GHOST) if gdal_ghost else 0)
    header = (header + 7) // 8 * 8
    ifd_off = []
    pos = header
    for ifd in ifds:
        ifd_off.append(pos)
        pos += ifd.size()
        pos = (pos + 15) // 16 * 16
    return ifds, grids, ifd_off, pos


def _write_header_and_ifds(fh, ifds, ifd_off, bigtiff, gdal_ghost) -> None:
    fh.seek(0)
    fh.write(struct.pack("<2sHHHQ", b"II", 43, 8, 0, ifd_off[0]) if bigtiff else struct.pack("<2sHI", b"II", 42, ifd_off[0]))
    if gdal_ghost:
        fh.write(GDAL_GHOST)
    for i, ifd in enumerate(ifds):
        fh.seek(ifd_off[i])
        fh.write(ifd.serialise(ifd_off[i], ifd_off[i + 1] if i + 1 < len(ifds) else 0))


def encode_tile_rows(band: np.ndarray, dt, bs: int, fill, pred: int, comp: int, level: int, pool) -> list:
    """Compressed blobs of the tiles of `band` (a whole number of tile rows, except the last row of the raster), row
    major; edge tiles are padded to full blocks."""
    def enc(tile):
        raw = apply_predictor(tile, pred)
        if comp == COMPRESSION_ZSTD:
            return zstd_compress(raw, level)
        if comp == COMPRESSION_DEFLATE:
            return zlib.compress(raw, max(1, min(9, int(level))))
        return raw

    h, w = band.shape
    tiles = []
    for r0 in range(0, h, bs):
        for c0 in range(0, w, bs):
            t = band[r0:r0 + bs, c0:c0 + bs]
            if t.shape != (bs, bs):
                full = np.full((bs, bs), fill, dtype=dt)
                full[: t.shape[0], : t.shape[1]] = t
                t = full
            tiles.append(np.ascontiguousarray(t))
    return list(pool.map(enc, tiles))


def write_tiff_pyramid(path: str, levels: Sequence, *, nodata=None, transform=None, epsg=None, blocksize: int = 512,
                       compress: str = "zstd", level: int = 1, bigtiff: bool = True, num_threads: Optional[int] = None,
                       row_reader=None, gdal_ghost: bool = True) -> dict:
    """Write levels[0] (full resolution) and levels[1:] (overviews, each a NumPy array or any object with .shape /
    .dtype whose rows `row_reader(level_obj, r0, r1)` returns as a NumPy array) as one tiled (Big)TIFF."""
    comp = {"zstd": COMPRESSION_ZSTD, "deflate": COMPRESSION_DEFLATE, "none": COMPRESSION_NONE}[compress]
    if row_reader is None:
        row_reader = lambda a, r0, r1: np.asarray(a[r0:r1])
    dt = np.dtype(str(levels[0].dtype).replace("torch.", ""))
    if dt not in (np.dtype("uint8"), np.dtype("int16"), np.dtype("float32")):
        raise ValueError(f"unsupported sample type {dt}")
    pred = predictor_for_dtype(dt) if comp != COMPRESSION_NONE else 1
    bs = int(blocksize)
    fill = np.nan if dt.kind == "f" else (0 if nodata is None else nodata)
    ifds, grids, ifd_off, data_start = _make_ifds([(int(lv.shape[0]), int(lv.shape[1])) for lv in levels], dt, comp, pred,
                                                  bs, bigtiff, nodata, transform, epsg, gdal_ghost)

    def encode_tile(tile: np.ndarray) -> bytes:
        raw = apply_predictor(tile, pred)
        if comp == COMPRESSION_ZSTD:
            return zstd_compress(raw, level)
        if comp == COMPRESSION_DEFLATE:
            return zlib.compress(raw, max(1, min(9, int(level))))
        return raw

    nthreads = int(num_threads or min(32, (os.cpu_count() or 1)))
    stats = {"levels": [], "bytes": 0}
    with open(path, "wb") as fh, ThreadPoolExecutor(max_workers=nthreads) as pool:
        fh.write(b"\0" * data_start)
        cur = data_start
        for li in range(len(levels) - 1, -1, -1):          # smallest overview first, full resolution last
            h, w, ty, tx = grids[li]
            offs, cnts = [], []
            for r in range(ty):
                r0, r1 = r * bs, min(h, (r + 1) * bs)
                band = row_reader(levels[li], r0, r1)
                if band.dtype != dt:
                    raise ValueError("all levels must have the dtype of level 0")
                tiles = []
                for c in range(tx):
                    c0, c1 = c * bs, min(w, (c + 1) * bs)
                    t = band[:, c0:c1]
                    if t.shape != (bs, bs):                # edge tiles are padded to full blocks
                        full = np.full((bs, bs), fill, dtype=dt)
                        full[: t.shape[0], : t.shape[1]] = t
                        t = full
                    tiles.append(np.ascontiguousarray(t))
                for blob in pool.map(encode_tile, tiles):
                    if gdal_ghost:       # leader: size as uint32; trailer: the last 4 bytes once more
                        fh.write(struct.pack("<I", len(blob)))
                        cur += 4
                    fh.write(blob)
                    offs.append(cur)
                    cnts.append(len(blob))
                    cur += len(blob)
                    if gdal_ghost:
                        fh.write(blob[-4:])
                        cur += len(blob[-4:])
            off_t = T_LONG8 if bigtiff else T_LONG
            ifds[li].set_payload(TAG_TILEOFFS, off_t, offs)
            ifds[li].set_payload(TAG_TILECOUNTS, off_t, cnts)
            stats["levels"].append({"level": li, "shape": (h, w), "tiles": ty * tx, "bytes": int(sum(cnts))})
        if not bigtiff and cur >= 2 ** 32:
            raise ValueError("file exceeds 4 GiB: use bigtiff=True")
        _write_header_and_ifds(fh, ifds, ifd_off, bigtiff, gdal_ghost)
        stats["bytes"] = cur
    stats["levels"].reverse()
    stats.update(compression=compress, predictor=pred, blocksize=bs, bigtiff=bool(bigtiff))
    return stats


def overview_levels(shape, overview_count: int = 8) -> int:
    """Number of overview levels: OVERVIEW_COUNT (8) capped where a level would drop below one pixel."""
    h, w = int(shape[0]), int(shape[1])
    n = 0
    while n < int(overview_count) and (h > 1 or w > 1):
        h, w = (h + 1) // 2, (w + 1) // 2
        n += 1
    return n


def write_cog(path: str, raster, *, nodata="auto", transform=None, epsg=None, blocksize: int = 512, level: int = 1,
              overview_count: int = 8, compress: str = "zstd", bigtiff: bool = True, num_threads: Optional[int] = None) -> dict:
    """Device raster (torch CUDA tensor / CuPy / __cuda_array_interface__; uint8, int16 or float32) -> COG file with
    the reference's creation options.  The AVERAGE overview cascade runs on the GPU; every level is copied to the
    host one 512-row band at a time through a pinned buffer."""
    import torch
    from .. import _device as _dev
    from .. import kernels as _k
    t = _dev.as_tensor(raster)
    if t.ndim != 2 or t.dtype not in (torch.uint8, torch.int16, torch.float32):
        raise ValueError("write_cog: 2-D uint8 / int16 / float32 device raster expected")
    if nodata == "auto":
        nodata = float("nan") if t.dtype == torch.float32 else 0     # output_nodata_for_dtype (io/output_encoding.py)
    lv = [t]
    for _ in range(overview_levels(t.shape, overview_count)):
        lv.append(_k.overview_average(lv[-1], None if t.dtype == torch.float32 else nodata))
    bs = int(blocksize)
    pin = torch.empty((bs, int(t.shape[1])), dtype=t.dtype, pin_memory=True)

    def rows(a, r0, r1):
        view = pin[: r1 - r0, : a.shape[1]]
        view.copy_(a[r0:r1], non_blocking=False)
        return view.numpy().copy()

    return write_tiff_pyramid(path, lv, nodata=nodata, transform=transform, epsg=epsg, blocksize=bs, compress=compress,
                              level=level, bigtiff=bigtiff, num_threads=num_threads, row_reader=rows)


__all__ = ["write_cog", "write_tiff_pyramid", "overview_levels", "predictor_for_dtype", "apply_predictor", "undo_predictor",
           "zstd_compress", "zstd_decompress"]
