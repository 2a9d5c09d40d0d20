"""Minimal (Big)TIFF / GeoTIFF reader for single-band rasters (the input side of SURVEY.md 8f rank 3:
load_input_dataarray / read_tile_window of the reference read through rasterio, which is not in this image).
Supports tiles or strips, no compression / DEFLATE / ZSTD, predictors 1-3, uint8 / int16 / uint16 / float32,
and the IFD chain (overview levels).  NoData (GDAL_NODATA), pixel scale / tie point and the EPSG code are returned
as metadata; `nodata_to_nan=True` applies the reference's mask rule (isclose(x, nodata, atol=1e-6) | isnan)."""
from __future__ import annotations

import struct
import zlib
from typing import Optional

import numpy as np

from .cog_writer import undo_predictor, zstd_decompress

_SIZES = {1: 1, 2: 1, 3: 2, 4: 4, 5: 8, 6: 1, 7: 1, 8: 2, 9: 4, 10: 8, 11: 4, 12: 8, 16: 8, 17: 8, 18: 8}
_FMT = {1: "B", 2: "c", 3: "H", 4: "I", 6: "b", 7: "B", 8: "h", 9: "i", 11: "f", 12: "d", 16: "Q", 17: "q", 18: "Q"}


def _read_ifds(fh):
    head = fh.read(16)
    if head[:2] != b"II":
        raise ValueError("only little-endian TIFF files are supported")
    magic = struct.unpack("<H", head[2:4])[0]
    big = magic == 43
    if magic not in (42, 43):
        raise ValueError("not a TIFF file")
    off = struct.unpack("<Q", head[8:16])[0] if big else struct.unpack("<I", head[4:8])[0]
    ifds = []
    while off:
        fh.seek(off)
        n = struct.unpack("<Q", fh.read(8))[0] if big else struct.unpack("<H", fh.read(2))[0]
        esz = 20 if big else 12
        raw = fh.read(n * esz + (8 if big else 4))
        tags = {}
        for i in range(n):
            e = raw[i * esz:(i + 1) * esz]
            tag, typ = struct.unpack("<HH", e[:4])
            count = struct.unpack("<Q", e[4:12])[0] if big else struct.unpack("<I", e[4:8])[0]
            val = e[12:20] if big else e[8:12]
            nbytes = count * _SIZES.get(typ, 1)
            if nbytes > len(val):
                pos = struct.unpack("<Q", val)[0] if big else struct.unpack("<I", val)[0]
                here = fh.tell()
                fh.seek(pos)
                data = fh.read(nbytes)
                fh.seek(here)
            else:
                data = val[:nbytes]
            if typ == 2:
                tags[tag] = data.rstrip(b"\0").decode("ascii", "replace")
            elif typ == 5:
                tags[tag] = struct.unpack("<" + "I" * (2 * count), data)
            else:
                tags[tag] = struct.unpack("<" + _FMT[typ] * count, data)
        ifds.append(tags)
        off = struct.unpack("<Q", raw[n * esz:])[0] if big else struct.unpack("<I", raw[n * esz:])[0]
    return ifds, big


def _dtype_of(tags) -> np.dtype:
    bits = tags.get(258, (1,))[0]
    fmt = tags.get(339, (1,))[0]
    key = (fmt, bits)
    table = {(1, 8): "uint8", (1, 16): "uint16", (2, 16): "int16", (3, 32): "float32", (2, 32): "int32", (1, 32): "uint32"}
    if key not in table:
        raise ValueError(f"unsupported sample format {key}")
    return np.dtype(table[key])


def _decode(blob: bytes, comp: int, nbytes: int) -> bytes:
    if comp == 1:
        return blob
    if comp in (8, 32946):
        return zlib.decompress(blob)
    if comp == 50000:
        return zstd_decompress(blob, nbytes)
    raise ValueError(f"unsupported compression {comp}")


def read_geotiff(path: str, level: int = 0, *, nodata_to_nan: bool = False, window=None):
    """-> (array, meta).  level 0 = full resolution, 1.. = overview IFDs.  window = (row0, rows, col0, cols)."""
    with open(path, "rb") as fh:
        ifds, big = _read_ifds(fh)
        tags = ifds[level]
        if tags.get(277, (1,))[0] != 1:
            raise ValueError("single-band rasters only")
        w, h = int(tags[256][0]), int(tags[257][0])
        dt = _dtype_of(tags)
        comp = int(tags.get(259, (1,))[0])
        pred = int(tags.get(317, (1,))[0])
        r0, nr, c0, nc = (0, h, 0, w) if window is None else [int(v) for v in window]
        out = np.empty((nr, nc), dtype=dt)
        if 322 in tags:   # tiled
            tw, tl = int(tags[322][0]), int(tags[323][0])
            offs, cnts = tags[324], tags[325]
            tx = (w + tw - 1) // tw
            for ty_i in range(r0 // tl, (r0 + nr - 1) // tl + 1):
                for tx_i in range(c0 // tw, (c0 + nc - 1) // tw + 1):
                    k = ty_i * tx + tx_i
                    fh.seek(offs[k])
                    raw = _decode(fh.read(cnts[k]), comp, tw * tl * dt.itemsize)
                    tile = undo_predictor(raw, (tl, tw), dt, pred)
                    ya, yb = max(r0, ty_i * tl), min(r0 + nr, (ty_i + 1) * tl, h)
                    xa, xb = max(c0, tx_i * tw), min(c0 + nc, (tx_i + 1) * tw, w)
                    out[ya - r0:yb - r0, xa - c0:xb - c0] = tile[ya - ty_i * tl:yb - ty_i * tl, xa - tx_i * tw:xb - tx_i * tw]
        else:             # strips
            rps = int(tags.get(278, (h,))[0])
            offs, cnts = tags[273], tags[279]
            for s_i in range(r0 // rps, (r0 + nr - 1) // rps + 1):
                rows = min(rps, h - s_i * rps)
                fh.seek(offs[s_i])
                raw = _decode(fh.read(cnts[s_i]), comp, rows * w * dt.itemsize)
                strip = undo_predictor(raw, (rows, w), dt, pred)
                ya, yb = max(r0, s_i * rps), min(r0 + nr, s_i * rps + rows)
                out[ya - r0:yb - r0] = strip[ya - s_i * rps:yb - s_i * rps, c0:c0 + nc]
    main = ifds[0]
    nod: Optional[float] = None
    if 42113 in main:
        try:
            nod = float(main[42113])
        except ValueError:
            nod = None
    meta = {"shape": (h, w), "dtype": str(dt), "levels": len(ifds), "bigtiff": big, "compression": comp, "predictor": pred,
            "tile": (int(tags[323][0]), int(tags[322][0])) if 322 in tags else None, "nodata": nod,
            "pixel_scale": main.get(33550), "tiepoint": main.get(33922), "epsg": None,
            "newsubfiletype": [int(i.get(254, (0,))[0]) for i in ifds]}
    keys = main.get(34735)
    if keys:
        for i in range(4, len(keys), 4):
            if keys[i] in (2048, 3072) and keys[i + 1] == 0:
                meta["epsg"] = int(keys[i + 3])
    if main.get(33550) and main.get(33922):
        sx, sy = main[33550][0], main[33550][1]
        tp = main[33922]
        meta["transform"] = (tp[3] - tp[0] * sx, sx, 0.0, tp[4] + tp[1] * sy, 0.0, -sy)
    if nodata_to_nan:
        a = out.astype(np.float32)
        if nod is not None and nod == nod:
            a[np.isclose(a, np.float32(nod), rtol=0.0, atol=1e-6)] = np.nan   # core/tile_processor.py:185-196
        out = a
    return out, meta


__all__ = ["read_geotiff"]
