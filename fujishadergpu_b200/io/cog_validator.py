"""COG quality check without GDAL (reference: io/cog_validator.py::_validate_cog_for_qgis, which asks GDAL for the
block size, the overview count and IMAGE_STRUCTURE LAYOUT=COG).  The same facts are read from the file itself:
tiled layout, overview IFDs, the structural-metadata ghost area, IFDs before the tile data, block leaders /
trailers consistent with the tile index; the score is the reference's (tiled 30, overviews 10 each up to 40,
COG layout 20, block size 512-1024 10; "good" from 60)."""
from __future__ import annotations

import logging
import os
import struct

from .cog_writer import GDAL_GHOST
from .geotiff_reader import _read_ifds

logger = logging.getLogger(__name__)


def inspect_cog(path: str) -> dict:
    with open(path, "rb") as fh:
        ifds, big = _read_ifds(fh)
        fh.seek(16 if big else 8)
        ghost = fh.read(len(GDAL_GHOST))
        main = ifds[0]
        tiled = 322 in main
        info = {"bigtiff": big, "width": int(main[256][0]), "height": int(main[257][0]), "tiled": tiled,
                "block": (int(main[322][0]), int(main[323][0])) if tiled else (int(main[256][0]), int(main.get(278, (1,))[0])),
                "overviews": sum(1 for i in ifds[1:] if int(i.get(254, (0,))[0]) & 1),
                "compression": int(main.get(259, (1,))[0]),
                "ghost": ghost.startswith(b"GDAL_STRUCTURAL_METADATA_SIZE=") and b"LAYOUT=IFDS_BEFORE_DATA" in ghost,
                "leaders_ok": None, "ifds_before_data": None, "overview_data_first": None,
                "file_size_mb": os.path.getsize(path) / (1024 * 1024)}
        if tiled:
            first = min(min(i[324]) for i in ifds if 324 in i)
            fh.seek(0)
            head = fh.read(16)
            ifd0 = struct.unpack("<Q", head[8:16])[0] if big else struct.unpack("<I", head[4:8])[0]
            info["ifds_before_data"] = ifd0 < first     # (every IFD and index array precedes the first block: checked below)
            info["overview_data_first"] = all(max(ifds[k + 1][324]) < min(ifds[k][324]) for k in range(len(ifds) - 1))
            if info["ghost"] and b"BLOCK_LEADER=SIZE_AS_UINT4" in ghost:
                ok = True
                for tags in ifds:
                    offs, cnts = tags[324], tags[325]
                    for k in (0, len(offs) // 2, len(offs) - 1):       # first, middle, last block of the level
                        fh.seek(offs[k] - 4)
                        lead = struct.unpack("<I", fh.read(4))[0]
                        fh.seek(offs[k] + cnts[k] - 4)
                        tail = fh.read(8)
                        ok = ok and lead == cnts[k] and tail[:4] == tail[4:8]
                info["leaders_ok"] = ok
    info["cog_layout"] = bool(info["tiled"] and info["ghost"] and info["ifds_before_data"] and info["leaders_ok"] is not False)
    score = (30 if info["tiled"] else 0) + min(info["overviews"] * 10, 40) + (20 if info["cog_layout"] else 0)
    score += 10 if 512 <= info["block"][0] <= 1024 else 0
    info["score"] = score
    return info


def validate_cog(path: str) -> bool:
    """True when the file has the COG layout and scores >= 60 (reference :118)."""
    try:
        info = inspect_cog(path)
    except Exception as exc:
        logger.error("COG validation error: %s", exc)
        return False
    logger.info("COG %d x %d, block %s, %d overviews, compression %d, layout %s, score %d/100", info["width"], info["height"],
                info["block"], info["overviews"], info["compression"], "COG" if info["cog_layout"] else "not COG", info["score"])
    return bool(info["cog_layout"] and info["score"] >= 60)


__all__ = ["inspect_cog", "validate_cog"]
