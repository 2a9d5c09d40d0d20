"""Signed per-axis pixel scales in metres from GeoTIFF metadata (reference: io/raster_info.py:12-99).
A geographic raster (degrees) is converted at its centre latitude with the WGS84 series the reference uses
everywhere, so radii given in pixels mean the same thing on both backends; the DEM itself is never rescaled."""
from __future__ import annotations

import math
from typing import Optional, Tuple

# metres per degree: cosine series in the latitude (WGS84), coefficients of the reference (:21-32)
_LAT_SERIES = ((0, 111132.92), (2, -559.82), (4, 1.175), (6, -0.0023))
_LON_SERIES = ((1, 111412.84), (3, -93.5), (5, 0.118))


def meters_per_degree(lat_deg: float) -> Tuple[float, float]:
    """(metres per degree of longitude, metres per degree of latitude) at a latitude."""
    phi = math.radians(float(lat_deg))
    m_lat = 0.0
    for k, c in _LAT_SERIES:
        m_lat = m_lat + c * (math.cos(k * 1.0 * phi) if k else 1.0)
    m_lon = 0.0
    for k, c in _LON_SERIES:
        m_lon = m_lon + c * math.cos(k * 1.0 * phi)
    return max(1e-6, float(m_lon)), float(m_lat)     # the longitude scale must stay positive at the poles


def is_geographic_epsg(epsg: Optional[int]) -> bool:
    """EPSG codes of geographic 2-D coordinate systems (4000-4999: WGS84 4326, JGD2011 6668 is the exception below)."""
    if epsg is None:
        return False
    code = int(epsg)
    if code in (4087, 4088, 4936, 4978):       # projected (World Equidistant Cylindrical) / geocentric codes of that range
        return False
    return 4000 <= code < 5000 or code in (6668, 6318, 7844)    # JGD2011, NAD83(2011), GDA2020


def metric_pixel_scales(transform, epsg: Optional[int], shape) -> Tuple[float, float, float, bool, Optional[float]]:
    """(scale_x, scale_y, mean |scale|, is_geographic, centre latitude) from a GDAL geotransform
    (x0, dx, 0, y0, 0, dy), an EPSG code and the raster shape -- metric_pixel_scales_from_metadata (:38-99)."""
    if transform is None:
        return 1.0, -1.0, 1.0, False, None
    x0, dx, _rx, y0, _ry, dy = [float(v) for v in transform]
    h, w = int(shape[0]), int(shape[1])
    if epsg is None:
        left, right = sorted((x0, x0 + dx * w))
        bottom, top = sorted((y0, y0 + dy * h))
        lonlat = -180.0 <= left <= 180.0 and -180.0 <= right <= 180.0 and -90.0 <= bottom <= 90.0 and -90.0 <= top <= 90.0
        if lonlat and 0.0 < abs(dx) <= 1.0 and 0.0 < abs(dy) <= 1.0:
            raise ValueError("Raster has no CRS and its extent/pixel size look geographic; "
                             "cannot safely interpret degree-sized pixels as meters. Assign a CRS first.")
    if is_geographic_epsg(epsg):
        lat_c = 0.5 * (y0 + (y0 + dy * h))
        m_lon, m_lat = meters_per_degree(lat_c)
        sx = math.copysign(abs(dx) * m_lon, dx if dx != 0 else 1.0)
        sy = math.copysign(abs(dy) * m_lat, dy if dy != 0 else -1.0)
        return float(sx), float(sy), 0.5 * (abs(sx) + abs(sy)), True, float(lat_c)
    sx = dx if dx != 0 else 1.0
    sy = dy if dy != 0 else -1.0
    return float(sx), float(sy), 0.5 * (abs(sx) + abs(sy)), False, None


__all__ = ["meters_per_degree", "metric_pixel_scales", "is_geographic_epsg"]
