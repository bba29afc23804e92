"""ImageJ ROI archives (``segmentation/cellpose/imagej_rois/global_coords_rois.zip``) -> polygons.

The reference reads them with ``roifile.roiread`` and uses ``roi.subpixel_coordinates[:, ::-1]`` as the
(y, x) polygon of a cell in global micrometres (PD:4060-4105).  ``roifile`` is not a dependency here; this is
a reader of the published ImageJ ``.roi`` layout (``ij/io/RoiDecoder.java``), big-endian:

    0  "Iout"      4 version (int16)   6 type (uint8)   8 top, 10 left, 12 bottom, 14 right (int16)
    16 n_coordinates (uint16)          50 options (int16; 128 = SUB_PIXEL_RESOLUTION)      60 header2 offset
    64 x[n], y[n] int16 relative to (left, top);  then, with sub-pixel resolution, x[n], y[n] float32 absolute

Only outline types carry a vertex list (polygon 0, freeline 4, polyline 5, freehand 7, traced 8, angle 9,
point 10); rectangles (1) are expanded to their four corners; other types are skipped like an unusable ROI.
"""

from __future__ import annotations

import struct
import zipfile
from pathlib import Path

import numpy as np

_VERTEX_TYPES = {0, 4, 5, 7, 8, 9, 10}
_SUB_PIXEL_RESOLUTION = 128


def parse_roi(data: bytes) -> np.ndarray | None:
    """(n, 2) float64 vertices in (x, y) order -- roifile's ``subpixel_coordinates`` convention -- or None."""
    if len(data) < 64 or data[:4] != b"Iout":
        return None
    version = struct.unpack(">h", data[4:6])[0]
    roi_type = data[6]
    top, left, bottom, right = struct.unpack(">hhhh", data[8:16])
    n = struct.unpack(">H", data[16:18])[0]
    options = struct.unpack(">h", data[50:52])[0]
    if roi_type == 1:  # rectangle
        if options & _SUB_PIXEL_RESOLUTION and version >= 223:
            xd, yd, wd, hd = struct.unpack(">ffff", data[18:34])
            x0, y0, x1, y1 = xd, yd, xd + wd, yd + hd
        else:
            x0, y0, x1, y1 = float(left), float(top), float(right), float(bottom)
        return np.array([[x0, y0], [x1, y0], [x1, y1], [x0, y1]], dtype=np.float64)
    if roi_type not in _VERTEX_TYPES or n == 0:
        return None
    base = 64
    if options & _SUB_PIXEL_RESOLUTION and version >= 222 and len(data) >= base + 4 * n + 8 * n:
        off = base + 4 * n
        xs = np.frombuffer(data, dtype=">f4", count=n, offset=off).astype(np.float64)
        ys = np.frombuffer(data, dtype=">f4", count=n, offset=off + 4 * n).astype(np.float64)
    else:
        if len(data) < base + 4 * n:
            return None
        xs = np.frombuffer(data, dtype=">i2", count=n, offset=base).astype(np.float64) + left
        ys = np.frombuffer(data, dtype=">i2", count=n, offset=base + 2 * n).astype(np.float64) + top
    return np.stack([xs, ys], axis=1)


def read_roi_zip(path: str | Path) -> list[np.ndarray | None]:
    """One entry per ``.roi`` member, in archive order (the reference numbers cells by that order)."""
    out = []
    with zipfile.ZipFile(path) as zf:
        for name in zf.namelist():
            if name.lower().endswith(".roi"):
                out.append(parse_roi(zf.read(name)))
    return out


def encode_polygon_roi(xy: np.ndarray, name: str = "") -> bytes:
    """Polygon ROI with sub-pixel float32 vertices (what Cellpose / roifile write); used by tests and tools."""
    xy = np.asarray(xy, dtype=np.float64)
    n = int(xy.shape[0])
    left, top = int(np.floor(xy[:, 0].min())), int(np.floor(xy[:, 1].min()))
    right, bottom = int(np.ceil(xy[:, 0].max())), int(np.ceil(xy[:, 1].max()))
    head = bytearray(64)
    head[:4] = b"Iout"
    struct.pack_into(">h", head, 4, 228)
    head[6] = 0  # polygon
    struct.pack_into(">hhhh", head, 8, top, left, bottom, right)
    struct.pack_into(">H", head, 16, n)
    struct.pack_into(">h", head, 50, _SUB_PIXEL_RESOLUTION)
    xi = np.clip(np.round(xy[:, 0]) - left, -32768, 32767).astype(">i2").tobytes()
    yi = np.clip(np.round(xy[:, 1]) - top, -32768, 32767).astype(">i2").tobytes()
    xf = xy[:, 0].astype(">f4").tobytes()
    yf = xy[:, 1].astype(">f4").tobytes()
    return bytes(head) + xi + yi + xf + yf


def write_roi_zip(path: str | Path, polygons_xy: list[np.ndarray]) -> None:
    with zipfile.ZipFile(path, "w") as zf:
        for i, xy in enumerate(polygons_xy):
            zf.writestr(f"{i:05d}.roi", encode_polygon_roi(xy))
