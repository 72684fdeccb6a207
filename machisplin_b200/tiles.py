"""Host-side mirror of ``machisplin.tiles.create`` / ``machisplin.tiles.merge`` (V73:1165-1256, 1392-1548).

Index arithmetic only (what the reference does in R on the host); the blend itself runs in
``mb_tiles_merge`` on the device.  Tiles are numbered like the reference: row-major from the SOUTH-WEST
tile (V73:1192-1197)."""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Sequence

import numpy as np

from .engine import Engine, Geom, as_geom


def _cround(x: float) -> int:
    """C round(): half away from zero (terra's extent snapping)."""
    return int(math.floor(x + 0.5)) if x >= 0 else -int(math.floor(-x + 0.5))


def crop_window(geom: Geom, ext) -> tuple:
    """terra::crop(rast, ext) as a half-open cell window (r0, r1, c0, c1): snap = "near" on cell edges, clipped."""
    exmin, exmax, eymin, eymax = ext
    c0 = min(max(_cround((exmin - geom.xmin) / geom.rx), 0), geom.ncol)
    c1 = min(max(_cround((exmax - geom.xmin) / geom.rx), 0), geom.ncol)
    b0 = min(max(_cround((eymin - geom.ymin) / geom.ry), 0), geom.nrow)
    b1 = min(max(_cround((eymax - geom.ymin) / geom.ry), 0), geom.nrow)
    return (geom.nrow - b1, geom.nrow - b0, c0, c1)


@dataclass
class Tile:
    ext: tuple            # (xmin, xmax, ymin, ymax) as computed at V73:1195 (not snapped)
    win: tuple            # (r0, r1, c0, c1) cell window of the full grid (V73:1207 crop)
    geom: Geom            # geometry of the cropped raster
    points: np.ndarray    # indices of the input points inside the extent (V73:1242)


@dataclass
class TileSet:
    tiles: List[Tile]
    nC: int
    nR: int


def tiles_create(geom, points_xy, out_ncol: int = 3, out_nrow: int = 3, feather_d: float = 50) -> TileSet:
    """``machisplin.tiles.create(int.values, rast.in, out.ncol, out.nrow, feather.d)``: tile h, j gets the
    extent of its 1/ncol x 1/nrow share of the raster widened by feather.d / 2 pixels on every side."""
    geom = as_geom(geom)
    fd = feather_d / 2.0                                                       # V73:1170
    long_pix, lat_pix = geom.rx, geom.ry                                       # V73:1178-1179
    long_dist = (geom.xmax - geom.xmin) / out_ncol
    lat_dist = (geom.ymax - geom.ymin) / out_nrow
    pts = np.asarray(points_xy, dtype=np.float64).reshape(-1, 2)
    tiles = []
    for j in range(1, out_nrow + 1):
        for h in range(1, out_ncol + 1):
            e = (geom.xmin + ((long_dist * (h - 1)) - (long_pix * fd)), geom.xmin + ((long_dist * h) + (long_pix * fd)),
                 geom.ymin + ((lat_dist * (j - 1))) - (lat_pix * fd), geom.ymin + ((lat_dist * j)) + (lat_pix * fd))   # V73:1195
            win = crop_window(geom, e)
            inside = np.nonzero((pts[:, 0] >= e[0]) & (pts[:, 0] <= e[1]) & (pts[:, 1] >= e[2]) & (pts[:, 1] <= e[3]))[0]
            tg = Geom(geom.xmin + win[2] * geom.rx, geom.xmin + win[3] * geom.rx, geom.ymax - win[1] * geom.ry,
                      geom.ymax - win[0] * geom.ry, win[1] - win[0], win[3] - win[2])
            tiles.append(Tile(e, win, tg, inside))
    return TileSet(tiles, out_ncol, out_nrow)


def tiles_merge(engine: Engine, geom, tileset: TileSet, rasters: Sequence[np.ndarray]) -> np.ndarray:
    """``machisplin.tiles.merge(tiles, in.ncol, in.nrow)`` on the ``$final`` rasters of the tiles."""
    return engine.tiles_merge(as_geom(geom), [t.win for t in tileset.tiles], list(rasters), tileset.nC, tileset.nR)
