"""Oracle: mltps part 3/4 tiling (V73:649-895) and machisplin.tiles.create / .merge
(V73:1165-1256, 1392-1548) restated on plain arrays.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  PARITY UNPINNED for the terra
geometry primitives (crop / extend / mosaic / merge are terra C++, not in /root/reference);
they are restated from terra's documented semantics:

* ``crop(r, e)`` snaps ``e`` to the cell edges of ``r`` with snap="near" (C ``round`` of
  (coord - origin)/res) and intersects with the raster extent;
* ``extract(r, pts)`` returns the value of the cell containing the point;
* ``mosaic(fun="mean")`` is the NA-ignoring mean of all layers covering a cell,
  ``mosaic(fun="first")`` / ``merge`` take the first non-NA layer.

Documented deviation: the seam strip in the reference is cropped to
``ext(as.points(AAA))`` (V73:773-780, 830-836), a box that runs from the first to the last
*cell centre* of the overlap and therefore sits on a rounding tie of terra's edge snapping.
Which of the two edge columns survives depends on floating-point noise in terra and cannot be
determined without running it; here the strip is every cell where both neighbours are non-NA
and the fade runs between the min and max cell-centre coordinate of that set (V73:787-791).

A grid geometry is the tuple ``(xmin, xmax, ymin, ymax, nrow, ncol)``; rasters are
(nrow, ncol) float64 arrays in terra cell order (row 0 = north), NaN = NA.  A window is the
half-open index box ``(r0, r1, c0, c1)``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

from . import tps as otps


def _cround(x: float) -> int:
    """C ``round``: half away from zero (terra SpatRaster::align, snap = "near")."""
    return int(math.floor(x + 0.5)) if x >= 0 else -int(math.floor(-x + 0.5))


def res(geom):
    xmin, xmax, ymin, ymax, nrow, ncol = geom
    return (xmax - xmin) / ncol, (ymax - ymin) / nrow


def crop_window(geom, ext):
    """terra::crop(rast, ext) as an index window.  ext = (xmin, xmax, ymin, ymax)."""
    xmin, xmax, ymin, ymax, nrow, ncol = geom
    rx, ry = res(geom)
    c0 = _cround((ext[0] - xmin) / rx)
    c1 = _cround((ext[1] - xmin) / rx)
    b0 = _cround((ext[2] - ymin) / ry)          # rows counted from the south edge
    b1 = _cround((ext[3] - ymin) / ry)
    c0, c1 = max(0, min(ncol, c0)), max(0, min(ncol, c1))
    b0, b1 = max(0, min(nrow, b0)), max(0, min(nrow, b1))
    return (nrow - b1, nrow - b0, c0, c1)


def cell_of_points(geom, xy):
    """terra cellFromXY: column/row of the cell containing each point (-1 when outside);
    x == xmax belongs to the last column and y == ymin to the last row."""
    xmin, xmax, ymin, ymax, nrow, ncol = geom
    rx, ry = res(geom)
    xy = np.asarray(xy, dtype=np.float64)
    col = np.floor((xy[:, 0] - xmin) / rx).astype(np.int64)
    row = np.floor((ymax - xy[:, 1]) / ry).astype(np.int64)
    col[xy[:, 0] == xmax] = ncol - 1
    row[xy[:, 1] == ymin] = nrow - 1
    bad = (xy[:, 0] < xmin) | (xy[:, 0] > xmax) | (xy[:, 1] < ymin) | (xy[:, 1] > ymax)
    col[bad] = -1
    row[bad] = -1
    return row, col


def knot_coordinates(geom, xy):
    """V73:127-145: the LONG / LAT columns used as TPS coordinates are the *cell-centre*
    coordinates of the cell each input point falls in (extract of the LONG/LAT rasters)."""
    row, col = cell_of_points(geom, xy)
    x, y = otps.cell_centres(geom, row, col)
    return np.column_stack([x, y]), row, col


# ---------------------------------------------------------------------------------------
# mltps internal tiling, V73:649-681
# ---------------------------------------------------------------------------------------
@dataclass
class TileLayout:
    nRx: int
    nCx: int
    fit_ext: list      # per tile (xmin, xmax, ymin, ymax), order j (south->north) outer, h (west->east) inner
    keep_ext: list
    fit_win: list      # index windows
    keep_win: list


def mltps_tile_layout(geom, tile_px=1500, fit_halo=0.2, keep_halo=0.025) -> TileLayout:
    xmin, xmax, ymin, ymax, nrow, ncol = geom
    nRx = int(math.ceil(nrow / tile_px))            # V73:656-657
    nCx = int(math.ceil(ncol / tile_px))            # V73:660-661
    longDist = (xmax - xmin) / nCx                  # V73:663
    latDist = (ymax - ymin) / nRx                   # V73:665
    fit_ext, keep_ext = [], []
    for j in range(1, nRx + 1):
        for h in range(1, nCx + 1):
            fit_ext.append((xmin + ((longDist * (h - 1)) - (longDist * fit_halo)),
                            xmin + ((longDist * h) + (longDist * fit_halo)),
                            ymin + ((latDist * (j - 1))) - (latDist * fit_halo),
                            ymin + ((latDist * j)) + (latDist * fit_halo)))        # V73:673
            keep_ext.append((xmin + ((longDist * (h - 1)) - (longDist * keep_halo)),
                             xmin + ((longDist * h) + (longDist * keep_halo)),
                             ymin + ((latDist * (j - 1))) - (latDist * keep_halo),
                             ymin + ((latDist * j)) + (latDist * keep_halo)))      # V73:680
    fit_win = [crop_window(geom, e) for e in fit_ext]
    keep_win = []
    for e, fw in zip(keep_ext, fit_win):
        kw = crop_window(geom, e)                   # crop(pred, d): same grid, inside the fit window
        keep_win.append((max(kw[0], fw[0]), min(kw[1], fw[1]), max(kw[2], fw[2]), min(kw[3], fw[3])))
    return TileLayout(nRx, nCx, fit_ext, keep_ext, fit_win, keep_win)


# ---------------------------------------------------------------------------------------
# seam feather + mosaic, V73:739-747, 759-895 and V73:1399-1546
# ---------------------------------------------------------------------------------------
def _place(shape, win, vals):
    out = np.full(shape, np.nan)
    out[win[0]:win[1], win[2]:win[3]] = vals
    return out


def _intersect(a, b):
    w = (max(a[0], b[0]), min(a[1], b[1]), max(a[2], b[2]), min(a[3], b[3]))
    return w if (w[1] > w[0] and w[3] > w[2]) else None


def _seam_blend(geom, winA, A, winB, B, axis):
    """One feathered strip.  Tile A is 'tile 1' of the reference (west tile for a vertical seam,
    south tile for a horizontal one): weight 1 - t for A and t for B, t = (coord - min)/(max - min)
    over the strip (V73:787-798, 856-867).  Returns (window, values) or None."""
    w = _intersect(winA, winB)
    if w is None:
        return None
    a = A[w[0] - winA[0]:w[1] - winA[0], w[2] - winA[2]:w[3] - winA[2]]
    b = B[w[0] - winB[0]:w[1] - winB[0], w[2] - winB[2]:w[3] - winB[2]]
    both = ~np.isnan(a) & ~np.isnan(b)               # AAA <- A + B ; as.points drops NA
    if not both.any():
        return None
    rows = np.nonzero(both.any(axis=1))[0]
    cols = np.nonzero(both.any(axis=0))[0]
    r0, r1, c0, c1 = rows[0], rows[-1] + 1, cols[0], cols[-1] + 1
    a, b, both = a[r0:r1, c0:c1], b[r0:r1, c0:c1], both[r0:r1, c0:c1]
    win = (w[0] + r0, w[0] + r1, w[2] + c0, w[2] + c1)
    x, y = otps.cell_centres(geom, np.arange(win[0], win[1]), np.arange(win[2], win[3]))
    with np.errstate(invalid="ignore", divide="ignore"):
        if axis == "x":
            t = (x - x.min()) / (x.max() - x.min())
            t = np.broadcast_to(t[None, :], a.shape)
        else:
            t = (y - y.min()) / (y.max() - y.min())
            t = np.broadcast_to(t[:, None], a.shape)
        vals = b * t + a * (1.0 - t)                 # feath.ras <- we.R.d2 + we.R.d1
    return win, vals


def feather_merge(geom, wins, tiles, nCx, nRx):
    """Part 4 / tiles.merge on tile rasters given as (window, values) in the reference's order
    (row-major from the south-west, V73:670-681 / 1192-1197)."""
    _, _, _, _, nrow, ncol = geom
    shape = (nrow, ncol)
    tsum = np.zeros(shape)
    tcnt = np.zeros(shape)
    for win, v in zip(wins, tiles):                  # rast.mosaic <- mosaic(tiles, "mean")
        sl = (slice(win[0], win[1]), slice(win[2], win[3]))
        ok = ~np.isnan(v)
        tsum[sl] += np.where(ok, v, 0.0)
        tcnt[sl] += ok
    with np.errstate(invalid="ignore"):
        mosaic = np.where(tcnt > 0, tsum / np.maximum(tcnt, 1), np.nan)
    if nRx * nCx == 1:
        return mosaic                                # V73:893-895
    ssum = np.zeros(shape)
    scnt = np.zeros(shape)

    def add(strip):
        if strip is None:
            return
        win, v = strip
        sl = (slice(win[0], win[1]), slice(win[2], win[3]))
        ok = ~np.isnan(v)
        ssum[sl] += np.where(ok, v, 0.0)
        scnt[sl] += ok

    for j in range(1, nRx + 1):                      # vertical seams, V73:764-806
        for h in range(1, nCx):
            v = h + (j * nCx - nCx) - 1
            add(_seam_blend(geom, wins[v], tiles[v], wins[v + 1], tiles[v + 1], "x"))
    for j in range(1, nRx):                          # horizontal seams, V73:815-877
        for h in range(1, nCx + 1):
            v = h + j * nCx - nCx - 1
            g = v + nCx
            add(_seam_blend(geom, wins[v], tiles[v], wins[g], tiles[g], "y"))
    with np.errstate(invalid="ignore"):
        strips = np.where(scnt > 0, ssum / np.maximum(scnt, 1), np.nan)   # mosaic(strips, "mean")
    return np.where(np.isnan(strips), mosaic, strips)                       # mosaic(first) / merge


def tps_tiled_surface(geom, knots_xy, resid, tile_px=1500, fit_halo=0.2, keep_halo=0.025,
                      min_pts=10, lam=None, return_parts=False, interpolate=None):
    """mltps part 3 + 4 (V73:636-897): the TPS-of-residuals raster.

    ``knots_xy`` are the LONG/LAT columns (cell-centre coordinates, one row per input point that
    survived complete.cases, V73:154), ``resid`` = res.FINAL.  ``interpolate(fit, geom, r0, r1, c0, c1)``
    replaces the numpy statement of terra::interpolate (``oracle.tps.tps_interpolate``) - the tests of
    config-sized rasters pass the C loop of ``oracle/c`` (same arithmetic, OpenMP) here.
    """
    interp = interpolate or otps.tps_interpolate
    lay = mltps_tile_layout(geom, tile_px, fit_halo, keep_halo)
    if lay.nRx * lay.nCx == 1:                       # V73:748-753
        fit = otps.tps_fit(knots_xy, resid, lam=lam)
        out = interp(fit, geom, 0, geom[4], 0, geom[5])
        return (out, lay, [fit]) if return_parts else out
    krow, kcol = cell_of_points(geom, knots_xy)
    tiles, fits = [], []
    for fw, kw in zip(lay.fit_win, lay.keep_win):
        inside = (krow >= fw[0]) & (krow < fw[1]) & (kcol >= fw[2]) & (kcol < fw[3])
        if int(inside.sum()) < min_pts:              # V73:710-721
            tiles.append(np.zeros((kw[1] - kw[0], kw[3] - kw[2])))
            fits.append(None)
            continue
        fit = otps.tps_fit(knots_xy[inside], resid[inside], lam=lam)       # V73:722
        tiles.append(interp(fit, geom, kw[0], kw[1], kw[2], kw[3]))              # V73:726-728
        fits.append(fit)
    out = feather_merge(geom, lay.keep_win, tiles, lay.nCx, lay.nRx)
    return (out, lay, fits) if return_parts else out


# ---------------------------------------------------------------------------------------
# machisplin.tiles.create / machisplin.tiles.merge
# ---------------------------------------------------------------------------------------
def tiles_create(geom, points_xy, out_ncol=3, out_nrow=3, feather_d=50):
    """V73:1165-1256.  Returns per tile: extent, raster window, indices of the input points inside
    (terra::crop of points by the un-snapped extent, borders inclusive), and tile geometry."""
    xmin, xmax, ymin, ymax, nrow, ncol = geom
    fd = feather_d / 2.0                             # V73:1170
    long_pix = (xmax - xmin) / ncol                  # V73:1178
    lat_pix = (ymax - ymin) / nrow
    longDist = (xmax - xmin) / out_ncol
    latDist = (ymax - ymin) / out_nrow
    rx, ry = res(geom)
    pts = np.asarray(points_xy, dtype=np.float64)
    tiles = []
    for j in range(1, out_nrow + 1):
        for h in range(1, out_ncol + 1):
            e = (xmin + ((longDist * (h - 1)) - (long_pix * fd)),
                 xmin + ((longDist * h) + (long_pix * fd)),
                 ymin + ((latDist * (j - 1))) - (lat_pix * fd),
                 ymin + ((latDist * j)) + (lat_pix * fd))                   # V73:1195
            win = crop_window(geom, e)                                      # V73:1207
            inside = np.nonzero((pts[:, 0] >= e[0]) & (pts[:, 0] <= e[1]) &
                                (pts[:, 1] >= e[2]) & (pts[:, 1] <= e[3]))[0]   # V73:1242
            tgeom = (xmin + win[2] * rx, xmin + win[3] * rx,
                     ymax - win[1] * ry, ymax - win[0] * ry, win[1] - win[0], win[3] - win[2])
            tiles.append({"ext": e, "win": win, "points": inside, "geom": tgeom})
    return {"tiles": tiles, "nC": out_ncol, "nR": out_nrow}


def tiles_merge(geom, wins, rasters, in_ncol, in_nrow):
    """V73:1392-1548 on tile rasters placed at ``wins`` of the full grid."""
    return feather_merge(geom, wins, rasters, in_ncol, in_nrow)
