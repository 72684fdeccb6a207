"""Oracle: parts 2-5 of ``machisplin.mltps`` (V73:442-932) for one response column, on arrays,
with the six fitted models supplied as flat descriptors (fitting is out of scope and stays in R).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  PARITY UNPINNED.
"""
from __future__ import annotations

import numpy as np

from . import models as om
from . import tiles as otl
from . import tps as otps


def raster_features(geom, cov, row0=0, row1=None):
    """rast_stack <- c(covar.ras, LONG, LAT) (V73:127-138) as an (n, C+2) matrix for rows [row0,row1)."""
    C, nrow, ncol = cov.shape
    row1 = nrow if row1 is None else row1
    rows = np.arange(row0, row1)
    x, y = otps.cell_centres(geom, rows, np.arange(ncol))
    n = rows.size * ncol
    X = np.empty((n, C + 2))
    X[:, :C] = cov[:, row0:row1, :].reshape(C, n).T
    X[:, C] = np.tile(x, rows.size)
    X[:, C + 1] = np.repeat(y, ncol)
    return X


def point_features(geom, cov, row, col):
    """RAST_VAL <- extract(rast_stack, points) (V73:145)."""
    C = cov.shape[0]
    x, y = otps.cell_centres(geom, row, col)
    X = np.empty((row.size, C + 2))
    X[:, :C] = cov[:, row, col].T
    X[:, C] = x
    X[:, C + 1] = y
    return X


def ensemble_raster(geom, cov, models, kept, w, w_total, block_rows=64):
    """pred.elev (V73:447-619): weighted sum of terra::predict rasters / OptX.mfit.wt.tot."""
    C, nrow, ncol = cov.shape
    out = np.empty((nrow, ncol))
    for r0 in range(0, nrow, block_rows):
        r1 = min(nrow, r0 + block_rows)
        X = raster_features(geom, cov, r0, r1)
        out[r0:r1] = om.ensemble_predict(models, kept, w, w_total, X).reshape(r1 - r0, ncol)
    return out


def ensemble_residuals(X_pts, resp, models, kept, w, w_total):
    """res.FINAL (V73:477-482, ..., 620): sum_k w_k (resp - f_k(x_i)) / w_total."""
    acc = np.zeros(resp.shape[0])
    for letter, wk in zip(kept, w):
        acc = acc + (resp - om.predict_model(letter, models[letter], X_pts)) * wk
    return acc / w_total


def mltps_one_response(geom, cov, points_xy, resp, models, kept, w, w_total, tps=True,
                       tile_px=1500, lam=None, interpolate=None, ensemble=None):
    """Parts 2-5 for one response column.  Returns a dict mirroring the reference's ``l`` list
    (V73:919-930): final raster, residuals, r2 ensemble, r2 final, plus the intermediate rasters."""
    knots_xy, krow, kcol = otl.knot_coordinates(geom, points_xy)
    X_pts = point_features(geom, cov, krow, kcol)
    ok = ~np.isnan(X_pts).any(axis=1) & (krow >= 0)             # complete.cases, V73:154
    knots_xy, krow, kcol, X_pts, resp = knots_xy[ok], krow[ok], kcol[ok], X_pts[ok], np.asarray(resp, float)[ok]
    pred = (ensemble or ensemble_raster)(geom, cov, models, kept, w, w_total)
    res_final = ensemble_residuals(X_pts, resp, models, kept, w, w_total)
    rss_m = float(np.sum(res_final ** 2))                        # V73:625
    tss = float(np.sum((resp - resp.mean()) ** 2))               # V73:626
    rsq_model = 1.0 - rss_m / tss
    out = {"pred": pred, "res_final": res_final, "rsq_model": rsq_model, "knots_xy": knots_xy}
    if not tps:
        out["final"] = pred                                      # V73:951-952
        return out
    surf = otl.tps_tiled_surface(geom, knots_xy, res_final, tile_px=tile_px, lam=lam, interpolate=interpolate)
    final = pred + surf                                          # app(c(pred, TPS), sum): NA propagates
    f_actual = final[krow, kcol]                                 # V73:910
    rss_final = float(np.nansum((resp - f_actual) ** 2)) if not np.isnan(f_actual).any() \
        else float(np.sum((resp - f_actual) ** 2))
    rsq_final = 1.0 - rss_final / tss                            # V73:917
    out.update({"tps": surf, "rsq_final": rsq_final,
                "residuals": resp - f_actual,
                "final": final if rsq_final > rsq_model else pred,   # V73:925-930
                "tps_kept": bool(rsq_final > rsq_model)})
    return out
