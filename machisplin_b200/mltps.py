"""Host-side mirror of parts 1b-5 of ``machisplin.mltps()`` for one response column (V73:326-932).

What stays in R in the reference's world (and is *not* here): fitting the six models (V73:220-320, 447-614)
and the L-BFGS-B driver.  This module takes the fitted models as flat descriptors
(``include/machisplin_b200.h``) and does the host glue the reference does in R
around the hot path - the weight rule, the residuals that feed the TPS, the R^2 rule of part 5 - with every
per-cell and per-knot computation going through the engine (C ABI -> CUDA)."""
from __future__ import annotations

from typing import Optional

import numpy as np

from .engine import Engine, as_geom

MODEL_LETTERS_ALL = "bgnmrv"      # brt gam nn mars rf svm (V73:340-362)
MODEL_LETTERS_SMOOTH = "gnmv"     # smooth.outputs.only = TRUE (V73:365-392)


def select_models(p, letters: str = MODEL_LETTERS_ALL):
    """V73:336-393: keep model k iff round(p_k, 2) > 0.05 * sum(p); its weight is round(p_k, 2); the
    divisor stays the UNROUNDED total over all candidates (V73:337, 619-620 - no renormalisation)."""
    p = np.asarray(p, dtype=np.float64)
    assert p.size == len(letters)
    tot = float(np.sum(p))
    kept, w = "", []
    for pk, letter in zip(p, letters):
        rk = float(np.round(pk, 2))
        if rk > 0.05 * tot:
            kept += letter
            w.append(rk)
    return kept, np.array(w), tot


def rss_objective_from_gram(G: np.ndarray):
    """V73:329-331 / 369-371 as a quadratic form: fit(k) = k'Gk / (sum k)^2 with G = R'R from ``Engine.gram``."""
    def fit(k):
        k = np.asarray(k, dtype=np.float64)
        return float(k @ G @ k) / float(np.sum(k)) ** 2
    return fit


def knot_cells(geom, points_xy):
    """terra::extract picks the containing cell; the knots of the TPS are the CELL-CENTRE coordinates of the
    points (LONG / LAT rasters, V73:127-145).  Returns (knots_xy, row, col); row = -1 outside the raster."""
    g = as_geom(geom)
    pts = np.asarray(points_xy, dtype=np.float64).reshape(-1, 2)
    col = np.floor((pts[:, 0] - g.xmin) / g.rx).astype(np.int64)
    row = np.floor((g.ymax - pts[:, 1]) / g.ry).astype(np.int64)
    col[pts[:, 0] == g.xmax] = g.ncol - 1
    row[pts[:, 1] == g.ymin] = g.nrow - 1
    out = (pts[:, 0] < g.xmin) | (pts[:, 0] > g.xmax) | (pts[:, 1] < g.ymin) | (pts[:, 1] > g.ymax)
    row[out] = col[out] = -1
    xy = np.column_stack([g.xmin + (col + 0.5) * g.rx, g.ymax - (row + 0.5) * g.ry])
    return xy, row.astype(np.int32), col.astype(np.int32)


def mltps_response(engine: Engine, geom, cov: np.ndarray, points_xy, resp, models: dict, p, letters: str = MODEL_LETTERS_ALL,
                   tps: bool = True, tile_px: int = 1500, lam: Optional[float] = None) -> dict:
    """Parts 1b (weight rule) to 5 for one response: returns the reference's per-layer list (V73:919-930,
    946-952): ``final`` raster, ``residuals``, ``summary`` (kept letters, weights, R^2 values)."""
    g = as_geom(geom)
    cov = np.ascontiguousarray(cov, dtype=np.float32)
    C = cov.shape[0]
    kept, w, wt = select_models(p, letters)
    knots_xy, krow, kcol = knot_cells(g, points_xy)
    X = np.empty((len(krow), C + 2))
    inside = krow >= 0
    X[:] = np.nan
    X[inside, :C] = cov[:, krow[inside], kcol[inside]].T                      # RAST_VAL <- extract(rast_stack, points), V73:145
    X[:, C:] = knots_xy
    ok = inside & ~np.isnan(X).any(axis=1)                                    # complete.cases, V73:154
    X, knots_xy, krow, kcol = X[ok], knots_xy[ok], krow[ok], kcol[ok]
    resp = np.asarray(resp, dtype=np.float64)[ok]
    ens = engine.ensemble_create(g, models, kept, w, wt, C + 2)
    # res.FINAL = sum_k w_k (resp - f_k) / w_total = (sum_k w_k) resp / w_total - ensemble(x)   (V73:477-482 ... 620)
    res_final = resp * (float(np.sum(w)) / wt) - engine.ensemble_predict_points(ens, X)
    tss = float(np.sum((resp - resp.mean()) ** 2))                           # V73:626
    rsq_model = 1.0 - float(np.sum(res_final ** 2)) / tss                    # V73:625-627
    summary = {"kept": kept, "weights": w, "w_total": wt, "rsq_model": rsq_model}
    if not tps:
        final, _ = engine.mltps_predict(g, ens, cov, None, None)
        return {"final": final, "residuals": res_final, "summary": summary}   # V73:946-952
    final, spline = engine.mltps_predict(g, ens, cov, knots_xy, res_final, lam=lam, tile_px=tile_px)
    f_actual = final[krow, kcol]                                             # V73:910
    rsq_final = 1.0 - float(np.sum((resp - f_actual) ** 2)) / tss            # V73:911-917
    summary.update({"rsq_final": rsq_final, "tps_kept": bool(rsq_final > rsq_model),
                    "lambda": None if spline is None else spline.lam})
    if not rsq_final > rsq_model:                                            # V73:925-930: keep the better of the two
        # the reference has already overwritten l$residuals with resp - f.actual (V73:913) and never restores it:
        # with tps = TRUE the residuals are the TPS-corrected ones even when the TPS raster is discarded
        final, _ = engine.mltps_predict(g, ens, cov, None, None)
    return {"final": final, "residuals": resp - f_actual, "summary": summary}
