"""GPU parity at BASELINE-config scale ("Next round" item 1): the reference's bundled workload (C1) through the engine,
the GCV fit at 5 000 knots against the oracle, a full C2 raster, user tiles end to end, and part 5 (a7).

Every comparison is CUDA path (through the C ABI) vs ``oracle/`` on the same inputs; tolerances are relative to max|ref| over the
raster and sit inside the north-star bound of 1e-5.  Large grids use the oracle's C loops (``oracle.cbind``: the same float64
pair loop as ``oracle.tps.tps_interpolate``, OpenMP) so that the checker finishes in seconds."""
import os

import numpy as np
import pytest

from machisplin_b200 import mltps as mm, synth, tiles as mt
from machisplin_b200.engine import Geom
from oracle import cbind, mltps as oml, models as om, tiles as otl, tps as otps

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def relerr(a, b):
    assert np.array_equal(np.isnan(a), np.isnan(b)), "NA masks differ"
    m = ~np.isnan(b)
    return float(np.max(np.abs(a[m] - b[m])) / np.max(np.abs(b[m])))


# ------------------------------------------------------------------------------------------------------------------------
# C1: the reference's own example data (README.md:67-91): data-raw/sampling.csv x inst/extdata rasters
# ------------------------------------------------------------------------------------------------------------------------
def test_c1_bundled_tps_surface_full_raster(engine):
    """mltps part 3-4 (V73:649-895) on the bundled geometry: 3264 x 2476 cells, 813 points, 2 x 3 internal tiles at the
    1500-px rule with 190 ... 237 knots each; response bio_1.  Near-interpolating fits (sum|c| ~ 1e8 for values of O(200)):
    the hardest precision case the survey found (H1)."""
    z = np.load(os.path.join(GOLD, "bundled_c1.npz"))
    geom = tuple(float(v) for v in z["geom"][:4]) + (int(z["geom"][4]), int(z["geom"][5]))
    knots, y = z["knots_xy"], z["points"][:, 2]
    ref, lay, fits = otl.tps_tiled_surface(geom, knots, y, tile_px=1500, return_parts=True, interpolate=cbind.interpolate_c)
    assert (lay.nRx, lay.nCx) == (2, 3) and [f.knots_xy.shape[0] for f in fits] == list(z["tile_counts"])
    got, sp = engine.mltps_predict(Geom(*geom), None, None, knots, y, tile_px=1500)
    assert sp is None                                     # tiled mode returns no global spline
    assert got.shape == (2476, 3264)
    assert relerr(got, ref) < 2e-6
    # the dedicated entry point of part 3-4 gives the same raster
    got2 = engine.tiles_tps(Geom(*geom), knots, y, tile_px=1500)
    np.testing.assert_array_equal(got, got2)


def _c1_decimated():
    z = np.load(os.path.join(GOLD, "bundled_c1.npz"))
    c = np.load(os.path.join(GOLD, "bundled_c1_cov4.npz"))
    geom = tuple(float(v) for v in c["geom"][:4]) + (int(c["geom"][4]), int(c["geom"][5]))
    cov = np.stack([c[k].astype(np.float32) for k in ("alt", "slope", "twi")])
    cov[np.stack([c[k] for k in ("alt", "slope", "twi")]) == -32768] = np.nan
    return geom, cov, z["points"]


def test_c1_bundled_mltps_on_the_real_covariates(engine):
    """machisplin.mltps parts 1b-5 for bio_1 (V73:326-932) on the reference's rasters (4 x decimated fixture: 619 x 816 cells of
    alt / slope / TWI - rough INT16 terrain fields, not the smooth synthetic planes), six models fitted on the real
    training table, TPS of the residuals with the 2 x 3 tile layout of the full-size raster (tile_px = 375)."""
    geom, cov, pts = _c1_decimated()
    krow, kcol = otl.cell_of_points(geom, pts[:, :2])
    kxy, _, _ = otl.knot_coordinates(geom, pts[:, :2])
    X = np.column_stack([cov[:, krow, kcol].T.astype(np.float64), kxy])
    ok = ~np.isnan(X).any(axis=1)
    resp = pts[:, 2]
    models = synth.make_models(Geom(*geom), 3, 0, 5, rf_trees=500, gbm_trees=1000, table=(X[ok], resp[ok]))
    p = np.array([0.31, 0.22, 0.17, 0.12, 0.10, 0.08])
    kept, w, wt = om.select_models(p)
    ref = oml.mltps_one_response(geom, cov, pts[:, :2], resp, models, kept, w, wt, tps=True, tile_px=375,
                                 interpolate=cbind.interpolate_c, ensemble=cbind.ensemble_raster_c)
    got = mm.mltps_response(engine, Geom(*geom), cov, pts[:, :2], resp, models, p, tile_px=375)
    assert got["summary"]["kept"] == kept
    assert got["summary"]["tps_kept"] == ref["tps_kept"]
    assert abs(got["summary"]["rsq_model"] - ref["rsq_model"]) < 1e-9
    assert abs(got["summary"]["rsq_final"] - ref["rsq_final"]) < 1e-6
    assert relerr(got["final"], ref["final"]) < 5e-6
    scale = np.max(np.abs(ref["final"][~np.isnan(ref["final"])]))
    assert np.max(np.abs(got["residuals"] - ref["residuals"])) < 5e-6 * scale


def test_forest_kernel_on_the_real_covariates_is_exact(engine):
    """The tile-pruned forest walk on rough real terrain (slope / TWI change by several units from cell to cell, so most
    trees fork inside a tile): every surviving comparison is the reference's own, the raster matches the C oracle."""
    geom, cov, pts = _c1_decimated()
    krow, kcol = otl.cell_of_points(geom, pts[:, :2])
    kxy, _, _ = otl.knot_coordinates(geom, pts[:, :2])
    X = np.column_stack([cov[:, krow, kcol].T.astype(np.float64), kxy])
    ok = ~np.isnan(X).any(axis=1)
    models = synth.make_models(Geom(*geom), 3, 0, 6, kept="rb", rf_trees=500, gbm_trees=1000, table=(X[ok], pts[ok, 2]))
    for kept in ("r", "b", "rb"):
        w = np.array([0.6, 0.4][:len(kept)])
        ens = engine.ensemble_create(Geom(*geom), models, kept, w, 1.0, 5)
        got = engine.ensemble_eval(ens, cov)
        ref = cbind.ensemble_eval(models, kept, w, 1.0, cov, geom)
        assert relerr(got, ref) < 2e-7
        ens.free()


# ------------------------------------------------------------------------------------------------------------------------
# a1 at config scale: fields::Tps GCV fit with 5 000 knots (BASELINE config 3)
# ------------------------------------------------------------------------------------------------------------------------
def test_fit_gcv_5000_knots_matches_oracle(engine):
    """lambda, eff.df, c, d of the two-stage in-house GCV fit against the oracle's LAPACK eigendecomposition at the knot count
    of BASELINE config 3 (the oracle fit takes ~10 s of CPU)."""
    cfg = synth.CONFIGS["c3"]
    geom = synth.make_geom(cfg["nrow"], cfg["ncol"])
    xy, _, _ = synth.make_knots(geom, cfg["knots"], cfg["seed"])
    y = synth.residual_field(xy, cfg["seed"])
    ref = otps.tps_fit(xy, y)
    sp = engine.tps_fit(xy, y)
    assert abs(sp.lam - ref.lam) <= 1e-6 * ref.lam
    assert abs(sp.eff_df - ref.eff_df) <= 1e-6 * ref.eff_df
    assert np.max(np.abs(sp.c - ref.c)) <= 1e-7 * np.max(np.abs(ref.c))
    np.testing.assert_allclose(sp.d, ref.d, rtol=1e-7, atol=1e-9 * np.abs(ref.d).max())
    eta, _ = sp.decomposition()
    np.testing.assert_allclose(eta, ref.eta, rtol=1e-6, atol=1e-12 * ref.eta.max())
    # and the surface of the oracle's coefficients vs the GPU's own fit + fast evaluator on a row block of the config's grid
    win = (4000, 4016, 0, cfg["ncol"])
    want = cbind.tps_eval(ref, geom.as_tuple(), win)
    got = engine.tps_eval(sp, geom, window=win, method="fast")
    assert relerr(got, want) < 2e-6


# ------------------------------------------------------------------------------------------------------------------------
# C2: 2048 x 2048, 1 000 knots, TPS residual surface only - the whole raster
# ------------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tile_px", [0, 1500])
def test_c2_full_raster(engine, tile_px):
    """BASELINE config 2 in both TPS modes: one global spline (tile_px = 0, the nRx*nCx == 1 branch V73:748-753) and the
    reference's own rule for a 2048^2 raster (2 x 2 tiles of 1500 px, V73:649-895)."""
    cfg = synth.CONFIGS["c2"]
    geom = synth.make_geom(cfg["nrow"], cfg["ncol"])
    xy, _, _ = synth.make_knots(geom, cfg["knots"], cfg["seed"])
    y = synth.residual_field(xy, cfg["seed"])
    ref = otl.tps_tiled_surface(geom.as_tuple(), xy, y, tile_px=tile_px or 10 ** 9, interpolate=cbind.interpolate_c)
    got, sp = engine.mltps_predict(geom, None, None, xy, y, tile_px=tile_px)
    assert (sp is not None) == (tile_px == 0)
    assert relerr(got, ref) < 2e-6


# ------------------------------------------------------------------------------------------------------------------------
# machisplin.tiles.create -> per-tile mltps -> machisplin.tiles.merge (V73:1165-1256, 1392-1548)
# ------------------------------------------------------------------------------------------------------------------------
def test_user_tiles_end_to_end(engine):
    """README Example 4 shape: cut raster + points into 2 x 2 user tiles with a 50-px feather halo, run mltps parts 2-5 on every
    tile (own fits, internal tiling at tile_px), merge.  The oracle composes the same three steps."""
    geom = synth.make_geom(640, 900)
    C = 3
    cov = synth.covariate_planes(geom, C)
    xy, _, _ = synth.make_knots(geom, 1200, 3)
    resp = 200.0 + 40 * synth.residual_field(xy, 3)
    models = synth.make_models(geom, C, 600, 4, kept="gnmv")
    p = np.array([0.4, 0.3, 0.2, 0.1])
    kept, w, wt = om.select_models(p, letters="gnmv")
    ts = mt.tiles_create(geom, xy, 2, 2, feather_d=50)
    tc = otl.tiles_create(geom.as_tuple(), xy, 2, 2, feather_d=50)
    got_tiles, ref_tiles = [], []
    for t, o in zip(ts.tiles, tc["tiles"]):
        assert t.win == tuple(o["win"]) and np.array_equal(t.points, o["points"])
        sub = cov[:, t.win[0]:t.win[1], t.win[2]:t.win[3]]
        r = mm.mltps_response(engine, t.geom, sub, xy[t.points], resp[t.points], models, p, letters="gnmv", tile_px=300)
        q = oml.mltps_one_response(o["geom"], sub, xy[o["points"]], resp[o["points"]], models, kept, w, wt, tile_px=300,
                                   interpolate=cbind.interpolate_c, ensemble=cbind.ensemble_raster_c)
        assert r["summary"]["tps_kept"] == q["tps_kept"]
        got_tiles.append(r["final"])
        ref_tiles.append(q["final"])
    got = mt.tiles_merge(engine, geom, ts, got_tiles)
    ref = otl.tiles_merge(geom.as_tuple(), [o["win"] for o in tc["tiles"]], ref_tiles, 2, 2)
    assert relerr(got, ref) < 5e-6


# ------------------------------------------------------------------------------------------------------------------------
# a7: part 5 (V73:906-930)
# ------------------------------------------------------------------------------------------------------------------------
def test_gather_cells_matches_numpy(engine):
    import torch
    rng = np.random.default_rng(8)
    nrow, ncol, stride = 300, 421, 448
    ras = torch.from_numpy(rng.standard_normal((nrow, stride))).to("cuda:0")
    row = rng.integers(0, nrow, 1000).astype(np.int32)
    col = rng.integers(0, ncol, 1000).astype(np.int32)
    row[:4], col[:4] = [-1, 0, nrow, 5], [-1, ncol, 3, -7]              # points outside the raster: extract gives NA
    got = engine.gather_cells_dev(ras.data_ptr(), stride, nrow, ncol, row, col)
    want = ras.cpu().numpy()[np.clip(row, 0, nrow - 1), np.clip(col, 0, ncol - 1)]
    want[:4] = np.nan
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("tps_helps", [True, False])
def test_part5_keeps_the_better_r2(engine, tps_helps):
    """V73:910-930: final = ensemble + TPS only if its R^2 at the points beats the ensemble's; the residuals are the
    TPS-corrected ones either way (V73:913 overwrites l$residuals before the test).
    A smoothing spline never increases the RSS at its own knots, so the TPS can only lose through the reference's weight quirk
    (V73:337, 619-620): the divisor is the UNROUNDED total over ALL candidates, so with a dropped model the kept weights sum to
    s < 1, res.FINAL = s (resp - f) and ensemble + TPS ~ s resp at the points - off by (1 - s) resp."""
    geom = synth.make_geom(192, 256)
    C = 2
    cov = synth.covariate_planes(geom, C, nan_frac=0.0)
    xy, krow, kcol = synth.make_knots(geom, 400, 12)
    X = np.column_stack([cov[:, krow, kcol].T.astype(np.float64), xy])
    coef = np.array([200.0, 0.02, -0.01, 3.0, -2.0])
    lin = coef[0] + X @ coef[1:]
    rng = np.random.default_rng(4)
    if tps_helps:
        models, letters, p = {"g": {"coef": coef}}, "g", np.array([1.0])
        resp = lin + 2.0 * np.sin(6 * xy[:, 0]) * np.cos(4 * xy[:, 1])      # smooth spatial residual: the TPS recovers it
    else:
        nn = synth.make_models(geom, C, 300, 9, kept="n")["n"]
        models, letters, p = {"g": {"coef": coef}, "n": nn}, "gn", np.array([0.96, 0.04])   # nnet dropped: 0.04 < 0.05 * 1.0
        resp = lin + 0.01 * rng.standard_normal(lin.size)                   # the ensemble alone is (almost) exact
    kept, w, wt = om.select_models(p, letters=letters)
    assert kept == "g"
    ref = oml.mltps_one_response(geom.as_tuple(), cov, xy, resp, models, kept, w, wt, tile_px=1500)
    got = mm.mltps_response(engine, geom, cov, xy, resp, models, p, letters=letters, tile_px=1500)
    assert ref["tps_kept"] == tps_helps
    assert got["summary"]["tps_kept"] == tps_helps
    assert abs(got["summary"]["rsq_model"] - ref["rsq_model"]) < 1e-9
    assert abs(got["summary"]["rsq_final"] - ref["rsq_final"]) < 1e-6 * max(1.0, abs(ref["rsq_final"]))
    assert relerr(got["final"], ref["final"]) < 2e-6
    scale = np.max(np.abs(ref["final"]))
    assert np.max(np.abs(got["residuals"] - ref["residuals"])) < 2e-6 * scale
    if not tps_helps:
        x, yv = otps.cell_centres(geom.as_tuple(), np.arange(geom.nrow), np.arange(geom.ncol))
        ens_only = 0.96 * (coef[0] + cov[0] * coef[1] + cov[1] * coef[2] + coef[3] * x[None, :] + coef[4] * yv[:, None])
        assert np.max(np.abs(got["final"] - ens_only)) < 1e-9 * scale      # the final raster is the ensemble alone
