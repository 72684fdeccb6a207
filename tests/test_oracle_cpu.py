"""CPU tests of the oracle itself: analytic known-answer tests of the thin-plate smoothing spline
(the reference has no tests and ships no expected outputs - SURVEY.md 4, 8c), the committed golden
fixture generated from the reference's bundled inputs, and the tiling / weight-rule restatements."""
import os

import numpy as np
import pytest

from oracle import cbind, models as om, tiles as otl, tps as otps

GOLD = os.path.join(os.path.dirname(__file__), "golden", "bundled_c1.npz")


@pytest.fixture(scope="module")
def cloud():
    rng = np.random.default_rng(0)
    n = 400
    xy = rng.uniform(0, 1, (n, 2)) * [3.0, 2.0] + [10, -5]
    y = np.sin(2 * xy[:, 0]) * np.cos(3 * xy[:, 1]) + 0.1 * rng.standard_normal(n)
    return xy, y, otps.tps_fit(xy, y)


def test_system_identities(cloud):
    xy, y, f = cloud
    s = f.knots_s
    T = np.column_stack([np.ones(len(s)), s])
    K = otps.rad_cov(s, s)
    assert np.abs(T.T @ f.c).max() < 1e-8 * np.abs(f.c).sum()                 # T'c = 0
    assert np.abs(K @ f.c + f.lam * f.c + T @ f.d - y).max() < 1e-10          # (K + lambda I)c + Td = y
    p = otps.tps_predict_points(f, xy)
    assert np.abs(p - (y - f.lam * f.c)).max() < 1e-10                        # f(x_i) = y_i - lambda c_i
    assert abs(f.eff_df - (3 + np.sum(f.eta / (f.eta + f.lam)))) < 1e-9       # trA
    assert np.all(np.diff(f.gcv_grid[:, 0]) > 0) and np.all(np.diff(f.gcv_grid[:, 1]) < 0)   # trA monotone in lambda
    assert not f.gcv_at_endpoint and f.gcv_grid.shape == (200, 3)
    assert abs(f.gcv_grid[:, 1].max() - 0.95 * len(s)) < 1e-3 and abs(f.gcv_grid[:, 1].min() - 3.001) < 1e-4


def test_affine_is_reproduced_exactly(cloud):
    xy, _, _ = cloud
    ya = 2.0 + 3.0 * xy[:, 0] - 0.5 * xy[:, 1]
    f = otps.tps_fit(xy, ya, lam=1e-3)
    assert np.abs(f.c).max() < 1e-7
    q = np.array([[11.0, -4.0], [12.5, -3.2]])
    assert np.allclose(otps.tps_predict_points(f, q), 2.0 + 3.0 * q[:, 0] - 0.5 * q[:, 1], atol=1e-9)


def test_lambda_limits(cloud):
    xy, y, _ = cloud
    f0 = otps.tps_fit(xy, y, lam=1e-12)                                       # lambda -> 0 interpolates
    assert np.abs(otps.tps_predict_points(f0, xy) - y).max() < 1e-5
    finf = otps.tps_fit(xy, y, lam=1e9)                                       # lambda -> inf: least-squares plane
    A = np.column_stack([np.ones(len(xy)), xy])
    plane = A @ np.linalg.lstsq(A, y, rcond=None)[0]
    assert np.abs(otps.tps_predict_points(finf, xy) - plane).max() < 1e-5


def test_axis_rescale_invariance(cloud):
    """scale.type = "range" makes the fit invariant to per-axis affine maps of the coordinates."""
    xy, y, f = cloud
    xy2 = xy * [7.0, 0.01] + [100.0, 3.0]
    f2 = otps.tps_fit(xy2, y)
    assert abs(f2.lam - f.lam) < 1e-9 * f.lam
    assert np.abs(f2.c - f.c).max() < 1e-6 * np.abs(f.c).max()


def test_replicates_are_pooled(cloud):
    xy, y, _ = cloud
    xy2 = np.vstack([xy, xy[:50]])
    y2 = np.concatenate([y, y[:50] + 0.2])
    f = otps.tps_fit(xy2, y2)
    assert f.knots_s.shape[0] == len(xy) and f.n_obs == len(xy2)
    assert abs(f.pure_ss - 50 * 2 * 0.1 ** 2) < 1e-9 and f.weights.sum() == len(xy2)


def test_knot_on_cell_centre_is_finite():
    geom = (0.0, 1.0, 0.0, 1.0, 32, 32)
    rng = np.random.default_rng(1)
    cells = rng.choice(32 * 32, 40, replace=False)
    x, yv = otps.cell_centres(geom, cells // 32, cells % 32)
    xy = np.column_stack([x, yv])
    f = otps.tps_fit(xy, rng.standard_normal(40), lam=1e-3)
    ras = otps.tps_interpolate(f, geom)
    assert np.isfinite(ras).all()
    assert np.allclose(ras[cells // 32, cells % 32], otps.tps_predict_points(f, xy), atol=1e-12)


def test_c_restatement_matches_numpy(cloud):
    xy, y, f = cloud
    geom = (10.0, 13.0, -5.0, -3.0, 70, 90)
    a = otps.tps_interpolate(f, geom, 5, 60, 7, 80)
    b = cbind.tps_eval(f, geom, (5, 60, 7, 80))
    assert np.abs(a - b).max() < 1e-11 * np.abs(a).max()


# ---- golden fixture from the reference's bundled inputs -----------------------------------------
def test_bundled_data_tiling_and_fits():
    z = np.load(GOLD)
    geom = tuple(float(v) if i < 4 else int(v) for i, v in enumerate(z["geom"]))
    assert geom[4:] == (2476, 3264)
    knots_xy, krow, kcol = otl.knot_coordinates(geom, z["points"][:, :2])
    assert np.array_equal(krow, z["krow"]) and np.array_equal(kcol, z["kcol"])
    assert np.unique(knots_xy, axis=0).shape[0] == 813                        # no duplicates
    lay = otl.mltps_tile_layout(geom)
    assert (lay.nRx, lay.nCx) == (2, 3)                                       # V73:651-661
    counts = []
    for fw in lay.fit_win:
        counts.append(int(((krow >= fw[0]) & (krow < fw[1]) & (kcol >= fw[2]) & (kcol < fw[3])).sum()))
    # numbers derived independently in the survey from V73:651-673
    assert counts == [190, 220, 199, 200, 237, 204] == list(z["tile_counts"])
    assert np.array_equal(np.array(lay.fit_win), z["fit_win"]) and np.array_equal(np.array(lay.keep_win), z["keep_win"])
    # one tile refit against the pinned oracle values
    fw = lay.fit_win[4]
    ins = (krow >= fw[0]) & (krow < fw[1]) & (kcol >= fw[2]) & (kcol < fw[3])
    f = otps.tps_fit(knots_xy[ins], z["points"][ins, 2])
    pin = z["tile_fits"][4]
    assert abs(f.lam - pin[0]) < 1e-8 * pin[0] and abs(f.eff_df - pin[1]) < 1e-6
    assert np.allclose(f.d, pin[2:5], rtol=1e-7) and abs(np.abs(f.c).sum() - pin[5]) < 1e-6 * pin[5]


# ---- tiling / feather ---------------------------------------------------------------------------------
def test_feather_is_a_partition_of_unity():
    geom = (0.0, 3.0, 0.0, 2.0, 200, 300)
    lay = otl.mltps_tile_layout(geom, tile_px=100)
    tiles = [np.full((w[1] - w[0], w[3] - w[2]), 5.0) for w in lay.keep_win]
    out = otl.feather_merge(geom, lay.keep_win, tiles, lay.nCx, lay.nRx)
    assert np.allclose(out, 5.0) and not np.isnan(out).any()


def test_feather_weights_on_a_vertical_seam():
    geom = (0.0, 2.0, 0.0, 1.0, 10, 200)
    lay = otl.mltps_tile_layout(geom, tile_px=100)
    assert (lay.nRx, lay.nCx) == (1, 2)
    a, b = lay.keep_win
    tiles = [np.zeros((a[1] - a[0], a[3] - a[2])), np.ones((b[1] - b[0], b[3] - b[2]))]
    out = otl.feather_merge(geom, lay.keep_win, tiles, 2, 1)
    strip = out[0, b[2]:a[3]]
    assert strip[0] == 0.0 and strip[-1] == 1.0 and np.allclose(np.diff(strip), 1.0 / (len(strip) - 1))
    assert np.all(out[0, :b[2]] == 0) and np.all(out[0, a[3]:] == 1)


def test_horizontal_seam_uses_the_southern_tile_as_tile_one():
    geom = (0.0, 1.0, 0.0, 2.0, 200, 10)
    lay = otl.mltps_tile_layout(geom, tile_px=100)
    assert (lay.nRx, lay.nCx) == (2, 1)
    south, north = lay.keep_win                      # order: j = 1 (south) first, V73:670
    assert south[0] > north[0]
    tiles = [np.zeros((south[1] - south[0], 10)), np.ones((north[1] - north[0], 10))]
    out = otl.feather_merge(geom, lay.keep_win, tiles, 1, 2)
    strip = out[south[0]:north[1], 0]
    assert strip[0] == 1.0 and strip[-1] == 0.0      # north end -> north tile, south end -> south tile


def test_crop_window_snaps_to_nearest_edge():
    geom = (0.0, 10.0, 0.0, 5.0, 50, 100)
    assert otl.crop_window(geom, (0.26, 9.74, 0.0, 5.0)) == (0, 50, 3, 97)
    assert otl.crop_window(geom, (-3.0, 20.0, 1.04, 3.96)) == (10, 40, 0, 100)


def test_tiles_create_matches_v73_arithmetic():
    geom = (0.0, 3.0, 0.0, 3.0, 300, 300)
    pts = np.array([[0.1, 0.1], [1.24, 0.5], [1.26, 0.5], [2.9, 2.9]])
    tc = otl.tiles_create(geom, pts, out_ncol=3, out_nrow=3, feather_d=50)
    t0 = tc["tiles"][0]
    assert np.allclose(t0["ext"], (-0.25, 1.25, -0.25, 1.25))                 # +- feather.d/2 pixels, V73:1170,1195
    assert t0["win"] == (175, 300, 0, 125)                                    # south-west tile first
    assert list(t0["points"]) == [0, 1] and list(tc["tiles"][1]["points"]) == [1, 2]
    assert list(tc["tiles"][8]["points"]) == [3]


# ---- ensemble weight rule / RSS objective --------------------------------------------------------------
def test_weight_rule_and_rss():
    p = np.array([0.5, 0.024, 0.026, 0.3, 0.0, 0.149])
    kept, w, tot = om.select_models(p)
    cut = 0.05 * p.sum()
    assert kept == "bmv" and np.allclose(w, [0.5, 0.3, 0.15]) and tot == p.sum() and cut > 0.03
    rng = np.random.default_rng(2)
    R = rng.standard_normal((900, 6))
    k = rng.uniform(0, 1, 6)
    assert abs(om.rss_objective(k, R) - om.rss_from_gram(k, om.gram(R))) < 1e-9
    assert abs(om.rss_objective(3 * k, R) - om.rss_objective(k, R)) < 1e-9    # homogeneous of degree 0


# ---- independent anchors: third-party implementations of the same mathematics ---------------------------
def test_tps_solve_matches_scipy_rbf_interpolator():
    """scipy.interpolate.RBFInterpolator(kernel='thin_plate_spline', degree=1) solves the same
    smoothing-spline system with phi = r^2 log r = 8 pi * fields' Rad.cov, so smoothing = 8 pi lambda gives
    the same surface.  Pins the oracle's (K + lambda I) c + T d = y restatement and predict.Krig on an
    implementation we did not write (parity with R fields itself stays unpinned)."""
    from scipy.interpolate import RBFInterpolator
    rng = np.random.default_rng(42)
    xy = rng.uniform(0, 1, (400, 2)) * [3.0, 1.5] + [10.0, -4.0]
    y = np.sin(2 * xy[:, 0]) * np.cos(3 * xy[:, 1]) + 0.05 * rng.standard_normal(400)
    pts = rng.uniform(0, 1, (500, 2)) * [3.0, 1.5] + [10.0, -4.0]
    for lam in (1e-2, 1e-4, 1e-7):
        fit = otps.tps_fit(xy, y, lam=lam)
        s = (xy - fit.center) / fit.scale                     # Krig's transformx, scale.type = "range"
        rbf = RBFInterpolator(s, y, kernel="thin_plate_spline", smoothing=8 * np.pi * lam, degree=1)
        ref = rbf((pts - fit.center) / fit.scale)
        got = otps.tps_predict_points(fit, pts)
        assert np.max(np.abs(got - ref)) < 1e-7 * np.max(np.abs(ref))


def test_model_descriptors_match_scikit_learn_predict():
    """the flat randomForest / gbm / ksvm descriptors + the oracle's predictors reproduce the predict() of
    the scikit-learn estimators they were exported from (same tree walks, same RBF expansion)."""
    from machisplin_b200 import synth
    from oracle import models as om
    geom = synth.make_geom(96, 128)
    m = synth.make_models(geom, 3, 400, 3, kept="brv", rf_trees=20, gbm_trees=30, keep_estimators=True)
    sk = m["_sk"]
    rng = np.random.default_rng(1)
    X = sk["X"][rng.integers(0, len(sk["X"]), 300)] + rng.normal(0, 0.3, (300, 5)) * sk["X"].std(0)
    X[:, :3] = X[:, :3].astype(np.float32)                    # raster values are float32
    np.testing.assert_allclose(om.predict_rf(m["r"], X), sk["r"].predict(X), rtol=1e-12)
    np.testing.assert_allclose(om.predict_gbm(m["b"], X), sk["b"].predict(X), rtol=1e-10)
    v = m["v"]
    ref = sk["v"].predict((X - v["x_center"]) / v["x_scale"]) * v["y_scale"] + v["y_center"]
    np.testing.assert_allclose(om.predict_svm_exact(v, X), ref, rtol=1e-9)


def test_smooth_model_descriptors_match_scikit_learn_predict():
    """gam (purely parametric formula = a linear model, V73:600-606) and nnet(size = 10, linout = TRUE) (V73:463-470): the oracle's
    predictors on descriptors exported from fitted scikit-learn estimators reproduce those estimators' predict() - LinearRegression
    and MLPRegressor(logistic hidden layer, identity output), weights re-ordered into nnet's `wts` layout (per hidden unit: bias,
    inputs 1..P; then output bias, hidden 1..H)."""
    import warnings
    from sklearn.linear_model import LinearRegression
    from sklearn.neural_network import MLPRegressor
    from oracle import models as om
    rng = np.random.default_rng(5)
    n, P, H = 400, 5, 10
    X = rng.standard_normal((n, P)) * [3.0, 0.5, 10.0, 1.0, 1.0] + [100.0, 2.0, -40.0, -77.0, -6.0]
    y = 0.02 * X[:, 0] - 1.5 * X[:, 1] + np.tanh(0.1 * X[:, 2] + 4.0) + 0.1 * rng.standard_normal(n)
    Xq = X[:150] + 0.2 * rng.standard_normal((150, P))
    lr = LinearRegression().fit(X, y)
    g = {"coef": np.concatenate([[lr.intercept_], lr.coef_])}
    np.testing.assert_allclose(om.predict_gam(g, Xq), lr.predict(Xq), rtol=1e-12, atol=1e-12)
    # nnet is fitted on the response scaled to [0, 1] and un-scaled after predict (V73:455-470)
    mn, mx = y.min(), y.max()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mlp = MLPRegressor(hidden_layer_sizes=(H,), activation="logistic", solver="lbfgs", max_iter=60, random_state=0)
        mlp.fit((X - X.mean(0)) / X.std(0), (y - mn) / (mx - mn))
    # fold the input standardisation into the first layer: the reference feeds unscaled covariates
    W1 = mlp.coefs_[0] / X.std(0)[:, None]                               # (P, H)
    b1 = mlp.intercepts_[0] - (X.mean(0) / X.std(0)) @ mlp.coefs_[0]
    wts = np.concatenate([np.column_stack([b1, W1.T]).ravel(), mlp.intercepts_[1], mlp.coefs_[1].ravel()])
    nn = {"wts": wts, "H": H, "max2": mx - mn, "min": mn}
    ref = mlp.predict((Xq - X.mean(0)) / X.std(0)) * (mx - mn) + mn
    np.testing.assert_allclose(om.predict_nnet(nn, Xq), ref, rtol=1e-10, atol=1e-10)


def test_gcv_criterion_against_the_brute_force_hat_matrix():
    """Independent of the eigen-decomposition the oracle (and the engine) search on: the smoother matrix A(lambda) from the
    full block system [[K + lambda I, T], [T', 0]], trA = trace(A), GCV = (RSS / n) / (1 - trA / n)^2 (Krig.fgcv with cost 1).
    The oracle's eff.df and GCV value agree with it, and the lambda it selects minimises the brute-force criterion."""
    rng = np.random.default_rng(21)
    n = 160
    xy = rng.uniform(0, 1, (n, 2)) * [2.0, 1.0] + [3.0, 40.0]
    y = np.sin(3 * xy[:, 0]) * np.cos(5 * xy[:, 1]) + 0.15 * rng.standard_normal(n)
    fit = otps.tps_fit(xy, y)
    s = fit.knots_s
    K = otps.rad_cov(s, s)
    T = np.column_stack([np.ones(n), s])

    def brute(lam):
        S = np.block([[K + lam * np.eye(n), T], [T.T, np.zeros((3, 3))]])
        sol = np.linalg.solve(S, np.vstack([np.eye(n), np.zeros((3, n))]))
        A = K @ sol[:n] + T @ sol[n:]                  # fitted values = A y
        r = y - A @ y
        tra = np.trace(A)
        return float(r @ r) / n / (1.0 - tra / n) ** 2, tra

    D = np.concatenate([np.zeros(3), 1.0 / fit.eta])
    g_star, tra_star = brute(fit.lam)
    assert abs(tra_star - fit.eff_df) < 1e-7 * fit.eff_df
    assert abs(g_star - otps.gcv_value(fit.lam, D, fit.u, fit.n_obs, fit.pure_ss)) < 1e-8 * g_star
    assert not fit.gcv_at_endpoint
    for f in (0.5, 0.7, 0.85, 0.95, 1.05, 1.15, 1.3, 2.0):
        assert brute(fit.lam * f)[0] >= g_star * (1 - 1e-9), f
