"""GPU parity: TPS fit + evaluation through the C ABI vs the float64 oracle."""
import numpy as np
import pytest

from machisplin_b200 import synth
from oracle import tps as otps

pytestmark = pytest.mark.gpu

TOL = 1e-5   # north_star: <= 1e-5 relative on predicted cell values (relative to max |ref| over the raster)


def relerr(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def _case(nrow, ncol, n, seed, lam=None):
    geom = synth.make_geom(nrow, ncol)
    xy, _, _ = synth.make_knots(geom, n, seed)
    y = synth.residual_field(xy, seed)
    fit = otps.tps_fit(xy, y, lam=lam)
    return geom, xy, y, fit


@pytest.mark.parametrize("method", ["direct", "fast"])
@pytest.mark.parametrize("shape", [(256, 256), (200, 333), (97, 64)])
def test_eval_from_oracle_coefficients(engine, method, shape):
    geom, xy, y, fit = _case(shape[0], shape[1], 300, 5)
    sp = engine.spline_create(fit.knots_xy, fit.c, fit.d, fit.center, fit.scale)
    got = engine.tps_eval(sp, geom, method=method)
    ref = otps.tps_interpolate(fit, geom.as_tuple())
    assert relerr(got, ref) < (1e-11 if method == "direct" else 1e-6)


@pytest.mark.parametrize("p", [8, 10, 12, 14, 16])
def test_fast_eval_orders(engine, p):
    geom, xy, y, fit = _case(512, 512, 1000, 11)
    sp = engine.spline_create(fit.knots_xy, fit.c, fit.d, fit.center, fit.scale)
    engine.set_fast_eval_params(cheb_p=p)
    try:
        got = engine.tps_eval(sp, geom, method="fast")
    finally:
        engine.set_fast_eval_params()
    ref = otps.tps_interpolate(fit, geom.as_tuple())
    bound = {8: 3e-6, 10: 2e-7, 12: 1e-8, 14: 1e-9, 16: 1e-9}[p]
    assert relerr(got, ref) < bound


def test_fast_eval_near_interpolating(engine):
    """lambda -> 0: sum|c| / max|f| ~ 1e7; the automatic order selection must still hold 1e-5."""
    geom, xy, y, fit = _case(512, 512, 1000, 11, lam=1e-8)
    sp = engine.spline_create(fit.knots_xy, fit.c, fit.d, fit.center, fit.scale)
    got = engine.tps_eval(sp, geom, method="fast")
    ref = otps.tps_interpolate(fit, geom.as_tuple())
    assert relerr(got, ref) < 1e-6


def test_window_eval_matches_full(engine):
    geom, xy, y, fit = _case(300, 280, 200, 3)
    sp = engine.spline_create(fit.knots_xy, fit.c, fit.d, fit.center, fit.scale)
    win = (37, 251, 19, 200)
    ref = otps.tps_interpolate(fit, geom.as_tuple(), *win)
    for method in ("direct", "fast"):
        got = engine.tps_eval(sp, geom, window=win, method=method)
        assert got.shape == ref.shape
        assert relerr(got, ref) < 1e-6


def test_r_zero_cells_are_finite(engine):
    """knots sit on cell centres (V73:145): d2 = 0 hits the 1e-20 clamp of radfun."""
    geom, xy, y, fit = _case(64, 64, 60, 2)
    sp = engine.spline_create(fit.knots_xy, fit.c, fit.d, fit.center, fit.scale)
    for method in ("direct", "fast"):
        got = engine.tps_eval(sp, geom, method=method)
        assert np.isfinite(got).all()


@pytest.mark.parametrize("n", [60, 500, 1500])
def test_fit_gcv_matches_oracle(engine, n):
    geom = synth.make_geom(1024, 1024)
    xy, _, _ = synth.make_knots(geom, n, 21 + n)
    y = synth.residual_field(xy, 21 + n)
    ref = otps.tps_fit(xy, y)
    sp = engine.tps_fit(xy, y)
    assert abs(sp.lam - ref.lam) <= 1e-6 * ref.lam
    assert abs(sp.eff_df - ref.eff_df) <= 1e-6 * ref.eff_df
    np.testing.assert_allclose(sp.d, ref.d, rtol=1e-7, atol=1e-9 * np.abs(ref.d).max())
    assert np.max(np.abs(sp.c - ref.c)) <= 1e-7 * np.max(np.abs(ref.c))
    eta, (dg, of, zh) = sp.decomposition()
    np.testing.assert_allclose(eta, ref.eta, rtol=1e-7, atol=1e-13 * ref.eta.max())
    # the tridiagonal form is orthogonally similar to Q2'KQ2: same spectrum, same |z|
    tri = np.diag(dg) + np.diag(of, 1) + np.diag(of, -1)
    np.testing.assert_allclose(np.linalg.eigvalsh(tri)[::-1], ref.eta, rtol=1e-7, atol=1e-13 * ref.eta.max())
    assert abs(np.linalg.norm(zh) - np.linalg.norm(ref.u[3:])) <= 1e-10 * np.linalg.norm(zh)
    # predictions at the knots: f(x_i) = y_i - lambda c_i
    f = engine.tps_predict_points(sp, xy)
    assert np.max(np.abs(f - (y - sp.lam * sp.c))) < 1e-9 * max(1.0, np.abs(y).max())


@pytest.mark.parametrize("lam", [1e-2, 1e-5])
def test_fit_fixed_lambda_cholesky(engine, lam):
    geom = synth.make_geom(1024, 1024)
    xy, _, _ = synth.make_knots(geom, 700, 77)
    y = synth.residual_field(xy, 77)
    ref = otps.tps_fit(xy, y, lam=lam)
    sp = engine.tps_fit(xy, y, lam=lam)
    assert np.max(np.abs(sp.c - ref.c)) <= 1e-7 * np.max(np.abs(ref.c))
    np.testing.assert_allclose(sp.d, ref.d, rtol=1e-7, atol=1e-9 * np.abs(ref.d).max())


def test_fit_replicates_pooled(engine):
    geom = synth.make_geom(256, 256)
    xy, _, _ = synth.make_knots(geom, 120, 9)
    y = synth.residual_field(xy, 9)
    xy2 = np.vstack([xy, xy[:30]])
    y2 = np.concatenate([y, y[:30] + 0.05])
    ref = otps.tps_fit(xy2, y2)
    sp = engine.tps_fit(xy2, y2)
    assert sp.np == ref.knots_xy.shape[0] == 120
    assert abs(sp.lam - ref.lam) <= 1e-6 * ref.lam
    # knot order differs only by a permutation: compare through predictions
    pts = xy[:50] + 1e-3
    assert np.max(np.abs(engine.tps_predict_points(sp, pts) - otps.tps_predict_points(ref, pts))) < 1e-8


def test_fit_multi_response_shares_decomposition(engine):
    geom = synth.make_geom(512, 512)
    xy, _, _ = synth.make_knots(geom, 400, 13)
    Y = synth.residual_field(xy, 13, L=4)
    sps = engine.tps_fit(xy, Y)
    for k, sp in enumerate(sps):
        ref = otps.tps_fit(xy, Y[:, k])
        assert abs(sp.lam - ref.lam) <= 1e-6 * ref.lam
        assert np.max(np.abs(sp.c - ref.c)) <= 1e-7 * np.max(np.abs(ref.c))


def test_errors_are_reported_not_thrown(engine):
    import machisplin_b200._lib as L
    xy = np.zeros((10, 2))
    with pytest.raises(L.MbError):
        engine.tps_fit(xy, np.zeros(10))          # zero range
    with pytest.raises(L.MbError):
        engine.tps_fit(np.column_stack([np.arange(10.0), np.arange(10.0)]), np.zeros(10))   # collinear


@pytest.mark.parametrize("mode", [1, 2])
def test_leaf_precision_paths(engine, mode):
    """float64-only and forced mixed-precision leaf paths against the oracle (automatic mode picks per launch)."""
    geom, xy, y, fit = _case(768, 1024, 1200, 31)
    sp = engine.spline_create(fit.knots_xy, fit.c, fit.d, fit.center, fit.scale)
    engine.set_param("eval_precision", mode)
    try:
        got = engine.tps_eval(sp, geom, method="fast")
    finally:
        engine.set_param("eval_precision", 0)
    ref = otps.tps_interpolate(fit, geom.as_tuple())
    assert relerr(got, ref) < (3e-7 if mode == 1 else 1e-6)


def test_mixed_path_is_refused_when_inaccurate(engine):
    """near-interpolating spline on a coarse grid: the device-side estimate must fall back to float64."""
    geom, xy, y, fit = _case(96, 96, 400, 12, lam=1e-9)
    sp = engine.spline_create(fit.knots_xy, fit.c, fit.d, fit.center, fit.scale)
    got = engine.tps_eval(sp, geom, method="fast")
    ref = otps.tps_interpolate(fit, geom.as_tuple())
    assert relerr(got, ref) < 1e-6


@pytest.mark.parametrize("n", [35, 200, 1100])
def test_fit_gcv_tridiagonal_paths_agree(engine, n):
    """The default GCV fit (two-stage tridiagonalisation: compact-WY band reduction + bulge chasing, then Sturm bisection
    and the Cholesky at the selected lambda) against its alternates: the one-stage Householder reduction as one
    persistent kernel (sytrd_mode = 1) or one kernel per phase (sytrd_mode = 2), and the cuSOLVER eigenvector path
    (eigen_impl = 1)."""
    geom = synth.make_geom(1024, 1024)
    xy, _, _ = synth.make_knots(geom, n, 300 + n)
    y = synth.residual_field(xy, 300 + n)
    sp0 = engine.tps_fit(xy, y)
    try:
        engine.set_param("sytrd_mode", 3)
        sp3 = engine.tps_fit(xy, y)
        engine.set_param("sytrd_mode", 1)
        spp = engine.tps_fit(xy, y)
        engine.set_param("sytrd_mode", 2)
        sp2 = engine.tps_fit(xy, y)
        engine.set_param("sytrd_mode", 0)
        engine.set_param("eigen_impl", 1)
        sp1 = engine.tps_fit(xy, y)
    finally:
        engine.set_param("sytrd_mode", 0)
        engine.set_param("eigen_impl", 0)
    # every path is deterministic (no atomics on data, fixed summation orders): repeating the default reproduces it bit
    # for bit, and the two schedules of the one-stage algorithm are bit-identical to each other
    assert sp3.lam == sp0.lam
    np.testing.assert_array_equal(sp3.c, sp0.c)
    np.testing.assert_array_equal(sp3.decomposition()[0], sp0.decomposition()[0])
    assert sp2.lam == spp.lam
    np.testing.assert_array_equal(sp2.c, spp.c)
    np.testing.assert_array_equal(sp2.decomposition()[0], spp.decomposition()[0])
    for alt in (spp, sp1):
        assert abs(alt.lam - sp0.lam) <= 1e-6 * sp0.lam
        assert np.max(np.abs(alt.c - sp0.c)) <= 1e-7 * np.max(np.abs(sp0.c))
        np.testing.assert_allclose(alt.decomposition()[0], sp0.decomposition()[0], rtol=1e-7,
                                   atol=1e-13 * sp0.decomposition()[0].max())


@pytest.mark.parametrize("n", [20, 36, 37, 70, 165, 600])
def test_fit_two_stage_edge_sizes(engine, n):
    """Two-stage tridiagonalisation at the sizes where its structure changes: no panel at all (m <= 33: bulge chasing on
    the dense matrix read as a band), one short panel (m = 34), panels with fewer rows than columns, ragged last row block,
    three right-hand sides carried through both stages.  Reference: the one-stage reduction."""
    geom = synth.make_geom(512, 512)
    xy, _, _ = synth.make_knots(geom, n, 900 + n)
    y = synth.residual_field(xy, 900 + n)
    Y = np.stack([y, y[::-1].copy(), y * y], axis=1)
    got = engine.tps_fit(xy, Y)
    try:
        engine.set_param("sytrd_mode", 1)
        ref = engine.tps_fit(xy, Y)
    finally:
        engine.set_param("sytrd_mode", 0)
    for g, r in zip(got, ref):
        assert abs(g.lam - r.lam) <= 1e-8 * r.lam
        assert np.max(np.abs(g.c - r.c)) <= 1e-8 * np.max(np.abs(r.c))
        eg, (dg, og, zg) = g.decomposition()
        er, (dr, orr, zr) = r.decomposition()
        np.testing.assert_allclose(eg, er, rtol=0, atol=1e-13 * er.max())
        assert abs(np.linalg.norm(zg) - np.linalg.norm(zr)) <= 1e-12 * np.linalg.norm(zr)


@pytest.mark.parametrize("mode", [0, 2])
def test_leaf_dense_knots_overflow_near_lists(engine, mode):
    """Many knots per leaf box (a 3 x 3 neighbourhood holds far more than the 15 inline entries of a leaf record):
    the overflow near list, partial boxes at both window edges (97 x 161 cells) and, with mode 2, the forced
    mixed-precision kernel."""
    geom, xy, y, fit = _case(97, 161, 700, 41, lam=3e-3)
    sp = engine.spline_create(fit.knots_xy, fit.c, fit.d, fit.center, fit.scale)
    engine.set_param("eval_precision", mode)
    try:
        got = engine.tps_eval(sp, geom, method="fast")
        win = (5, 90, 33, 150)
        gotw = engine.tps_eval(sp, geom, window=win, method="fast")
    finally:
        engine.set_param("eval_precision", 0)
    ref = otps.tps_interpolate(fit, geom.as_tuple())
    tol = 2e-6 if mode == 0 else TOL
    assert relerr(got, ref) < tol
    assert np.max(np.abs(gotw - ref[win[0]:win[1], win[2]:win[3]])) < tol * np.max(np.abs(ref))


@pytest.mark.parametrize("strip", ["x", "y"])
def test_leaf_elongated_knot_cloud(engine, strip):
    """Knots confined to a strip: scale.type = "range" stretches one axis, so cells are far from square in scaled
    coordinates and the leaf boxes are 32 columns x bh rows with bh = 4 (strip along x) or ~ 107 (strip along y):
    the generic row loop and the mirror-row collapse of k_leaf_stream for bh != 32."""
    geom = synth.make_geom(256, 512)
    xy, _, _ = synth.make_knots(geom, 6000, 47)
    if strip == "x":
        keep = (xy[:, 1] > 0.1) & (xy[:, 1] < 0.2)          # y range 0.1 of a 0.5-high raster, x range 1
    else:
        keep = (xy[:, 0] > 0.2) & (xy[:, 0] < 0.5)          # x range 0.3, y range 0.5
    xy = xy[keep][:400]
    y = synth.residual_field(xy, 47)
    fit = otps.tps_fit(xy, y, lam=1e-3)
    sp = engine.spline_create(fit.knots_xy, fit.c, fit.d, fit.center, fit.scale)
    ref = otps.tps_interpolate(fit, geom.as_tuple())
    for mode in (0, 2):
        engine.set_param("eval_precision", mode)
        try:
            got = engine.tps_eval(sp, geom, method="fast")
        finally:
            engine.set_param("eval_precision", 0)
        assert relerr(got, ref) < (2e-6 if mode == 0 else TOL)


@pytest.mark.parametrize("n", [20, 36, 37, 70, 165, 600, 1100])
def test_coefficients_from_band_form(engine, n):
    """(M + lambda I)^-1 z from the band form of the two-stage reduction (block band Cholesky + back-transformation by the stored
    panel reflectors, tools/proto_two_stage.py coefficients_from_band; the default) against the dense Cholesky (coef_impl = 2):
    same lambda (the search does not depend on it), coefficients to 1e-9, three responses."""
    geom = synth.make_geom(512, 512)
    xy, _, _ = synth.make_knots(geom, n, 700 + n)
    y = synth.residual_field(xy, 700 + n)
    Y = np.stack([y, y[::-1].copy(), y * y], axis=1)
    try:
        engine.set_param("coef_impl", 1)              # the band form also where the default would refuse it (cond > 1e8)
        got = engine.tps_fit(xy, Y)
        engine.set_param("coef_impl", 2)
        ref = engine.tps_fit(xy, Y)
    finally:
        engine.set_param("coef_impl", 0)
    for g, r in zip(got, ref):
        assert g.lam == r.lam
        assert np.max(np.abs(g.c - r.c)) <= 1e-9 * np.max(np.abs(r.c))
        assert np.max(np.abs(g.d - r.d)) <= 1e-9 * max(np.max(np.abs(r.d)), 1e-300)


@pytest.mark.parametrize("n", [36, 70, 200, 1100])
def test_chase_variants_are_bit_identical(engine, n):
    """The three bulge-chase kernels - flags with three warps (2), flags with watcher / publisher warps (1), tagged band elements
    with loader warps (3, the default of a stand-alone fit) - do the same arithmetic in the same order: bit-identical lambda,
    coefficients and eigenvalues, for one and for three responses."""
    geom = synth.make_geom(512, 512)
    xy, _, _ = synth.make_knots(geom, n, 800 + n)
    y = synth.residual_field(xy, 800 + n)
    Y = np.stack([y, y[::-1].copy(), y * y], axis=1)
    if n == 200:                                   # 20 responses: more right-hand sides than one batch of tagged loads
        Y = np.concatenate([Y] + [np.roll(Y, 7 * (j + 1), axis=0) * (1.0 + 0.1 * j) for j in range(6)], axis=1)[:, :20]
    res = {}
    try:
        for impl in (2, 1, 3, 0):
            engine.set_param("sbr_chase_impl", impl)
            res[impl] = (engine.tps_fit(xy, y), engine.tps_fit(xy, Y))
    finally:
        engine.set_param("sbr_chase_impl", 0)
    ref, ref3 = res[2]
    for impl in (1, 3, 0):
        got, got3 = res[impl]
        assert got.lam == ref.lam
        np.testing.assert_array_equal(got.c, ref.c)
        np.testing.assert_array_equal(got.decomposition()[0], ref.decomposition()[0])
        for g, r in zip(got3, ref3):
            assert g.lam == r.lam
            np.testing.assert_array_equal(g.c, r.c)
