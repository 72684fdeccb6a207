"""GPU parity: per-cell ensemble evaluation, Gram reduction (a6) vs the oracle."""
import numpy as np
import pytest

from machisplin_b200 import synth
from oracle import cbind, models as om, mltps as omt, tps as otps

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def case():
    geom = synth.make_geom(160, 224)
    C = 4
    models = synth.make_models(geom, C, 700, 5, kept="bgnmrv", rf_trees=50, gbm_trees=80)
    cov = synth.covariate_planes(geom, C)
    return geom, C, models, cov


def _cmp(got, ref, tol):
    assert np.array_equal(np.isnan(got), np.isnan(ref)), "NA mask differs"
    m = ~np.isnan(ref)
    err = np.max(np.abs(got[m] - ref[m])) / np.max(np.abs(ref[m]))
    assert err < tol, err
    return err


@pytest.mark.parametrize("letter", list("bgnmrv"))
def test_single_model_raster(engine, case, letter):
    geom, C, models, cov = case
    ens = engine.ensemble_create(geom, models, letter, [1.0], 1.0, C + 2)
    got = engine.ensemble_eval(ens, cov)
    ref = cbind.ensemble_eval(models, letter, [1.0], 1.0, cov, geom.as_tuple())
    # trees decide on identical comparisons -> only leaf-value float rounding; gam/nnet/earth are float64
    tol = {"b": 2e-7, "r": 2e-7, "g": 1e-12, "n": 1e-12, "m": 1e-12, "v": 5e-6}[letter]
    _cmp(got, ref, tol)


def test_tree_decisions_bit_exact(engine, case):
    """every RF / GBM comparison must take the same branch as the float64 reference comparison:
    with integer-valued leaves the float32 leaf storage is exact, so the result must be identical."""
    geom, C, models, cov = case
    m2 = {"r": dict(models["r"]), "b": dict(models["b"])}
    rng = np.random.default_rng(0)
    m2["r"]["nodepred"] = rng.integers(-50, 50, models["r"]["nodepred"].shape).astype(np.float64)
    sc = models["b"]["splitcode"].copy()
    leaf = models["b"]["splitvar"] == -1
    sc[leaf] = rng.integers(-8, 8, int(leaf.sum()))
    m2["b"]["splitcode"] = sc
    m2["b"]["initF"] = 3.0
    for letter in "rb":
        ens = engine.ensemble_create(geom, m2, letter, [1.0], 1.0, C + 2)
        got = engine.ensemble_eval(ens, cov)
        ref = cbind.ensemble_eval(m2, letter, [1.0], 1.0, cov, geom.as_tuple())
        m = ~np.isnan(ref)
        assert np.array_equal(np.isnan(got), np.isnan(ref))
        if letter == "b":
            assert np.array_equal(got[m], ref[m])
        else:   # leaves are stored as float32(value - offset): one wrong branch would move the mean by >= 1/ntree
            assert np.max(np.abs(got[m] - ref[m])) < 1e-5 < 0.1 / m2["r"]["ntree"]


def test_full_ensemble_plus_tps(engine, case):
    geom, C, models, cov = case
    kept, w, wt = synth.ensemble_weights("bgnmrv")
    xy, _, _ = synth.make_knots(geom, 250, 8)
    y = synth.residual_field(xy, 8)
    fit = otps.tps_fit(xy, y)
    sp = engine.spline_create(fit.knots_xy, fit.c, fit.d, fit.center, fit.scale)
    ens = engine.ensemble_create(geom, models, kept, w, wt, C + 2)
    got = engine.ensemble_eval(ens, cov, spline=sp)
    surf = cbind.tps_eval(fit, geom.as_tuple())
    ref = cbind.ensemble_eval(models, kept, w, wt, cov, geom.as_tuple(), tps=surf)
    _cmp(got, ref, 2e-6)
    # same through a precomputed surface and on a window
    win = (17, 131, 40, 199)
    got_w = engine.ensemble_eval(ens, cov, tps_surface=surf[win[0]:win[1], win[2]:win[3]], window=win)
    _cmp(got_w, ref[win[0]:win[1], win[2]:win[3]], 2e-6)


def test_smooth_only_and_weight_rule(engine, case):
    geom, C, models, cov = case
    p = np.array([0.41, 0.03, 0.333, 0.9])            # g n m v raw optimiser output
    kept, w, wt = om.select_models(p, om.MODEL_LETTERS_SMOOTH)
    assert kept == "gmv" and np.allclose(w, [0.41, 0.33, 0.9]) and abs(wt - p.sum()) < 1e-15
    ens = engine.ensemble_create(geom, models, kept, w, wt, C + 2)
    got = engine.ensemble_eval(ens, cov)
    ref = cbind.ensemble_eval(models, kept, w, wt, cov, geom.as_tuple())
    _cmp(got, ref, 2e-6)


def test_gbm_alone_predicts_na_cells(engine, case):
    """gbm routes NA through MissingNode and returns numbers where every other model gives NA."""
    geom, C, models, cov = case
    ens = engine.ensemble_create(geom, models, "b", [1.0], 1.0, C + 2)
    got = engine.ensemble_eval(ens, cov)
    assert np.isnan(cov[0]).any() and not np.isnan(got).any()
    ref = cbind.ensemble_eval(models, "b", [1.0], 1.0, cov, geom.as_tuple())
    assert np.max(np.abs(got - ref)) / np.max(np.abs(ref)) < 2e-7


def test_point_predictions(engine, case):
    geom, C, models, cov = case
    kept, w, wt = synth.ensemble_weights("bgnmrv")
    ens = engine.ensemble_create(geom, models, kept, w, wt, C + 2)
    _, row, col = synth.make_knots(geom, 300, 3)
    X = omt.point_features(geom.as_tuple(), cov.astype(np.float64), row, col)
    got = engine.ensemble_predict_points(ens, X)
    ref = om.ensemble_predict(models, kept, w, wt, X)
    _cmp(got, ref, 1e-10)


@pytest.mark.parametrize("K", [4, 6])
def test_gram(engine, K):
    rng = np.random.default_rng(K)
    R = rng.standard_normal((45000, K)) * rng.uniform(0.5, 20, K)
    G = engine.gram(R)
    ref = om.gram(R)
    assert np.max(np.abs(G - ref)) <= 1e-12 * np.max(np.abs(ref))
    k = rng.uniform(0, 1, K)
    assert abs(om.rss_from_gram(k, G) - om.rss_objective(k, R)) <= 1e-10 * om.rss_objective(k, R)


# ---- the one-call path (mb_mltps_predict*) and the tunable kernel variants ---------------------------
def test_mltps_predict_one_call_matches_oracle(engine, case):
    """parts 2-5 in one call: ensemble kernels beside the GCV fit, fused TPS + combine pass."""
    geom, C, models, cov = case
    kept, w, wt = synth.ensemble_weights("bgnmrv")
    xy, _, _ = synth.make_knots(geom, 250, 8)
    y = synth.residual_field(xy, 8)
    fit = otps.tps_fit(xy, y)
    ens = engine.ensemble_create(geom, models, kept, w, wt, C + 2)
    got, sp = engine.mltps_predict(geom, ens, cov, xy, y)
    assert abs(sp.lam - fit.lam) <= 1e-6 * fit.lam
    surf = cbind.tps_eval(fit, geom.as_tuple())
    ref = cbind.ensemble_eval(models, kept, w, wt, cov, geom.as_tuple(), tps=surf)
    _cmp(got, ref, 2e-6)
    # tps = FALSE and ensemble-less variants
    got2, sp2 = engine.mltps_predict(geom, ens, cov, None, None)
    assert sp2 is None
    _cmp(got2, cbind.ensemble_eval(models, kept, w, wt, cov, geom.as_tuple()), 2e-6)
    got3, _ = engine.mltps_predict(geom, None, None, xy, y)
    assert np.max(np.abs(got3 - surf)) < 1e-6 * np.max(np.abs(surf))


def test_mltps_predict_tiled_mode(engine, case):
    """tile_px > 0 and more than one tile: the reference's internal tiling (V73:649-895) feeds the combine."""
    from oracle import tiles as otl
    geom, C, models, cov = case
    kept, w, wt = synth.ensemble_weights("gnm")
    xy, _, _ = synth.make_knots(geom, 600, 18)
    y = synth.residual_field(xy, 18)
    ens = engine.ensemble_create(geom, models, kept, w, wt, C + 2)
    got, sp = engine.mltps_predict(geom, ens, cov, xy, y, tile_px=100)
    assert sp is None
    surf = otl.tps_tiled_surface(geom.as_tuple(), xy, y, tile_px=100)
    ref = cbind.ensemble_eval(models, kept, w, wt, cov, geom.as_tuple(), tps=surf)
    _cmp(got, ref, 2e-6)


@pytest.mark.parametrize("win", [(3, 150, 5, 217), (0, 8, 0, 32), (17, 18, 40, 41), (120, 160, 200, 224)])
def test_forest_on_ragged_windows(engine, case, win):
    """the two-level tile-pruned forest kernel on windows that cut its 32 x 8 tiles and 8 x 4 warp blocks anywhere (same
    comparisons as the reference, only the order of the float64 sum over trees differs)."""
    geom, C, models, cov = case
    ens = engine.ensemble_create(geom, models, "rb", [0.6, 0.4], 1.0, C + 2)
    got = engine.ensemble_eval(ens, cov, window=win)
    ref = cbind.ensemble_eval(models, "rb", [0.6, 0.4], 1.0, cov, geom.as_tuple())[win[0]:win[1], win[2]:win[3]]
    _cmp(got, ref, 2e-7)


def test_forest_on_rough_raster(engine, case):
    """white-noise covariates defeat the tile pruning (every tree forks in every tile): same answer."""
    geom, C, models, cov = case
    rng = np.random.default_rng(4)
    rough = (cov + rng.normal(0, 60, cov.shape)).astype(np.float32)
    ens = engine.ensemble_create(geom, models, "rb", [0.5, 0.5], 1.0, C + 2)
    got = engine.ensemble_eval(ens, rough)
    ref = cbind.ensemble_eval(models, "rb", [0.5, 0.5], 1.0, rough, geom.as_tuple())
    _cmp(got, ref, 2e-7)


def test_fused_part5_on_ragged_windows(engine, case):
    """The accumulator-fused grid-evaluation kernel (TPS + part 5) on windows whose size and origin are not
    multiples of the 32 x 32 leaf box: the padded accumulator layout and the bulk-copied tiles must line up."""
    geom, C, models, cov = case
    kept, w, wt = synth.ensemble_weights("gnmrv")
    xy, _, _ = synth.make_knots(geom, 300, 28)
    y = synth.residual_field(xy, 28)
    fit = otps.tps_fit(xy, y, lam=2e-3)
    sp = engine.spline_create(fit.knots_xy, fit.c, fit.d, fit.center, fit.scale)
    ens = engine.ensemble_create(geom, models, kept, w, wt, C + 2)
    surf = cbind.tps_eval(fit, geom.as_tuple())
    ref = cbind.ensemble_eval(models, kept, w, wt, cov, geom.as_tuple(), tps=surf)
    for win in [(0, 160, 0, 224), (7, 150, 13, 201), (64, 97, 32, 65), (159, 160, 0, 224)]:
        got = engine.ensemble_eval(ens, cov, spline=sp, window=win)
        _cmp(got, ref[win[0]:win[1], win[2]:win[3]], 2e-6)


def test_tiled_mode_is_independent_of_lane_concurrency(engine, case):
    """mltps part 3 runs its tiles on four worker lanes; the raster must be bit-identical to the one-lane order
    (a user stream forces the sequential path)."""
    import torch
    geom, C, models, cov = case
    xy, _, _ = synth.make_knots(geom, 600, 18)
    y = synth.residual_field(xy, 18)
    a = engine.tiles_tps(geom, xy, y, tile_px=60)
    out = torch.empty((geom.nrow, geom.ncol), dtype=torch.float64, device="cuda:0")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        engine.tiles_tps_dev(geom, xy, y, out.data_ptr(), tile_px=60, stream=s.cuda_stream)
    s.synchronize()
    np.testing.assert_array_equal(a, out.cpu().numpy())


def test_forest_with_more_trees_than_one_chunk(engine, case):
    """k_ens_trees prunes 2 048 trees per pass (kTreeChunk): 50 rf + 2 400 gbm trees (the 80 fitted trees thirty times over - the
    descriptor's child indices are local to a tree) run as two passes whose tile constants, fork lists and per-cell sums have to
    add up to the plain walk of the oracle; integer leaves make the gbm part exact."""
    geom, C, models, cov = case
    b = models["b"]
    rep = 30
    nt = len(b["tree_off"]) - 1
    sizes = np.diff(b["tree_off"])
    m2 = {"r": models["r"],
          "b": {"initF": b["initF"], "tree_off": np.concatenate([[0], np.cumsum(np.tile(sizes, rep))]).astype(np.int32),
                "splitvar": np.tile(b["splitvar"], rep), "splitcode": np.tile(b["splitcode"], rep),
                "left": np.tile(b["left"], rep), "right": np.tile(b["right"], rep), "missing": np.tile(b["missing"], rep)}}
    assert len(m2["b"]["tree_off"]) - 1 == rep * nt and models["r"]["ntree"] + rep * nt > 2048
    for kept in ("b", "rb"):
        ws = [1.0 / len(kept)] * len(kept)
        ens = engine.ensemble_create(geom, m2, kept, ws, 1.0, C + 2)
        got = engine.ensemble_eval(ens, cov)
        sub = (5, 133, 9, 200)
        got_w = engine.ensemble_eval(ens, cov, window=sub)
        ref = cbind.ensemble_eval(m2, kept, ws, 1.0, cov, geom.as_tuple())
        _cmp(got, ref, 2e-7)
        _cmp(got_w, ref[sub[0]:sub[1], sub[2]:sub[3]], 2e-7)
    # integer leaves: the float32 leaf storage is exact, so 2 400 trees in two passes must reproduce the reference bit for bit
    rng = np.random.default_rng(1)
    sc = m2["b"]["splitcode"].copy()
    leaf = m2["b"]["splitvar"] == -1
    sc[leaf] = rng.integers(-8, 8, int(leaf.sum()))
    m3 = {"b": dict(m2["b"], splitcode=sc, initF=3.0)}
    ens = engine.ensemble_create(geom, m3, "b", [1.0], 1.0, C + 2)
    got = engine.ensemble_eval(ens, cov)
    ref = cbind.ensemble_eval(m3, "b", [1.0], 1.0, cov, geom.as_tuple())
    ok = ~np.isnan(ref)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    assert np.array_equal(got[ok], ref[ok])


@pytest.mark.parametrize("impl", [1, 2, 3])
@pytest.mark.parametrize("kept", ["v", "gnmv", "bgnmrv"])
def test_svm_kernel_variants(engine, case, kept, impl):
    """The ksvm kernels against the oracle: dot products on the tensor pipe with FP16 split operands (impl 3, two HMMA.16816 per
    16 cells x 8 support vectors; the default for P <= 8) or 3 x TF32 (impl 1, three HMMA.1688), and packed FP32 (impl 2,
    k_ens_svm; the only one for P > 8).  Same tolerance, same NA mask, ragged window."""
    geom, C, models, cov = case
    ws = [1.0 / len(kept)] * len(kept)
    ref = cbind.ensemble_eval(models, kept, ws, 1.0, cov, geom.as_tuple())
    try:
        engine.set_param("svm_impl", impl)
        ens = engine.ensemble_create(geom, models, kept, ws, 1.0, C + 2)
        got = engine.ensemble_eval(ens, cov)
        sub = (3, 150, 5, 201)
        got_w = engine.ensemble_eval(ens, cov, window=sub)
    finally:
        engine.set_param("svm_impl", 0)
    _cmp(got, ref, 5e-6)
    _cmp(got_w, ref[sub[0]:sub[1], sub[2]:sub[3]], 5e-6)


@pytest.mark.parametrize("overlap,tma,per_sm", [(1, 1, 0), (1, 2, 1), (2, 1, 0), (2, 2, 0)])
@pytest.mark.parametrize("kept", ["rv", "bgnmrv", "gnmv"])
def test_ensemble_kernel_schedules(engine, case, kept, overlap, tma, per_sm):
    """The forest kernel and the tensor-pipe ksvm kernel side by side on two streams (each adds its sum to the zeroed accumulator)
    or one after the other; covariate tiles of the ksvm kernel by TMA tensor copy or by plain loads; persistent or chunked tile
    schedule - against the oracle, on the full grid and on a window that cuts the 32 x 8 tiles."""
    geom, C, models, cov = case
    ws = [1.0 / len(kept)] * len(kept)
    ref = cbind.ensemble_eval(models, kept, ws, 1.0, cov, geom.as_tuple())
    try:
        engine.set_param("ens_overlap", overlap)
        engine.set_param("ens_tma", tma)
        engine.set_param("svm_ctas_per_sm", per_sm)
        ens = engine.ensemble_create(geom, models, kept, ws, 1.0, C + 2)
        got = engine.ensemble_eval(ens, cov)
        again = engine.ensemble_eval(ens, cov)
        wins = [(3, 150, 5, 201), (159, 160, 0, 224), (64, 97, 32, 65), (7, 150, 13, 201)]
        got_w = [engine.ensemble_eval(ens, cov, window=sub) for sub in wins]
    finally:
        engine.set_param("ens_overlap", 0)
        engine.set_param("ens_tma", 0)
        engine.set_param("svm_ctas_per_sm", 0)
    _cmp(got, ref, 5e-6)
    np.testing.assert_array_equal(got, again)                 # two commutative adds per cell: the race does not show
    for sub, g in zip(wins, got_w):
        _cmp(g, ref[sub[0]:sub[1], sub[2]:sub[3]], 5e-6)


def test_ensemble_on_a_raster_tma_cannot_describe(engine):
    """ncol = 203: the row stride is not a multiple of 16 bytes, so no tensor map exists and the tile ring is filled by plain
    loads; 3 covariates (P = 5 < 8: padded ksvm fragments); NA cells; last tile column and row ragged."""
    geom = synth.make_geom(77, 203)
    C = 3
    models = synth.make_models(geom, C, 500, 8, kept="bgnmrv", rf_trees=40, gbm_trees=60)
    cov = synth.covariate_planes(geom, C)
    kept, w, wt = synth.ensemble_weights("bgnmrv")
    ref = cbind.ensemble_eval(models, kept, w, wt, cov, geom.as_tuple())
    ens = engine.ensemble_create(geom, models, kept, w, wt, C + 2)
    got = engine.ensemble_eval(ens, cov)
    _cmp(got, ref, 5e-6)


def test_sm_partitions_do_not_change_the_raster(engine):
    """mb_mltps_predict with CUDA green contexts (stage 1 of the GCV fit on the fit partition, the forest kernel on the ensemble
    partition from the start, the rest of the chain behind stage 1) against the deferred schedule (gc_split = -1): same kernels,
    same order of the additions - bit-identical raster and lambda.  1 600 knots: above the size where the partitions switch on."""
    geom = synth.make_geom(384, 512)
    C = 4
    cov = synth.covariate_planes(geom, C)
    xy, _, _ = synth.make_knots(geom, 1600, 77)
    resid = synth.residual_field(xy, 77)
    models = synth.make_models(geom, C, 600, 77, kept="bgnmrv", rf_trees=60, gbm_trees=90)
    kept, w, wt = synth.ensemble_weights("bgnmrv")
    ens = engine.ensemble_create(geom, models, kept, w, wt, C + 2)
    try:
        engine.set_param("gc_split", -1)
        ref, sp_ref = engine.mltps_predict(geom, ens, cov, xy, resid)
    finally:
        engine.set_param("gc_split", 0)
    got, sp = engine.mltps_predict(geom, ens, cov, xy, resid)
    again, _ = engine.mltps_predict(geom, ens, cov, xy, resid)
    assert sp.lam == sp_ref.lam
    np.testing.assert_array_equal(got, ref)
    np.testing.assert_array_equal(again, ref)

