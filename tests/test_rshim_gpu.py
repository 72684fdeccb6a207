"""The R .Call shim (r/mb_shim.c) driven through fake SEXPs (tests/r_stub) down to the CUDA kernels: what an R session would get
from mbR_gram / mbR_tps_fit / mbR_tps_eval / mbR_ensemble_create / mbR_mltps_predict, checked against the oracle."""
import numpy as np
import pytest

from machisplin_b200 import synth
from oracle import cbind, models as om, tps as otps
from tests.rshim_harness import RStub, models_to_r

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rs():
    return RStub()


@pytest.fixture(scope="module")
def mb(rs):
    ctx = rs.call("mbR_init", rs.integer(0))
    yield ctx
    rs.finalize(ctx)


def r_grid(rs, g):
    return rs.real([g.xmin, g.xmax, g.ymin, g.ymax, g.nrow, g.ncol])


def test_gram(rs, mb):
    R = np.random.default_rng(1).standard_normal((700, 6))
    G = rs.as_numpy(rs.call("mbR_gram", mb, rs.real(R, matrix=True)))
    assert G.shape == (6, 6)
    np.testing.assert_allclose(G, R.T @ R, rtol=1e-12, atol=1e-12)


def test_tps_fit_and_eval(rs, mb):
    geom = synth.make_geom(96, 128)
    xy, _, _ = synth.make_knots(geom, 200, 31)
    y = synth.residual_field(xy, 31)
    sp = rs.call("mbR_tps_fit", mb, rs.real(xy, matrix=True), rs.real(y), rs.real([-1.0]))
    assert rs.lib.stub_prot(sp) == mb                     # the handle keeps its context alive (finalizer order)
    v = rs.as_numpy(rs.call("mbR_tps_eval", mb, sp, r_grid(rs, geom))).reshape(geom.nrow, geom.ncol)
    ref = otps.tps_interpolate(otps.tps_fit(xy, y), geom.as_tuple())
    assert np.max(np.abs(v - ref)) < 2e-6 * np.max(np.abs(ref))
    rs.finalize(sp)
    with pytest.raises(RuntimeError, match="machisplin_b200"):       # n <= 3 observations: an R error, not a crash
        rs.call("mbR_tps_fit", mb, rs.real(xy[:3], matrix=True), rs.real(y[:3]), rs.real([-1.0]))


def test_mltps_predict_takes_planes_as_terra_values_lays_them_out(rs, mb):
    """r/machisplin_b200.R packs as.numeric(terra::values(stack)) - an ncell x nlyr matrix in column-major order - as float32:
    that IS [C][nrow][ncol].  A cell-interleaved packing (the t() bug of round 1) would scramble every covariate."""
    geom = synth.make_geom(64, 96)
    C = 3
    cov = synth.covariate_planes(geom, C)
    values = cov.reshape(C, -1).T                                       # terra::values(): ncell x nlyr
    packed = np.asfortranarray(values).ravel(order="F").astype(np.float32).tobytes()   # as.numeric(values) -> writeBin(size = 4)
    xy, _, _ = synth.make_knots(geom, 150, 2)
    res = synth.residual_field(xy, 2)
    models = synth.make_models(geom, C, 300, 2, kept="gnmv")
    kept, w, wt = om.select_models(np.array([0.4, 0.3, 0.2, 0.1]), letters="gnmv")
    ens = rs.call("mbR_ensemble_create", mb, r_grid(rs, geom), models_to_r(rs, models, C + 2), rs.string(kept), rs.real(w),
                  rs.real([wt]))
    out = rs.call("mbR_mltps_predict", mb, r_grid(rs, geom), ens, rs.raw(packed), rs.integer(C), rs.real(xy, matrix=True),
                  rs.real(res), rs.real([-1.0]), rs.integer(1500))
    got = rs.as_numpy(out).reshape(geom.nrow, geom.ncol)
    fit = otps.tps_fit(xy, res)
    ref = cbind.ensemble_eval(models, kept, w, wt, cov, geom.as_tuple(), tps=cbind.tps_eval(fit, geom.as_tuple()))
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    m = ~np.isnan(ref)
    assert np.max(np.abs(got[m] - ref[m])) < 5e-6 * np.max(np.abs(ref[m]))
    rs.finalize(ens)
