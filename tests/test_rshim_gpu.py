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


def test_dot_c_twins():
    """The .C()-loadable entry points (all-pointer arguments, status out-parameter) called the way R's .C()
    calls them: mbC_gram, mbC_tps_surface (global and tiled), mbC_tiles_merge, the error path."""
    import ctypes as C
    from machisplin_b200 import _lib
    from oracle import tiles as otl
    lib = _lib.load()
    i32 = lambda v: (C.c_int32 * 1)(v)
    st = i32(99)
    R = np.asfortranarray(np.random.default_rng(3).standard_normal((500, 4)))
    G = np.empty((4, 4))
    lib.mbC_gram(R.ctypes.data_as(_lib.PD), i32(500), i32(4), G.ctypes.data_as(_lib.PD), st)
    assert st[0] == 0
    np.testing.assert_allclose(G, R.T @ R, rtol=1e-12)
    geom = synth.make_geom(120, 160)
    xy, _, _ = synth.make_knots(geom, 300, 77)
    y = synth.residual_field(xy, 77)
    g6 = np.array(geom.as_tuple(), dtype=np.float64)
    xyf = np.asfortranarray(xy)
    for tile_px in (0, 60):
        out = np.empty((geom.nrow, geom.ncol))
        lam_in, lam_out = (C.c_double * 1)(-1.0), (C.c_double * 1)(0.0)
        lib.mbC_tps_surface(xyf.ctypes.data_as(_lib.PD), y.ctypes.data_as(_lib.PD), i32(300), g6.ctypes.data_as(_lib.PD), lam_in,
                            i32(tile_px), out.ctypes.data_as(_lib.PD), lam_out, st)
        assert st[0] == 0
        ref = otl.tps_tiled_surface(geom.as_tuple(), xy, y, tile_px=tile_px or 10 ** 9)
        assert np.max(np.abs(out - ref)) < 2e-6 * np.max(np.abs(ref))
        assert np.isnan(lam_out[0]) == (tile_px > 0)
    lib.mbC_tps_surface(xyf.ctypes.data_as(_lib.PD), y.ctypes.data_as(_lib.PD), i32(3), g6.ctypes.data_as(_lib.PD), lam_in, i32(0),
                        out.ctypes.data_as(_lib.PD), lam_out, st)
    assert st[0] < 0                                           # n <= 3: status, no crash
    buf = C.create_string_buffer(b" " * 200)
    p = (C.c_char_p * 1)(C.addressof(buf))
    lib.mbC_last_error(p)
    assert b"more than 3" in buf.value
    tc = otl.tiles_create(geom.as_tuple(), np.zeros((0, 2)), out_ncol=2, out_nrow=2, feather_d=20)
    wins = np.array([t["win"] for t in tc["tiles"]], dtype=np.int32)
    rng = np.random.default_rng(1)
    tiles = [rng.standard_normal((w[1] - w[0], w[3] - w[2])) for w in wins]
    flat = np.concatenate([t.ravel() for t in tiles])
    merged = np.empty((geom.nrow, geom.ncol))
    lib.mbC_tiles_merge(g6.ctypes.data_as(_lib.PD), i32(2), i32(2), wins.ctypes.data_as(_lib.PI32), flat.ctypes.data_as(_lib.PD),
                        merged.ctypes.data_as(_lib.PD), st)
    assert st[0] == 0
    ref = otl.tiles_merge(geom.as_tuple(), [tuple(w) for w in wins], tiles, 2, 2)
    np.testing.assert_allclose(merged, ref, rtol=1e-13, atol=1e-13)
    lib.mbC_shutdown()


def test_tiles_merge_and_the_multi_gpu_wrappers(rs, mb):
    """mbR_tiles_merge against the oracle (V73:1392-1548); mbR_comm_unique_id / mbR_comm_init with one rank, then
    mbR_gram_allreduce and mbR_mltps_predict_shard - the calls an R process per GPU makes - give what the single-process calls give."""
    from oracle import tiles as otl
    geom = synth.make_geom(120, 170)
    tc = otl.tiles_create(geom.as_tuple(), np.zeros((0, 2)), out_ncol=2, out_nrow=2, feather_d=20)
    wins = np.array([t["win"] for t in tc["tiles"]], dtype=np.int32)
    rng = np.random.default_rng(1)
    tiles = [rng.standard_normal((w[1] - w[0], w[3] - w[2])) for w in wins]
    lst = rs.named_list({f"t{k}": rs.real(t.ravel()) for k, t in enumerate(tiles)})
    out = rs.call("mbR_tiles_merge", mb, r_grid(rs, geom), rs.integer(2), rs.integer(2), rs.integer(wins.ravel()), lst)
    got = rs.as_numpy(out).reshape(geom.nrow, geom.ncol)
    ref = otl.tiles_merge(geom.as_tuple(), [tuple(w) for w in wins], tiles, 2, 2)
    np.testing.assert_allclose(got, ref, rtol=1e-13, atol=1e-13)
    # one-rank communicator through the shim (a second GPU is not needed to exercise the wrappers)
    ctx2 = rs.call("mbR_init", rs.integer(0))
    uid = rs.call("mbR_comm_unique_id")
    assert rs.as_numpy(uid).size == 128
    rs.call("mbR_comm_init", ctx2, rs.integer(1), rs.integer(0), uid)
    R = rng.standard_normal((500, 6))
    G = rs.as_numpy(rs.call("mbR_gram_allreduce", ctx2, rs.real(R, matrix=True)))
    np.testing.assert_allclose(G, R.T @ R, rtol=1e-12, atol=1e-12)
    C = 3
    g2 = synth.make_geom(64, 96)
    cov = synth.covariate_planes(g2, C)
    xy, _, _ = synth.make_knots(g2, 150, 2)
    res = synth.residual_field(xy, 2)
    models = synth.make_models(g2, C, 300, 2, kept="gnmv")
    kept, w, wt = om.select_models(np.array([0.4, 0.3, 0.2, 0.1]), letters="gnmv")
    packed = cov.astype(np.float32).tobytes()
    ens = rs.call("mbR_ensemble_create", ctx2, r_grid(rs, g2), models_to_r(rs, models, C + 2), rs.string(kept), rs.real(w), rs.real([wt]))
    a = rs.as_numpy(rs.call("mbR_mltps_predict_shard", ctx2, r_grid(rs, g2), ens, rs.raw(packed), rs.integer(C), rs.real(xy, matrix=True),
                            rs.real(res), rs.integer(len(res)), rs.real([-1.0]), rs.integer(0)))
    b = rs.as_numpy(rs.call("mbR_mltps_predict", ctx2, r_grid(rs, g2), ens, rs.raw(packed), rs.integer(C), rs.real(xy, matrix=True),
                            rs.real(res), rs.real([-1.0]), rs.integer(0)))
    np.testing.assert_array_equal(a, b)
    rs.finalize(ens)
    rs.finalize(ctx2)

