"""The .Call shim (r/mb_shim.c) compiles against the stub R API and exports every wrapper INTEGRATION.md names; the wrappers
that need no GPU (GeoTIFF in / out) run here.  The GPU ones are in test_rshim_gpu.py."""
import os

import numpy as np
import pytest

from tests.rshim_harness import RStub

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
WRAPPERS = ["mbR_init", "mbR_tps_fit", "mbR_tps_eval", "mbR_ensemble_create", "mbR_mltps_predict", "mbR_gram",
            "mbR_read_stack", "mbR_write_raster", "mbR_tiles_merge", "mbR_tiles_owned_window", "mbR_comm_unique_id", "mbR_comm_init",
            "mbR_gram_allreduce", "mbR_mltps_predict_shard"]


@pytest.fixture(scope="module")
def rs():
    return RStub()


def test_shim_builds_and_exports_every_wrapper(rs):
    for w in WRAPPERS:
        assert hasattr(rs.lib, w), w
    src = open(os.path.join(os.path.dirname(GOLD), "..", "r", "machisplin_b200.R")).read()
    for w in WRAPPERS:
        assert f'"{w}"' in src, f"{w} is not called from r/machisplin_b200.R"


def test_read_stack_gives_planes_in_terra_cell_order(rs):
    """mbR_read_stack: list(grid, cov) with cov = float32 planes [C][nrow][ncol] - the layout mbR_mltps_predict takes and the one
    as.numeric(terra::values(stack)) produces (ncell x nlyr, column-major; no transpose - ADVICE r1)."""
    path = os.path.join(GOLD, "twi_2x2_tiles.tif")
    z = np.load(os.path.join(GOLD, "twi_2x2_tiles.npz"))
    want = z["values"].astype(np.float32)
    want[z["values"] == int(z["nodata"])] = np.nan
    out = rs.call("mbR_read_stack", rs.strings([path, path]))     # character vector of two paths (the same file twice)
    grid = rs.as_numpy(rs.list_elt(out, 0))
    cov = rs.as_numpy(rs.list_elt(out, 1)).view(np.float32)
    nrow, ncol = int(grid[4]), int(grid[5])
    assert (nrow, ncol) == want.shape and cov.size == 2 * nrow * ncol
    planes = cov.reshape(2, nrow, ncol)
    for k in range(2):
        np.testing.assert_array_equal(np.isnan(planes[k]), np.isnan(want))
        np.testing.assert_array_equal(planes[k][~np.isnan(want)], want[~np.isnan(want)])


def test_write_raster_round_trip_and_error_path(rs, tmp_path):
    from machisplin_b200 import geotiff
    rng = np.random.default_rng(0)
    v = rng.standard_normal((40, 70))
    v[3, 5] = np.nan
    p = str(tmp_path / "o.tif")
    rs.call("mbR_write_raster", rs.string(p), rs.real([10.0, 17.0, -4.0, 0.0, 40, 70]), rs.real(v.ravel()), rs.integer(4326))
    g, back = geotiff.read_raster(p)
    assert (g.nrow, g.ncol, g.xmin, g.ymax) == (40, 70, 10.0, 0.0)
    np.testing.assert_array_equal(back, v.astype(np.float32))
    # a failing C-ABI call becomes an R error (longjmp out of the wrapper), not a crash or a silent NULL
    with pytest.raises(RuntimeError, match="machisplin_b200"):
        rs.call("mbR_write_raster", rs.string("/nonexistent-dir/o.tif"), rs.real([10.0, 17.0, -4.0, 0.0, 40, 70]),
                rs.real(v.ravel()), rs.integer(0))


def test_owned_window_through_the_shim(rs):
    """mbR_tiles_owned_window needs no GPU: the windows of a 2 x 2 lattice partition the raster."""
    from machisplin_b200 import synth, tiles as mtiles
    geom = synth.make_geom(240, 310)
    ts = mtiles.tiles_create(geom, np.zeros((0, 2)), 2, 2, feather_d=30)
    wins = np.array([t.win for t in ts.tiles], dtype=np.int32)            # ntiles x 4, row-major = R's 4 x ntiles column-major
    grid = rs.real([geom.xmin, geom.xmax, geom.ymin, geom.ymax, geom.nrow, geom.ncol])
    cover = np.zeros((geom.nrow, geom.ncol), dtype=int)
    for t in range(4):
        o = rs.as_numpy(rs.call("mbR_tiles_owned_window", grid, rs.integer(2), rs.integer(2), rs.integer(wins.ravel()), rs.integer(t)))
        cover[o[0]:o[1], o[2]:o[3]] += 1
    assert np.all(cover == 1)
    with pytest.raises(RuntimeError, match="machisplin_b200"):
        rs.call("mbR_tiles_owned_window", grid, rs.integer(2), rs.integer(2), rs.integer(wins.ravel()), rs.integer(4))

