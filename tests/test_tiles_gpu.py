"""GPU parity: mltps part 3/4 tiling + seam feather and machisplin.tiles.merge."""
import numpy as np
import pytest

from machisplin_b200 import synth
from oracle import tiles as otl

pytestmark = pytest.mark.gpu


def relerr(a, b):
    assert np.array_equal(np.isnan(a), np.isnan(b))
    m = ~np.isnan(b)
    return float(np.max(np.abs(a[m] - b[m])) / np.max(np.abs(b[m])))


@pytest.mark.parametrize("shape,tile_px", [((300, 420), 150), ((257, 190), 100), ((128, 300), 150), ((200, 200), 1500)])
@pytest.mark.parametrize("method", ["direct", "fast"])
def test_tiled_tps_surface(engine, shape, tile_px, method):
    geom = synth.make_geom(*shape)
    xy, _, _ = synth.make_knots(geom, 900, 17)
    y = synth.residual_field(xy, 17)
    ref = otl.tps_tiled_surface(geom.as_tuple(), xy, y, tile_px=tile_px)
    got = engine.tiles_tps(geom, xy, y, tile_px=tile_px, method=method)
    assert relerr(got, ref) < (1e-8 if method == "direct" else 2e-6)


def test_sparse_tile_becomes_zero_tile(engine):
    """< 10 knots in a fit box -> the tile is zeros (V73:710-721)."""
    geom = synth.make_geom(200, 400)
    xy, _, _ = synth.make_knots(geom, 400, 23)
    xy = xy[xy[:, 0] < 0.45]                     # empty the east half
    y = synth.residual_field(xy, 23)
    ref = otl.tps_tiled_surface(geom.as_tuple(), xy, y, tile_px=100)
    got = engine.tiles_tps(geom, xy, y, tile_px=100, method="fast")
    assert relerr(got, ref) < 2e-6
    assert np.all(got[:, 350:] == 0.0)


@pytest.mark.parametrize("nc,nr", [(2, 2), (3, 2), (2, 1), (1, 3), (1, 1)])
def test_tiles_merge_with_na(engine, nc, nr):
    geom = synth.make_geom(240, 310)
    rng = np.random.default_rng(nc * 10 + nr)
    tc = otl.tiles_create(geom.as_tuple(), np.zeros((0, 2)), out_ncol=nc, out_nrow=nr, feather_d=30)
    wins, rasters = [], []
    yy, xx = np.mgrid[0:geom.nrow, 0:geom.ncol]
    base = np.sin(xx / 40.0) + np.cos(yy / 55.0)
    hole = (xx - 150) ** 2 + (yy - 120) ** 2 < 30 ** 2
    for k, t in enumerate(tc["tiles"]):
        w = t["win"]
        r = (base + 0.1 * k + 0.01 * rng.standard_normal(base.shape))[w[0]:w[1], w[2]:w[3]].copy()
        r[hole[w[0]:w[1], w[2]:w[3]]] = np.nan
        if k == 0:
            r[:5, :] = np.nan                    # ragged NA edge inside an overlap
        wins.append(w)
        rasters.append(r)
    ref = otl.tiles_merge(geom.as_tuple(), wins, rasters, nc, nr)
    got = engine.tiles_merge(geom, wins, rasters, nc, nr)
    assert relerr(got, ref) < 1e-13


def test_tiles_merge_dev_matches_host_entry_point(engine):
    """machisplin.tiles.merge on device-resident tiles (what the multi-GPU path calls after the NCCL tile gather)."""
    import torch
    from machisplin_b200 import tiles as mt
    geom = synth.make_geom(240, 310)
    rng = np.random.default_rng(3)
    pts = rng.uniform([geom.xmin, geom.ymin], [geom.xmax, geom.ymax], (50, 2))
    ts = mt.tiles_create(geom, pts, 2, 2, feather_d=30)
    rasters = [rng.standard_normal((t.geom.nrow, t.geom.ncol)) for t in ts.tiles]
    ref = mt.tiles_merge(engine, geom, ts, rasters)
    dev = [torch.from_numpy(r).to("cuda:0") for r in rasters]
    out = torch.empty((geom.nrow, geom.ncol), dtype=torch.float64, device="cuda:0")
    engine.tiles_merge_dev(geom, [t.win for t in ts.tiles], [d.data_ptr() for d in dev], 2, 2, out.data_ptr())
    torch.cuda.synchronize()
    np.testing.assert_array_equal(out.cpu().numpy(), ref)


def test_tiled_surface_with_heterogeneous_tiles(engine):
    """Tiles whose knot clouds have very different aspect ratios get different leaf-box heights (hence different
    shared-memory footprints of the grid-evaluation kernel) while they run concurrently on the tile lanes."""
    geom = synth.make_geom(300, 420)
    rng = np.random.default_rng(91)
    xy_all, _, _ = synth.make_knots(geom, 9000, 91)
    third = (geom.xmax - geom.xmin) / 3
    left = xy_all[(xy_all[:, 0] < third) & (np.abs(xy_all[:, 1] % 0.18 - 0.09) < 0.012)][:260]      # thin horizontal bands
    mid = xy_all[(xy_all[:, 0] >= third) & (xy_all[:, 0] < 2 * third)][:300]                        # isotropic
    right = xy_all[(xy_all[:, 0] >= 2 * third) & (np.abs(xy_all[:, 0] - 2.5 * third) < 0.02)][:260]  # thin vertical band
    xy = np.vstack([left, mid, right])
    y = synth.residual_field(xy, 91)
    ref = otl.tps_tiled_surface(geom.as_tuple(), xy, y, tile_px=150, lam=2e-3)
    for _ in range(3):                                     # the lanes race differently every time
        got = engine.tiles_tps(geom, xy, y, tile_px=150, lam=2e-3)
        assert relerr(got, ref) < 2e-6


def _merge_case(geom, nc, nr, feather_d=30, seed=0):
    rng = np.random.default_rng(seed + nc * 10 + nr)
    tc = otl.tiles_create(geom.as_tuple(), np.zeros((0, 2)), out_ncol=nc, out_nrow=nr, feather_d=feather_d)
    yy, xx = np.mgrid[0:geom.nrow, 0:geom.ncol]
    base = np.sin(xx / 40.0) + np.cos(yy / 55.0)
    hole = (xx - 150) ** 2 + (yy - 120) ** 2 < 30 ** 2
    wins, rasters = [], []
    for k, t in enumerate(tc["tiles"]):
        w = t["win"]
        r = (base + 0.1 * k + 0.01 * rng.standard_normal(base.shape))[w[0]:w[1], w[2]:w[3]].copy()
        r[hole[w[0]:w[1], w[2]:w[3]]] = np.nan
        if k == 0:
            r[:5, :] = np.nan                    # ragged NA edge inside an overlap: moves a seam box
        wins.append(tuple(int(v) for v in w))
        rasters.append(r)
    return wins, rasters


@pytest.mark.parametrize("nc,nr", [(2, 2), (3, 2), (2, 1), (1, 3), (1, 1), (4, 2)])
def test_sharded_merge_equals_the_gathered_merge(engine, nc, nr):
    """mb_tiles_merge_shard_dev without a communicator (every tile local): the owned windows partition the raster and each holds
    exactly the cells mb_tiles_merge computes (bit-identical), which itself matches the oracle (V73:1392-1548)."""
    import torch
    geom = synth.make_geom(240, 310)
    wins, rasters = _merge_case(geom, nc, nr)
    ref = engine.tiles_merge(geom, wins, rasters, nc, nr)
    dev = torch.device("cuda", 0)
    tiles = {t: torch.from_numpy(r).to(dev) for t, r in enumerate(rasters)}
    own = {t: engine.tiles_owned_window(geom, wins, nc, nr, t) for t in range(nc * nr)}
    outs = {t: torch.full((o[1] - o[0], o[3] - o[2]), -1.0, dtype=torch.float64, device=dev) for t, o in own.items()}
    engine.tiles_merge_shard_dev(geom, wins, {t: v.data_ptr() for t, v in tiles.items()}, nc, nr,
                                 {t: v.data_ptr() for t, v in outs.items()})
    torch.cuda.synchronize()
    got = np.full((geom.nrow, geom.ncol), -2.0)
    cover = np.zeros((geom.nrow, geom.ncol), dtype=int)
    for t, o in own.items():
        got[o[0]:o[1], o[2]:o[3]] = outs[t].cpu().numpy()
        cover[o[0]:o[1], o[2]:o[3]] += 1
    assert np.all(cover == 1)
    np.testing.assert_array_equal(got, ref)
    assert relerr(got, otl.tiles_merge(geom.as_tuple(), wins, rasters, nc, nr)) < 1e-13
