"""CPU tests of the host-side mirror of the reference interface (machisplin_b200/{tiles,mltps,parallel}.py):
index arithmetic and sharding logic only - no compute call (that needs the GPU)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from machisplin_b200 import mltps as pm, parallel as par, synth, tiles as pt   # noqa: E402
from oracle import models as om, tiles as otl                                    # noqa: E402


def test_tiles_create_matches_oracle_and_v73():
    geom = synth.make_geom(3264, 2476)
    rng = np.random.default_rng(0)
    pts = rng.uniform([geom.xmin, geom.ymin], [geom.xmax, geom.ymax], (813, 2))
    for nc, nr, fd in [(3, 3, 50), (2, 4, 50), (5, 1, 20), (1, 1, 50)]:
        ts = pt.tiles_create(geom, pts, nc, nr, fd)
        ref = otl.tiles_create(geom.as_tuple(), pts, nc, nr, fd)
        assert ts.nC == nc and ts.nR == nr and len(ts.tiles) == nc * nr
        for t, r in zip(ts.tiles, ref["tiles"]):
            assert t.win == tuple(r["win"])
            np.testing.assert_allclose(t.ext, r["ext"], rtol=0, atol=0)
            assert np.array_equal(t.points, r["points"])
            assert t.geom.as_tuple() == pytest.approx(r["geom"], abs=0)
    # tile 1 is the south-west tile (V73:1192-1197); neighbours overlap by feather.d pixels
    ts = pt.tiles_create(geom, pts, 3, 3, 50)
    assert ts.tiles[0].win[1] == geom.nrow and ts.tiles[0].win[2] == 0
    assert ts.tiles[0].win[3] - ts.tiles[1].win[2] == 50


def test_crop_window_snaps_like_terra():
    geom = synth.make_geom(100, 200)
    r = geom.rx
    assert pt.crop_window(geom, (10.4 * r, 20.6 * r, 0.0, 5.5 * r)) == (100 - 6, 100, 10, 21)
    assert pt.crop_window(geom, (-1.0, 2.0, -1.0, 2.0)) == (0, 100, 0, 200)          # clipped to the raster


def test_weight_rule_and_quadratic_objective():
    p = np.array([0.41, 0.03, 0.333, 0.9, 0.0449, 0.21])
    assert pm.select_models(p)[0] == om.select_models(p)[0]
    np.testing.assert_array_equal(pm.select_models(p)[1], om.select_models(p)[1])
    assert pm.select_models(p)[2] == om.select_models(p)[2]
    rng = np.random.default_rng(3)
    R = rng.standard_normal((500, 6))
    fit = pm.rss_objective_from_gram(R.T @ R)
    k = rng.uniform(0, 1, 6)
    assert abs(fit(k) - om.rss_objective(k, R)) < 1e-10 * om.rss_objective(k, R)


def test_knot_cells_are_cell_centres():
    geom = synth.make_geom(64, 96)
    xy = np.array([[geom.xmin, geom.ymax], [geom.xmax, geom.ymin], [0.5 * geom.xmax, 0.5 * geom.ymax], [2.0, 2.0]])
    k, row, col = pm.knot_cells(geom, xy)
    assert list(row) == [0, 63, 32, -1] and list(col) == [0, 95, 48, -1]
    ref_k, ref_r, ref_c = otl.knot_coordinates(geom.as_tuple(), xy[:3])
    np.testing.assert_allclose(k[:3], ref_k)
    assert np.array_equal(row[:3], ref_r) and np.array_equal(col[:3], ref_c)


def test_sharding_covers_everything_once():
    for world in (1, 2, 3, 8):
        owned = sorted(t for r in range(world) for t in par.tiles_of_rank(121, world, r))
        assert owned == list(range(121))
        blocks = par.row_blocks(8192 + 17, world)
        assert blocks[0][0] == 0 and blocks[-1][1] == 8192 + 17
        assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
        assert all(b[0] % 32 == 0 for b in blocks)
        rows = np.concatenate([np.arange(90001)[par.shard_rows(90001, world, r)] for r in range(world)])
        assert np.array_equal(rows, np.arange(90001))


def test_row_block_geometry_reproduces_the_cell_centres():
    """A rank's block geometry (parallel.block_geom) gives the cells of its rows the LONG / LAT of the full raster."""
    geom = synth.make_geom(8192 + 17, 4096)
    for (r0, r1) in par.row_blocks(geom.nrow, 8):
        b = par.block_geom(geom, r0, r1)
        assert (b.nrow, b.ncol, b.xmin, b.xmax) == (r1 - r0, geom.ncol, geom.xmin, geom.xmax)
        rows = np.arange(r0, r1)
        y_full = geom.ymax - (rows + 0.5) * geom.ry
        y_blk = b.ymax - (rows - r0 + 0.5) * b.ry
        assert np.max(np.abs(y_full - y_blk)) < 4e-16 * max(1.0, abs(geom.ymax))
        assert abs(b.rx - geom.rx) == 0 and abs(b.ry - geom.ry) < 1e-18


WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, os.environ["MB_ROOT"])
import torch.distributed as dist
from machisplin_b200 import parallel as par
rank, world = par.init("gloo")
rng = np.random.default_rng(7)
R = rng.standard_normal((9001, 6)) * rng.uniform(0.5, 20, 6)      # same matrix on every rank
sl = par.shard_rows(R.shape[0], world, rank)
G = par.allreduce_gram(R[sl].T @ R[sl])                            # per-rank Gram (mb_gram on a GPU box) + all-reduce
assert np.allclose(G, R.T @ R, rtol=1e-12), "Gram all-reduce"
assert par.max_over_ranks(10.0 + rank) == 10.0 + world - 1
mine = [(t, np.full((2, 2), float(t))) for t in par.tiles_of_rank(7, world, rank)]
got = par.gather_tiles(mine, 7, dst=0)
if rank == 0:
    assert [int(g[0, 0]) for g in got] == list(range(7))
else:
    assert got is None
# tile-border blend, step 1: tensors travel point-to-point to the merging rank (NCCL on GPUs, gloo here)
import torch
shapes = [(3, 4), (2, 5)]
tile = torch.full(shapes[rank], float(rank + 1), dtype=torch.float64)
tl = par.gather_tiles_device(tile, shapes, dst=0)
if rank == 0:
    assert [tuple(t.shape) for t in tl] == shapes and float(tl[1][1, 4]) == 2.0 and float(tl[0][0, 0]) == 1.0
else:
    assert tl is None
# the library's own NCCL communicator: rank 0 draws the id through the C ABI, the host ships it (here: gloo), every rank
# would call mb_comm_init with the same 128 bytes (needs a GPU; the recording engine below stands in for that one call)
import ctypes as C
from machisplin_b200 import _lib
class Rec:
    def comm_unique_id(self):
        buf = C.create_string_buffer(128)
        assert _lib.load().mb_comm_unique_id(buf) == 0, _lib.load().mb_last_error()
        return buf.raw
    def comm_init(self, world, rank, uid):
        self.got = (world, rank, uid)
rec = Rec()
assert par.comm_init(rec) == (rank, world)
ids = [None] * world
dist.all_gather_object(ids, rec.got[2])
assert rec.got[:2] == (world, rank) and len(rec.got[2]) == 128 and ids[0] == ids[1] and any(ids[0])
dist.barrier()
dist.destroy_process_group()
sys.stdout.write(f"rank{rank}ok\n"); sys.stdout.flush()
'''


def test_two_rank_gloo_gram_allreduce_and_tile_gather(tmp_path):
    """world_size 2 over gloo on CPU: the N > 1 plumbing of SURVEY.md 8e (sharded CV-residual Gram + tile ownership)."""
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MB_ROOT=ROOT, OMP_NUM_THREADS="1")
    import socket
    with socket.socket() as s:                       # a free port: back-to-back runs must not collide in TIME_WAIT
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2, r.stdout


@pytest.mark.parametrize("shape,nc,nr,fd", [((240, 310), 2, 2, 30), ((240, 310), 4, 2, 50), ((1000, 700), 3, 3, 50), ((64, 900), 8, 1, 10),
                                            ((500, 500), 1, 1, 50)])
def test_owned_windows_partition_the_raster(shape, nc, nr, fd):
    """mb_tiles_owned_window (host arithmetic of the sharded tiles.merge): the windows the tiles own are disjoint, cover every
    cell, lie inside their tile's window, and are cut through the middle of the overlap zones of machisplin.tiles.create."""
    from machisplin_b200 import _lib, engine as eng_mod, synth, tiles as mtiles
    lib = _lib.load()
    geom = synth.make_geom(*shape)
    ts = mtiles.tiles_create(geom, np.zeros((0, 2)), nc, nr, feather_d=fd)
    wins = [t.win for t in ts.tiles]
    cover = np.zeros(shape, dtype=int)
    for t, w in enumerate(wins):
        o = eng_mod.tiles_owned_window(lib, geom, wins, nc, nr, t)
        assert w[0] <= o[0] < o[1] <= w[1] and w[2] <= o[2] < o[3] <= w[3]
        cover[o[0]:o[1], o[2]:o[3]] += 1
        h, j = t % nc, t // nc
        if h + 1 < nc:                                   # the boundary halves the overlap with the eastern neighbour
            assert o[3] == (w[3] + wins[t + 1][2]) // 2
        if j + 1 < nr:                                   # tile rows count from the south: the northern neighbour is t + nc
            assert o[0] == (w[0] + wins[t + nc][1]) // 2
    assert np.all(cover == 1)
    with pytest.raises(_lib.MbError):
        eng_mod.tiles_owned_window(lib, geom, wins, nc, nr, nc * nr)
