"""Multi-GPU path behind the C ABI (comm.cu): the library's own NCCL communicator, the Gram / R^2 all-reduce, the spline
broadcast and mb_mltps_predict_shard*.  One GPU: a 1-rank communicator (every collective still goes through NCCL).
Two GPUs (gpurun --gpus 2): two processes, one raster cut into two row blocks, compared with the unsharded result."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from machisplin_b200 import parallel as par, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _case(nrow=256, ncol=320, knots=500, C=3, kept="gnmrv"):
    geom = synth.make_geom(nrow, ncol)
    cov = synth.covariate_planes(geom, C)
    xy, krow, kcol = synth.make_knots(geom, knots, 41)
    resid = synth.residual_field(xy, 41)
    models = synth.make_models(geom, C, 400, 41, kept=kept, rf_trees=40, gbm_trees=60)
    kept, w, wt = synth.ensemble_weights(kept)
    return geom, cov, xy, resid, models, kept, w, wt


def test_one_rank_communicator():
    import machisplin_b200 as mb
    eng = mb.Engine(0)
    assert "NCCL via" in eng.comm_backend()
    eng.comm_init(1, 0, eng.comm_unique_id())
    assert (eng.rank, eng.world) == (0, 1)
    R = np.random.default_rng(2).standard_normal((3001, 6))
    np.testing.assert_allclose(eng.gram_allreduce(R), R.T @ R, rtol=1e-12)
    np.testing.assert_array_equal(eng.allreduce([3.5, -1.0], op="max"), [3.5, -1.0])
    geom, cov, xy, resid, models, kept, w, wt = _case()
    ens = eng.ensemble_create(geom, models, kept, w, wt, cov.shape[0] + 2)
    ref, sp_ref = eng.mltps_predict(geom, ens, cov, xy, resid)
    got, sp = eng.mltps_predict_shard(geom, ens, cov, xy, resid, len(resid), root=0)
    assert sp.lam == sp_ref.lam
    np.testing.assert_array_equal(got, ref)
    sp2 = eng.spline_bcast(sp, len(resid), root=0)            # root of a 1-rank broadcast keeps its handle
    assert sp2 is sp
    eng.comm_destroy()
    eng.close()


WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, os.environ["MB_ROOT"])
import torch, torch.distributed as dist
import machisplin_b200 as mb
from machisplin_b200 import parallel as par, synth
sys.path.insert(0, os.path.join(os.environ["MB_ROOT"], "tests"))
from test_comm_gpu import _case
rank, world, local = par.env_rank()
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = mb.Engine(local)
assert par.comm_init(eng) == (rank, world) and eng.world == world
geom, cov, xy, resid, models, kept, w, wt = _case()
C = cov.shape[0]
R = np.random.default_rng(2).standard_normal((3001, 6))
G = eng.gram_allreduce(R[par.shard_rows(3001, world, rank)])
assert np.allclose(G, R.T @ R, rtol=1e-12)
assert eng.allreduce([float(rank)], op="max")[0] == world - 1
r0, r1 = par.row_blocks(geom.nrow, world)[rank]
bg = par.block_geom(geom, r0, r1)
ens = eng.ensemble_create(bg, models, kept, w, wt, C + 2)
# the root alone holds the points; the other ranks pass nothing but the count
got, sp = eng.mltps_predict_shard(bg, ens, cov[:, r0:r1], xy if rank == 0 else None, resid if rank == 0 else None, len(resid), root=0)
lam = eng.allreduce([sp.lam], op="max")[0]
assert lam == sp.lam, "every rank evaluates the same spline"
if rank == 0:
    full = mb.Engine(local)                                   # no communicator: the unsharded path on the same GPU
    ens_f = full.ensemble_create(geom, models, kept, w, wt, C + 2)
    ref, sp_ref = full.mltps_predict(geom, ens_f, cov, xy, resid)
    assert sp_ref.lam == sp.lam
    np.save(os.environ["MB_OUT"], ref)
dist.barrier()
ref = np.load(os.environ["MB_OUT"])[r0:r1]
assert np.array_equal(np.isnan(got), np.isnan(ref))
m = ~np.isnan(ref)
err = np.max(np.abs(got[m] - ref[m])) / np.max(np.abs(ref[m]))
assert err < 1e-9, err
# a spline fitted on rank 1 travels to rank 0
sp1 = eng.tps_fit(xy[:200], resid[:200]) if rank == 1 else None
sp1 = eng.spline_bcast(sp1, 200, root=1)
pts = xy[200:260]
v = eng.tps_predict_points(sp1, pts)
vs = eng.allreduce(np.concatenate([v, -v]), op="max")
assert np.array_equal(vs[:60], v) and np.array_equal(vs[60:], -v), "both ranks hold the same spline"
dist.barrier()
eng.comm_destroy()
dist.destroy_process_group()
sys.stdout.write(f"rank{rank}ok err={err:.2e}\n"); sys.stdout.flush()
'''


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs (NCCL refuses two ranks on one device): run under gpurun --gpus 2")
def test_two_ranks_share_one_raster(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MB_ROOT=ROOT, MB_OUT=str(tmp_path / "ref.npy"))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("ok") == 2, r.stdout


WORKER_MERGE = r'''
import os, sys
import numpy as np
sys.path.insert(0, os.environ["MB_ROOT"])
import torch, torch.distributed as dist
import machisplin_b200 as mb
from machisplin_b200 import parallel as par, synth
sys.path.insert(0, os.path.join(os.environ["MB_ROOT"], "tests"))
from test_tiles_gpu import _merge_case
rank, world, local = par.env_rank()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
eng = mb.Engine(local)
par.comm_init(eng)
geom = synth.make_geom(240, 310)
for nc, nr in [(2, 1), (2, 2), (3, 2), (4, 2)]:
    wins, rasters = _merge_case(geom, nc, nr)
    nt = nc * nr
    mine = [t for t in range(nt) if t % world == rank]
    tiles = {t: torch.from_numpy(rasters[t]).to(dev) for t in mine}
    own = {t: eng.tiles_owned_window(geom, wins, nc, nr, t) for t in mine}
    outs = {t: torch.full((o[1] - o[0], o[3] - o[2]), -1.0, dtype=torch.float64, device=dev) for t, o in own.items()}
    eng.tiles_merge_shard_dev(geom, wins, {t: v.data_ptr() for t, v in tiles.items()}, nc, nr, {t: v.data_ptr() for t, v in outs.items()})
    torch.cuda.synchronize()
    full = mb.Engine(local)                                   # no communicator: the gathered merge of ALL tiles on this GPU
    ref = full.tiles_merge(geom, wins, rasters, nc, nr)
    full.close()
    for t, o in own.items():
        got = outs[t].cpu().numpy()
        assert np.array_equal(got, ref[o[0]:o[1], o[2]:o[3]], equal_nan=True), (nc, nr, t)
dist.barrier()
eng.comm_destroy()
dist.destroy_process_group()
sys.stdout.write(f"rank{rank}ok\n"); sys.stdout.flush()
'''


@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs (NCCL refuses two ranks on one device): run under gpurun --gpus 2")
def test_two_ranks_blend_tile_borders_without_a_gather(tmp_path):
    """machisplin.tiles.merge with the tiles on two ranks (tile t on rank t % 2): seam strips by ncclSend / ncclRecv, seam boxes by
    ncclAllReduce(min), every rank blends the cells its tiles own - bit-identical to the gathered merge."""
    script = tmp_path / "worker_merge.py"
    script.write_text(WORKER_MERGE)
    env = dict(os.environ, MB_ROOT=ROOT)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("ok") == 2, r.stdout
