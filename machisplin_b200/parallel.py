"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests).

The path shards with no data-path collective: cells are independent given the spline and
model descriptors, so ranks own whole tiles (``machisplin.tiles.create`` boundaries) or contiguous row blocks.
The only reduction is the K x K Gram of the cross-validation residuals (V73:329-333) when the residual rows
are sharded; timing is reported as the max over ranks."""
from __future__ import annotations

import os
from typing import List, Sequence, Tuple

import numpy as np


def env_rank() -> Tuple[int, int, int]:
    """(rank, world, local_rank) from the torchrun environment (1 process = 1 GPU)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend: str | None = None, device=None):
    """Initialise the default process group from the environment; no-op for world size 1."""
    import torch
    import torch.distributed as dist
    rank, world, _ = env_rank()
    if world == 1 or dist.is_initialized():
        return rank, world
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
    dist.init_process_group(backend, **kw)
    return rank, world


def tiles_of_rank(n_tiles: int, world: int, rank: int) -> List[int]:
    """Round-robin tile ownership: tile t belongs to rank t mod world (tiles are numbered from the SW corner)."""
    return [t for t in range(n_tiles) if t % world == rank]


def row_blocks(nrow: int, world: int, align: int = 32) -> List[Tuple[int, int]]:
    """Contiguous row blocks [r0, r1) per rank for the global-spline mode: balanced, boundaries aligned to the
    leaf-box height so that no box is split between ranks.  Empty blocks are (r, r)."""
    units = (nrow + align - 1) // align
    out, start = [], 0
    for r in range(world):
        cnt = units // world + (1 if r < units % world else 0)
        r0, r1 = min(nrow, start * align), min(nrow, (start + cnt) * align)
        out.append((r0, r1))
        start += cnt
    return out


def block_geom(geom, r0: int, r1: int):
    """Extent + shape of rows [r0, r1) of a raster: what a rank passes as ``g_block`` to ``mb_mltps_predict_shard*`` and creates
    its ensemble handle for (same xmin / xmax, so LONG / LAT of a cell are those of the full raster)."""
    from .engine import Geom, as_geom
    g = as_geom(geom)
    return Geom(g.xmin, g.xmax, g.ymax - r1 * g.ry, g.ymax - r0 * g.ry, r1 - r0, g.ncol)


def comm_init(engine, backend_group=None):
    """Give ``engine`` the library's own NCCL communicator over the ranks of the torch process group: rank 0 draws the
    unique id (``mb_comm_unique_id``), the HOST side ships its 128 bytes (here: ``torch.distributed``; an R host would use a
    file or a socket), every rank calls ``mb_comm_init``.  Data-path collectives then run inside the library."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0, 1
    rank, world = dist.get_rank(), dist.get_world_size()
    box = [engine.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=backend_group)
    engine.comm_init(world, rank, box[0])
    return rank, world


def shard_rows(n: int, world: int, rank: int) -> slice:
    """Contiguous shard of n cross-validation residual rows."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return slice(lo, lo + base + (1 if rank < extra else 0))


def allreduce_gram(G_local: np.ndarray, device=None) -> np.ndarray:
    """SUM of the per-rank Gram matrices (<= 8 x 8 doubles): G = sum_r R_r' R_r = R'R."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return np.array(G_local, dtype=np.float64)
    t = torch.as_tensor(np.ascontiguousarray(G_local, dtype=np.float64))
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def max_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_tiles(local: Sequence[Tuple[int, np.ndarray]], n_tiles: int, dst: int = 0):
    """Collect (tile index, raster) pairs on rank ``dst`` (the rank that runs ``tiles_merge`` and writes the
    output).  Returns the list ordered by tile index on ``dst``, None elsewhere."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        got = dict(local)
        return [got[t] for t in range(n_tiles)]
    buf = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object(list(local), buf, dst=dst)
    if dist.get_rank() != dst:
        return None
    got = {t: r for part in buf for (t, r) in part}
    assert len(got) == n_tiles, "a tile is missing: ownership map and results disagree"
    return [got[t] for t in range(n_tiles)]


def gather_tiles_device(local_tile, tile_shapes: Sequence[Tuple[int, int]], dst: int = 0):
    """Tile-border blend, step 1 (V73:1392-1548 across GPUs): every rank owns the ``$final`` raster of tile
    ``rank`` as a float64 torch tensor on its GPU; rank ``dst`` receives all of them over NCCL (point-to-point
    on NVLink, device to device - nothing is staged on the host) and then runs ``Engine.tiles_merge_dev``.
    Returns the list of tile tensors (tile order) on ``dst``, None elsewhere."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [local_tile]
    rank, world = dist.get_rank(), dist.get_world_size()
    assert len(tile_shapes) == world, "one tile per rank"
    if rank != dst:
        dist.send(local_tile.contiguous(), dst=dst)
        return None
    tiles = []
    for r in range(world):
        if r == dst:
            tiles.append(local_tile)
        else:
            buf = torch.empty(tuple(tile_shapes[r]), dtype=local_tile.dtype, device=local_tile.device)
            dist.recv(buf, src=r)
            tiles.append(buf)
    return tiles
