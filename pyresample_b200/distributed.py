"""Multi-GPU execution of the kd-tree resampling path: one process per GPU, target rows sharded.

The reference already processes the target in independent row blocks (``segments``,
pyresample/kd_tree.py:343-366 via geometry._get_slice, geometry.py:3052-3068); that is the shard axis here.
Rank ``r`` of ``world`` owns rows ``_get_slice(world, shape)[r]`` of the target, generates the coordinates of
those rows itself (no input traffic) and needs no exchange with other ranks during the query.  The source
swath (lon, lat, data) is broadcast once from rank 0 with ``torch.distributed`` (NCCL over NVLink on GPUs,
gloo in the CPU tests); every rank then builds the same deterministic index.  An optional all-gather
assembles the full grid on every rank.
"""
from __future__ import annotations

import numpy as np

from . import geometry


def shard_rows(height, world_size, rank):
    """(row0, nrows) of ``rank``: contiguous blocks of ceil(height / world_size) rows, like geometry._get_slice."""
    blocks = [s[0] for s in geometry._get_slice(world_size, (height, 1))]
    if rank >= len(blocks):
        return height, 0
    sl = blocks[rank]
    return sl.start, sl.stop - sl.start


def shard_plan(height, world_size):
    """[(row0, nrows)] for every rank (ranks beyond the last block get zero rows)."""
    return [shard_rows(height, world_size, r) for r in range(world_size)]


def broadcast_source(tensors, src=0, group=None):
    """Broadcast the source tensors (lon, lat, data, ...) from rank ``src`` in place; returns them."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return tensors
    for t in tensors:
        dist.broadcast(t, src=src, group=group)
    return tensors


def all_gather_rows(local_block, height, group=None):
    """Assemble the (height, ...) result from every rank's row block (blocks may differ in size)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_block
    world = dist.get_world_size(group)
    plan = shard_plan(height, world)
    max_rows = max(n for _, n in plan)
    pad = torch.zeros((max_rows,) + tuple(local_block.shape[1:]), dtype=local_block.dtype, device=local_block.device)
    pad[:local_block.shape[0]] = local_block
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:n] for p, (_, n) in zip(parts, plan)], dim=0)


def sharded_rows_apply(block_fn, height, group=None, gather=False):
    """Run ``block_fn(row0, nrows) -> tensor (nrows, ...)`` for this rank's row block; optionally all-gather."""
    import torch.distributed as dist
    rank, world = 0, 1
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    row0, nrows = shard_rows(height, world, rank)
    out = block_fn(row0, nrows)
    if gather:
        return all_gather_rows(out, height, group)
    return out


def resample_nearest_sharded(source_geo_def, data, target_geo_def, radius_of_influence, fill_value=0,
                             reduce_data=True, group=None, gather=False):
    """``kd_tree.resample_nearest`` for this rank's row block of ``target_geo_def`` (CUDA tensors in/out).

    Returns the local block ``(nrows, width[, C])`` or, with ``gather=True``, the full grid on every rank.
    """
    from . import kd_tree
    target = geometry.as_geo_def(target_geo_def)

    def block(row0, nrows):
        return kd_tree._resample(source_geo_def, data, target, 'nn', radius_of_influence, neighbours=1,
                                 fill_value=fill_value, reduce_data=reduce_data, _rows=(row0, nrows))

    return sharded_rows_apply(block, target.shape[0], group, gather)


def concat_row_blocks(blocks):
    """Host-side concatenation of per-rank row blocks (RowAppendableArray semantics, utils/row_appendable_array.py)."""
    return np.concatenate(blocks, axis=0)
