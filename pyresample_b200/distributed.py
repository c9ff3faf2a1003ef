"""Multi-GPU execution of the kd-tree resampling path: one process per GPU, target rows sharded.

The reference already processes the target in independent row blocks (``segments``,
pyresample/kd_tree.py:343-366 via geometry._get_slice, geometry.py:3052-3068); rows are the shard axis here too.

* **Row ownership.**  ``shard_rows`` gives rank ``r`` the contiguous block ``_get_slice(world, shape)[r]``
  (the reference's segment arithmetic).  ``cyclic_rows`` deals stripes of ``stripe`` rows round-robin
  (rank r owns global rows g with ``(g // stripe) % world == r``): a swath usually covers only part of a
  grid, and stripes give every rank the same share of the covered rows.  Either way a rank generates the
  coordinates of its rows itself from the *global* pixel generator (global row index, identical rounding).
* **Source.**  Broadcast once from rank 0 (``torch.distributed``: NCCL over NVLink on GPUs, gloo in the CPU
  tests).  Every rank then indexes only the source points that can be within the radius of one of ITS rows
  (coverage mask, ``prs_target_coverage`` -> ``cover`` of ``prs_index_build``): index build and query both shrink
  with the number of ranks, and no rank exchanges anything during the query.
* **Result.**  Each rank holds its rows; ``all_gather_rows`` / ``assemble_rows`` put the grid back together.
"""
from __future__ import annotations

import numpy as np

from . import geometry

DEFAULT_STRIPE = 64  # rows per stripe of the block-cyclic distribution (a multiple of the 4-row query tiles)


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


def cyclic_rows(height, world_size, rank, stripe=DEFAULT_STRIPE):
    """Row specification ``(row0, nrows, row_block, row_stride)`` of ``rank`` under the block-cyclic distribution."""
    if world_size == 1:
        return 0, height, 0, 0
    return rank * stripe, len(cyclic_global_rows(height, world_size, rank, stripe)), stripe, world_size * stripe


def cyclic_global_rows(height, world_size, rank, stripe=DEFAULT_STRIPE):
    """Global row index of every local row of ``rank`` (ascending)."""
    g = np.arange(height)
    return g[(g // stripe) % world_size == rank]


def broadcast_source(tensors, src=0, group=None):
    """Broadcast the source tensors (lon, lat, data, ...) from rank ``src`` in place; returns them."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return tensors
    for t in tensors:
        dist.broadcast(t, src=src, group=group)
    return tensors


def _world(group=None):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def all_gather_rows(local_block, height, group=None, stripe=None):
    """Assemble the (height, ...) result on every rank from the ranks' row sets (which may differ in size).

    ``stripe=None``: contiguous blocks of ``shard_rows``; otherwise the block-cyclic rows of ``cyclic_rows``.
    """
    import torch
    import torch.distributed as dist
    rank, world = _world(group)
    if world == 1:
        return local_block
    if stripe is None:
        counts = [n for _, n in shard_plan(height, world)]
    else:
        counts = [len(cyclic_global_rows(height, world, r, stripe)) for r in range(world)]
    max_rows = max(counts)
    pad = torch.zeros((max_rows,) + tuple(local_block.shape[1:]), dtype=local_block.dtype, device=local_block.device)
    pad[:local_block.shape[0]] = local_block
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    if stripe is None:
        return torch.cat([p[:n] for p, n in zip(parts, counts)], dim=0)
    out = torch.empty((height,) + tuple(local_block.shape[1:]), dtype=local_block.dtype, device=local_block.device)
    for r, (p, n) in enumerate(zip(parts, counts)):
        rows = torch.as_tensor(cyclic_global_rows(height, world, r, stripe), device=local_block.device)
        out[rows] = p[:n]
    return out


def assemble_rows(blocks, height, stripe=None):
    """Host-side assembly of per-rank numpy blocks (RowAppendableArray semantics for contiguous blocks,
    pyresample/utils/row_appendable_array.py:23-56)."""
    world = len(blocks)
    if stripe is None or world == 1:
        return np.concatenate(blocks, axis=0)
    out = np.empty((height,) + tuple(blocks[0].shape[1:]), dtype=blocks[0].dtype)
    for r, b in enumerate(blocks):
        out[cyclic_global_rows(height, world, r, stripe)] = b
    return out


def sharded_rows_apply(block_fn, height, group=None, gather=False):
    """Run ``block_fn(row0, nrows) -> tensor (nrows, ...)`` for this rank's contiguous row block; optionally all-gather."""
    rank, world = _world(group)
    row0, nrows = shard_rows(height, world, rank)
    out = block_fn(row0, nrows)
    if gather:
        return all_gather_rows(out, height, group)
    return out


def resample_nearest_sharded(source_geo_def, data, target_geo_def, radius_of_influence, fill_value=0,
                             reduce_data=True, group=None, gather=False, stripe=DEFAULT_STRIPE, rank=None, world=None):
    """``kd_tree.resample_nearest`` for this rank's rows of ``target_geo_def`` (CUDA tensors in/out).

    ``stripe`` rows are dealt round-robin (``stripe=None``: contiguous blocks).  Returns the local rows
    ``(nrows, width[, C])`` or, with ``gather=True``, the full grid on every rank.  ``rank`` / ``world`` default
    to the process group's.
    """
    from . import kd_tree
    target = geometry.as_geo_def(target_geo_def)
    if rank is None or world is None:
        rank, world = _world(group)
    height = target.shape[0]
    rows = shard_rows(height, world, rank) if stripe is None else cyclic_rows(height, world, rank, stripe)
    out = kd_tree._resample(source_geo_def, data, target, 'nn', radius_of_influence, neighbours=1,
                            fill_value=fill_value, reduce_data=reduce_data, _rows=rows)
    if gather:
        return all_gather_rows(out, height, group, stripe)
    return out


class _UploadAndBroadcast(object):
    """Data arrival for kd_tree._nearest_pipelined on several ranks: ``src`` uploads the rows from its host memory on
    the side stream, then they are broadcast (NCCL) - both while every rank classifies and fills its own rows."""

    def __init__(self, host_rows, shape, src, group):
        self.host_rows, self.shape, self.src, self.group = host_rows, tuple(shape), src, group
        self.row_bytes = self.shape[1]
        self.start_early = True

    def start(self, main, up):
        import torch
        rank, _ = _world(self.group)
        self.ready = None
        if rank == self.src:
            up.wait_stream(main)
            with torch.cuda.stream(up):
                from . import kd_tree
                self.dev = kd_tree._dev(self.host_rows)
                self.ready = torch.cuda.Event()
                self.ready.record(up)
        else:
            self.dev = torch.empty(self.shape, dtype=torch.uint8, device="cuda")

    def finish(self, main):
        import torch.distributed as dist
        if self.ready is not None:
            main.wait_event(self.ready)
        dist.broadcast(self.dev, src=self.src, group=self.group)
        return self.dev

    def finish_without_start(self):
        """For a rank that has nothing to resample but must take part in the broadcast."""
        import torch
        main = torch.cuda.current_stream()
        self.start(main, main)
        return self.finish(main)


def resample_nearest_from_host(lons, lats, data, target_geo_def, radius_of_influence, n_source, coord_dtype,
                               data_dtype, channels, fill_value=0, src=0, group=None, stripe=DEFAULT_STRIPE):
    """Multi-rank ``resample_nearest`` whose source lives in the HOST memory of rank ``src``: ``lons``/``lats``
    (``n_source`` points of ``coord_dtype``) and ``data`` (``(n_source[, channels])`` of ``data_dtype``) are numpy
    arrays there and ``None`` on the other ranks, which only know the sizes and dtypes.

    The coordinates are uploaded and broadcast first; every rank then builds its index and fills its rows while
    the data follow (upload on ``src`` + broadcast), reads back the bands of rows that are already final, searches
    the rest and reads it back: each rank returns its own rows ``(nrows, width[, channels])`` as a host array.
    Same result as ``resample_nearest_sharded`` + a device-to-host copy.
    """
    import torch
    import torch.distributed as dist
    from . import _cabi, kd_tree
    rank, world = _world(group)
    target = geometry.as_geo_def(target_geo_def)
    height = target.shape[0]
    rows = shard_rows(height, world, rank) if stripe is None else cyclic_rows(height, world, rank, stripe)
    coord_dtype, data_dtype = np.dtype(coord_dtype), np.dtype(data_dtype)
    row_bytes = channels * data_dtype.itemsize
    ll = torch.empty((2, n_source), dtype=getattr(torch, coord_dtype.name), device="cuda")
    host_rows = None
    if rank == src:
        ll[0].copy_(torch.from_numpy(np.ascontiguousarray(lons, dtype=coord_dtype).ravel()), non_blocking=True)
        ll[1].copy_(torch.from_numpy(np.ascontiguousarray(lats, dtype=coord_dtype).ravel()), non_blocking=True)
        host_rows = kd_tree._byte_rows(np.ma.getdata(data))
    if world > 1:
        dist.broadcast(ll, src=src, group=group)
    swath = geometry.SwathDefinition(ll[0], ll[1])
    source = kd_tree._Source(swath, target, True, radius_of_influence, 1, need_ranks=False,
                             cover_rows=rows if world > 1 else None)
    fill_row_host = kd_tree._byte_rows(np.full((1, channels), fill_value, dtype=data_dtype))
    tail = (channels,) if channels > 1 else ()
    if source.index is None:  # no valid source point at all: every pixel is fill (kd_tree.py:336-341)
        if world > 1:  # stay in step with the collective of the other ranks
            _UploadAndBroadcast(host_rows, (n_source, row_bytes), src, group).finish_without_start()
        return np.full((rows[1], target.shape[1]) + tail, fill_value, dtype=data_dtype)
    tg, n_t, keep = kd_tree._make_target(source, target, True, radius_of_influence, rows=rows)
    if n_t == 0:
        if world > 1:
            _UploadAndBroadcast(host_rows, (n_source, row_bytes), src, group).finish_without_start()
        return np.empty((0, target.shape[1]) + tail, dtype=data_dtype)
    arrival = _UploadAndBroadcast(host_rows, (n_source, row_bytes), src, group) if world > 1 \
        else kd_tree._HostUpload(host_rows)
    L = _cabi.lib()
    res = kd_tree._nearest_pipelined(L, source, tg, n_t, arrival, fill_row_host, radius_of_influence)
    del keep
    res = res.view(data_dtype)
    nrows = n_t // target.shape[1]
    return res.reshape((nrows, target.shape[1]) + tail)


def concat_row_blocks(blocks):
    """Host-side concatenation of contiguous per-rank row blocks."""
    return np.concatenate(blocks, axis=0)
