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


# ------------------------------------------------------------------------------------------ shared host memory
class SharedHostArray(object):
    """A numpy array in POSIX shared memory that every rank of the node maps and page-locks for CUDA.

    With the source swath and the result grid in such arrays a multi-GPU call needs no host-side funnel: every
    rank uploads ITS share of the source through its own PCIe link (the shares are exchanged over NVLink) and
    writes ITS rows of the result straight into the one assembled grid.  ``create=True`` on one rank, then
    ``SharedHostArray(shape, dtype, name=...)`` on the others (:func:`share_host_array` does both)."""

    def __init__(self, shape, dtype, name=None, create=False):
        from multiprocessing import shared_memory
        self.shape, self.dtype = tuple(int(v) for v in shape), np.dtype(dtype)
        nbytes = max(1, int(np.prod(self.shape)) * self.dtype.itemsize)
        self._shm = shared_memory.SharedMemory(name=name, create=create, size=nbytes)
        if not create:
            # only the creating rank unlinks the segment: keep this process' resource tracker from trying as well
            try:
                from multiprocessing import resource_tracker
                resource_tracker.unregister(self._shm._name, "shared_memory")
            except Exception:  # pragma: no cover - tracker internals differ between Python versions
                pass
        self.name = self._shm.name
        self.array = np.ndarray(self.shape, dtype=self.dtype, buffer=self._shm.buf)
        self._owner = create
        self._registered = False
        self.nbytes = nbytes

    def register(self):
        """Page-lock the mapping for this process' CUDA context (cudaHostRegister): copies to and from the array
        are then asynchronous and run at full PCIe speed."""
        if not self._registered:
            import torch
            rc = torch.cuda.cudart().cudaHostRegister(self.array.ctypes.data, self.nbytes, 0)
            if int(rc) != 0:
                raise RuntimeError("cudaHostRegister failed with error %d" % int(rc))
            self._registered = True
        return self

    def tensor(self):
        import torch
        return torch.from_numpy(self.array)

    def close(self):
        if self._registered:
            import torch
            torch.cuda.cudart().cudaHostUnregister(self.array.ctypes.data)
            self._registered = False
        self.array = None
        self._shm.close()
        if self._owner:
            try:
                self._shm.unlink()
            except FileNotFoundError:  # pragma: no cover
                pass


def share_host_array(array_or_none, shape, dtype, src=0, group=None):
    """A :class:`SharedHostArray` holding ``array_or_none`` of rank ``src`` (or uninitialised memory when that is
    None), attached and registered on every rank."""
    import torch.distributed as dist
    rank, world = _world(group)
    shared = None
    names = [None]
    if rank == src:
        shared = SharedHostArray(shape, dtype, create=True)
        if array_or_none is not None:
            shared.array[...] = np.asarray(array_or_none).reshape(shape)
        names = [shared.name]
    if world > 1:
        dist.broadcast_object_list(names, src=src, group=group)
        if rank != src:
            shared = SharedHostArray(shape, dtype, name=names[0])
        dist.barrier(group=group)
    return shared.register()


def _my_chunk(n, world, rank):
    chunk = -(-n // world)
    a = min(n, rank * chunk)
    return chunk, a, min(n, a + chunk)


def resample_nearest_shared(lons, lats, data, out, target_geo_def, radius_of_influence, fill_value=0, group=None,
                            stripe=DEFAULT_STRIPE):
    """Multi-rank ``resample_nearest`` from and to shared host memory (:class:`SharedHostArray` on every rank).

    ``lons`` / ``lats``: ``(n_source,)`` coordinates, ``data``: ``(n_source, channels)`` rows, ``out``:
    ``(height, width, channels)`` result grid.  Every rank uploads 1/world of the source through its own PCIe link
    and the shares are exchanged with an all-gather (NCCL over NVLink); the coordinates go first, so that the index
    build overlaps the data's upload and exchange; every rank then resamples its stripes of rows and copies them to
    their place in ``out``.  After the closing barrier the assembled grid is complete in ``out.array`` on every
    rank.  Same result as ``kd_tree.resample_nearest`` of the whole target."""
    import ctypes

    import torch
    import torch.distributed as dist
    from . import _cabi, kd_tree
    rank, world = _world(group)
    target = geometry.as_geo_def(target_geo_def)
    height, width = target.shape
    n = int(lons.array.size)
    channels = int(data.array.shape[1]) if data.array.ndim > 1 else 1
    rows = shard_rows(height, world, rank) if stripe is None else cyclic_rows(height, world, rank, stripe)
    chunk, a, b = _my_chunk(n, world, rank)
    main = torch.cuda.current_stream()
    up, _ = kd_tree._side_streams()
    cdt = getattr(torch, lons.dtype.name)
    ddt = getattr(torch, data.dtype.name)
    # coordinates: this rank's share of (lon, lat), then everybody's
    mine = torch.zeros((2, chunk), dtype=cdt, device="cuda")
    if b > a:
        mine[0, :b - a].copy_(lons.tensor().reshape(-1)[a:b], non_blocking=True)
        mine[1, :b - a].copy_(lats.tensor().reshape(-1)[a:b], non_blocking=True)
    if world > 1:
        both = torch.empty((world, 2, chunk), dtype=cdt, device="cuda")
        dist.all_gather_into_tensor(both.view(-1), mine.view(-1), group=group)
        lon_t = both[:, 0, :].reshape(-1)[:n].contiguous()
        lat_t = both[:, 1, :].reshape(-1)[:n].contiguous()
    else:
        lon_t, lat_t = mine[0, :n], mine[1, :n]
    # data rows on the side stream, while the index is built from the coordinates
    up.wait_stream(main)
    with torch.cuda.stream(up):
        rows_mine = torch.zeros((chunk, channels), dtype=ddt, device="cuda")
        if b > a:
            rows_mine[:b - a].copy_(data.tensor().reshape(n, channels)[a:b], non_blocking=True)
        if world > 1:
            rows_all = torch.empty((world * chunk, channels), dtype=ddt, device="cuda")
            dist.all_gather_into_tensor(rows_all.view(-1), rows_mine.view(-1), group=group)
        else:
            rows_all = rows_mine
        arrived = torch.cuda.Event()
        arrived.record(up)
    swath = geometry.SwathDefinition(lon_t, lat_t)
    source = kd_tree._Source(swath, target, True, radius_of_influence, 1, need_ranks=False,
                             cover_rows=rows if world > 1 else None)
    tg, n_t, keep = kd_tree._make_target(source, target, True, radius_of_influence, rows=rows)
    fill_row = torch.full((channels,), fill_value, dtype=ddt, device="cuda")
    out_dev = torch.empty((max(n_t, 1), channels), dtype=ddt, device="cuda")
    main.wait_event(arrived)
    if n_t > 0:
        _cabi.check(_cabi.lib().prs_resample_nearest(source.handle, ctypes.byref(tg), float(radius_of_influence),
                                                    kd_tree._ptr(rows_all), channels * rows_all.element_size(),
                                                    kd_tree._ptr(fill_row), kd_tree._ptr(out_dev), kd_tree._stream()))
        # each stripe of rows is one contiguous piece of the assembled grid
        grid = out.tensor().reshape(height, width * channels)
        local = out_dev.reshape(-1, width * channels)
        if world == 1:
            g_rows = np.arange(height)
        elif stripe is not None:
            g_rows = cyclic_global_rows(height, world, rank, stripe)
        else:
            g_rows = np.arange(rows[0], rows[0] + rows[1])
        starts = np.flatnonzero(np.diff(g_rows, prepend=g_rows[0] - 2) != 1)
        ends = np.append(starts[1:], len(g_rows))
        for l0, l1 in zip(starts, ends):
            g0 = int(g_rows[l0])
            grid[g0:g0 + (l1 - l0)].copy_(local[l0:l1], non_blocking=True)
    del keep
    main.synchronize()
    if world > 1:
        dist.barrier(group=group)
    return out.array
