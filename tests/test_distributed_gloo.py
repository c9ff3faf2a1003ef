"""world_size-2 gloo test of the multi-GPU host logic on CPU: row sharding, source broadcast, all-gather.

The per-rank compute is the oracle here (the CUDA kernels need a GPU); what is under test is that rank r owning
rows shard_rows(H, W, r) with the *global* pixel-centre generator reproduces the unsharded result bit for bit.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from helpers import jittered_swath, laea_area
    from oracle import kd_tree_ref, proj_ref
    from oracle.knn_ref import KDTreeRef
    from pyresample_b200 import distributed
    area = laea_area(90, 77)
    # source exists on rank 0 only and is broadcast once
    if rank == 0:
        lon, lat = jittered_swath(120, 100, np.float64, seed=21)
        data = np.random.default_rng(22).standard_normal((lon.size, 2))
        tensors = [torch.from_numpy(lon.copy()), torch.from_numpy(lat.copy()), torch.from_numpy(data)]
    else:
        tensors = [torch.empty((120, 100), dtype=torch.float64), torch.empty((120, 100), dtype=torch.float64),
                   torch.empty((12000, 2), dtype=torch.float64)]
    lon_t, lat_t, data_t = distributed.broadcast_source(tensors)
    lon, lat, data = lon_t.numpy(), lat_t.numpy(), data_t.numpy()
    tree = KDTreeRef(kd_tree_ref.transform_lonlats(lon.ravel(), lat.ravel()))

    def block(row0, nrows):
        tlon, tlat = proj_ref.area_lonlats(area.proj_dict, area.area_extent, area.width, area.height,
                                           data_slice=slice(row0, row0 + nrows))
        q = kd_tree_ref.transform_lonlats(tlon.ravel(), tlat.ravel())
        _, idx = tree.query(q, k=1, distance_upper_bound=50000)
        out = data[np.where(idx == tree.n, 0, idx)]
        out[idx == tree.n] = -1
        return torch.from_numpy(out.reshape(nrows, area.width, 2))

    local = distributed.sharded_rows_apply(block, area.height)
    row0, nrows = distributed.shard_rows(area.height, world, rank)
    assert local.shape == (nrows, area.width, 2)
    full = distributed.sharded_rows_apply(block, area.height, gather=True)
    assert full.shape == (area.height, area.width, 2)
    np.save(os.path.join(out_dir, "full_%d.npy" % rank), full.numpy())
    if rank == 0:
        np.savez(os.path.join(out_dir, "src.npz"), lon=lon, lat=lat, data=data)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_row_sharding(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import laea_area
    from oracle import kd_tree_ref
    from pyresample_b200 import geometry
    src = np.load(tmp_path / "src.npz")
    swath = geometry.SwathDefinition(src["lon"], src["lat"])
    want = kd_tree_ref.resample_nearest(swath, src["data"], laea_area(90, 77), 50000, fill_value=-1, reduce_data=False)
    for r in range(world):
        got = np.load(tmp_path / ("full_%d.npy" % r))
        assert np.array_equal(got, want)


def test_shard_rows_uneven():
    from pyresample_b200 import distributed
    assert distributed.shard_plan(10, 4) == [(0, 3), (3, 3), (6, 3), (9, 1)]
    assert distributed.shard_plan(2, 4) == [(0, 1), (1, 1), (2, 0), (2, 0)]


def _shared_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pyresample_b200 import distributed
    # rank 0 owns the array, everybody maps it (register() needs CUDA and is not called here)
    names = [None]
    shared = None
    if rank == 0:
        shared = distributed.SharedHostArray((6, 5), np.float32, create=True)
        shared.array[...] = np.arange(30, dtype=np.float32).reshape(6, 5)
        names = [shared.name]
    dist.broadcast_object_list(names, src=0)
    if rank != 0:
        shared = distributed.SharedHostArray((6, 5), np.float32, name=names[0])
    dist.barrier()
    assert shared.array[3, 2] == 17.0
    # every rank writes the rows it owns under the block-cyclic plan: together they cover the grid exactly once
    g = distributed.cyclic_global_rows(6, world, rank, stripe=2)
    shared.array[g] = 100 + rank
    dist.barrier()
    full = shared.array.copy()
    np.save(os.path.join(out_dir, "shared_%d.npy" % rank), full)
    chunk, a, b = distributed._my_chunk(11, world, rank)
    assert chunk == 6 and (a, b) == ((0, 6) if rank == 0 else (6, 11))
    dist.barrier()
    shared.close()
    dist.destroy_process_group()


def test_shared_host_array_two_ranks(tmp_path):
    world = 2
    mp.spawn(_shared_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    a, b = np.load(tmp_path / "shared_0.npy"), np.load(tmp_path / "shared_1.npy")
    assert np.array_equal(a, b)
    assert np.array_equal(a[:, 0], [100, 100, 101, 101, 100, 100])
