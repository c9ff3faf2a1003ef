"""Row sharding on one GPU: the ranks of a world are run one after the other (no process group) and their row sets
are assembled; the result must equal the unsharded run bit for bit.  This exercises the block-cyclic row mapping,
the global pixel generator and - above all - the coverage culling of the per-rank index (a source point wrongly
dropped would change an output pixel)."""
import numpy as np
import pytest

from helpers import jittered_swath, laea_area

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world,stripe", [(2, 8), (3, 4), (4, None), (8, 64)])
@pytest.mark.parametrize("area_kind", ["laea", "eqc"])
def test_sharded_equals_unsharded(cuda_lib, world, stripe, area_kind):
    import torch

    from pyresample_b200 import distributed, geometry, kd_tree
    lon, lat = jittered_swath(400, 300, np.float32, seed=41, lon0=-20, lat0=30, dlon=50, dlat=35)
    data = torch.from_numpy(np.random.default_rng(42).standard_normal((lon.size, 3)).astype(np.float32)).cuda()
    swath = geometry.SwathDefinition(torch.from_numpy(lon).cuda(), torch.from_numpy(lat).cuda())
    if area_kind == "laea":
        area = laea_area(333, 301, half=2500000.0, lat_0=48, lon_0=5)
    else:
        area = geometry.AreaDefinition("w", "w", "eqc", "+proj=eqc +datum=WGS84", 720, 361,
                                       (-20037508.342789244, -10018754.171394622, 20037508.342789244,
                                        10018754.171394622))
    radius = 30000
    full = kd_tree.resample_nearest(swath, data, area, radius, fill_value=-1).cpu().numpy()
    blocks = []
    n_indexed = []
    for rank in range(world):
        blk = distributed.resample_nearest_sharded(swath, data, area, radius, fill_value=-1, stripe=stripe,
                                                   rank=rank, world=world)
        blocks.append(blk.cpu().numpy())
        rows = distributed.shard_rows(area.shape[0], world, rank) if stripe is None else \
            distributed.cyclic_rows(area.shape[0], world, rank, stripe)
        src = kd_tree._Source(swath, area, True, radius, 1, need_ranks=False, cover_rows=rows)
        n_indexed.append(src.index.n_indexed if src.index is not None else 0)
    got = distributed.assemble_rows(blocks, area.shape[0], stripe)
    assert got.shape == full.shape
    assert np.array_equal(got, full)
    # the per-rank indices really are smaller than the full one
    full_src = kd_tree._Source(swath, area, True, radius, 1, need_ranks=False)
    assert max(n_indexed) < full_src.index.n_indexed
    assert sum(n_indexed) >= full_src.index.n_indexed * 0.5


def test_cyclic_row_plan():
    from pyresample_b200 import distributed
    for h, w, s in ((3600, 8, 64), (361, 3, 4), (10, 4, 64), (1000, 2, 8)):
        seen = np.concatenate([distributed.cyclic_global_rows(h, w, r, s) for r in range(w)])
        assert np.array_equal(np.sort(seen), np.arange(h))
        for r in range(w):
            row0, nrows, blk, stride = distributed.cyclic_rows(h, w, r, s)
            g = distributed.cyclic_global_rows(h, w, r, s)
            local = np.arange(nrows)
            assert np.array_equal(row0 + (local // blk) * stride + local % blk, g)


def test_resample_nearest_from_host_single_rank(cuda_lib):
    """The pipelined multi-rank entry point, run on one rank, equals the plain numpy API."""
    from pyresample_b200 import distributed, geometry, kd_tree
    rng = np.random.default_rng(71)
    lon = rng.uniform(-30, 30, 90000).astype(np.float32)
    lat = rng.uniform(-20, 35, 90000).astype(np.float32)
    data = rng.normal(size=(90000, 2)).astype(np.float32)
    area = geometry.AreaDefinition("g", "g", "g", "+proj=eqc +datum=WGS84", 900, 500,
                                   (-10018754.17, -6000000.0, 10018754.17, 7000000.0))
    got = distributed.resample_nearest_from_host(lon, lat, data, area, 40000, lon.size, lon.dtype, data.dtype, 2,
                                                 fill_value=-1)
    want = kd_tree.resample_nearest(geometry.SwathDefinition(lon, lat), data, area, 40000, fill_value=-1)
    assert got.shape == want.shape and np.array_equal(got, want)
    assert (want != -1).any()


def test_resample_nearest_shared_single_rank(cuda_lib):
    """Source and result in shared, page-locked host memory (the multi-rank end-to-end path), run on one rank."""
    from pyresample_b200 import distributed, geometry, kd_tree
    rng = np.random.default_rng(72)
    lon = rng.uniform(-30, 30, 70000).astype(np.float32)
    lat = rng.uniform(-20, 35, 70000).astype(np.float32)
    data = rng.normal(size=(70000, 3)).astype(np.float32)
    area = geometry.AreaDefinition("g", "g", "g", "+proj=eqc +datum=WGS84", 800, 450,
                                   (-10018754.17, -6000000.0, 10018754.17, 7000000.0))
    s_lon = distributed.share_host_array(lon, lon.shape, lon.dtype)
    s_lat = distributed.share_host_array(lat, lat.shape, lat.dtype)
    s_data = distributed.share_host_array(data, data.shape, data.dtype)
    s_out = distributed.share_host_array(None, area.shape + (3,), data.dtype)
    try:
        got = distributed.resample_nearest_shared(s_lon, s_lat, s_data, s_out, area, 40000, fill_value=-1).copy()
        want = kd_tree.resample_nearest(geometry.SwathDefinition(lon, lat), data, area, 40000, fill_value=-1)
        assert got.shape == want.shape and np.array_equal(got, want)
        assert (want != -1).any()
    finally:
        for s in (s_lon, s_lat, s_data, s_out):
            s.close()


@pytest.mark.parametrize("proj,world,stripe", [("geos", 8, 16), ("geos", 3, 8), ("laea", 4, 8), ("stere", 2, None)])
def test_area_source_cropped_to_the_ranks_coverage(cuda_lib, proj, world, stripe):
    """An AREA source of a sharded nearest run is not transformed where the rank cannot need it
    (prs_area_lonlats_covered): the index holds exactly the records it holds without the crop, the pixels that were
    transformed carry the same coordinates, and the assembled result equals the unsharded one."""
    import torch

    from pyresample_b200 import distributed, geometry, kd_tree
    if proj == "geos":
        src = geometry.AreaDefinition("fd", "fd", "geos", "+proj=geos +h=35786023 +lon_0=-75 +sweep=x +ellps=GRS80",
                                      2800, 2800, (-5434894.9, -5434894.9, 5434894.9, 5434894.9))
        tgt = geometry.AreaDefinition("ll", "ll", "longlat", "+proj=longlat +ellps=GRS80", 1800, 1280,
                                      (-156.0, -80.0, 6.0, 80.0))
        radius = 8000.0
    elif proj == "laea":
        src = laea_area(2400, 2300, half=3000000.0, lat_0=50, lon_0=10)
        tgt = geometry.AreaDefinition("w", "w", "eqc", "+proj=eqc +datum=WGS84", 720, 361,
                                      (-20037508.342789244, -10018754.171394622, 20037508.342789244, 10018754.171394622))
        radius = 30000.0
    else:
        src = geometry.AreaDefinition("ps", "ps", "stere", "+proj=stere +lat_0=90 +lat_ts=70 +lon_0=-45 +ellps=WGS84",
                                      2400, 2400, (-3000000.0, -3000000.0, 3000000.0, 3000000.0))
        tgt = laea_area(1000, 800, half=2500000.0, lat_0=75, lon_0=-30)
        radius = 12000.0
    data = torch.from_numpy(np.random.default_rng(7).standard_normal(src.size).astype(np.float32)).cuda()
    full = kd_tree.resample_nearest(src, data, tgt, radius, fill_value=-1).cpu().numpy()
    assert (full != -1).mean() > 0.05
    lon_full, lat_full, _ = kd_tree._geo_lonlats_device(src)
    blocks, dropped = [], 0
    for rank in range(world):
        rows = distributed.shard_rows(tgt.shape[0], world, rank) if stripe is None else \
            distributed.cyclic_rows(tgt.shape[0], world, rank, stripe)
        plain = kd_tree._Source(src, tgt, True, radius, 1, need_ranks=False, cover_rows=rows)
        crop = kd_tree._Source(src, tgt, True, radius, 1, need_ranks=False, cover_rows=rows, crop_to_cover=True)
        assert crop.index.n_indexed == plain.index.n_indexed
        there = ~torch.isnan(crop.lon_t)
        assert torch.equal(crop.lon_t[there], lon_full[there]) and torch.equal(crop.lat_t[there], lat_full[there])
        dropped += int((~there & ~torch.isnan(lon_full) & (lon_full.abs() <= 180)).sum())
        blocks.append(kd_tree._resample(src, data, tgt, 'nn', radius, neighbours=1, fill_value=-1, _rows=rows).cpu().numpy())
    got = distributed.assemble_rows(blocks, tgt.shape[0], stripe)
    assert np.array_equal(got, full)
    n_valid = int((~torch.isnan(lon_full) & (lon_full.abs() <= 180)).sum())
    assert dropped > 0.05 * (world - 1) * n_valid  # the crop really skips part of the source (how much: stripe spacing)
