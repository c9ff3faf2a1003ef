"""BASELINE.json's full-size configurations through size-independent properties (the oracle cannot run them in
seconds): two independent code paths must agree bit for bit, sharded == unsharded, distances sorted, every
reported neighbour within the radius and really at the reported distance."""
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c2(cuda_lib):
    import torch

    import workloads
    from pyresample_b200 import geometry
    cfg = workloads.make_config(2)
    cfg["lon_t"] = torch.from_numpy(cfg["lon"]).cuda()
    cfg["lat_t"] = torch.from_numpy(cfg["lat"]).cuda()
    cfg["data_t"] = torch.from_numpy(cfg["data"]).cuda()
    cfg["swath_t"] = geometry.SwathDefinition(cfg["lon_t"], cfg["lat_t"])
    return cfg


def test_c2_fused_equals_info_plus_gather(c2):
    """resample_nearest (fused kernels) == get_neighbour_info + get_sample_from_neighbour_info (separate kernels)."""
    from pyresample_b200 import geometry, kd_tree
    swath = geometry.SwathDefinition(c2["lon"], c2["lat"])
    area, radius = c2["area"], c2["radius"]
    fused = kd_tree.resample_nearest(c2["swath_t"], c2["data_t"], area, radius, fill_value=-7).cpu().numpy()
    vii, voi, ia, da = kd_tree.get_neighbour_info(swath, area, radius, neighbours=1)
    assert vii.all() and voi.all() and ia.dtype == np.uint32
    two_step = kd_tree.get_sample_from_neighbour_info('nn', area.shape, c2["data"], vii, voi, ia, fill_value=-7)
    assert np.array_equal(fused, two_step, equal_nan=True)
    found = ia < vii.sum()
    assert 0.01 < found.mean() < 0.25  # the swath covers a few percent of the globe
    assert np.all(da[found] < radius) and np.all(np.isinf(da[~found]))
    # the reported neighbour really is at the reported distance (recomputed on the host in float32)
    from oracle import kd_tree_ref, proj_ref
    sel = np.flatnonzero(found)[::997]
    rows, cols = np.unravel_index(sel, area.shape)
    tx, ty = proj_ref.proj_vectors(area.area_extent, area.width, area.height, np.float32)
    tlon, tlat = proj_ref.inverse(area.proj_dict, tx[cols], ty[rows])
    q = kd_tree_ref.transform_lonlats(tlon.astype(np.float32), tlat.astype(np.float32))
    p = kd_tree_ref.transform_lonlats(c2["lon"].ravel()[ia[sel]], c2["lat"].ravel()[ia[sel]])
    d = np.sqrt((((q - p) ** 2)[:, 0] + ((q - p) ** 2)[:, 1]) + ((q - p) ** 2)[:, 2])
    np.testing.assert_allclose(da[sel], d, rtol=2e-4)


def test_c2_sharded_equals_unsharded(c2):
    from pyresample_b200 import distributed, kd_tree
    area, radius = c2["area"], c2["radius"]
    full = kd_tree.resample_nearest(c2["swath_t"], c2["data_t"], area, radius).cpu().numpy()
    blocks = [distributed.resample_nearest_sharded(c2["swath_t"], c2["data_t"], area, radius, rank=r, world=8)
              .cpu().numpy() for r in range(8)]
    got = distributed.assemble_rows(blocks, area.shape[0], distributed.DEFAULT_STRIPE)
    assert np.array_equal(got, full, equal_nan=True)


def test_c2_k8_sorted_and_consistent(c2):
    """k = 8 on the full grid: distances ascending, first column identical to the k = 1 search, no duplicates."""
    from pyresample_b200 import geometry, kd_tree
    swath = geometry.SwathDefinition(c2["lon"], c2["lat"])
    area, radius = c2["area"], c2["radius"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _, _, ia8, da8 = kd_tree.get_neighbour_info(swath, area, radius, neighbours=8)
    _, _, ia1, da1 = kd_tree.get_neighbour_info(swath, area, radius, neighbours=1)
    assert np.array_equal(ia8[:, 0], ia1) and np.array_equal(da8[:, 0], da1)
    fin = np.isfinite(da8)
    d = np.where(fin, da8, np.float32(3e38))
    assert np.all(d[:, 1:] >= d[:, :-1])
    n = c2["lon"].size
    some = ia8[fin.all(axis=1)][::101]
    assert all(len(set(r)) == 8 for r in some) and some.max() < n
    assert np.all(ia8[~fin] == n)


# ------------------------------------------------------------------------------------------------ C3 (full size)
@pytest.fixture(scope="module")
def c3(cuda_lib):
    import torch

    import workloads
    from pyresample_b200 import geometry
    cfg = workloads.make_config(3)
    cfg["data_t"] = torch.from_numpy(cfg["data"]).cuda()
    cfg["swath_t"] = geometry.SwathDefinition(torch.from_numpy(cfg["lon"]).cuda(), torch.from_numpy(cfg["lat"]).cuda())
    return cfg


def test_c3_fused_gauss_equals_info_plus_weighted_gather(c3):
    """resample_gauss k=8, 16 channels on the 4000x4000 polar grid: the fused epilogue and the two-step path
    (prs_query, then prs_gather_weighted on the stored info) accumulate in the same order -> same bits."""
    import torch

    from pyresample_b200 import geometry, kd_tree
    area, radius, k = c3["area"], c3["radius"], c3["k"]
    sig = [c3["sigma"]] * c3["channels"]
    funcs = [kd_tree._make_gauss(s) for s in sig]
    fused = kd_tree._resample(c3["swath_t"], c3["data_t"], area, 'custom', radius, neighbours=k, weight_funcs=funcs,
                              fill_value=-1)
    assert fused.dtype == torch.float64 and tuple(fused.shape) == (4000, 4000, 16)
    swath = geometry.SwathDefinition(c3["lon"], c3["lat"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        vii, voi, ia, da = kd_tree.get_neighbour_info(swath, area, radius, neighbours=k)
    assert voi.all() and ia.shape == (16000000, 8) and da.dtype == np.float32
    n_valid = int(vii.sum())
    fin = np.isfinite(da)
    assert np.array_equal(fin, ia < n_valid)
    assert 0.05 < fin[:, 0].mean() < 0.6
    d = np.where(fin, da, np.float32(3e38))
    assert np.all(d[:, 1:] >= d[:, :-1]) and np.all(da[fin] < radius)
    full_rows = ia[fin.all(axis=1)][::1009]
    assert all(len(set(r)) == 8 for r in full_rows)
    # channels 0-3 through the two-step path (the full 16-channel float64 grid would be 2 GB of host memory twice)
    two = kd_tree.get_sample_from_neighbour_info('custom', area.shape, c3["data"][:, :4], vii, voi, ia,
                                                 distance_array=da, weight_funcs=funcs[:4], fill_value=-1)
    got = fused[:, :, :4].cpu().numpy()
    assert np.array_equal(got, two, equal_nan=True)
    # pixels without any neighbour are fill, pixels with one are not (unless the data are NaN there)
    none = ~fin[:, 0].reshape(area.shape)
    assert np.all(got[none] == -1)


def test_c3_sharded_equals_unsharded(c3):
    from pyresample_b200 import distributed, kd_tree
    area, radius, k = c3["area"], c3["radius"], c3["k"]
    funcs = [kd_tree._make_gauss(c3["sigma"])] * c3["channels"]
    full = kd_tree._resample(c3["swath_t"], c3["data_t"], area, 'custom', radius, neighbours=k, weight_funcs=funcs)
    for world in (8,):
        blocks = []
        for r in range(world):
            rows = distributed.cyclic_rows(area.shape[0], world, r)
            blocks.append(kd_tree._resample(c3["swath_t"], c3["data_t"], area, 'custom', radius, neighbours=k,
                                            weight_funcs=funcs, _rows=rows).cpu().numpy())
        got = distributed.assemble_rows(blocks, area.shape[0], distributed.DEFAULT_STRIPE)
        assert np.array_equal(got, full.cpu().numpy(), equal_nan=True)


# ------------------------------------------------------------------------------------------------ C5 (full size)
def test_c5_geos_source_full_size():
    """ABI-like 5424x5424 geos source (off-disk pixels invalid) -> 10000x10000 lat/lon grid, nearest:
    fused == two-step, sharded (8 ranks) == unsharded, reported neighbours really are at the reported distance."""
    import torch

    import workloads
    from oracle import kd_tree_ref, proj_ref
    from pyresample_b200 import distributed, kd_tree
    cfg = workloads.make_config(5)
    src, area, radius = cfg["source_area"], cfg["area"], cfg["radius"]
    data_t = torch.from_numpy(cfg["data"]).cuda()
    fused = kd_tree.resample_nearest(src, data_t, area, radius, fill_value=-7).cpu().numpy()
    vii, voi, ia, da = kd_tree.get_neighbour_info(src, area, radius, neighbours=1)
    assert voi.all() and 0.6 < vii.mean() < 0.85  # ~75 % of a full-disk frame is on the Earth
    two = kd_tree.get_sample_from_neighbour_info('nn', area.shape, cfg["data"], vii, voi, ia, fill_value=-7)
    assert np.array_equal(fused, two, equal_nan=True)
    blocks = [distributed.resample_nearest_sharded(src, data_t, area, radius, fill_value=-7, rank=r, world=8)
              .cpu().numpy() for r in range(8)]
    assert np.array_equal(distributed.assemble_rows(blocks, area.shape[0], distributed.DEFAULT_STRIPE), fused,
                          equal_nan=True)
    found = ia < vii.sum()
    assert found.mean() > 0.5 and np.all(da[found] < radius) and np.all(np.isinf(da[~found]))
    # spot check against the oracle's projections: source pixel centre (geos inverse) and target pixel centre
    sel = np.flatnonzero(found)[::100003]
    src_flat = np.flatnonzero(vii)[ia[sel]]
    sx, sy = proj_ref.proj_vectors(src.area_extent, src.width, src.height, np.float32)
    srow, scol = np.unravel_index(src_flat, src.shape)
    slon, slat = proj_ref.inverse(src.proj_dict, sx[scol], sy[srow])
    tx, ty = proj_ref.proj_vectors(area.area_extent, area.width, area.height, np.float32)
    trow, tcol = np.unravel_index(sel, area.shape)
    tlon, tlat = proj_ref.inverse(area.proj_dict, tx[tcol], ty[trow])
    p = kd_tree_ref.transform_lonlats(slon.astype(np.float32), slat.astype(np.float32))
    q = kd_tree_ref.transform_lonlats(tlon.astype(np.float32), tlat.astype(np.float32))
    d = np.sqrt((((q - p) ** 2)[:, 0] + ((q - p) ** 2)[:, 1]) + ((q - p) ** 2)[:, 2])
    np.testing.assert_allclose(da[sel], d, rtol=0, atol=2.0)  # float32 ECEF: 0.5 m ulp at 6.4e6 m
