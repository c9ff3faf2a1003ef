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
