"""Pin the CPU oracle against the reference's own known-answer tests, golden grids and projection literals.

Runs without a GPU.  The same cases run through the CUDA product in tests/test_gpu_reference_cases.py.
"""
import numpy as np
import pytest

import reference_cases as rc
from oracle import kd_tree_ref, proj_ref
from oracle.knn_ref import KDTreeRef, brute_query


@pytest.mark.parametrize("case", rc.ALL_CASES, ids=lambda f: f.__name__)
def test_reference_case_through_oracle(case):
    case(kd_tree_ref)


def test_proj_stere_oblique_golden():
    # pyresample/test/test_geometry/test_area.py:293-306
    proj = {'a': '6378144.0', 'b': '6356759.0', 'lat_0': '50.00', 'lat_ts': '50.00', 'lon_0': '8.00', 'proj': 'stere'}
    ext = [-1370912.72, -909968.64000000001, 1029087.28, 1490031.3600000001]
    lon, lat = proj_ref.area_lonlats(proj, ext, 800, 800)
    np.testing.assert_allclose(lon[400, 400], 5.5028467120975835, rtol=1e-14)
    np.testing.assert_allclose(lat[400, 400], 52.566998432390619, rtol=1e-14)
    cart = kd_tree_ref.transform_lonlats(lon.ravel(), lat.ravel())
    exp = 5872039989466.8457031
    assert abs(cart.sum() - exp) < 1e-7 * exp


def test_proj_laea_polar_and_latlong_golden():
    # test_area.py:748-761 (ease_sh corners) and :589-600 (latlong)
    lon, lat = proj_ref.area_lonlats('+proj=laea +lat_0=-90 +lon_0=0 +a=6371228.0 +units=m',
                                     (-5326849.0625, -5326849.0625, 5326849.0625, 5326849.0625), 425, 425)
    got = [(lon[0, 0], lat[0, 0]), (lon[0, -1], lat[0, -1]), (lon[-1, -1], lat[-1, -1]), (lon[-1, 0], lat[-1, 0])]
    exp = [(-45.0, -17.713517415148853), (45.000000000000014, -17.71351741514884),
           (135.0, -17.713517415148825), (-135.00000000000003, -17.71351741514884)]
    np.testing.assert_allclose(got, exp)
    lon, lat = proj_ref.area_lonlats({'proj': 'latlong'}, (-180, -90, 180, 90), 360, 180)
    assert lon[0, 0] == -179.5 and lat[0, 0] == 89.5


def test_cartesian_golden():
    # pyresample/test/test_spatial_mp.py:31-66
    got = kd_tree_ref.transform_lonlats(np.array([0., 10., 25.]), np.array([0., 10., 25.]))
    want = np.array([[6370997., 0., 0.], [6178887.9339746, 1089504.6535337, 1106312.0189715],
                     [5233097.4664751, 2440233.4244856, 2692499.6776952]])
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-3)
    assert kd_tree_ref.transform_lonlats(np.array([1, 2]), np.array([3, 4])).dtype == np.float64
    assert kd_tree_ref.transform_lonlats(np.float32([1, 2]), np.float32([3, 4])).dtype == np.float32


def test_reduce_boundary_golden():
    # pyresample/test/test_data_reduce.py:70-82 (test_reduce_boundary): cross sum 20685125.0
    from pyresample_b200 import geometry
    area = geometry.AreaDefinition('areaD', 'Europe (3km, HRV, VTC)', 'areaD',
                                   {'a': '6378144.0', 'b': '6356759.0', 'lat_0': '50.00', 'lat_ts': '50.00',
                                    'lon_0': '8.00', 'proj': 'stere'}, 800, 800,
                                   [-1370912.72, -909968.64000000001, 1029087.28, 1490031.3600000001])
    data = np.fromfunction(lambda y, x: (y + x), (1000, 1000))
    lons = np.fromfunction(lambda y, x: -180 + (360.0 / 1000) * x, (1000, 1000))
    lats = np.fromfunction(lambda y, x: -90 + (180.0 / 1000) * y, (1000, 1000))
    blons, blats = kd_tree_ref.get_boundary_lonlats(area)
    valid = kd_tree_ref.get_valid_index_from_lonlat_boundaries(blons, blats, lons.ravel(), lats.ravel(), 7000)
    assert data.ravel()[valid].sum() == 20685125.0


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("k", [1, 5])
def test_kdtree_ref_against_c_brute_force(dtype, k):
    """The cKDTree-backed oracle and the exhaustive C checker agree bit for bit (indices and distances)."""
    rng = np.random.default_rng(12)
    src = kd_tree_ref.transform_lonlats(rng.uniform(0, 20, 6000).astype(dtype), rng.uniform(40, 60, 6000).astype(dtype))
    src[100:140] = src[:40]  # exact duplicates: ties must go to the lower index
    qry = kd_tree_ref.transform_lonlats(rng.uniform(-1, 21, 3000).astype(dtype), rng.uniform(39, 61, 3000).astype(dtype))
    mask = rng.random(6000) < 0.1
    for m in (None, mask):
        d1, i1 = KDTreeRef(src).query(qry, k=k, distance_upper_bound=60000.0, mask=m)
        d2, i2 = brute_query(src, qry, k, 60000.0, mask=m)
        assert i1.dtype == np.uint32 and d1.dtype == dtype
        assert np.array_equal(i1, i2)
        assert np.array_equal(d1, d2)
    # pykdtree call contract (future/resamplers test_nearest.py:311-373): missing -> n / inf, mask skips points
    d, i = KDTreeRef(src).query(qry[:5] * 2, k=1, distance_upper_bound=10.0)
    assert (i == 6000).all() and np.isinf(d).all()


def test_nearest_resampler_restatement_matches_reference_literals():
    """oracle/nearest_ref.py against test_resamplers/test_nearest.py:55-228 (KDTreeNearestXarrayResampler)."""
    from oracle import nearest_ref
    from pyresample_b200 import geometry
    import reference_cases as rc
    fx = rc.nearest_resampler_fixtures(geometry)
    swath_1d = geometry.SwathDefinition(*fx["swath_1d"])
    grid = geometry.CoordinateDefinition(*fx["coord_def_2d_float32"])
    res = nearest_ref.resample(swath_1d, grid, fx["data_1d"], (0,), mask=np.isnan(fx["data_1d"]),
                               radius_of_influence=100000)
    np.testing.assert_allclose(res, rc.NEAREST_1D_EXPECTED)
    res = nearest_ref.resample(swath_1d, grid, np.array([1, 2, 3]), (0,), fill_value=255, radius_of_influence=100000)
    np.testing.assert_equal(res, rc.NEAREST_1D_INT_EXPECTED)
    euro = geometry.SwathDefinition(*fx["euro_lonlats"])
    anti = geometry.SwathDefinition(*fx["antimeridian_lonlats"])
    src, tgt, pm = fx["area_stere_source"], fx["area_stere_target"], fx["area_lonlat_pm180_target"]
    d2, d3 = fx["data_2d"], fx["data_3d"]
    S = rc.NEAREST_RESAMPLER_SUMS
    mask = nearest_ref.compute_data_mask(d2, (0, 1))
    assert float(np.nansum(nearest_ref.resample(euro, tgt, d2, (0, 1), mask=mask, radius_of_influence=50000))) \
        == S["swath_2d_to_area"]
    assert float(np.nansum(nearest_ref.resample(src, tgt, d2, (0, 1), radius_of_influence=50000))) == S["area_to_area"]
    assert float(np.nansum(nearest_ref.resample(anti, pm, d2, (0, 1), mask=mask, radius_of_influence=50000))) \
        == S["swath_pm180"]
    res = nearest_ref.resample(src, tgt, d2, (0, 1))
    assert res.shape == (80, 85) and float(np.nansum(res)) == S["area_no_roi"]
    radius = nearest_ref.compute_radius_of_influence(src, tgt, fail_source=True, fail_target=True)
    assert radius == 10000
    assert float(np.nansum(nearest_ref.resample(src, tgt, d2, (0, 1), radius_of_influence=radius))) \
        == S["area_no_roi_no_geocentric"]
    res3 = nearest_ref.resample(src, tgt, d3, (0, 1), radius_of_influence=50000)
    assert res3.shape == (80, 85, 3) and float(np.nansum(res3)) == S["area_3d"]


def test_image_container_literals_pin_the_geos_source_path():
    """test_image.py:130-148: full-disk geos 3712x3712 source (off-earth pixels are inf) -> stere / geos targets."""
    from pyresample_b200 import geometry
    import reference_cases as rc
    area_def, msg_area, msg_resize = rc.image_fixtures(geometry)
    data = np.fromfunction(lambda y, x: y * x * 10 ** -6, (3712, 3712))
    res = kd_tree_ref.resample_nearest(msg_area, data.ravel(), area_def, 50000, segments=1)
    assert abs(res.sum() - rc.IMAGE_NEAREST_SUM) < 5e-8
    res = kd_tree_ref.resample_nearest(msg_area, data.ravel(), msg_resize, 50000, segments=1)
    assert abs(res.sum() - rc.IMAGE_NEAREST_RESIZE_SUM) < 5e-8


def test_xarray_resampler_nn_literals():
    """test_kd_tree.py:897-995 (XArrayResamplerNN, area -> area) through the oracle's restatement."""
    from oracle import nearest_ref
    from pyresample_b200 import geometry
    import reference_cases as rc
    area_def, src, d2, d3 = rc.xarray_nn_fixtures(geometry)
    S = rc.XARRAY_NN_SUMS
    flat2 = d2.reshape(src.shape)            # the reference ravels the (50, 10) data against the (10, 50) area
    flat3 = d3.reshape(src.shape + (3,))
    assert np.nansum(nearest_ref.resample(src, area_def, flat2, (0, 1), radius_of_influence=50000)) == S["area_to_area"]
    assert np.nansum(nearest_ref.resample(src, area_def, flat2, (0, 1))) == S["area_no_roi"]
    assert np.nansum(nearest_ref.resample(src, area_def, flat2, (0, 1), radius_of_influence=10000)) \
        == S["area_no_roi_no_geocentric"]
    assert np.nansum(nearest_ref.resample(src, area_def, flat3, (0, 1), radius_of_influence=50000)) == S["area_3d"]
