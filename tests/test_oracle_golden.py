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


# ------------------------------------------------------------------------------------------------ bilinear (f2)
def _bilinear_fixtures():
    """setUpClass of pyresample/test/test_bilinear.py:42-103."""
    from pyresample_b200 import geometry
    target = geometry.AreaDefinition('areaD', 'Europe (3km, HRV, VTC)', 'areaD',
                                     {'a': '6378144.0', 'b': '6356759.0', 'lat_0': '50.00', 'lat_ts': '50.00',
                                      'lon_0': '8.00', 'proj': 'stere'}, 4, 4,
                                     [-1370912.72, -909968.64000000001, 1029087.28, 1490031.3600000001])
    lons, lats = np.meshgrid(np.linspace(-25., 40., num=100), np.linspace(45., 75., num=100))
    return target, geometry.SwathDefinition(lons=lons, lats=lats), geometry.SwathDefinition(lons=np.ravel(lons),
                                                                                          lats=np.ravel(lats))


def test_bilinear_fractional_distance_literals():
    """test_bilinear.py:152-205 (irregular / parallel cases) and :207-219 (division by zero)."""
    from oracle import bilinear_ref as br
    irregular = (np.array([[-1., 1.]]), np.array([[1., 2.]]), np.array([[-2., -1.]]), np.array([[2., -4.]]))
    vert_parallel = (np.array([[-1., 1.]]), np.array([[1., 2.]]), np.array([[-1., -1.]]), np.array([[1., -2.]]))
    both_parallel = (np.array([[-1., 1.]]), np.array([[1., 1.]]), np.array([[-1., -1.]]), np.array([[1., -1.]]))
    res = br.distances_irregular(irregular, 0., 0.)
    assert res[0] == 0.375 and res[1] == 0.5
    res = br.distances_irregular(vert_parallel, 0., 0.)
    assert res[0] == 0.5 and np.isnan(res[1])
    assert tuple(br.distances_uprights_parallel(vert_parallel, 0., 0.)) == (0.5, 0.5)
    assert tuple(br.distances_parallelogram(both_parallel[:3], 0., 0.)) == (0.5, 0.5)
    out = np.array([[0.]])
    for pts, want in ((irregular, (0.375, 0.5)), (both_parallel, (0.5, 0.5)), (vert_parallel, (0.5, 0.5))):
        t, s = br.fractional_distances(pts, out, out)
        assert (t, s) == want
    corner_points = [np.array([[-64.9936752319336, -5.140199184417725]]), np.array([[-64.98487091064453, -5.142156600952148]]),
                     np.array([[-64.98683166503906, -5.151054859161377]]), np.array([[-64.97802734375, -5.153012275695801]])]
    t, s = br.fractional_distances(corner_points, np.array([-64.985]), np.array([-5.145]))
    np.testing.assert_allclose(t, [0.30769689])
    np.testing.assert_allclose(s, [0.74616628])
    assert br.solve_quadratic(1, 0, 0) == 0.0 and np.isnan(br.solve_quadratic(1, 2, 1))
    assert br.solve_quadratic(1, 2, 1, min_val=-2.) == -1.0
    nan_pts = (np.array([[np.nan, np.nan]]),) + irregular[1:]
    assert all(np.isnan(v) for v in br.calc_abc(nan_pts, 0.0, 0.0))


def test_bilinear_info_and_sampling_literals():
    """test_bilinear.py:240-383: corners, t / s of pixel 5, NaN pixels, sample sums, masked output."""
    import warnings
    from oracle import bilinear_ref as br
    target, source, source_1d = _bilinear_fixtures()
    info = br.bil_info(source, target, 50e3, neighbours=32, reduce_data=True)
    pts = info["corner_points"]
    assert all(p.shape == (16, 2) for p in pts) and info["index_array"].shape == (16, 4)
    assert np.sum(~np.isnan(np.sum(pts[0] + pts[1] + pts[2] + pts[3], axis=1))) == 10  # :240-262
    for reduce_data in (False, True):  # :264-296
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            t, s, _, _ = br.get_bil_info(source, target, 50e5, neighbours=32, reduce_data=reduce_data)
        assert abs(t[5] - 0.730659147133) < 1e-5 and abs(s[5] - 0.310314173004) < 1e-5
        assert np.isnan(t[12:]).all() and np.isnan(s[12:]).all()
        assert np.all((t[:12] >= 0) & (t[:12] <= 1) & (s[:12] >= 0) & (s[:12] <= 1))
    data1 = np.ones((100, 100))
    t, s, vii, ia = br.get_bil_info(source, target, 50e5, neighbours=32)
    assert abs(br.get_sample_from_bil_info(data1.ravel(), t, s, vii, ia).ravel()[5] - 1.) < 1e-7  # :298-341
    assert abs(br.get_sample_from_bil_info(2 * data1.ravel(), t, s, vii, ia).ravel()[5] - 2.) < 1e-7
    res = br.get_sample_from_bil_info((data1 + 9.5).ravel(), t, s, vii, ia, output_shape=target.shape)
    assert res.shape == target.shape and np.isnan(res).sum() == 4
    t1, s1, vii1, ia1 = br.get_bil_info(source_1d, target, 50e5, neighbours=32)  # :343-357
    res = br.get_sample_from_bil_info((data1 + 9.5).ravel(), t1, s1, vii1, ia1)
    assert abs(np.nanmin(res) - 10.5) < 1e-7 and abs(np.nanmax(res) - 10.5) < 1e-7 and np.isnan(res).sum() == 4
    res = br.resample_bilinear(data1, source, target, 50e5, neighbours=32)  # :359-383
    assert res.shape == target.shape and res.sum() == 12 and (res == 0).sum() == 4
    res = br.resample_bilinear(data1, source, target, 50e5, neighbours=32, fill_value=None)
    assert hasattr(res, 'mask') and target.size - res.mask.sum() == 12
    res = br.resample_bilinear(np.dstack((data1, 2 * data1)), source, target)
    assert res.shape == target.shape + (2,)
