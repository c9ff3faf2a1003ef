"""The reference's own known-answer tests for the kd-tree path, written once and run against BOTH
implementations: the CPU oracle (tests/test_oracle_golden.py, no GPU) and the CUDA product
(tests/test_gpu_reference_cases.py).  Every case cites pyresample/test/<file>:<lines> and keeps its literals.

``kd`` is a module-like object with the ``pyresample.kd_tree`` API; geometry classes come from
``pyresample_b200.geometry`` (host-side containers; the oracle reads them duck-typed).
"""
import os
import warnings

import numpy as np

from helpers import stere_area
from pyresample_b200 import geometry

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _grids():
    return np.load(os.path.join(GOLDEN, "kd_tree_mask_grids.npz"))


def _mask(name):
    return np.unpackbits(_grids()[name])[:640000].reshape(800, 800).astype(bool)


def _swath_50x10():
    lons = np.fromfunction(lambda y, x: 3 + x, (50, 10))
    lats = np.fromfunction(lambda y, x: 75 - y, (50, 10))
    return geometry.SwathDefinition(lons=lons, lats=lats)


def _swath_5000x100():
    lons = np.fromfunction(lambda y, x: 3 + (10.0 / 100) * x, (5000, 100))
    lats = np.fromfunction(lambda y, x: 75 - (50.0 / 5000) * y, (5000, 100))
    return geometry.SwathDefinition(lons=lons, lats=lats)


def fwhm2sigma(fwhm):
    """pyresample/utils/__init__.py:161-174."""
    return fwhm / (2 * np.sqrt(np.log(2)))


class catch(object):
    def __enter__(self):
        self._cm = warnings.catch_warnings(record=True)
        self.w = self._cm.__enter__()
        warnings.simplefilter("always")
        return self.w

    def __exit__(self, *a):
        return self._cm.__exit__(*a)


# ------------------------------------------------------------------ test_kd_tree.py:58-114 (base cases)
def case_base_scalars(kd):
    tdata = np.array([1, 2, 3])
    tswath = geometry.SwathDefinition(lons=np.array([11.280789, 12.649354, 12.080402]),
                                      lats=np.array([56.011037, 55.629675, 55.641535]))
    tgrid = geometry.CoordinateDefinition(lons=np.array([12.562036]), lats=np.array([55.715613]))
    res = kd.resample_nearest(tswath, tdata.ravel(), tgrid, 100000, reduce_data=False, segments=1)
    assert res[0] == 2
    with catch() as w:
        res = kd.resample_gauss(tswath, tdata.ravel(), tgrid, 50000, 25000, reduce_data=False, segments=1)
        assert len(w) == 1 and 'Searching' in str(w[0].message)
    np.testing.assert_almost_equal(res[0], 2.2020729, 5)

    def wf(dist):
        return 1 - dist / 100000.0
    with catch() as w:
        res = kd.resample_custom(tswath, tdata.ravel(), tgrid, 50000, wf, reduce_data=False, segments=1)
        assert len(w) == 1 and 'Searching' in str(w[0].message)
    np.testing.assert_almost_equal(res[0], 2.4356757, 5)
    sigma = fwhm2sigma(41627.730557884883)
    with catch() as w:
        res, stddev, count = kd.resample_gauss(tswath, tdata, tgrid, 100000, sigma, with_uncert=True)
        assert any('Searching' in str(_w.message) for _w in w)
    np.testing.assert_almost_equal(res[0], 2.20206560694, 5)
    np.testing.assert_almost_equal(stddev[0], 0.707115076173, 5)
    assert count[0] == 3
    with catch() as w:
        res, stddev, counts = kd.resample_custom(tswath, tdata, tgrid, 100000, wf, with_uncert=True)
        assert any('Searching' in str(_w.message) for _w in w)
    np.testing.assert_almost_equal(res[0], 2.32193149, 5)
    np.testing.assert_almost_equal(stddev[0], 0.81817972, 5)
    assert counts[0] == 3


# ------------------------------------------------------------------ test_kd_tree.py:116-274 (nearest)
def case_nearest_cross_sums(kd):
    area_def = stere_area()
    swath_def = _swath_50x10()
    data = np.fromfunction(lambda y, x: y * x, (50, 10))
    assert kd.resample_nearest(swath_def, data.ravel(), area_def, 50000, segments=1).sum() == 15874591.0  # :116-125
    assert kd.resample_nearest(swath_def, data.ravel(), area_def, 50000, segments=2).sum() == 15874591.0  # :216-225
    assert kd.resample_nearest(swath_def, data.ravel(), area_def, 50000, nprocs=2, segments=1).sum() == 15874591.0
    cdata = np.fromfunction(lambda y, x: y + complex("j") * x, (50, 10), dtype=np.complex128)
    res = kd.resample_nearest(swath_def, cdata.ravel(), area_def, 50000, segments=1)  # :127-137
    assert np.issubdtype(res.dtype, np.complex128)
    assert res.sum().real == 3530219.0 and res.sum().imag == 688723.0
    data_multi = np.column_stack((data.ravel(), data.ravel(), data.ravel()))
    assert kd.resample_nearest(swath_def, data_multi, area_def, 50000, segments=1).sum() == 3 * 15874591.0  # :251-261
    res = kd.resample_nearest(swath_def, np.dstack((data, data, data)), area_def, 50000, segments=1)  # :263-274
    assert res.sum() == 3 * 15874591.0 and res.shape == (800, 800, 3)
    res = kd.resample_nearest(swath_def, data.ravel(), area_def, 50000, segments=1)
    remap = kd.resample_nearest(area_def, res.ravel(), swath_def, 5000, segments=1)  # :227-238
    assert remap.sum() == 22275.0


def case_nearest_masked_swath_target(kd):
    # :139-155
    data = np.fromfunction(lambda y, x: y * x, (50, 10))
    lons = np.fromfunction(lambda y, x: 3 + x, (50, 10))
    lats = np.fromfunction(lambda y, x: 75 - y, (50, 10))
    mask = np.ones_like(lons, dtype=np.bool_)
    mask[::2, ::2] = False
    swath_def = geometry.SwathDefinition(lons=np.ma.masked_array(lons, mask=mask),
                                         lats=np.ma.masked_array(lats, mask=False))
    res = kd.resample_nearest(swath_def, data.ravel(), swath_def, 50000, segments=3)
    assert res.sum() == 12000


def case_nearest_1d_and_empty(kd):
    area_def = stere_area()
    data800 = np.fromfunction(lambda x, y: x * y, (800, 800))
    sw1d = geometry.SwathDefinition(lons=np.fromfunction(lambda x: 3 + x / 100., (500,)),
                                    lats=np.fromfunction(lambda x: 75 - x / 10., (500,)))
    res = kd.resample_nearest(area_def, data800.ravel(), sw1d, 50000, segments=1)  # :157-167
    assert res.shape == (500,) and res.sum() == 35821299.0
    data = np.fromfunction(lambda y, x: y * x, (50, 10))
    lons = np.fromfunction(lambda y, x: 165 + x, (50, 10))
    lats = np.fromfunction(lambda y, x: 75 - y, (50, 10))
    swath_def = geometry.SwathDefinition(lons=lons, lats=lats)
    assert kd.resample_nearest(swath_def, data.ravel(), area_def, 50000, segments=1).sum() == 0  # :169-178
    data_multi = np.column_stack((data.ravel(), data.ravel(), data.ravel()))
    assert kd.resample_nearest(swath_def, data_multi, area_def, 50000, segments=1).shape == (800, 800, 3)  # :180-190
    res = kd.resample_nearest(swath_def, data_multi, area_def, 50000, segments=1, fill_value=None)  # :192-202
    assert res.shape == (800, 800, 3)
    res = kd.resample_nearest(swath_def, data.ravel(), area_def, 50000, segments=1, fill_value=None)  # :204-214
    assert res.mask.sum() == res.size


# ------------------------------------------------------------------ test_kd_tree.py:276-486 (gauss / custom)
def case_gauss_and_custom(kd):
    area_def = stere_area()
    swath_def = _swath_50x10()
    data = np.fromfunction(lambda y, x: y * x, (50, 10))
    res = kd.resample_gauss(swath_def, data.ravel(), area_def, 50000, 25000, fill_value=-1, segments=1)  # :276-285
    np.testing.assert_almost_equal(res.sum(), 15387753.9852, 3)
    swath_def = _swath_5000x100()
    data = np.fromfunction(lambda y, x: (y + x) * 10 ** -5, (5000, 100))
    with catch() as w:
        res = kd.resample_gauss(swath_def, data.ravel(), area_def, 50000, 25000, segments=1)  # :287-317
        assert len(w) == 1 and 'Possible more' in str(w[0].message)
    np.testing.assert_almost_equal(res.sum(), 4872.8100353517921)
    data6 = np.fromfunction(lambda y, x: (y + x) * 10 ** -6, (5000, 100))
    data_multi = np.column_stack((data6.ravel(), data6.ravel(), data6.ravel()))
    with catch() as w:
        res = kd.resample_gauss(swath_def, data_multi, area_def, 50000, [25000, 15000, 10000], segments=1)  # :319-335
        assert len(w) == 1 and 'Possible more' in str(w[0].message)
    np.testing.assert_almost_equal(res.sum(), 1461.8429990248171)
    with catch() as w:
        res, stddev, counts = kd.resample_gauss(swath_def, data_multi, area_def, 50000, [25000, 15000, 10000],
                                                segments=1, with_uncert=True)  # :337-370
        assert len(w) >= 1 and any('Possible more' in str(x.message) for x in w)
    np.testing.assert_almost_equal(res.sum(), 1461.8429990248171)
    for i, e in enumerate([0.44621800779801657, 0.44363137712896705, 0.43861019464274459]):
        np.testing.assert_almost_equal(stddev[:, :, i].sum(), e)
    np.testing.assert_almost_equal(counts.sum(), 4934802.0)

    def wf1(dist):
        return 1 - dist / 100000.0

    def wf2(dist):
        return 1

    def wf3(dist):
        return np.cos(dist) ** 2
    with catch() as w:
        res = kd.resample_custom(swath_def, data.ravel(), area_def, 50000, wf1, segments=1)  # :428-459
        assert len(w) == 1 and 'Possible more' in str(w[0].message)
    np.testing.assert_almost_equal(res.sum(), 4872.8100347930776)
    with catch() as w:
        res = kd.resample_custom(swath_def, data_multi, area_def, 50000, [wf1, wf2, wf3], segments=1)  # :461-486
        assert len(w) == 1 and 'Possible more' in str(w[0].message)
    np.testing.assert_almost_equal(res.sum(), 1461.8428378742638)
    # :663-709 - from precomputed neighbour info, twice (input must not be mutated)
    with catch() as w:
        vii, voi, ia, da = kd.get_neighbour_info(swath_def, area_def, 50000, segments=1)
        assert len(w) == 1 and 'Possible more' in str(w[0].message)
    for _ in range(2):
        res = kd.get_sample_from_neighbour_info('custom', (800, 800), data_multi, vii, voi, ia, da,
                                                weight_funcs=[wf1, wf2, wf3])
        np.testing.assert_almost_equal(res.sum(), 1461.8428378742638)


# ------------------------------------------------------------------ test_kd_tree.py:488-661 (masks, dtypes)
def case_masked_and_fill(kd):
    area_def = stere_area()
    swath_def = _swath_50x10()
    data = np.ones((50, 10))
    data[:, 5:] = 2
    mask = np.ones((50, 10))
    mask[:, :5] = 0
    masked_data = np.ma.array(data, mask=mask)
    g = _grids()
    res = kd.resample_nearest(swath_def, masked_data.ravel(), area_def, 50000, segments=1)  # :488-504
    assert np.array_equal(_mask("mask_test_nearest_mask"), res.mask)
    assert np.array_equal(g["mask_test_nearest_data"], res.data)
    res = kd.resample_gauss(swath_def, masked_data.ravel(), area_def, 50000, 25000, segments=1)  # :519-538
    assert np.array_equal(_mask("mask_test_mask"), res.mask)
    np.testing.assert_almost_equal(res.data.sum(), g["mask_test_data"].sum(), 3)
    ones = np.ones((50, 10))
    res = kd.resample_nearest(swath_def, ones.ravel(), area_def, 50000, fill_value=None, segments=1)  # :540-550
    assert np.array_equal(res.mask, _mask("mask_test_fill_value"))
    res = kd.resample_nearest(swath_def, ones.astype('int').ravel(), area_def, 50000, fill_value=None, segments=1)
    assert np.array_equal(res.mask, _mask("mask_test_fill_value"))  # :552-562
    res = kd.resample_nearest(swath_def, masked_data.ravel(), area_def, 50000, fill_value=None, segments=1)  # :564-581
    assert np.array_equal(res.mask, _mask("mask_test_full_fill"))
    # :583-609 (the golden mask file is missing from the reference mount; the cross-sum literal remains)
    mask1 = np.ones((50, 10))
    mask1[:, :5] = 0
    mask2 = np.ones((50, 10))
    mask2[:, 5:] = 0
    mask3 = np.ones((50, 10))
    mask3[:25, :] = 0
    data_multi = np.column_stack((data.ravel(), data.ravel(), data.ravel()))
    mask_multi = np.column_stack((mask1.ravel(), mask2.ravel(), mask3.ravel()))
    masked_multi = np.ma.array(data_multi, mask=mask_multi)
    res = kd.resample_nearest(swath_def, masked_multi, area_def, 50000, fill_value=None, segments=1)
    np.testing.assert_almost_equal(res.sum(), 357140.0)
    # :506-517 area -> 1-D swath with masked data
    data800 = np.ones((800, 800))
    data800[:400, :] = 2
    mask800 = np.ones((800, 800))
    mask800[400:, :] = 0
    sw1d = geometry.SwathDefinition(lons=np.fromfunction(lambda x: 3 + x / 100., (500,)),
                                    lats=np.fromfunction(lambda x: 75 - x / 10., (500,)))
    res = kd.resample_nearest(area_def, np.ma.array(data800, mask=mask800).ravel(), sw1d, 50000, segments=1)
    assert res.mask.sum() == 112


def case_from_sample_dtypes(kd):
    area_def = stere_area()
    swath_def = _swath_50x10()
    # :611-621 float32 swath against a float64 grid definition
    grid_def = geometry.GridDefinition(np.fromfunction(lambda y, x: 3 + x, (50, 10)),
                                       np.fromfunction(lambda y, x: 75 - y, (50, 10)))
    swath32 = geometry.SwathDefinition(lons=np.fromfunction(lambda y, x: 3 + x, (50, 10)).astype('f4'),
                                       lats=np.fromfunction(lambda y, x: 75 - y, (50, 10)).astype('f4'))
    vii, voi, ia, da = kd.get_neighbour_info(swath32, grid_def, 50000, neighbours=1, segments=1)
    assert da.dtype == np.float32
    vii, voi, ia, da = kd.get_neighbour_info(swath_def, area_def, 50000, neighbours=1, segments=1)
    data = np.fromfunction(lambda y, x: y * x, (50, 10))
    res = kd.get_sample_from_neighbour_info('nn', (800, 800), data.ravel(), vii, voi, ia)  # :623-637
    assert res.sum() == 15874591.0
    for dtype in [np.uint16, np.float32]:  # :639-661
        d = data.astype(dtype)
        res = kd.get_sample_from_neighbour_info('nn', (800, 800), d.ravel(), vii, voi, ia, fill_value=dtype(0.0))
        assert res.dtype == dtype and res.sum() == 15874591.0


# ------------------------------------------------------------------ test_swath.py:48-82 (real SSMIS swath)
def case_ssmis_self_map(kd):
    sw = np.load(os.path.join(GOLDEN, "ssmis_swath.npz"))["data"]
    lons, lats, tb37v = sw[:, 0].astype(np.float64), sw[:, 1].astype(np.float64), sw[:, 2].astype(np.float64)
    fvalue = -10000000000.0
    ok = (lons != fvalue) * (lats != fvalue) * (tb37v != fvalue)
    lons, lats, tb37v = lons[ok], lats[ok], tb37v[ok]
    swath_def = geometry.SwathDefinition(lons=lons, lats=lats)
    with catch() as w:
        res = kd.resample_gauss(swath_def, tb37v.copy(), swath_def, radius_of_influence=70000, sigmas=56500)
        assert len(w) == 1 and 'Possible more' in str(w[0].message)
    np.testing.assert_almost_equal(res.sum() / 100., 668848.0, 0)
    data = np.column_stack((tb37v, tb37v, tb37v))
    with catch() as w:
        res = kd.resample_gauss(swath_def, data, swath_def, radius_of_influence=70000, sigmas=[56500, 56500, 56500])
        assert len(w) == 1
    for c in range(3):
        np.testing.assert_almost_equal(res[:, c].sum() / 100., 668848.0, 0)


# ------------------------------------------------------------------ error behaviour (kd_tree.py:167-178, 241-247, 503-510, 612-690)
def case_errors(kd):
    import pytest
    swath_def = _swath_50x10()
    area_def = stere_area()
    data = np.ones((50, 10))
    with pytest.raises(TypeError):
        kd.resample_gauss(swath_def, data.ravel(), area_def, 50000, "sigma")
    with pytest.raises(TypeError):
        kd.resample_custom(swath_def, data.ravel(), area_def, 50000, "not a function")
    with pytest.raises(TypeError):
        kd.resample_nearest(swath_def, data.ravel(), area_def, "far", reduce_data=False)
    with pytest.raises(TypeError):
        kd.get_neighbour_info(swath_def, area_def, 50000, neighbours=2.5)
    with pytest.raises(ValueError):
        kd.resample_nearest(swath_def, np.ones(7), area_def, 50000)
    vii, voi, ia, da = kd.get_neighbour_info(swath_def, area_def, 50000, neighbours=1, segments=1)
    with pytest.raises(TypeError):
        kd.get_sample_from_neighbour_info('bilinear', (800, 800), data.ravel(), vii, voi, ia)
    with pytest.raises(ValueError):
        kd.get_sample_from_neighbour_info('custom', (800, 800), data.ravel(), vii, voi, ia)
    with catch():
        vii, voi, ia2, da2 = kd.get_neighbour_info(swath_def, area_def, 50000, neighbours=2, segments=1)
    with pytest.raises(ValueError):
        kd.get_sample_from_neighbour_info('nn', (800, 800), data.ravel(), vii, voi, ia2)


ALL_CASES = [case_base_scalars, case_nearest_cross_sums, case_nearest_masked_swath_target, case_nearest_1d_and_empty,
             case_gauss_and_custom, case_masked_and_fill, case_from_sample_dtypes, case_ssmis_self_map, case_errors]


# --------------------------------------------------------------- KDTreeNearestXarrayResampler
def nearest_resampler_fixtures(geometry):
    """The geometries and arrays of pyresample/test/conftest.py:31-34,118-135,204-340 as plain numpy objects."""
    def test_longitude(start, stop, shape, dtype=np.float32):
        if start > 0 > stop:
            stop += 360.0
        row = np.linspace(start, stop, num=shape[1]).astype(dtype)
        arr = np.repeat([row], shape[0], axis=0)
        if stop > 360.0:
            arr[arr > 360.0] -= 360
        return arr

    def test_latitude(start, stop, shape, dtype=np.float32):
        col = np.linspace(start, stop, num=shape[0]).astype(dtype).reshape((shape[0], 1))
        return np.repeat(col, shape[1], axis=1)

    shape = (50, 10)
    fx = {}
    fx["euro_lonlats"] = (test_longitude(3.0, 12.0, shape), test_latitude(75.0, 26.0, shape))
    alons = test_longitude(172.0, 190.0, shape)
    alons[alons > 180.0] -= 360.0
    fx["antimeridian_lonlats"] = (alons, test_latitude(25.0, 33.0, shape))
    stere = dict(a='6378144.0', b='6356759.0', proj='stere')
    extent = (-1370912.72, -909968.64000000001, 1029087.28, 1490031.3600000001)
    fx["area_stere_source"] = geometry.AreaDefinition("s", "s", "s", dict(stere, lat_0='52.00', lat_ts='52.00', lon_0='5.00'),
                                                      10, 50, extent)
    fx["area_stere_target"] = geometry.AreaDefinition("t", "t", "t", dict(stere, lat_0='50.00', lat_ts='50.00', lon_0='8.00'),
                                                      85, 80, extent)
    fx["area_lonlat_pm180_target"] = geometry.AreaDefinition(
        "p", "p", "p", {'proj': 'longlat', 'pm': '180.0', 'datum': 'WGS84', 'no_defs': None}, 85, 80,
        (-20.0, 20.0, 20.0, 35.0))
    fx["coord_def_2d_float32"] = (
        np.array([[11.5, 12.562036, 12.9]] * 4, dtype=np.float32),
        np.array([[55.715613, 55.715613, 55.715613], [55.715613, 55.715613, 55.715613],
                  [55.715613, np.nan, 55.715613], [55.715613, 55.715613, 55.715613]], dtype=np.float32))
    fx["swath_1d"] = (np.array([11.280789, 12.649354, 12.080402]), np.array([56.011037, 55.629675, 55.641535]))
    fx["data_1d"] = np.array([1., 2., 3.], dtype=np.float32)
    fx["data_2d"] = np.fromfunction(lambda y, x: y * x, shape, dtype=np.float32)
    fx["data_3d"] = np.fromfunction(lambda y, x, b: y * x * b, (50, 10, 3), dtype=np.float32)
    return fx


NEAREST_RESAMPLER_SUMS = {  # test_resamplers/test_nearest.py:96-228
    "swath_2d_to_area": 167913.0, "area_to_area": 303048.0, "swath_pm180": 115591.0, "area_no_roi": 952386.0,
    "area_no_roi_no_geocentric": 20666.0, "area_3d": 909144.0}
NEAREST_1D_EXPECTED = np.array([[1., 2., 2.], [1., 2., 2.], [1., np.nan, 2.], [1., 2., 2.]])  # :67-72
NEAREST_1D_INT_EXPECTED = np.array([[1, 2, 2], [1, 2, 2], [1, 255, 2], [1, 2, 2]])           # :88-93


# --------------------------------------------------------------- image containers / linesamples
def image_fixtures(geometry):
    """pyresample/test/test_image.py:32-75."""
    area_def = geometry.AreaDefinition('areaD', 'Europe (3km, HRV, VTC)', 'areaD',
                                       {'a': '6378144.0', 'b': '6356759.0', 'lat_0': '50.00', 'lat_ts': '50.00',
                                        'lon_0': '8.00', 'proj': 'stere'}, 800, 800,
                                       [-1370912.72, -909968.64000000001, 1029087.28, 1490031.3600000001])

    def msg(n):
        return geometry.AreaDefinition('msg_full', 'Full globe MSG image 0 degrees', 'msg_full',
                                       {'a': '6378169.0', 'b': '6356584.0', 'h': '35785831.0', 'lon_0': '0',
                                        'proj': 'geos'}, n, n,
                                       [-5568742.4000000004, -5568742.4000000004, 5568742.4000000004,
                                        5568742.4000000004])
    return area_def, msg(3712), msg(928)


IMAGE_NEAREST_SUM = 399936.70287099993        # test_image.py:130-138
IMAGE_NEAREST_RESIZE_SUM = 2212023.0175830    # test_image.py:140-148


# --------------------------------------------------------------- XArrayResamplerNN literals (test_kd_tree.py:743-995)
def xarray_nn_fixtures(geometry):
    """The class fixtures of TestXArrayResamplerNN: note src_area_2d is 50 wide x 10 high while the data are (50, 10) -
    the reference only ravels, so the sizes match (500) and the literals below depend on exactly that."""
    stere = {'a': '6378144.0', 'b': '6356759.0', 'proj': 'stere'}
    extent = [-1370912.72, -909968.64000000001, 1029087.28, 1490031.3600000001]
    area_def = geometry.AreaDefinition('areaD', 'Europe (3km, HRV, VTC)', 'areaD',
                                       dict(stere, lat_0='50.00', lat_ts='50.00', lon_0='8.00'), 800, 800, extent)
    src_area_2d = geometry.AreaDefinition('areaD_src', 'Europe (3km, HRV, VTC)', 'areaD',
                                          dict(stere, lat_0='52.00', lat_ts='52.00', lon_0='5.00'), 50, 10, extent)
    data_2d = np.fromfunction(lambda y, x: y * x, (50, 10))
    data_3d = np.fromfunction(lambda y, x, b: y * x * b, (50, 10, 3))
    return area_def, src_area_2d, data_2d, data_3d


XARRAY_NN_SUMS = {"area_to_area": 27706753.0, "area_no_roi": 87281406.0, "area_no_roi_no_geocentric": 1855928.0,
                  "area_3d": 83120259.0}
