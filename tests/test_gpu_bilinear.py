"""Bilinear resampling on the GPU against the oracle's restatement of pyresample/bilinear/_base.py and
the reference's own literals (pyresample/test/test_bilinear.py:139-383)."""
import warnings

import numpy as np
import pytest

from helpers import jittered_swath, laea_area, stere_area
from oracle import bilinear_ref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bil(cuda_lib):
    from pyresample_b200 import bilinear
    return bilinear


def _fixtures():
    from pyresample_b200 import geometry
    target = geometry.AreaDefinition('areaD', 'Europe (3km, HRV, VTC)', 'areaD',
                                     {'a': '6378144.0', 'b': '6356759.0', 'lat_0': '50.00', 'lat_ts': '50.00',
                                      'lon_0': '8.00', 'proj': 'stere'}, 4, 4,
                                     [-1370912.72, -909968.64000000001, 1029087.28, 1490031.3600000001])
    lons, lats = np.meshgrid(np.linspace(-25., 40., num=100), np.linspace(45., 75., num=100))
    return target, geometry.SwathDefinition(lons=lons, lats=lats), geometry.SwathDefinition(lons=np.ravel(lons),
                                                                                          lats=np.ravel(lats))


def test_reference_literals_through_the_product(bil):
    target, source, source_1d = _fixtures()
    data1 = np.ones((100, 100))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for reduce_data in (False, True):  # test_bilinear.py:264-296
            t, s, _, _ = bil.get_bil_info(source, target, 50e5, neighbours=32, reduce_data=reduce_data)
            assert abs(t[5] - 0.730659147133) < 1e-5 and abs(s[5] - 0.310314173004) < 1e-5
            assert np.isnan(t[12:]).all() and np.isnan(s[12:]).all()
            assert np.all((t[:12] >= 0) & (t[:12] <= 1) & (s[:12] >= 0) & (s[:12] <= 1))
        t, s, vii, ia = bil.get_bil_info(source, target, 50e5, neighbours=32)  # :298-341
        assert abs(bil.get_sample_from_bil_info(data1.ravel(), t, s, vii, ia).ravel()[5] - 1.) < 1e-7
        assert abs(bil.get_sample_from_bil_info(2 * data1.ravel(), t, s, vii, ia).ravel()[5] - 2.) < 1e-7
        res = bil.get_sample_from_bil_info((data1 + 9.5).ravel(), t, s, vii, ia, output_shape=target.shape)
        assert res.shape == target.shape and np.isnan(res).sum() == 4
        res = bil.get_sample_from_bil_info(np.ma.masked_all(data1.shape).ravel(), t, s, vii, ia)
        assert not hasattr(res, 'mask')
        t1, s1, vii1, ia1 = bil.get_bil_info(source_1d, target, 50e5, neighbours=32)  # :343-357
        res = bil.get_sample_from_bil_info((data1 + 9.5).ravel(), t1, s1, vii1, ia1)
        assert abs(np.nanmin(res) - 10.5) < 1e-7 and abs(np.nanmax(res) - 10.5) < 1e-7 and np.isnan(res).sum() == 4
        res = bil.resample_bilinear(data1, source, target, 50e5, neighbours=32)  # :359-383
        assert res.shape == target.shape and res.sum() == 12 and (res == 0).sum() == 4
        res = bil.resample_bilinear(data1, source, target, 50e5, neighbours=32, fill_value=None)
        assert hasattr(res, 'mask') and target.size - res.mask.sum() == 12
        res = bil.resample_bilinear(np.dstack((data1, 2 * data1)), source, target)
        assert res.shape == target.shape + (2,)
    resampler = bil.NumpyBilinearResampler(source, target, 50e5, neighbours=32, epsilon=0)  # :385-410
    res = resampler.resample(data1)
    assert res.shape == target.shape and res.sum() == 12 and (res == 0).sum() == 4
    assert resampler.slices_x.shape == (16, 4) and resampler.mask_slices.shape == (16, 4) and not resampler.mask_slices.any()
    resampler = bil.NumpyBilinearResampler(source, target, 50e3)  # :412-425
    assert resampler.bilinear_t is None and resampler._neighbours == 32 and resampler._reduce_data
    resampler._create_empty_bil_info()
    assert resampler.bilinear_t.shape == (16,) and resampler._index_array.shape == (16, 4)
    assert resampler._index_array.dtype == np.int32 and resampler._valid_input_index.dtype == bool


AREAS = {
    "stere_oblique_ell": lambda g: stere_area(),
    "laea_sphere": lambda g: laea_area(170, 150),
    "eqc": lambda g: g.AreaDefinition("e", "e", "e", "+proj=eqc +datum=WGS84", 300, 200,
                                      (-500000., 4200000., 3600000., 7400000.)),
    "merc": lambda g: g.AreaDefinition("m", "m", "m", "+proj=merc +lon_0=0 +ellps=WGS84", 260, 240,
                                       (-400000., 4800000., 3500000., 10500000.)),
    "geos": lambda g: g.AreaDefinition("geos", "geos", "geos", "+proj=geos +h=35785831 +lon_0=0 +a=6378169 +b=6356583.8",
                                       300, 300, (-1500000., 2500000., 2000000., 4500000.)),
}


@pytest.mark.parametrize("name", sorted(AREAS))
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_bil_info_and_resample_against_oracle(bil, name, dtype):
    from pyresample_b200 import geometry
    lon, lat = jittered_swath(150, 130, dtype, seed=91)
    source = geometry.SwathDefinition(lon, lat)
    area = AREAS[name](geometry)
    rng = np.random.default_rng(92)
    data = rng.standard_normal(lon.shape)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        t, s, vii, ia = bil.get_bil_info(source, area, 60e3, neighbours=32)
        want = bilinear_ref.bil_info(source, area, 60e3, neighbours=32)
    assert np.array_equal(vii, want["valid_input_index"])
    ot, os_, oia = want["bilinear_t"], want["bilinear_s"], want["index_array"]
    assert t.shape == ot.shape and ia.shape == oia.shape
    # the corners are picked from neighbours sorted by distance: float64 coordinates differ by ulps between CUDA and
    # libm, which can only matter at exact-tie scale; everything else must agree to rounding
    same = (ia == oia).all(axis=1)
    assert same.mean() > 0.999
    assert np.array_equal(np.isnan(t[same]), np.isnan(ot[same])) and np.array_equal(np.isnan(s[same]), np.isnan(os_[same]))
    fin = same & np.isfinite(ot) & np.isfinite(os_)
    assert fin.mean() > 0.1
    np.testing.assert_allclose(t[fin], ot[fin], rtol=0, atol=2e-7)
    np.testing.assert_allclose(s[fin], os_[fin], rtol=0, atol=2e-7)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = bil.resample_bilinear(data, source, area, 60e3, neighbours=32, fill_value=None)
        ref = bilinear_ref.resample_bilinear(data, source, area, 60e3, neighbours=32, fill_value=None)
    sm = same.reshape(area.shape)
    assert np.array_equal(np.ma.getmaskarray(got)[sm], np.ma.getmaskarray(ref)[sm])
    ok = sm & ~np.ma.getmaskarray(ref)
    np.testing.assert_allclose(np.ma.getdata(got)[ok], np.ma.getdata(ref)[ok], rtol=0, atol=1e-5)
