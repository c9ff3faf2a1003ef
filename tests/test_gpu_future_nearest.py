"""KDTreeNearestXarrayResampler on the GPU: the reference's own tests
(pyresample/test/test_resamplers/test_nearest.py:55-305) replayed with numpy arrays and a minimal DataArray stand-in
(xarray is not installable in this image), literals and all, plus full-array comparison with oracle/nearest_ref.py."""
import numpy as np
import pytest

import reference_cases as rc
from oracle import nearest_ref
from test_gpu_xarray_nn import FakeDataArray

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods(cuda_lib):
    from pyresample_b200 import future_nearest, geometry, registry
    return future_nearest, geometry, registry


@pytest.fixture(scope="module")
def fx(mods):
    return rc.nearest_resampler_fixtures(mods[1])


def _xr(arr, dims, **kw):
    return FakeDataArray(arr, dims=dims, **kw)


def test_nearest_swath_1d_to_grid(mods, fx):
    """test_nearest.py:55-94: 1-D swath with named dimension -> 2-D grid; float data with mask, integer data."""
    fn, geometry, _ = mods
    swath = geometry.SwathDefinition(_xr(fx["swath_1d"][0], ('my_dim1',)), _xr(fx["swath_1d"][1], ('my_dim1',)))
    grid = geometry.CoordinateDefinition(*fx["coord_def_2d_float32"])
    data = _xr(fx["data_1d"], ('my_dim1',))
    resampler = fn.KDTreeNearestXarrayResampler(swath, grid)
    res = resampler.resample(data, mask_area=np.isnan(fx["data_1d"]), radius_of_influence=100000)
    assert isinstance(res, FakeDataArray) and res.dims == ('y', 'x')
    np.testing.assert_allclose(res.values, rc.NEAREST_1D_EXPECTED)
    ints = _xr(np.array([1, 2, 3]), ('my_dim1',))
    res = resampler.resample(ints, fill_value=255, radius_of_influence=100000)
    assert res.values.dtype == ints.dtype
    np.testing.assert_equal(res.values, rc.NEAREST_1D_INT_EXPECTED)
    res = resampler.resample(_xr(np.array([1, 2, 3], dtype=np.uint8), ('my_dim1',)), radius_of_influence=100000)
    np.testing.assert_equal(res.values, rc.NEAREST_1D_INT_EXPECTED)  # NaN fill -> dtype max (nearest.py:289-293)


def test_nearest_reference_cross_sums(mods, fx):
    """test_nearest.py:96-160, 212-228 - the reference's literals, and the whole arrays against the oracle."""
    fn, geometry, _ = mods
    S = rc.NEAREST_RESAMPLER_SUMS
    euro = geometry.SwathDefinition(_xr(fx["euro_lonlats"][0], ('y', 'x')), _xr(fx["euro_lonlats"][1], ('y', 'x')))
    euro_np = geometry.SwathDefinition(*fx["euro_lonlats"])
    anti = geometry.SwathDefinition(_xr(fx["antimeridian_lonlats"][0], ('y', 'x')),
                                    _xr(fx["antimeridian_lonlats"][1], ('y', 'x')))
    anti_np = geometry.SwathDefinition(*fx["antimeridian_lonlats"])
    src, tgt, pm = fx["area_stere_source"], fx["area_stere_target"], fx["area_lonlat_pm180_target"]
    d2 = _xr(fx["data_2d"], ('y', 'x'), attrs={"name": "d2"})
    nomask = nearest_ref.compute_data_mask(fx["data_2d"], (0, 1))

    res = fn.KDTreeNearestXarrayResampler(euro, tgt).resample(d2, radius_of_influence=50000)
    assert res.values.shape == tgt.shape and float(np.nansum(res.values)) == S["swath_2d_to_area"]
    want = nearest_ref.resample(euro_np, tgt, fx["data_2d"], (0, 1), mask=nomask, radius_of_influence=50000)
    assert np.array_equal(res.values, want, equal_nan=True)
    assert res.attrs["area"] is tgt and res.attrs["name"] == "d2" and len(res.coords["x"]) == 85

    resampler = fn.KDTreeNearestXarrayResampler(src, tgt)
    resampler.precompute(radius_of_influence=50000)
    res = resampler.resample(d2, radius_of_influence=50000)
    assert len(resampler._internal_cache) == 1  # the precomputed neighbours were reused
    assert float(np.nansum(res.values)) == S["area_to_area"]
    assert np.array_equal(res.values, nearest_ref.resample(src, tgt, fx["data_2d"], (0, 1), radius_of_influence=50000),
                          equal_nan=True)

    res = fn.KDTreeNearestXarrayResampler(anti, pm).resample(d2, radius_of_influence=50000)
    assert float(np.nansum(res.values)) == S["swath_pm180"]
    want = nearest_ref.resample(anti_np, pm, fx["data_2d"], (0, 1), mask=nomask, radius_of_influence=50000)
    assert np.array_equal(res.values, want, equal_nan=True)

    resampler = fn.KDTreeNearestXarrayResampler(src, tgt)
    resampler.precompute()
    res = resampler.resample(d2)  # radius from the geocentric resolutions of both areas
    assert float(np.nansum(res.values)) == S["area_no_roi"] and res.values.shape == tgt.shape
    assert abs(resampler._compute_radius_of_influence() - nearest_ref.compute_radius_of_influence(src, tgt)) < 1e-3

    d3 = _xr(fx["data_3d"], ('y', 'x', 'bands'), coords={'bands': ['r', 'g', 'b']})
    resampler = fn.KDTreeNearestXarrayResampler(src, tgt)
    resampler.precompute(radius_of_influence=50000)
    res = resampler.resample(d3, radius_of_influence=50000)
    assert list(res.coords['bands']) == ['r', 'g', 'b'] and res.dims == ('y', 'x', 'bands')
    assert float(np.nansum(res.values)) == S["area_3d"] and res.values.shape[:2] == tgt.shape


def test_nearest_no_geocentric_resolution_falls_back_to_10km(mods, fx, monkeypatch):
    """test_nearest.py:162-182."""
    fn, geometry, _ = mods
    src, tgt = fx["area_stere_source"], fx["area_stere_target"]

    def boom(*a, **k):
        raise RuntimeError

    monkeypatch.setattr(src, "geocentric_resolution", boom, raising=False)
    monkeypatch.setattr(tgt, "geocentric_resolution", boom, raising=False)
    res = fn.KDTreeNearestXarrayResampler(src, tgt).resample(_xr(fx["data_2d"], ('y', 'x')))
    assert np.nansum(res.values) == rc.NEAREST_RESAMPLER_SUMS["area_no_roi_no_geocentric"]


def test_nearest_object_types(mods, fx):
    """test_nearest.py:184-210: the result is the kind of object that came in (numpy in -> numpy out)."""
    fn, geometry, _ = mods
    src, tgt = fx["area_stere_source"], fx["area_stere_target"]
    res = fn.KDTreeNearestXarrayResampler(src, tgt).resample(fx["data_2d"])
    assert type(res) is np.ndarray and res.shape == tgt.shape
    assert np.nansum(res) == rc.NEAREST_RESAMPLER_SUMS["area_no_roi"]
    with pytest.raises(ValueError, match="not 2D"):
        fn.KDTreeNearestXarrayResampler(src, tgt).resample(fx["data_3d"])


def test_nearest_invalid_usage(mods, fx):
    """test_nearest.py:231-305."""
    fn, geometry, _ = mods
    src, tgt = fx["area_stere_source"], fx["area_stere_target"]
    resampler = fn.KDTreeNearestXarrayResampler(src, tgt)
    with pytest.raises(ValueError, match='.*dimensions do not match.*'):
        resampler.resample(_xr(fx["data_2d"], ('my_dim_y', 'my_dim_x')))
    with pytest.raises(ValueError, match='.*dimensions do not match.*'):
        resampler.resample(_xr(fx["data_3d"], ('my_dim_y', 'my_dim_x', 'bands')))
    euro = geometry.SwathDefinition(_xr(fx["euro_lonlats"][0], ('y', 'x')), _xr(fx["euro_lonlats"][1], ('y', 'x')))
    with pytest.raises(ValueError, match='.*dimensions do not match.*'):
        fn.KDTreeNearestXarrayResampler(euro, tgt).resample(_xr(fx["data_2d"], ('my_dim_y', 'my_dim_x')))
    with pytest.raises(ValueError, match='.*not consecutive.*'):
        resampler.resample(_xr(np.zeros((50, 3, 10), np.float32), ('y', 'bands', 'x')))
    with pytest.raises(ValueError, match='.*not equal to the shape.*'):
        resampler.resample(_xr(np.zeros((40, 10), np.float32), ('y', 'x')))
    with pytest.raises(ValueError, match=".*'mask'.*shape.*"):
        resampler.precompute(mask=np.zeros((40, 10), dtype=bool))
    with pytest.raises(ValueError, match="2 dimensions"):
        fn.KDTreeNearestXarrayResampler(src, geometry.SwathDefinition(np.zeros(4), np.zeros(4)))
    with pytest.raises(NotImplementedError):
        resampler.get_sample_from_neighbor_info(_xr(fx["data_2d"], ('y', 'x')), None, None, neighbors=2)


def test_nearest_data_mask_steers_the_neighbours(mods, fx):
    """A pixel whose data are invalid is skipped by the search (nearest.py:395-413, 518-534) and cached by mask."""
    fn, geometry, _ = mods
    euro = geometry.SwathDefinition(_xr(fx["euro_lonlats"][0], ('y', 'x')), _xr(fx["euro_lonlats"][1], ('y', 'x')))
    euro_np = geometry.SwathDefinition(*fx["euro_lonlats"])
    tgt = fx["area_stere_target"]
    data = fx["data_2d"].copy() + 1
    data[10:30, 2:8] = np.nan
    resampler = fn.KDTreeNearestXarrayResampler(euro, tgt)
    res = resampler.resample(_xr(data, ('y', 'x')), radius_of_influence=120000)
    mask = nearest_ref.compute_data_mask(data, (0, 1))
    assert np.array_equal(resampler.compute_data_mask(_xr(data, ('y', 'x'))), mask) and mask.sum() == 120
    want = nearest_ref.resample(euro_np, tgt, data, (0, 1), mask=mask, radius_of_influence=120000)
    assert np.array_equal(res.values, want, equal_nan=True)
    unmasked = resampler.resample(_xr(data, ('y', 'x')), mask_area=False, radius_of_influence=120000)
    assert np.isnan(unmasked.values).sum() > np.isnan(res.values).sum()  # NaN pixels were picked as neighbours
    assert len(resampler._internal_cache) == 2
    resampler.resample(_xr(data, ('y', 'x')), radius_of_influence=120000)
    assert len(resampler._internal_cache) == 2
    # integer data: the mask is "== _FillValue" (default: dtype max)
    idata = np.arange(500, dtype=np.int16).reshape(50, 10)
    idata[0, 0] = np.iinfo(np.int16).max
    assert resampler.compute_data_mask(_xr(idata, ('y', 'x'))).sum() == 1
    assert resampler.compute_data_mask(_xr(idata, ('y', 'x'), attrs={'_FillValue': 7})).sum() == 1


def test_registry(mods, fx):
    """registry.py:33-151."""
    fn, geometry, registry = mods
    assert registry.list_resamplers() == ["nearest"]
    r = registry.create_resampler(fx["area_stere_source"], fx["area_stere_target"])
    assert isinstance(r, fn.KDTreeNearestXarrayResampler) and r.version == "0.1"
    with pytest.raises(ValueError, match="already registered"):
        registry.register_resampler("nearest", fn.KDTreeNearestXarrayResampler)
    registry.register_resampler("mine", fn.KDTreeNearestXarrayResampler)
    assert registry.list_resamplers() == ["mine", "nearest"]
    assert isinstance(registry.create_resampler(fx["area_stere_source"], fx["area_stere_target"], "mine"),
                      fn.Resampler)
    registry.unregister_resampler("mine")
    with pytest.raises(KeyError):
        registry.create_resampler(fx["area_stere_source"], fx["area_stere_target"], "mine")
