"""Source cropping (pyresample/slicer.py): the slices must enclose everything the resampling touches -
resampling the cropped source gives the grid of the uncropped one."""
import numpy as np
import pytest

from helpers import jittered_swath, laea_area

pytestmark = pytest.mark.gpu


def test_swath_and_area_slices_enclose_the_target(cuda_lib):
    from pyresample_b200 import geometry, kd_tree, slicer
    radius = 30000
    target = laea_area(120, 100, half=400000.0, lat_0=52, lon_0=15)
    # swath much larger than the target
    lon, lat = jittered_swath(400, 360, np.float64, seed=95, lon0=-10, lat0=35, dlon=55, dlat=35)
    data = np.random.default_rng(96).standard_normal(lon.shape)
    swath = geometry.SwathDefinition(lon, lat)
    col_sl, row_sl = slicer.create_slicer(swath, target, margin=radius).get_slices()
    assert (row_sl.stop - row_sl.start) * (col_sl.stop - col_sl.start) < 0.5 * lon.size
    cropped = geometry.SwathDefinition(lon[row_sl, col_sl], lat[row_sl, col_sl])
    full = kd_tree.resample_nearest(swath, data, target, radius, fill_value=np.nan)
    part = kd_tree.resample_nearest(cropped, data[row_sl, col_sl], target, radius, fill_value=np.nan)
    assert np.array_equal(full, part, equal_nan=True) and np.isfinite(full).mean() > 0.9
    # area source: a geostationary full disc cropped to the same target
    disc = geometry.AreaDefinition("fd", "fd", "fd", "+proj=geos +h=35785831 +lon_0=0 +a=6378169 +b=6356583.8", 928, 928,
                                   (-5570248.4773, -5567248.0742, 5567248.0742, 5570248.4773))
    ddata = np.random.default_rng(97).standard_normal(disc.shape)
    col_sl, row_sl = slicer.get_area_slices(disc, target, margin=radius)
    assert (row_sl.stop - row_sl.start) < 300 and (col_sl.stop - col_sl.start) < 300
    full = kd_tree.resample_nearest(disc, ddata, target, radius, fill_value=np.nan)
    sub = disc[row_sl, col_sl] if hasattr(disc, "__getitem__") else None
    if sub is not None:
        part = kd_tree.resample_nearest(sub, ddata[row_sl, col_sl], target, radius, fill_value=np.nan)
        assert np.array_equal(full, part, equal_nan=True)
    # no overlap at all
    far = laea_area(50, 50, half=200000.0, lat_0=-60, lon_0=-150)
    with pytest.raises(slicer.IncompatibleAreas):
        slicer.create_slicer(swath, far, margin=radius).get_slices()
