"""ImageContainerNearest / linesample helpers: pyresample/test/test_image.py:130-215 replayed on the
GPU path with the reference's literals."""
import warnings

import numpy as np
import pytest

import reference_cases as rc
from oracle import kd_tree_ref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods(cuda_lib):
    from pyresample_b200 import geometry, image, utils
    return geometry, image, utils


@pytest.fixture(scope="module")
def areas(mods):
    return rc.image_fixtures(mods[0])


@pytest.fixture(scope="module")
def msg_data():
    return np.fromfunction(lambda y, x: y * x * 10 ** -6, (3712, 3712))


def _container(image, *args, **kwargs):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", FutureWarning)
        return image.ImageContainerNearest(*args, **kwargs)


def test_nearest_neighbour_and_resize(mods, areas, msg_data):
    geometry, image, utils = mods
    area_def, msg_area, msg_resize = areas
    with pytest.warns(FutureWarning):
        msg_con = image.ImageContainerNearest(msg_data, msg_area, 50000, segments=1)
    area_con = msg_con.resample(area_def)
    assert isinstance(area_con, image.ImageContainerNearest) and area_con.image_data.shape == (800, 800)
    assert abs(area_con.image_data.sum() - rc.IMAGE_NEAREST_SUM) < 5e-8
    want = kd_tree_ref.resample_nearest(msg_area, msg_data.ravel(), area_def, 50000, segments=1)
    assert np.array_equal(area_con.image_data, want)
    res = msg_con.resample(msg_resize).image_data
    assert abs(res.sum() - rc.IMAGE_NEAREST_RESIZE_SUM) < 5e-8


def test_nearest_neighbour_multi_and_linesample(mods, areas, msg_data):
    geometry, image, utils = mods
    area_def, msg_area, _ = areas
    data = np.dstack((msg_data, msg_data * 2))
    res = _container(image, data, msg_area, 50000, segments=1).resample(area_def).image_data
    assert abs(res[:, :, 0].sum() - rc.IMAGE_NEAREST_SUM) < 5e-8
    assert abs(res[:, :, 1].sum() - rc.IMAGE_NEAREST_SUM * 2) < 1e-7
    # test_image.py:167-189: the same through precomputed row / column index arrays
    with pytest.warns(FutureWarning):
        msg_con = image.ImageContainer(data, msg_area)
    rows, cols = utils.generate_nearest_neighbour_linesample_arrays(msg_area, area_def, 50000)
    assert rows.dtype == np.uint16 and cols.dtype == np.uint16 and rows.shape == (800, 800)
    sampled = msg_con.get_array_from_linesample(rows, cols)
    assert np.array_equal(sampled, res)
    masked = image.get_image_from_linesample(rows, cols, data[:, :, 0], fill_value=None)
    assert np.ma.isMaskedArray(masked) and masked.mask.sum() == (rows == 3712).sum()
    with pytest.raises(NotImplementedError):
        msg_con.resample(area_def)


def test_nearest_swath_and_errors(mods, areas):
    geometry, image, utils = mods
    area_def = areas[0]
    data = np.fromfunction(lambda y, x: y * x, (50, 10))
    lons = np.fromfunction(lambda y, x: 3 + x, (50, 10))
    lats = np.fromfunction(lambda y, x: 75 - y, (50, 10))
    swath_def = geometry.SwathDefinition(lons=lons, lats=lats)
    res = _container(image, data, swath_def, 50000, segments=1).resample(area_def).image_data
    assert res.sum() == 15874591.0                                    # test_image.py:191-202
    res3 = _container(image, np.dstack(3 * (data,)), swath_def, 50000, segments=2).resample(area_def).image_data
    assert res3.sum() == 3 * 15874591.0                               # test_image.py:204-216
    with pytest.raises(TypeError):
        _container(image, [1, 2], swath_def, 50000)
    with pytest.raises(ValueError, match="Size mismatch"):
        _container(image, np.zeros((50, 11)), swath_def, 50000)
    with pytest.raises(ValueError, match="Unexpected number of dimensions"):
        _container(image, np.zeros(50), swath_def, 50000)
    assert utils.fwhm2sigma(41627.730557884883) == pytest.approx(25000.0, rel=1e-6)
