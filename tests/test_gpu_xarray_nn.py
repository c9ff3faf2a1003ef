"""XArrayResamplerNN + query_no_distance / _my_index on the GPU, against the oracle's restatement of the block
kernels (future/resamplers/nearest.py:44-108) and the index-level pins of test_resamplers/test_nearest.py:308-373.

xarray is not installable in this image, so the class is driven through a minimal DataArray stand-in that offers
exactly the attributes the reference class touches (dims, data, coords, attrs, sizes, dtype, shape).
"""
import numpy as np
import pytest

from helpers import jittered_swath, laea_area
from oracle import kd_tree_ref
from oracle.knn_ref import KDTreeRef

pytestmark = pytest.mark.gpu


class FakeDataArray(object):
    def __init__(self, data, dims=None, coords=None, attrs=None):
        self.data = data
        self.dims = tuple(dims)
        self.coords = dict(coords or {})
        self.attrs = dict(attrs or {})

    @property
    def shape(self):
        return self.data.shape

    @property
    def ndim(self):
        return self.data.ndim

    def __getitem__(self, key):
        return self.data[key]

    @property
    def dtype(self):
        return self.data.dtype

    @property
    def sizes(self):
        return dict(zip(self.dims, self.data.shape))

    @property
    def values(self):
        return np.asarray(self.data.compute() if hasattr(self.data, "compute") else self.data)


@pytest.fixture()
def kd(cuda_lib, monkeypatch):
    from pyresample_b200 import kd_tree
    if kd_tree.DataArray is None:
        monkeypatch.setattr(kd_tree, "DataArray", FakeDataArray)
    return kd_tree


def _values(x):
    x = x.data if hasattr(x, "dims") else x
    return np.asarray(x.compute() if hasattr(x, "compute") else x)


def test_query_no_distance_index_pins(kd):
    # test_resamplers/test_nearest.py:311-331: one source pixel at (0, 0), 2x2 target
    from pyresample_b200 import xarray_nn
    slons, slats = np.array([0.0, 0.5]), np.array([0.0, 0.5])
    valid, tree = xarray_nn.build_tree(slons, slats)
    tlons = np.array([[0.0, 10.0], [20.0, 30.0]])
    tlats = np.array([[0.0, 10.0], [20.0, 30.0]])
    voi = np.ones((2, 2), dtype=bool)
    res = kd.query_no_distance(tlons, tlats, voi, neighbours=1, epsilon=0, radius=1000, kdtree=tree)
    assert res.dtype == np.int64 and res.shape == (2, 2, 1)
    assert np.array_equal(res[:, :, 0], [[0, -1], [-1, -1]])
    # :333-373: masking source pixel 0 makes pixel 1 the neighbour
    mask = np.array([True, False])
    res = kd.query_no_distance(tlons, tlats, voi, mask=mask, valid_input_index=valid, neighbours=1, epsilon=0,
                               radius=100000, kdtree=tree)
    assert res[0, 0, 0] == 1


def test_query_no_distance_matches_oracle(kd):
    from pyresample_b200 import xarray_nn
    lon, lat = jittered_swath(90, 70, np.float64, seed=31)
    lon[3, 4] = np.nan
    lat[10, 10] = 95.0
    rng = np.random.default_rng(32)
    mask = rng.random(lon.shape) < 0.2
    area = laea_area(64, 48)
    from oracle import proj_ref
    tlon, tlat = proj_ref.area_lonlats(area.proj_dict, area.area_extent, area.width, area.height)
    with np.errstate(invalid="ignore"):
        vii = (lon >= -180) & (lon <= 180) & (lat <= 90) & (lat >= -90)
    voi = np.ones(tlon.shape, dtype=bool)
    voi[0, :5] = False
    otree = KDTreeRef(kd_tree_ref.lonlat2xyz(lon, lat)[vii.ravel()].astype(np.float64))
    for k in (1, 4):
        for m in (None, mask):
            want = kd_tree_ref.query_no_distance(tlon, tlat, voi, mask=m, valid_input_index=vii, neighbours=k,
                                                 epsilon=0, radius=60000, kdtree=otree)
            valid, tree = xarray_nn.build_tree(lon, lat, mask=m)
            assert np.array_equal(valid, vii)
            got = kd.query_no_distance(tlon, tlat, voi, neighbours=k, epsilon=0, radius=60000, kdtree=tree)
            assert got.shape == want.shape and got.dtype == np.int64
            assert np.array_equal(got, want)


def test_my_index_matches_oracle(kd):
    rng = np.random.default_rng(33)
    vii = rng.random(500) < 0.8
    n_valid = int(vii.sum())
    ia = rng.integers(-1, n_valid, size=(20, 30))
    for shape, axis in (((500,), 0), ((3, 500), 1), ((500, 4), 0), ((2, 500, 3), 1)):
        data = rng.standard_normal(shape).astype(np.float32)
        vs = [slice(None)] * len(shape)
        vs[axis] = None
        want = kd_tree_ref._my_index(ia, vii, data, vii_slices=list(vs), ia_slices=list(vs), fill_value=np.nan)
        got = kd._my_index(ia, vii, data, vii_slices=list(vs), ia_slices=list(vs), fill_value=np.nan)
        assert got.shape == want.shape
        assert np.array_equal(got, want, equal_nan=True)
    idata = rng.integers(0, 200, size=(500,)).astype(np.uint8)
    want = kd_tree_ref._my_index(ia, vii, idata, vii_slices=[None], ia_slices=[None], fill_value=255)
    got = kd._my_index(ia, vii, idata, vii_slices=[None], ia_slices=[None], fill_value=255)
    assert got.dtype == np.uint8 and np.array_equal(got, want)


def test_xarray_resampler_nn_swath_to_area(kd):
    """Mirrors test_kd_tree.py:874-895 (2-D swath -> area) with the oracle as the truth."""
    from pyresample_b200 import geometry
    lon, lat = jittered_swath(80, 60, np.float64, seed=34)
    data = np.random.default_rng(35).standard_normal((80, 60))
    area = laea_area(50, 40)
    swath = geometry.SwathDefinition(FakeDataArray(lon, dims=('y', 'x')), FakeDataArray(lat, dims=('y', 'x')))
    swath_np = geometry.SwathDefinition(lon, lat)
    resampler = kd.XArrayResamplerNN(swath, area, radius_of_influence=50000, neighbours=1)
    vii, voi, ia, da_ = resampler.get_neighbour_info()
    assert da_ is None
    assert _values(ia).shape == (40, 50, 1) and _values(ia).dtype == np.int64
    res = resampler.get_sample_from_neighbour_info(FakeDataArray(data, dims=('y', 'x'), attrs={"name": "t"}))
    want = kd_tree_ref.resample_nearest(swath_np, data.ravel(), area, 50000, fill_value=np.nan, reduce_data=False)
    assert res.dims == ('y', 'x') and res.attrs == {"name": "t"}
    assert np.array_equal(_values(res), want, equal_nan=True)
    assert 'x' in res.coords and len(res.coords['y']) == 40
    # 3-D data with a leading band dimension (test_kd_tree.py:969-995) and integer data / fill (845-872)
    data3 = np.stack([data, data * 2, data + 1])
    res3 = resampler.get_sample_from_neighbour_info(FakeDataArray(data3, dims=('bands', 'y', 'x')))
    assert res3.dims == ('bands', 'y', 'x') and _values(res3).shape == (3, 40, 50)
    assert np.array_equal(_values(res3)[1], want * 2, equal_nan=True)
    idata = (np.abs(data) * 10).astype(np.uint8)
    resi = resampler.get_sample_from_neighbour_info(FakeDataArray(idata, dims=('y', 'x')), fill_value=255)
    assert _values(resi).dtype == np.uint8
    resi2 = resampler.get_sample_from_neighbour_info(FakeDataArray(idata, dims=('y', 'x')))  # NaN -> dtype max
    assert np.array_equal(_values(resi), _values(resi2))
    # k > 1: info only (kd_tree.py:1082-1084)
    r8 = kd.XArrayResamplerNN(swath, area, radius_of_influence=50000, neighbours=8)
    _, _, ia8, _ = r8.get_neighbour_info()
    assert _values(ia8).shape == (40, 50, 8)
    with pytest.raises(NotImplementedError):
        r8.get_sample_from_neighbour_info(FakeDataArray(data, dims=('y', 'x')))
    with pytest.raises(ValueError):
        kd.XArrayResamplerNN(swath, geometry.SwathDefinition(lon[0], lat[0]), 1000)
    with pytest.raises(ValueError):
        resampler.get_sample_from_neighbour_info(FakeDataArray(data, dims=('a', 'b')))
    # default radius from the geometries' geocentric resolution (kd_tree.py:949-968)
    auto = kd.XArrayResamplerNN(swath_np, area)
    assert 1000 < auto.radius_of_influence < 200000


def test_xarray_resampler_nn_reference_literals(kd, monkeypatch):
    """test_kd_tree.py:897-995 replayed on the GPU class: area -> area, with / without radius, 3-D data, and the
    dimension-name check that precedes them."""
    import reference_cases as rc
    from pyresample_b200 import geometry
    area_def, src, d2, d3 = rc.xarray_nn_fixtures(geometry)
    S = rc.XARRAY_NN_SUMS
    resampler = kd.XArrayResamplerNN(src, area_def, radius_of_influence=50000, neighbours=1)
    ninfo = resampler.get_neighbour_info()
    assert _values(ninfo[2]).shape == (800, 800, 1)
    with pytest.raises(ValueError):
        resampler.get_sample_from_neighbour_info(FakeDataArray(d2, dims=('my_dim_y', 'my_dim_x')))
    res = resampler.get_sample_from_neighbour_info(FakeDataArray(d2, dims=('y', 'x')))
    assert np.nansum(_values(res)) == S["area_to_area"]
    res3 = resampler.get_sample_from_neighbour_info(
        FakeDataArray(d3, dims=('y', 'x', 'bands'), coords={'bands': ['r', 'g', 'b']}))
    assert sorted(res3.coords['bands']) == ['b', 'g', 'r'] and np.nansum(_values(res3)) == S["area_3d"]
    auto = kd.XArrayResamplerNN(src, area_def, neighbours=1)
    auto.get_neighbour_info()
    assert np.nansum(_values(auto.get_sample_from_neighbour_info(FakeDataArray(d2, dims=('y', 'x'))))) \
        == S["area_no_roi"]

    def boom(*a, **k):
        raise RuntimeError

    monkeypatch.setattr(src, "geocentric_resolution", boom, raising=False)
    monkeypatch.setattr(area_def, "geocentric_resolution", boom, raising=False)
    fallback = kd.XArrayResamplerNN(src, area_def, neighbours=1)
    assert fallback.radius_of_influence == 10000
    fallback.get_neighbour_info()
    assert np.nansum(_values(fallback.get_sample_from_neighbour_info(FakeDataArray(d2, dims=('y', 'x'))))) \
        == S["area_no_roi_no_geocentric"]


def test_float32_swath_in_float64_tree_like_the_reference(kd):
    """kd_tree.py:970-981 transforms float32 lon/lat in float32 (lonlat2xyz on float32 arrays) and only then widens
    the ECEF coordinates to the float64 tree; the query points are float64.  The index must be built from exactly
    those coordinates: indices equal to the oracle's on a jittered float32 swath, k = 1 and 8."""
    from oracle import proj_ref
    from pyresample_b200 import xarray_nn
    lon, lat = jittered_swath(300, 200, np.float32, seed=61)
    area = laea_area(150, 130)
    tlon, tlat = proj_ref.area_lonlats(area.proj_dict, area.area_extent, area.width, area.height)
    with np.errstate(invalid="ignore"):
        vii = (lon >= -180) & (lon <= 180) & (lat <= 90) & (lat >= -90)
    xyz32 = kd_tree_ref.lonlat2xyz(lon, lat)
    assert xyz32.dtype == np.float32
    otree = KDTreeRef(xyz32.reshape(-1, 3)[vii.ravel()].astype(np.float64))
    voi = np.ones(tlon.shape, dtype=bool)
    valid, tree = xarray_nn.build_tree(lon, lat)
    assert np.array_equal(valid, vii)
    for k in (1, 8):
        want = kd_tree_ref.query_no_distance(tlon, tlat, voi, valid_input_index=vii, neighbours=k, epsilon=0,
                                             radius=40000, kdtree=otree)
        got = kd.query_no_distance(tlon, tlat, voi, neighbours=k, epsilon=0, radius=40000, kdtree=tree)
        assert np.array_equal(got, want)


def test_xarray_resampler_nn_computes_nothing_until_asked(kd):
    """test_kd_tree.py:908-919 (`assert_maximum_dask_computes(0)` around get_neighbour_info and
    get_sample_from_neighbour_info): nothing is computed until a result is used.  Counted with the module's own
    counter of deferred computations that ran (dask is not installed here: the results are LazyArrays)."""
    import reference_cases as rc
    from pyresample_b200 import _cabi, geometry, xarray_nn
    area_def, src, d2, _ = rc.xarray_nn_fixtures(geometry)
    L = _cabi.lib()
    ran0, launches0 = xarray_nn.COMPUTE_COUNT, L.prs_launch_count()
    resampler = kd.XArrayResamplerNN(src, area_def, radius_of_influence=50000, neighbours=1)
    ninfo = resampler.get_neighbour_info()
    assert ninfo[3] is None
    for val in ninfo[:3]:
        assert hasattr(val, "compute")  # lazy (dask array or LazyArray)
    res = resampler.get_sample_from_neighbour_info(FakeDataArray(d2, dims=('y', 'x')))
    assert hasattr(res.data, "compute")
    assert xarray_nn.COMPUTE_COUNT == ran0  # no deferred computation (index build, search, gather) has run yet
    total = np.nansum(_values(res))
    assert total == rc.XARRAY_NN_SUMS["area_to_area"]
    assert xarray_nn.COMPUTE_COUNT > ran0 and L.prs_launch_count() > launches0
    # a second use does not compute the tree / the query again
    ran1 = xarray_nn.COMPUTE_COUNT
    ia = _values(ninfo[2])
    assert ia.shape == (800, 800, 1) and ia.dtype == np.int64
    assert xarray_nn.COMPUTE_COUNT <= ran1 + 1


def test_query_resample_kdtree_matches_get_neighbour_info(kd):
    """kd_tree.py:983-1004: the class's own query step on a tree from _create_resample_kdtree."""
    import reference_cases as rc
    from pyresample_b200 import geometry
    area_def, src, _, _ = rc.xarray_nn_fixtures(geometry)
    resampler = kd.XArrayResamplerNN(src, area_def, radius_of_influence=50000, neighbours=1)
    vii, voi, ia, _ = resampler.get_neighbour_info()
    valid_input_index, tree = resampler._create_resample_kdtree()
    assert np.array_equal(_values(valid_input_index), _values(vii))
    tlons, tlats = area_def.get_lonlats()
    res, dist = resampler.query_resample_kdtree(tree, tlons, tlats, np.ones(tlons.shape, dtype=bool), None)
    assert dist is None  # the reference returns (index_array, None) as well
    assert np.array_equal(_values(res), _values(ia))
