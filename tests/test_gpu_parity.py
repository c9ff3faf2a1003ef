"""GPU parity tests: the CUDA path (through the C-ABI) against the CPU oracle on the same seeded inputs.

Bit-exact for indices, masks and byte gathers; float64 distances / weighted outputs within 1e-6 relative,
float32 within 1e-4 relative (BASELINE.json north_star).
"""
import ctypes
import warnings

import numpy as np
import pytest

from helpers import index_mismatch_report, jittered_swath, laea_area, stere_area
from oracle import kd_tree_ref, proj_ref
from oracle.knn_ref import brute_query

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def kd(cuda_lib):
    from pyresample_b200 import kd_tree
    return kd_tree


@pytest.fixture(scope="module")
def geometry():
    from pyresample_b200 import geometry
    return geometry


# ------------------------------------------------------------------------------------------------ coordinates
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_lonlat_to_xyz_matches_numpy_chain(kd, geometry, dtype):
    rng = np.random.default_rng(1)
    lon = rng.uniform(-180, 180, 200000).astype(dtype)
    lat = rng.uniform(-90, 90, 200000).astype(dtype)
    lon[:4] = [-180, 180, 0, 90]
    lat[:4] = [-90, 90, 0, 45]
    got = geometry.SwathDefinition(lon, lat).get_cartesian_coords()
    want = kd_tree_ref.transform_lonlats(lon, lat)
    assert got.dtype == want.dtype == dtype
    if dtype == np.float32:
        # numpy's float32 sin/cos SIMD kernels are reproduced operation for operation
        assert np.array_equal(got, want)
    else:
        assert np.max(np.abs(got - want)) < 1e-8  # <= 4 ulp of 6.37e6 in float64 (CUDA vs libm sin/cos)


def test_cartesian_known_answers(kd, geometry):
    # pyresample/test/test_spatial_mp.py:33-35
    lon = np.array([0., 10., 25.])
    lat = np.array([0., 10., 25.])
    got = geometry.SwathDefinition(lon, lat).get_cartesian_coords()
    want = np.array([[6370997., 0., 0.], [6178887.9339746, 1089504.6535337, 1106312.0189715],
                     [5233097.4664751, 2440233.4244856, 2692499.6776952]])
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-3)


AREAS = {
    "stere_oblique_ell": ({'a': '6378144.0', 'b': '6356759.0', 'lat_0': '50.00', 'lat_ts': '50.00', 'lon_0': '8.00',
                           'proj': 'stere'}, 400, 380, [-1370912.72, -909968.64, 1029087.28, 1490031.36]),
    "stere_polar_ell": ("+proj=stere +lat_0=90 +lat_ts=70 +lon_0=-45 +a=6378273 +b=6356889.449", 300, 300,
                        (-4000000., -4000000., 4000000., 4000000.)),
    "stere_south_ell": ("+proj=stere +lat_0=-90 +lat_ts=-71 +lon_0=0 +ellps=WGS84", 200, 210,
                        (-3000000., -3000000., 3000000., 3000000.)),
    "stere_sphere": ("+proj=stere +lat_0=45 +lon_0=10 +R=6371000", 150, 160, (-2000000., -2000000., 2000000., 2000000.)),
    "laea_sphere_oblique": ("+proj=laea +lat_0=52 +lon_0=15 +a=6371228 +units=m", 500, 500,
                            (-1000000., -1000000., 1000000., 1000000.)),
    "laea_north": ("+proj=laea +lat_0=90 +lon_0=0 +a=6371228.0 +units=m", 425, 425,
                   (-5326849.0625, -5326849.0625, 5326849.0625, 5326849.0625)),
    "laea_south": ("+proj=laea +lat_0=-90 +lon_0=0 +a=6371228.0 +units=m", 425, 425,
                   (-5326849.0625, -5326849.0625, 5326849.0625, 5326849.0625)),
    "laea_ell": ("+proj=laea +lat_0=52 +lon_0=10 +x_0=4321000 +y_0=3210000 +ellps=GRS80 +units=m", 300, 280,
                 (2500000., 1500000., 6500000., 5500000.)),
    "eqc_world": ("+proj=eqc +datum=WGS84", 720, 360,
                  (-20037508.342789244, -10018754.171394622, 20037508.342789244, 10018754.171394622)),
    "longlat": ("+proj=longlat +datum=WGS84", 360, 180, (-180, -90, 180, 90)),
    "merc": ("+proj=merc +lon_0=0 +ellps=WGS84", 300, 200, (-15000000., -8000000., 15000000., 8000000.)),
    "geos": ("+proj=geos +h=35785831 +lon_0=0 +a=6378169 +b=6356583.8", 464, 464,
             (-5570248.4773, -5567248.0742, 5567248.0742, 5570248.4773)),
    "geos_sweep_x": ("+proj=geos +h=35786023 +lon_0=-75 +sweep=x +ellps=GRS80", 339, 339,
                     (-5434894.885056, -5434894.885056, 5434894.885056, 5434894.885056)),
}


@pytest.mark.parametrize("name", sorted(AREAS))
def test_area_lonlats_match_oracle(kd, geometry, name):
    proj, w, h, ext = AREAS[name]
    area = geometry.AreaDefinition(name, name, name, proj, w, h, ext)
    lon, lat = area.get_lonlats()
    olon, olat = proj_ref.area_lonlats(area.proj_dict, ext, w, h)
    assert lon.shape == (h, w)
    assert np.array_equal(np.isfinite(lon), np.isfinite(olon))
    fin = np.isfinite(olon)
    dlon = np.abs(lon[fin] - olon[fin])
    dlon = np.minimum(dlon, np.abs(dlon - 360))
    assert dlon.max() < 1e-9, dlon.max()
    assert np.abs(lat[fin] - olat[fin]).max() < 1e-9
    # float32 request: cast of the float64 result (geometry.py:2637-2638), projection vectors in float32
    lon32, lat32 = area.get_lonlats(dtype=np.float32)
    olon32, olat32 = proj_ref.area_lonlats(area.proj_dict, ext, w, h, dtype=np.float32)
    assert lon32.dtype == np.float32
    neq = (lon32 != olon32) & fin | (lat32 != olat32) & fin
    assert neq.mean() < 1e-4  # double rounding at a float32 boundary only


def test_stere_known_answer(kd):
    # pyresample/test/test_geometry/test_area.py:293-306
    area = stere_area()
    lon, lat = area.get_lonlats()
    np.testing.assert_allclose(lon[400, 400], 5.5028467120975835)
    np.testing.assert_allclose(lat[400, 400], 52.566998432390619)
    cart = area.get_cartesian_coords()
    exp = 5872039989466.8457031
    assert abs(cart.sum() - exp) < 1e-7 * exp
    sl_lon, sl_lat = area.get_lonlats(data_slice=(slice(10, 20), slice(5, 50, 3)))
    assert np.array_equal(sl_lon, lon[10:20, 5:50:3])
    side_lon, _ = area.get_lonlats(data_slice=(slice(None, None, -1), 0))
    assert np.array_equal(side_lon, lon[::-1, 0])


# ------------------------------------------------------------------------------------------------ neighbour info
def _compare_info(got, want, rtol, allow_frac=0.0):
    vii, voi, ia, da = got
    ovii, ovoi, oia, oda = want
    assert np.array_equal(vii, ovii)
    assert np.array_equal(voi, ovoi)
    assert ia.shape == oia.shape and ia.dtype == oia.dtype == np.uint32
    assert da.dtype == oda.dtype
    ndiff, nties = index_mismatch_report(ia, oia, da, oda)
    assert ndiff - nties <= allow_frac * ia.size, (ndiff, nties, ia.size)
    same = (ia == oia)
    fin = np.isfinite(oda) & same
    assert np.array_equal(np.isinf(da), np.isinf(oda)) or allow_frac > 0
    np.testing.assert_allclose(da[fin], oda[fin], rtol=rtol, atol=0)
    return ndiff, nties


@pytest.mark.parametrize("k", [1, 8])
def test_neighbour_info_f64_swath_to_area(kd, geometry, k):
    lon, lat = jittered_swath(300, 240, np.float64, seed=3)
    swath = geometry.SwathDefinition(lon, lat)
    area = laea_area(260, 250)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = kd.get_neighbour_info(swath, area, 50000, neighbours=k)
        want = kd_tree_ref.get_neighbour_info(swath, area, 50000, neighbours=k)
    _compare_info(got, want, 1e-6)


@pytest.mark.parametrize("k", [1, 8])
def test_neighbour_info_f32_swath_to_area(kd, geometry, k):
    lon, lat = jittered_swath(300, 240, np.float32, seed=4)
    swath = geometry.SwathDefinition(lon, lat)
    area = laea_area(260, 250)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = kd.get_neighbour_info(swath, area, 50000, neighbours=k)
        want = kd_tree_ref.get_neighbour_info(swath, area, 50000, neighbours=k)
    assert got[3].dtype == np.float32
    # float32: target lon/lats may differ from the oracle's in the last bit for ~1e-7 of the pixels
    _compare_info(got, want, 1e-4, allow_frac=1e-4)


def test_query_against_c_brute_force_with_ties(kd, geometry):
    """Exact ties (lattice + duplicated points) must resolve to the lowest source index."""
    lon, lat = np.meshgrid(np.arange(0., 20., 0.5), np.arange(40., 60., 0.5))
    lon = np.concatenate([lon.ravel(), lon.ravel()[:300]])  # duplicates
    lat = np.concatenate([lat.ravel(), lat.ravel()[:300]])
    swath = geometry.SwathDefinition(lon, lat)
    tlon, tlat = np.meshgrid(np.arange(0.25, 19.5, 0.5), np.arange(40.25, 59.5, 0.5))  # cell centres: 4-way ties
    target = geometry.SwathDefinition(tlon.ravel(), tlat.ravel())
    for k in (1, 4, 8):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            vii, voi, ia, da = kd.get_neighbour_info(swath, target, 80000, neighbours=k, reduce_data=False)
        src_xyz = kd_tree_ref.transform_lonlats(lon, lat)
        tgt_xyz = kd_tree_ref.transform_lonlats(tlon.ravel(), tlat.ravel())
        bd, bi = brute_query(src_xyz, tgt_xyz, k, 80000)
        # coordinates differ by <= a few ulp between CUDA and numpy in float64, so compare as sets of ties:
        # where the brute-force distances of two candidates are equal the lower index must come first
        ia = ia.reshape(len(ia), -1)
        bi = bi.reshape(len(bi), -1)
        bd = bd.reshape(len(bd), -1)
        da = da.reshape(len(da), -1)
        np.testing.assert_allclose(da[np.isfinite(bd)], bd[np.isfinite(bd)], rtol=1e-9)
        assert (ia == bi).mean() > 0.5
        # duplicates: the first 300 points have exact twins with identical coordinates -> must report the lower index
        twins = ia[:, 0] >= 1600
        assert not np.any(twins & (ia[:, 0] - 1600 < 300) & (da[:, 0] == 0))


def test_area_to_swath_and_grid_reduce(kd, geometry):
    """Area source (geos-like, with off-disk inf pixels) -> 1-D swath target, reduce_data on the target side."""
    proj, w, h, ext = AREAS["geos"]
    area = geometry.AreaDefinition("geos", "geos", "geos", proj, 232, 232, ext)
    rng = np.random.default_rng(5)
    tlon = rng.uniform(-70, 70, 5000)
    tlat = rng.uniform(-70, 70, 5000)
    tlon[:3] = [np.nan, 200., 10.]
    tlat[:3] = [0., 0., 95.]
    target = geometry.SwathDefinition(tlon, tlat)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = kd.get_neighbour_info(area, target, 60000, neighbours=1)
        want = kd_tree_ref.get_neighbour_info(area, target, 60000, neighbours=1)
    _compare_info(got, want, 1e-6)


def test_reduce_data_masks(kd, geometry):
    """valid_input_index for swath -> area with reduce_data, incl. pole-covering and dateline areas."""
    lon, lat = jittered_swath(200, 300, np.float64, seed=6, lon0=-180, lat0=-90, dlon=360, dlat=180, rot_deg=0)
    lon = np.clip(lon, -180, 180)
    lat = np.clip(lat, -90, 90)
    swath = geometry.SwathDefinition(lon, lat)
    for name in ("stere_oblique_ell", "laea_north", "laea_south", "eqc_world", "geos", "stere_polar_ell"):
        proj, w, h, ext = AREAS[name]
        area = geometry.AreaDefinition(name, name, name, proj, 64, 48, ext)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            got = kd.get_neighbour_info(swath, area, 30000, neighbours=1)
            want = kd_tree_ref.get_neighbour_info(swath, area, 30000, neighbours=1)
        assert np.array_equal(got[0], want[0]), name
        _compare_info(got, want, 1e-6)


# ------------------------------------------------------------------------------------------------ resampling
def test_reference_known_answers_through_product(kd, geometry):
    """Literals of pyresample/test/test_kd_tree.py replayed through the CUDA path."""
    area_def = stere_area()
    tdata = np.array([1, 2, 3])
    tswath = geometry.SwathDefinition(lons=np.array([11.280789, 12.649354, 12.080402]),
                                      lats=np.array([56.011037, 55.629675, 55.641535]))
    tgrid = geometry.CoordinateDefinition(lons=np.array([12.562036]), lats=np.array([55.715613]))
    res = kd.resample_nearest(tswath, tdata.ravel(), tgrid, 100000, reduce_data=False, segments=1)
    assert res[0] == 2  # :58-62
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        res = kd.resample_gauss(tswath, tdata.ravel(), tgrid, 50000, 25000, reduce_data=False, segments=1)
        assert len(w) == 1 and 'Searching' in str(w[0].message)
    assert res[0] == pytest.approx(2.2020729, abs=1e-5)  # :64-71
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        res = kd.resample_custom(tswath, tdata.ravel(), tgrid, 50000, lambda d: 1 - d / 100000.0, reduce_data=False,
                                 segments=1)
        assert len(w) == 1 and 'Searching' in str(w[0].message)
    assert res[0] == pytest.approx(2.4356757, abs=1e-5)  # :73-83
    sigma = 41627.730557884883 / (2 * np.sqrt(np.log(2)))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res, stddev, count = kd.resample_gauss(tswath, tdata, tgrid, 100000, sigma, with_uncert=True)
    assert res[0] == pytest.approx(2.20206560694, abs=1e-5)  # :85-99
    assert stddev[0] == pytest.approx(0.707115076173, abs=1e-5)
    assert count[0] == 3
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res, stddev, count = kd.resample_custom(tswath, tdata, tgrid, 100000, lambda d: 1 - d / 100000.0,
                                                with_uncert=True)
    assert res[0] == pytest.approx(2.32193149, abs=1e-5)  # :101-114
    assert stddev[0] == pytest.approx(0.81817972, abs=1e-5)
    assert count[0] == 3

    data = np.fromfunction(lambda y, x: y * x, (50, 10))
    lons = np.fromfunction(lambda y, x: 3 + x, (50, 10))
    lats = np.fromfunction(lambda y, x: 75 - y, (50, 10))
    swath_def = geometry.SwathDefinition(lons=lons, lats=lats)
    res = kd.resample_nearest(swath_def, data.ravel(), area_def, 50000, segments=1)
    assert res.sum() == 15874591.0  # :116-125
    res = kd.resample_nearest(swath_def, data.ravel(), area_def, 50000, segments=2)
    assert res.sum() == 15874591.0  # :216-225
    data_multi = np.column_stack((data.ravel(), data.ravel(), data.ravel()))
    res = kd.resample_nearest(swath_def, data_multi, area_def, 50000, segments=1)
    assert res.sum() == 3 * 15874591.0  # :251-261
    res = kd.resample_nearest(swath_def, np.dstack((data, data, data)), area_def, 50000, segments=1)
    assert res.sum() == 3 * 15874591.0 and res.shape == (800, 800, 3)  # :263-274
    cdata = np.fromfunction(lambda y, x: y + complex("j") * x, (50, 10), dtype=np.complex128)
    res = kd.resample_nearest(swath_def, cdata.ravel(), area_def, 50000, segments=1)
    assert np.issubdtype(res.dtype, np.complex128)
    assert res.sum().real == 3530219.0 and res.sum().imag == 688723.0  # :127-137
    res = kd.resample_gauss(swath_def, data.ravel(), area_def, 50000, 25000, fill_value=-1, segments=1)
    assert res.sum() == pytest.approx(15387753.9852, abs=1e-3)  # :276-285
    remap = kd.resample_nearest(area_def, kd.resample_nearest(swath_def, data.ravel(), area_def, 50000).ravel(),
                                swath_def, 5000, segments=1)
    assert remap.sum() == 22275.0  # :227-238
    # 1-D swath target from an area source  :157-167
    data800 = np.fromfunction(lambda x, y: x * y, (800, 800))
    sw1d = geometry.SwathDefinition(lons=np.fromfunction(lambda x: 3 + x / 100., (500,)),
                                    lats=np.fromfunction(lambda x: 75 - x / 10., (500,)))
    res = kd.resample_nearest(area_def, data800.ravel(), sw1d, 50000, segments=1)
    assert res.shape == (500,) and res.sum() == 35821299.0
    # no overlap  :169-214
    lons_e = np.fromfunction(lambda y, x: 165 + x, (50, 10))
    swath_e = geometry.SwathDefinition(lons=lons_e, lats=lats)
    assert kd.resample_nearest(swath_e, data.ravel(), area_def, 50000, segments=1).sum() == 0
    res = kd.resample_nearest(swath_e, data_multi, area_def, 50000, segments=1, fill_value=None)
    assert res.shape == (800, 800, 3) and res.mask.all()


def test_reference_gauss_5000x100(kd, geometry):
    area_def = stere_area()
    data = np.fromfunction(lambda y, x: (y + x) * 10 ** -5, (5000, 100))
    lons = np.fromfunction(lambda y, x: 3 + (10.0 / 100) * x, (5000, 100))
    lats = np.fromfunction(lambda y, x: 75 - (50.0 / 5000) * y, (5000, 100))
    swath_def = geometry.SwathDefinition(lons=lons, lats=lats)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        res = kd.resample_gauss(swath_def, data.ravel(), area_def, 50000, 25000, segments=1)
        assert len(w) == 1 and 'Possible more' in str(w[0].message)
    assert res.sum() == pytest.approx(4872.8100353517921, rel=1e-9)  # test_kd_tree.py:287-317
    data6 = np.fromfunction(lambda y, x: (y + x) * 10 ** -6, (5000, 100))
    data_multi = np.column_stack((data6.ravel(), data6.ravel(), data6.ravel()))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = kd.resample_gauss(swath_def, data_multi, area_def, 50000, [25000, 15000, 10000], segments=1)
    assert res.sum() == pytest.approx(1461.8429990248171, rel=1e-9)  # :319-335
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res, stddev, counts = kd.resample_gauss(swath_def, data_multi, area_def, 50000, [25000, 15000, 10000],
                                                segments=1, with_uncert=True)
    assert res.sum() == pytest.approx(1461.8429990248171, rel=1e-9)  # :337-370
    for i, e in enumerate([0.44621800779801657, 0.44363137712896705, 0.43861019464274459]):
        assert stddev[:, :, i].sum() == pytest.approx(e, rel=1e-6)
    assert counts.sum() == 4934802.0

    def wf1(dist):
        return 1 - dist / 100000.0

    def wf2(dist):
        return 1

    def wf3(dist):
        return np.cos(dist) ** 2
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = kd.resample_custom(swath_def, data.ravel(), area_def, 50000, wf1, segments=1)
    assert res.sum() == pytest.approx(4872.8100347930776, rel=1e-9)  # :428-459
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = kd.resample_custom(swath_def, data_multi, area_def, 50000, [wf1, wf2, wf3], segments=1)
    assert res.sum() == pytest.approx(1461.8428378742638, rel=1e-9)  # :461-486


@pytest.mark.parametrize("dtype,rtol", [(np.float64, 1e-6), (np.float32, 1e-4)])
def test_fused_paths_against_oracle(kd, geometry, dtype, rtol):
    lon, lat = jittered_swath(260, 200, dtype, seed=8)
    rng = np.random.default_rng(9)
    data = (rng.standard_normal((260 * 200, 5)) * 10 + 280).astype(np.float32)
    data[rng.random(data.shape) < 0.01] = np.nan
    swath = geometry.SwathDefinition(lon, lat)
    area = laea_area(210, 190)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = kd.resample_nearest(swath, data, area, 50000, fill_value=-5)
        want = kd_tree_ref.resample_nearest(swath, data, area, 50000, fill_value=-5)
    assert got.dtype == want.dtype and got.shape == want.shape
    neq = ~((got == want) | (np.isnan(got) & np.isnan(want)))
    assert neq.mean() <= (0 if dtype == np.float64 else 1e-4)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got, gs, gc = kd.resample_gauss(swath, data[:, :3].astype(dtype), area, 50000, [25000., 15000., 20000.],
                                        neighbours=8, fill_value=None, with_uncert=True)
        want, ws, wc = kd_tree_ref.resample_gauss(swath, data[:, :3].astype(dtype), area, 50000,
                                                  [25000., 15000., 20000.], neighbours=8, fill_value=None,
                                                  with_uncert=True)
    assert np.array_equal(np.ma.getmaskarray(got), np.ma.getmaskarray(want))
    g, w = np.ma.filled(got, 0), np.ma.filled(want, 0)
    ok = np.isclose(g, w, rtol=rtol, atol=1e-6, equal_nan=True)
    assert (~ok).mean() <= (0 if dtype == np.float64 else 1e-4)
    assert np.array_equal(np.ma.filled(gc, -1), np.ma.filled(wc, -1)) or dtype == np.float32
    gsf, wsf = np.ma.filled(gs, -1), np.ma.filled(ws, -1)
    ok = np.isclose(gsf, wsf, rtol=max(rtol, 1e-5), atol=1e-5, equal_nan=True)
    assert (~ok).mean() <= (0 if dtype == np.float64 else 1e-3)


def test_masked_data_and_sample_from_info(kd, geometry):
    lon, lat = jittered_swath(120, 100, np.float64, seed=10)
    rng = np.random.default_rng(11)
    vals = rng.standard_normal((120 * 100, 2))
    mask = rng.random(vals.shape) < 0.2
    data = np.ma.array(vals, mask=mask)
    swath = geometry.SwathDefinition(lon, lat)
    area = laea_area(90, 80)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = kd.resample_nearest(swath, data, area, 50000, fill_value=None)
        want = kd_tree_ref.resample_nearest(swath, data, area, 50000, fill_value=None)
        assert np.array_equal(got.mask, want.mask) and np.array_equal(got.filled(0), want.filled(0))
        got = kd.resample_gauss(swath, data, area, 50000, [20000., 30000.], neighbours=4, fill_value=None)
        want = kd_tree_ref.resample_gauss(swath, data, area, 50000, [20000., 30000.], neighbours=4, fill_value=None)
        assert np.array_equal(got.mask, want.mask)
        np.testing.assert_allclose(got.filled(0), want.filled(0), rtol=1e-6, atol=1e-9)
        info = kd.get_neighbour_info(swath, area, 50000, neighbours=4)
        oinfo = kd_tree_ref.get_neighbour_info(swath, area, 50000, neighbours=4)
    _compare_info(info, oinfo, 1e-6)
    wfs = [lambda d: 1 - d / 1e5, lambda d: np.cos(d / 3e4) ** 2]
    got = kd.get_sample_from_neighbour_info('custom', area.shape, vals, *info[:3], distance_array=info[3],
                                            weight_funcs=wfs, fill_value=-1, with_uncert=True)
    want = kd_tree_ref.get_sample_from_neighbour_info('custom', area.shape, vals, *oinfo[:3], distance_array=oinfo[3],
                                                      weight_funcs=wfs, fill_value=-1, with_uncert=True)
    for g, w in zip(got, want):
        np.testing.assert_allclose(np.ma.filled(g, -7), np.ma.filled(w, -7), rtol=1e-6, atol=1e-9)
    info1 = kd.get_neighbour_info(swath, area, 50000, neighbours=1)
    for arr in (vals, vals[:, 0], vals.astype(np.float32), (vals * 100).astype(np.uint16), (vals * 100).astype(np.int64)):
        got = kd.get_sample_from_neighbour_info('nn', area.shape, arr, *info1[:3], fill_value=7)
        want = kd_tree_ref.get_sample_from_neighbour_info('nn', area.shape, arr, *info1[:3], fill_value=7)
        assert got.dtype == want.dtype and np.array_equal(got, want)


def test_c_abi_error_reporting(cuda_lib):
    from pyresample_b200 import _cabi
    rc = cuda_lib.prs_index_build(None, None, 0, 0, None, None, None, None, 0, 1, 1.0, None, 0,
                                  ctypes.byref(ctypes.c_void_p()), None)
    assert rc != 0
    with pytest.raises(_cabi.PrsError):
        _cabi.check(rc)
