"""Edge cases of the GPU index: cube-face edges and corners, poles, the dateline, large radii, every top-k width,
duplicates, invalid coordinates, excluded points, tiny and degenerate inputs - always against the exhaustive C
checker or the oracle, indices bit for bit."""
import warnings

import numpy as np
import pytest

from oracle import kd_tree_ref
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


def _brute_info(lon, lat, tlon, tlat, k, radius, dtype):
    src = kd_tree_ref.transform_lonlats(lon.astype(dtype), lat.astype(dtype))
    tgt = kd_tree_ref.transform_lonlats(tlon.astype(dtype), tlat.astype(dtype))
    return brute_query(src, tgt, k, radius)


def _check_against_brute(kd, geometry, lon, lat, tlon, tlat, k, radius, dtype=np.float32):
    """float32: ECEF coordinates are bit-identical to numpy's, so indices and distances must match exactly."""
    swath = geometry.SwathDefinition(lon.astype(dtype), lat.astype(dtype))
    target = geometry.SwathDefinition(tlon.astype(dtype), tlat.astype(dtype))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        vii, voi, ia, da = kd.get_neighbour_info(swath, target, radius, neighbours=k, reduce_data=False)
    assert vii.all() and voi.all()
    bd, bi = _brute_info(lon, lat, tlon, tlat, k, radius, dtype)
    assert np.array_equal(ia, bi), (np.argwhere(ia != bi)[:5], k)
    if dtype == np.float32:
        assert np.array_equal(da, bd)
    else:  # CUDA and libm sin/cos differ by a few ulp in float64
        assert np.array_equal(np.isinf(da), np.isinf(bd))
        np.testing.assert_allclose(da[np.isfinite(bd)], bd[np.isfinite(bd)], rtol=1e-9, atol=1e-6)


# the 8 cube corners sit at lon = +-45/+-135, lat = +-35.264; edges run along lon = +-45/+-135 and between corners
CORNERS = [(lo, la) for lo in (45., 135., -45., -135.) for la in (35.26438968, -35.26438968)]


@pytest.mark.parametrize("k", [1, 3, 8])
def test_cube_corners_and_edges(kd, geometry, k):
    rng = np.random.default_rng(50)
    lons, lats, tlons, tlats = [], [], [], []
    for lo, la in CORNERS:
        lons.append(lo + rng.uniform(-1.5, 1.5, 1500))
        lats.append(la + rng.uniform(-1.5, 1.5, 1500))
        tlons.append(lo + rng.uniform(-1.2, 1.2, 600))
        tlats.append(la + rng.uniform(-1.2, 1.2, 600))
    for lo in (45., -135.):  # along two vertical cube edges
        lons.append(lo + rng.uniform(-0.3, 0.3, 3000))
        lats.append(rng.uniform(-30, 30, 3000))
        tlons.append(lo + rng.uniform(-0.3, 0.3, 1000))
        tlats.append(rng.uniform(-30, 30, 1000))
    lon, lat = np.concatenate(lons), np.concatenate(lats)
    tlon, tlat = np.concatenate(tlons), np.concatenate(tlats)
    for radius in (8000., 60000.):
        _check_against_brute(kd, geometry, lon, lat, tlon, tlat, k, radius)


@pytest.mark.parametrize("k", [1, 4])
def test_poles_and_dateline(kd, geometry, k):
    rng = np.random.default_rng(51)
    lon = np.concatenate([rng.uniform(-180, 180, 4000), rng.uniform(-180, 180, 4000),
                          rng.uniform(175, 180, 1500), rng.uniform(-180, -175, 1500), [180., -180., 0., 0.]])
    lat = np.concatenate([rng.uniform(86, 90, 4000), rng.uniform(-90, -86, 4000),
                          rng.uniform(-10, 10, 3000), [0., 0., 90., -90.]])
    tlon = np.concatenate([rng.uniform(-180, 180, 1500), rng.uniform(-180, 180, 1500), rng.uniform(178, 180, 500),
                           rng.uniform(-180, -178, 500), [180., -180., 13., -77.]])
    tlat = np.concatenate([rng.uniform(87, 90, 1500), rng.uniform(-90, -87, 1500), rng.uniform(-8, 8, 1000),
                           [0., 0., 90., -90.]])
    _check_against_brute(kd, geometry, lon, lat, tlon, tlat, k, 40000.)


@pytest.mark.parametrize("k", [1, 2, 5, 8, 13, 16, 27, 32])
def test_every_topk_width(kd, geometry, k):
    rng = np.random.default_rng(52)
    lon, lat = rng.uniform(0, 6, 5000), rng.uniform(50, 55, 5000)
    tlon, tlat = rng.uniform(-0.5, 6.5, 1200), rng.uniform(49.5, 55.5, 1200)
    _check_against_brute(kd, geometry, lon, lat, tlon, tlat, k, 25000.)
    _check_against_brute(kd, geometry, lon, lat, tlon, tlat, k, 25000., dtype=np.float64) if k in (1, 8) else None


def test_large_radius_sparse_source(kd, geometry):
    """Radius far larger than the cells and than the point spacing: block growth, phase 2 over several faces."""
    rng = np.random.default_rng(53)
    lon, lat = rng.uniform(-180, 180, 300), np.rad2deg(np.arcsin(rng.uniform(-1, 1, 300)))
    tlon, tlat = rng.uniform(-180, 180, 2000), np.rad2deg(np.arcsin(rng.uniform(-1, 1, 2000)))
    for radius in (500000., 2500000., 7000000.):
        for k in (1, 8):
            _check_against_brute(kd, geometry, lon, lat, tlon, tlat, k, radius)


def test_neighbours_more_than_points_and_tiny_inputs(kd, geometry):
    lon, lat = np.array([10.0, 10.001, 10.002]), np.array([50.0, 50.0, 50.001])
    tlon, tlat = np.array([10.0005, 40.0]), np.array([50.0, 50.0])
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        vii, voi, ia, da = kd.get_neighbour_info(geometry.SwathDefinition(lon, lat), geometry.SwathDefinition(tlon, tlat),
                                                 10000, neighbours=8, reduce_data=False)
        assert any("Searching" in str(x.message) for x in w)
    assert ia.shape == (2, 8) and ia.dtype == np.uint32
    assert set(ia[0, :3]) == {0, 1, 2} and (ia[0, 3:] == 3).all() and np.isinf(da[0, 3:]).all()
    assert (ia[1] == 3).all() and np.isinf(da[1]).all()
    # a single source point, a single target point
    vii, voi, ia, da = kd.get_neighbour_info(geometry.SwathDefinition(lon[:1], lat[:1]),
                                             geometry.SwathDefinition(tlon[:1], tlat[:1]), 10000, neighbours=1)
    assert ia.shape == (1,) and ia[0] == 0 and da[0] < 100


def test_invalid_coordinates_and_all_invalid(kd, geometry):
    rng = np.random.default_rng(54)
    lon, lat = rng.uniform(0, 5, 2000), rng.uniform(40, 45, 2000)
    lon[::7] = np.nan
    lat[3::11] = 91.0
    lon[5::13] = np.inf
    lon[1::17] = -180.0001
    tlon, tlat = rng.uniform(0, 5, 500), rng.uniform(40, 45, 500)
    tlon[::9] = np.nan
    tlat[4::10] = -90.5
    swath, target = geometry.SwathDefinition(lon, lat), geometry.SwathDefinition(tlon, tlat)
    for k in (1, 4):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            got = kd.get_neighbour_info(swath, target, 30000, neighbours=k)
            want = kd_tree_ref.get_neighbour_info(swath, target, 30000, neighbours=k)
        for g, w_ in zip(got[:3], want[:3]):
            assert np.array_equal(g, w_)
        np.testing.assert_allclose(got[3], want[3], rtol=1e-9)
    # every source point invalid -> the reference's empty info (kd_tree.py:553-563): int32 indices, distance 1.0
    bad = geometry.SwathDefinition(np.full(10, np.nan), np.full(10, 0.0))
    vii, voi, ia, da = kd.get_neighbour_info(bad, target, 30000, neighbours=1)
    ovii, ovoi, oia, oda = kd_tree_ref.get_neighbour_info(bad, target, 30000, neighbours=1)
    assert ia.dtype == oia.dtype == np.int32 and np.array_equal(ia, oia) and np.array_equal(da, oda)
    assert not vii.any() and voi.all()
    res = kd.resample_nearest(bad, np.arange(10.), target, 30000, fill_value=-3)
    assert np.array_equal(res, np.full(500, -3.0))
    with pytest.raises(ValueError):
        kd.resample_nearest(geometry.SwathDefinition(np.zeros(0), np.zeros(0)), np.zeros(0), target, 1000)


def test_radius_is_strict_and_zero_distance(kd, geometry):
    """A source point exactly at the query has distance 0; a candidate at d2 == dtype(r^2) is excluded."""
    lon, lat = np.array([10.0, 10.0, 11.0]), np.array([50.0, 50.0, 50.0])
    tgt = geometry.SwathDefinition(np.array([10.0, 10.5]), np.array([50.0, 50.0]))
    src = geometry.SwathDefinition(lon, lat)
    xyz = kd_tree_ref.transform_lonlats(lon, lat)
    txyz = kd_tree_ref.transform_lonlats(np.array([10.0, 10.5]), np.array([50.0, 50.0]))
    d = np.sqrt(((txyz[1] - xyz[0]) ** 2).sum())
    vii, voi, ia, da = kd.get_neighbour_info(src, tgt, float(d) * 1.01, neighbours=2, reduce_data=False)
    assert ia[0, 0] == 0 and ia[0, 1] == 1 and da[0, 0] == 0 and da[0, 1] == 0  # duplicates: lower index first
    bd, bi = brute_query(xyz, txyz, 2, float(d) * 1.01)
    assert np.array_equal(ia, bi)
    # radius slightly below the distance to every point: nothing is found for the second target
    vii, voi, ia, da = kd.get_neighbour_info(src, tgt, float(d) * 0.999, neighbours=1, reduce_data=False)
    bd, bi = brute_query(xyz, txyz, 1, float(d) * 0.999)
    assert np.array_equal(ia, bi)


def test_dense_duplicates_in_one_cell(kd, geometry):
    """Thousands of identical points fall into one cell; top-k must list the lowest indices."""
    lon = np.concatenate([np.full(3000, 12.5), np.random.default_rng(55).uniform(12, 13, 500)])
    lat = np.concatenate([np.full(3000, 42.25), np.random.default_rng(56).uniform(42, 42.5, 500)])
    tlon, tlat = np.array([12.5, 12.6]), np.array([42.25, 42.3])
    _check_against_brute(kd, geometry, lon, lat, tlon, tlat, 8, 20000., dtype=np.float64)


@pytest.mark.parametrize("min_run", [0, 1 << 40])
def test_pipelined_nearest_equals_plain(kd, geometry, monkeypatch, min_run):
    """The staged C-ABI calls (fill, then search) with overlapped transfers give exactly the one-call result."""
    rng = np.random.default_rng(57)
    lon = rng.uniform(-20, 40, 120000).astype(np.float32)
    lat = rng.uniform(10, 40, 120000).astype(np.float32)
    data = rng.normal(size=(120000, 3)).astype(np.float32)
    swath = geometry.SwathDefinition(lon, lat)
    area = geometry.AreaDefinition("g", "g", "g", "+proj=eqc +datum=WGS84", 1400, 611,
                                   (-10018754.17, -8000000.0, 10018754.17, 9000000.0))
    plain = kd.resample_nearest(swath, data, area, 30000, fill_value=-1)
    monkeypatch.setattr(kd, "_PIPELINE_MIN_BYTES", 0)
    monkeypatch.setattr(kd, "_PIPELINE_MIN_RUN_BYTES", min_run)
    piped = kd.resample_nearest(swath, data, area, 30000, fill_value=-1)
    assert piped.shape == plain.shape == (611, 1400, 3)
    assert np.array_equal(piped, plain)
    assert (plain != -1).any() and (plain[:8] == -1).all()
    want = kd_tree_ref.resample_nearest(swath, data, area, 30000, fill_value=-1)
    assert np.array_equal(piped, want)
    # masked result (fill_value=None) goes through the same path
    masked = kd.resample_nearest(swath, data[:, 0], area, 30000, fill_value=None)
    want_m = kd_tree_ref.resample_nearest(swath, data[:, 0], area, 30000, fill_value=None)
    assert np.array_equal(masked.mask, want_m.mask) and np.array_equal(masked.filled(0), want_m.filled(0))


def test_stacked_area_definition_as_target_and_source(kd, geometry):
    """A StackedAreaDefinition behaves like the explicit lon/lat grid of its stacked members (geometry.py:2966-2998)."""
    proj = "+proj=laea +lat_0=55 +lon_0=10 +a=6371228.0 +units=m"
    top = geometry.AreaDefinition("t", "t", "t", proj, 120, 40, (-600000, 0, 600000, 400000))
    far = geometry.AreaDefinition("f", "f", "f", proj, 120, 30, (-600000, -900000, 600000, -600000))
    st = geometry.StackedAreaDefinition(top, far)
    assert len(st.defs) == 2
    lons, lats = st.get_lonlats()
    want = [kd_tree_ref.get_lonlats(a) for a in (top, far)]
    np.testing.assert_allclose(lons, np.vstack([w[0] for w in want]), rtol=0, atol=1e-9)
    np.testing.assert_allclose(lats, np.vstack([w[1] for w in want]), rtol=0, atol=1e-9)
    part = st.get_lonlats(data_slice=(slice(35, 50), slice(None)))
    assert part[0].shape == (15, 120) and np.array_equal(part[0], lons[35:50])
    rng = np.random.default_rng(58)
    slon, slat = rng.uniform(-5, 25, 40000), rng.uniform(44, 60, 40000)
    data = rng.normal(size=40000)
    swath = geometry.SwathDefinition(slon, slat)
    res = kd.resample_nearest(swath, data, st, 20000, fill_value=None)
    grid = geometry.GridDefinition(np.vstack([w[0] for w in want]), np.vstack([w[1] for w in want]))
    ref = kd_tree_ref.resample_nearest(swath, data, grid, 20000, fill_value=None)
    assert res.shape == (70, 120)
    assert np.array_equal(res.mask, ref.mask) and np.array_equal(res.filled(0), ref.filled(0))
    # and as the source geometry (grid -> swath)
    src_data = rng.normal(size=(70, 120))
    tgt = geometry.SwathDefinition(rng.uniform(0, 20, 500), rng.uniform(48, 58, 500))
    back = kd.resample_nearest(st, src_data, tgt, 20000, fill_value=-9)
    ref_back = kd_tree_ref.resample_nearest(grid, src_data, tgt, 20000, fill_value=-9, reduce_data=False)
    assert np.array_equal(back, ref_back)


def test_device_neighbour_info_matches_get_sample(kd, geometry, tmp_path):
    """Neighbour info saved, reloaded and kept in HBM samples exactly like get_sample_from_neighbour_info."""
    from pyresample_b200 import neighbour_cache as nc
    rng = np.random.default_rng(59)
    lon, lat = rng.uniform(0, 20, 30000), rng.uniform(40, 60, 30000)
    lon[::50] = np.nan
    swath = geometry.SwathDefinition(lon, lat)
    area = geometry.AreaDefinition("a", "a", "a", "+proj=laea +lat_0=50 +lon_0=10 +a=6371228.0 +units=m", 300, 200,
                                   (-900000, -700000, 900000, 700000))
    info = kd.get_neighbour_info(swath, area, 30000, neighbours=1)
    path = tmp_path / "n.npz"
    nc.save_neighbour_info(path, *info)
    loaded = nc.load_neighbour_info(path)
    for a, b in zip(info, loaded):
        assert a.dtype == b.dtype and np.array_equal(a, b)
    dev = nc.DeviceNeighbourInfo.load(path)
    for data in (rng.normal(size=30000), rng.normal(size=(30000, 3)).astype(np.float32),
                 rng.integers(0, 200, 30000).astype(np.uint8)):
        for fill in (0, None):
            want = kd.get_sample_from_neighbour_info('nn', area.shape, data, *info[:3], fill_value=fill)
            got = dev.sample_nearest(data, area.shape, fill_value=fill)
            assert got.dtype == want.dtype and got.shape == want.shape
            assert np.array_equal(np.ma.getmaskarray(got), np.ma.getmaskarray(want))
            assert np.array_equal(np.ma.getdata(got), np.ma.getdata(want))
    ref = kd_tree_ref.resample_nearest(swath, np.arange(30000.), area, 30000, fill_value=-1)
    assert np.array_equal(dev.sample_nearest(np.arange(30000.), area.shape, fill_value=-1), ref)


def test_release_cached_memory_and_rebuild(kd, geometry, cuda_lib):
    """prs_release_cached_memory frees the build workspace / pool cache; the next build simply allocates again."""
    import torch
    rng = np.random.default_rng(60)
    lon, lat = rng.uniform(0, 10, 200000), rng.uniform(40, 50, 200000)
    swath = geometry.SwathDefinition(lon, lat)
    tgt = geometry.SwathDefinition(rng.uniform(0, 10, 1000), rng.uniform(40, 50, 1000))
    a = kd.get_neighbour_info(swath, tgt, 20000, neighbours=1)
    free0 = torch.cuda.mem_get_info()[0]
    assert cuda_lib.prs_release_cached_memory() == 0
    assert torch.cuda.mem_get_info()[0] >= free0
    b = kd.get_neighbour_info(swath, tgt, 20000, neighbours=1)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


@pytest.mark.parametrize("seed", range(12))
def test_randomised_sweep_against_brute_force(kd, geometry, seed):
    """Random source density, location, radius, k, dtype and target kind (swath or one of the area projections):
    indices must equal the exhaustive checker's bit for bit (area targets: via the oracle's restated lon/lats)."""
    from oracle import proj_ref
    rng = np.random.default_rng(1000 + seed)
    dtype = np.float32 if seed % 3 else np.float64
    lon0, lat0 = rng.uniform(-180, 180), rng.uniform(-80, 80)
    span = 10 ** rng.uniform(-0.5, 1.3)                 # 0.3 .. 20 degrees
    n = int(10 ** rng.uniform(2.5, 4.3))
    lon = ((lon0 + rng.uniform(-span, span, n) + 180) % 360) - 180
    lat = np.clip(lat0 + rng.uniform(-span, span, n), -90, 90)
    k = int(rng.choice([1, 1, 2, 4, 7, 8, 12, 16, 20, 32]))
    spacing_m = 2 * span * 111e3 / np.sqrt(n)
    radius = float(spacing_m * 10 ** rng.uniform(-0.3, 1.2))
    kind = seed % 4
    if kind == 0:
        tlon = ((lon0 + rng.uniform(-1.3 * span, 1.3 * span, 1500) + 180) % 360) - 180
        tlat = np.clip(lat0 + rng.uniform(-1.3 * span, 1.3 * span, 1500), -90, 90)
    else:
        half = 1.3 * span * 111e3
        proj = {1: "+proj=laea +lat_0=%.3f +lon_0=%.3f +a=6371228.0 +units=m" % (lat0, lon0),
                2: "+proj=stere +lat_0=%.3f +lon_0=%.3f +ellps=WGS84" % (lat0, lon0),
                3: "+proj=eqc +lon_0=%.3f +datum=WGS84" % lon0}[kind]
        y0 = 0.0 if kind != 3 else lat0 * 111e3
        area = geometry.AreaDefinition("r", "r", "r", proj, 41, 37, (-half, y0 - half, half, y0 + half), dtype=dtype)
        tlon, tlat = proj_ref.area_lonlats(area.proj_dict, area.area_extent, area.width, area.height, dtype=dtype)
        tlon, tlat = tlon.ravel(), tlat.ravel()
        ok = np.isfinite(tlon) & np.isfinite(tlat) & (np.abs(tlat) <= 90) & (np.abs(tlon) <= 180)
        tlon, tlat = tlon[ok], tlat[ok]
    _check_against_brute(kd, geometry, lon, lat, np.asarray(tlon, np.float64), np.asarray(tlat, np.float64), k,
                         radius, dtype=dtype)
    if kind != 0:
        # the same target generated in-kernel from the AreaDefinition must agree with the explicit lon/lat target
        swath = geometry.SwathDefinition(lon.astype(dtype), lat.astype(dtype))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            a = kd.get_neighbour_info(swath, area, radius, neighbours=k, reduce_data=False)
            want = kd_tree_ref.get_neighbour_info(swath, area, radius, neighbours=k, reduce_data=False)
        assert np.array_equal(a[1], want[1]) and np.array_equal(a[2], want[2])


def test_staged_upload_of_pageable_arrays(kd, monkeypatch):
    """Large ordinary numpy arrays go up through pinned staging buffers in chunks; the tensor must equal the array
    for every size around the chunk / buffer-count boundaries and for every dtype width."""
    monkeypatch.setattr(kd, "_STAGED_MIN_BYTES", 1)
    monkeypatch.setattr(kd, "_STAGE_CHUNK", 1 << 16)
    monkeypatch.setitem(kd._stage, "bufs", None)
    rng = np.random.default_rng(61)
    chunk = 1 << 16
    for nbytes in (1, chunk - 1, chunk, chunk + 1, 6 * chunk, 6 * chunk + 5, 13 * chunk + 12345):
        a = rng.integers(0, 255, nbytes, dtype=np.uint8)
        assert np.array_equal(kd._dev(a).cpu().numpy(), a)
    for dt in (np.float32, np.float64, np.int16, np.bool_):
        a = (rng.random((1237, 219)) * 100).astype(dt)
        t = kd._dev(a)
        assert tuple(t.shape) == a.shape
        assert np.array_equal(t.cpu().numpy().astype(dt), a)
    monkeypatch.setitem(kd._stage, "bufs", None)  # full-size buffers again for whoever comes next


def test_custom_weights_on_sparse_coverage(kd, geometry):
    """Python weight functions on a grid that is mostly outside the swath: only rows with a neighbour are evaluated
    on the host, the rest must still come out as fill / NaN / 0 like in the reference."""
    rng = np.random.default_rng(62)
    lon, lat = rng.uniform(8, 12, 20000), rng.uniform(48, 52, 20000)
    swath = geometry.SwathDefinition(lon, lat)
    area = geometry.AreaDefinition("w", "w", "w", "+proj=eqc +datum=WGS84", 300, 200,
                                   (-3000000.0, 2000000.0, 5000000.0, 8000000.0))
    data = rng.normal(size=20000)
    data2 = np.column_stack([data, data * 2 + 1])

    def wf(d):
        return 1 - d / 60000.0

    def wf2(d):
        return np.exp(-d ** 2 / 30000.0 ** 2)

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = kd.resample_custom(swath, data, area, 60000, wf, neighbours=6, fill_value=-7)
        want = kd_tree_ref.resample_custom(swath, data, area, 60000, wf, neighbours=6, fill_value=-7)
        assert (want == -7).mean() > 0.8
        np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-9)
        got = kd.resample_custom(swath, data2, area, 60000, [wf, wf2], neighbours=6, fill_value=None, with_uncert=True)
        want = kd_tree_ref.resample_custom(swath, data2, area, 60000, [wf, wf2], neighbours=6, fill_value=None,
                                           with_uncert=True)
        for g, w in zip(got, want):
            assert np.array_equal(np.ma.getmaskarray(g), np.ma.getmaskarray(w))
            np.testing.assert_allclose(np.ma.getdata(g)[~np.ma.getmaskarray(g)],
                                       np.ma.getdata(w)[~np.ma.getmaskarray(w)], rtol=1e-6, atol=1e-9)
        # a weight function that is not finite at 1: no shortcut, same answer as the reference
        def bad(d):
            return 1.0 / (d - 1.0)
        got = kd.resample_custom(swath, data, area, 60000, bad, neighbours=3, fill_value=-7)
        want = kd_tree_ref.resample_custom(swath, data, area, 60000, bad, neighbours=3, fill_value=-7)
        np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-9, equal_nan=True)


def test_concurrent_calls_from_threads(kd, geometry, monkeypatch):
    """SURVEY 8b threading: calls from several Python threads (each on its own CUDA stream) must not disturb each
    other - the build workspace, the staging buffers and the side streams are shared process-wide."""
    import threading

    import torch
    monkeypatch.setattr(kd, "_PIPELINE_MIN_BYTES", 0)      # take the pipelined path
    monkeypatch.setattr(kd, "_STAGED_MIN_BYTES", 1 << 16)  # and the staged upload
    area = geometry.AreaDefinition("t", "t", "t", "+proj=eqc +datum=WGS84", 700, 400,
                                   (-2000000.0, 3000000.0, 4000000.0, 7500000.0))
    jobs = []
    for s in range(4):
        rng = np.random.default_rng(70 + s)
        lon, lat = rng.uniform(-10 + 5 * s, 20 + 5 * s, 60000), rng.uniform(30, 60, 60000)
        data = rng.normal(size=(60000, 2)).astype(np.float32)
        jobs.append((geometry.SwathDefinition(lon.astype(np.float32), lat.astype(np.float32)), data))

    def work(job):
        swath, data = job
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            a = kd.resample_nearest(swath, data, area, 25000, fill_value=-1)
            b = kd.get_neighbour_info(swath, area, 25000, neighbours=4)
        return a, b

    expected = [work(j) for j in jobs]
    results = [None] * len(jobs)
    errors = []

    def runner(i):
        try:
            with torch.cuda.stream(torch.cuda.Stream()):
                for _ in range(3):
                    results[i] = work(jobs[i])
        except Exception as exc:  # pragma: no cover - reported below
            errors.append(exc)

    threads = [threading.Thread(target=runner, args=(i,)) for i in range(len(jobs))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for (a, b), (ea, eb) in zip(results, expected):
        assert np.array_equal(a, ea)
        for x, y in zip(b, eb):
            assert np.array_equal(x, y)


@pytest.mark.parametrize("dtype,channels", [(np.uint8, 1), (np.int16, 3), (np.float64, 7), (np.float64, 9),
                                            (np.complex64, 1)])
def test_pipelined_nearest_row_layouts(kd, geometry, monkeypatch, dtype, channels):
    """Row sizes that take the other store paths of the fill / search stages (1, 6, 56, 72 and 8 bytes per pixel)."""
    rng = np.random.default_rng(63)
    lon = rng.uniform(-15, 35, 80000).astype(np.float32)
    lat = rng.uniform(20, 45, 80000).astype(np.float32)
    shape = (80000,) if channels == 1 else (80000, channels)
    if np.issubdtype(dtype, np.complexfloating):
        data = (rng.normal(size=shape) + 1j * rng.normal(size=shape)).astype(dtype)
    elif np.issubdtype(dtype, np.integer):
        data = rng.integers(0, 100, size=shape).astype(dtype)
    else:
        data = rng.normal(size=shape).astype(dtype)
    swath = geometry.SwathDefinition(lon, lat)
    area = geometry.AreaDefinition("g", "g", "g", "+proj=eqc +datum=WGS84", 1003, 517,
                                   (-10018754.17, -8000000.0, 10018754.17, 9000000.0))
    plain = kd.resample_nearest(swath, data, area, 30000, fill_value=7)
    monkeypatch.setattr(kd, "_PIPELINE_MIN_BYTES", 0)
    monkeypatch.setattr(kd, "_PIPELINE_MIN_RUN_BYTES", 0)
    piped = kd.resample_nearest(swath, data, area, 30000, fill_value=7)
    assert piped.dtype == plain.dtype == data.dtype and piped.shape == plain.shape
    assert np.array_equal(piped, plain)
    want = kd_tree_ref.resample_nearest(swath, data, area, 30000, fill_value=7)
    assert np.array_equal(piped, want)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("edge_lon", [45.0, 135.0, -45.0, -135.0])
def test_source_ends_exactly_on_a_cube_edge(kd, geometry, dtype, edge_lon):
    """A regional lon/lat grid whose border column lies EXACTLY on a cube edge (|x| == |y| up to the last bit) with
    nothing beyond it: the index build sees that column on one face in its approximate selection pass and possibly on
    the other face in its exact pass.  Every point of the column must still be found from both sides of the edge."""
    step = 0.05 if edge_lon > 0 else -0.05
    lon1 = edge_lon - step * np.arange(240)[::-1]  # ... up to and including the edge, nothing beyond
    lat1 = np.arange(-30.0, 30.0, 0.05)
    lon, lat = np.meshgrid(lon1, lat1)
    rng = np.random.default_rng(77)
    lat = lat + rng.uniform(-0.01, 0.01, lat.shape)  # tie-free, but the border column keeps its exact longitude
    tlon = edge_lon + rng.uniform(-0.12, 0.12, 4000)
    tlat = rng.uniform(-29.5, 29.5, 4000)
    for k, radius in ((1, 9000.0), (4, 20000.0)):
        _check_against_brute(kd, geometry, lon.ravel(), lat.ravel(), tlon, tlat, k, radius, dtype)
    # the same with ONLY the border column as the source (one face's table may be a single strip of cells)
    lon_b = np.full(lat1.size, edge_lon)
    _check_against_brute(kd, geometry, lon_b, lat1 + 0.003, tlon, tlat, 2, 15000.0, dtype)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_source_ends_on_a_horizontal_cube_edge(kd, geometry, dtype):
    """Border row along the edge between an equatorial face and the north face (|z| == |x| at lon = 0)."""
    rng = np.random.default_rng(78)
    lon1 = np.arange(-20.0, 20.0, 0.05)
    lat_edge = 45.0  # at lon = 0: z == x
    lat1 = lat_edge - 0.05 * np.arange(200)[::-1]
    lon, lat = np.meshgrid(lon1, lat1)
    lon = lon + rng.uniform(-0.01, 0.01, lon.shape)
    tlon = rng.uniform(-1.0, 1.0, 3000)
    tlat = lat_edge + rng.uniform(-0.12, 0.12, 3000)
    _check_against_brute(kd, geometry, lon.ravel(), lat.ravel(), tlon, tlat, 3, 12000.0, dtype)


@pytest.mark.parametrize("proj,w,h,ext", [
    ("+proj=geos +h=35785831 +lon_0=0 +a=6378169 +b=6356583.8", 464, 464, (-5570248.4773, -5567248.0742, 5567248.0742, 5570248.4773)),
    ("+proj=stere +lat_0=90 +lat_ts=70 +lon_0=-45 +a=6378273 +b=6356889.449", 400, 380, (-3000000., -3000000., 3000000., 3000000.)),
    ("+proj=laea +lat_0=10 +lon_0=20 +ellps=GRS80", 500, 300, (-2500000., -1500000., 2500000., 1500000.)),
])
def test_tiles_culled_from_five_probe_pixels(kd, geometry, proj, w, h, ext):
    """Areas with a per-pixel inverse projection: tiles are first bounded from their corner and centre pixels only.
    A narrow band of source points crossing the grid diagonally makes most tiles dead and puts the band's edge
    through many others - a tile culled by mistake would lose neighbours.  Includes the geos grid whose corners are
    off the Earth disc (invalid pixels)."""
    area = geometry.AreaDefinition("a", "a", "a", proj, w, h, ext)
    olon, olat = area.get_lonlats()
    fin = np.isfinite(olon)
    rng = np.random.default_rng(81)
    # points along the grid's diagonal, +-40 km wide
    rows = np.linspace(0, h - 1, 30000).astype(int)
    cols = np.linspace(0, w - 1, 30000).astype(int)
    ok = fin[rows, cols]
    lon = olon[rows, cols][ok] + rng.uniform(-0.35, 0.35, ok.sum())
    lat = np.clip(olat[rows, cols][ok] + rng.uniform(-0.35, 0.35, ok.sum()), -89.9, 89.9)
    swath = geometry.SwathDefinition(lon, lat)
    for radius, k in ((4000., 1), (25000., 4)):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            got = kd.get_neighbour_info(swath, area, radius, neighbours=k)
            want = kd_tree_ref.get_neighbour_info(swath, area, radius, neighbours=k)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
        ga, wa = got[2].reshape(len(got[2]), -1), want[2].reshape(len(want[2]), -1)
        gd, wd = got[3].reshape(ga.shape), want[3].reshape(wa.shape)
        diff = ga != wa
        # float64 coordinates differ by a few ulp between CUDA and libm: only exact-tie-scale flips are possible
        assert not np.any(diff & ~np.isclose(gd, wd, rtol=1e-9, atol=1e-6)), (int(diff.sum()), radius, k)
        assert np.array_equal(np.isinf(gd), np.isinf(wd))
        assert 0 < np.isfinite(wd[:, 0]).mean() < 0.5
