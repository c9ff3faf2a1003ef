"""CPU-only tests of the host-side logic: PROJ-string handling, geometry containers, row sharding and the
boundary analysis of reduce_data (compared with the oracle's statement-for-statement restatement)."""
import ctypes
import math

import numpy as np
import pytest

from oracle import kd_tree_ref
from pyresample_b200 import distributed, geometry
from pyresample_b200.proj import PROJ_EQC, PROJ_GEOS, PROJ_LAEA, PROJ_LONGLAT, PROJ_STERE, ProjSpec, proj_to_dict


def test_proj_string_and_dict_agree():
    a = ProjSpec("+proj=stere +lat_0=90 +lat_ts=70 +lon_0=-45 +a=6378273 +b=6356889.449").cstruct
    b = ProjSpec({'proj': 'stere', 'lat_0': 90, 'lat_ts': '70', 'lon_0': -45.0, 'a': 6378273, 'b': '6356889.449'}).cstruct
    assert bytes(a) == bytes(b)
    assert a.kind == PROJ_STERE and a.ellipsoidal == 1
    assert abs(a.es - (1 - 6356889.449 ** 2 / 6378273 ** 2)) < 1e-15
    assert proj_to_dict("+proj=laea +lat_0=52 +units=m +no_defs")["no_defs"] is True


@pytest.mark.parametrize("text,kind,ell", [
    ("+proj=longlat +datum=WGS84", PROJ_LONGLAT, 1), ("+proj=eqc +datum=WGS84", PROJ_EQC, 0),
    ("+proj=laea +lat_0=90 +lon_0=0 +a=6371228.0 +units=m", PROJ_LAEA, 0),
    ("+proj=geos +h=35786023 +lon_0=-75 +sweep=x +ellps=GRS80", PROJ_GEOS, 1), ("EPSG:4326", PROJ_LONGLAT, 1)])
def test_proj_kinds(text, kind, ell):
    s = ProjSpec(text).cstruct
    assert s.kind == kind and s.ellipsoidal == ell


def test_proj_ellipsoid_matches_oracle_parser():
    from oracle import proj_ref
    for text in ("+proj=stere +ellps=WGS84 +lat_0=-90", "+proj=laea +R=6371000", "+proj=merc +a=6378137 +rf=298.257223563",
                 "+proj=geos +h=35785831 +a=6378169 +b=6356583.8", "+proj=laea +ellps=GRS80 +lat_0=52",
                 "+proj=stere +datum=WGS84 +lat_0=90", "+proj=eqc +a=6371228"):
        a, es = proj_ref.ellipsoid(proj_ref.parse_proj(text))
        s = ProjSpec(text).cstruct
        if s.kind != PROJ_EQC:
            assert s.a == a and s.es == pytest.approx(es, abs=1e-18), text


def test_unsupported_projection_raises():
    with pytest.raises(NotImplementedError):
        ProjSpec("+proj=lcc +lat_1=30 +lat_2=60")
    with pytest.raises(NotImplementedError):
        ProjSpec("EPSG:3857")


def test_area_definition_attributes():
    # pyresample/geometry.py:1577-1619 and test_geometry/test_area.py:666-698 (projection vectors)
    area = geometry.AreaDefinition('areaD', 'Europe', 'areaD', {'proj': 'stere', 'a': 6378144.0, 'b': 6356759.0,
                                                               'lat_0': 50, 'lat_ts': 50, 'lon_0': 8}, 800, 800,
                                   [-1370912.72, -909968.64000000001, 1029087.28, 1490031.3600000001])
    assert area.shape == (800, 800) and area.size == 640000 and area.ndim == 2
    assert area.pixel_size_x == pytest.approx(3000.0) and area.pixel_size_y == pytest.approx(3000.0)
    x, y = area.get_proj_vectors()
    assert x[0] == pytest.approx(-1370912.72 + 1500.0) and y[0] == pytest.approx(1490031.36 - 1500.0)
    xx, yy = area.get_proj_coords(data_slice=(slice(2, 5), slice(10, 20)))
    assert xx.shape == (3, 10) and np.array_equal(xx[0], x[10:20]) and np.array_equal(yy[:, 0], y[2:5])
    x32, _ = area.get_proj_vectors(dtype=np.float32)
    assert x32.dtype == np.float32
    assert np.array_equal(x32, np.arange(800, dtype=np.float32) * 3000.0 + float(area.pixel_upper_left[0]))


def test_coordinate_definitions():
    lons, lats = np.zeros((4, 5), dtype=np.float32), np.ones((4, 5), dtype=np.float32)
    sw = geometry.SwathDefinition(lons, lats)
    assert sw.shape == (4, 5) and sw.size == 20 and sw.ndim == 2 and sw.dtype == np.float32
    with pytest.raises(ValueError):
        geometry.SwathDefinition(lons, np.ones((4, 6), dtype=np.float32))
    with pytest.raises(ValueError):
        geometry.GridDefinition(np.zeros(5), np.zeros(5))
    assert geometry.is_coord(sw) and not geometry.is_griddish(sw) and geometry.is_swath(sw)
    grid = geometry.GridDefinition(lons, lats)
    assert geometry.is_griddish(grid) and geometry.is_coord(grid)
    b_lon, b_lat = sw.get_boundary_lonlats()
    assert b_lon.side1.shape == (5,) and b_lon.side2.shape == (4,)


def test_get_slice_and_shard_rows_match_reference_segments():
    # geometry._get_slice geometry.py:3052-3068
    assert [s[0] for s in geometry._get_slice(3, (10, 4))] == [slice(0, 4), slice(4, 8), slice(8, 10)]
    assert list(geometry._get_slice(2, (5,))) == [slice(0, 3), slice(3, 5)]
    assert list(kd_tree_ref.get_slice(3, (10, 4))) == list(geometry._get_slice(3, (10, 4)))
    for h in (1, 7, 8, 3600, 10000):
        for w in (1, 2, 3, 8):
            plan = distributed.shard_plan(h, w)
            assert sum(n for _, n in plan) == h
            assert all(plan[i][0] + plan[i][1] == plan[i + 1][0] for i in range(w - 1) if plan[i + 1][1])
            assert max(n for _, n in plan) == math.ceil(h / w)


@pytest.mark.parametrize("seed", range(6))
def test_reduce_box_matches_oracle_on_lonlat_grids(seed):
    """kd_tree._reduce_box (host part) + the comparison rules == data_reduce._get_valid_index."""
    from pyresample_b200 import kd_tree
    from pyresample_b200._cabi import (PRS_REDUCE_ALL, PRS_REDUCE_BOX, PRS_REDUCE_BOX_DATELINE, PRS_REDUCE_LAT_GE,
                                       PRS_REDUCE_LAT_LE)
    rng = np.random.default_rng(seed)
    lon0 = [(-30, 40), (150, 220), (-180, 180), (-10, 10), (0, 30), (100, 179)][seed]
    lat0 = [(30, 60), (-20, 20), (-90, 90), (-5, 5), (0, 40), (-89, -60)][seed]
    glon, glat = np.meshgrid(np.linspace(lon0[0], lon0[1], 40), np.linspace(lat0[1], lat0[0], 30))
    glon = ((glon + 180) % 360) - 180
    dtype = np.float32 if seed % 2 else np.float64
    grid = geometry.GridDefinition(glon.astype(dtype), glat.astype(dtype))
    lons = rng.uniform(-180, 180, 20000)
    lats = rng.uniform(-90, 90, 20000)
    radius = 50000
    box = kd_tree._reduce_box(grid, radius)
    blons, blats = kd_tree_ref.get_boundary_lonlats(grid)
    want = kd_tree_ref.get_valid_index_from_lonlat_boundaries(blons, blats, lons, lats, radius)
    if box.kind == PRS_REDUCE_ALL:
        got = np.ones(lons.size, dtype=bool)
    elif box.kind == PRS_REDUCE_LAT_GE:
        got = lats >= box.lat_min
    elif box.kind == PRS_REDUCE_LAT_LE:
        got = lats <= box.lat_max
    elif box.kind == PRS_REDUCE_BOX:
        got = (lats >= box.lat_min) & (lats <= box.lat_max) & (lons >= box.lon_min) & (lons <= box.lon_max)
    else:
        assert box.kind == PRS_REDUCE_BOX_DATELINE
        got = (lats >= box.lat_min) & (lats <= box.lat_max) & \
            (((lons >= box.lon_min) & (lons <= 180)) | ((lons <= box.lon_max) & (lons >= -180)))
    assert np.array_equal(got, np.asarray(want, dtype=bool))


def test_ctypes_target_struct_roundtrip():
    from pyresample_b200._cabi import prs_area_grid, prs_target
    tg = prs_target()
    tg.grid = prs_area_grid(7200, 3600, 5566.0, 5566.0, -2e7, 1e7)
    tg.row0, tg.nrows, tg.n = 450, 450, 450 * 7200
    assert tg.grid.width == 7200 and ctypes.sizeof(tg) % 8 == 0


def test_band_runs_of_the_pipelined_path():
    """Tile-row flags -> runs of rows to read back early (finished) or late (live); short finished runs merge."""
    from pyresample_b200.kd_tree import _band_runs
    live = np.array([0, 0, 0, 1, 1, 0, 1, 0, 0, 0], dtype=np.uint8)
    runs = _band_runs(live, 8, 77, 100, 0)
    assert runs == [(0, 24, False), (24, 40, True), (40, 48, False), (48, 56, True), (56, 77, False)]
    assert sum(r1 - r0 for r0, r1, _ in runs) == 77
    # a finished run of one band (8 rows * 100 bytes) is below a 1000-byte minimum -> merged with its neighbours
    runs = _band_runs(live, 8, 77, 100, 1000)
    assert runs == [(0, 24, False), (24, 56, True), (56, 77, False)]
    assert _band_runs(np.zeros(4, np.uint8), 8, 32, 10, 0) == [(0, 32, False)]
    assert _band_runs(np.ones(1, np.uint8), 8, 5, 10, 0) == [(0, 5, True)]


def test_stacked_area_definition_append_rules():
    """StackedAreaDefinition (pyresample/geometry.py:2923-2964): contiguous members merge, others stay apart."""
    proj = "+proj=laea +lat_0=60 +lon_0=10 +a=6371228.0 +units=m"
    top = geometry.AreaDefinition("t", "t", "t", proj, 100, 40, (-500000, 0, 500000, 400000))
    mid = geometry.AreaDefinition("m", "m", "m", proj, 100, 30, (-500000, -300000, 500000, 0))
    far = geometry.AreaDefinition("f", "f", "f", proj, 100, 20, (-500000, -900000, 500000, -700000))
    st = geometry.StackedAreaDefinition(top, mid, far)
    assert len(st.defs) == 2 and st.defs[0].height == 70 and st.defs[1].height == 20
    assert st.defs[0].area_extent == (-500000, -300000, 500000, 400000)
    assert st.shape == (90, 100) and st.size == 9000 and st.width == 100 and st.ndim == 2
    assert geometry.StackedAreaDefinition(top, mid).squeeze() is not None
    assert isinstance(geometry.StackedAreaDefinition(top, mid).squeeze(), geometry.AreaDefinition)
    other = geometry.AreaDefinition("o", "o", "o", "+proj=eqc +datum=WGS84", 100, 20, (-500000, -900000, 500000, -700000))
    with pytest.raises(NotImplementedError):
        st.append(other)
    narrow = geometry.AreaDefinition("n", "n", "n", proj, 50, 20, (-500000, -1300000, 0, -1100000))
    with pytest.raises(geometry.IncompatibleAreas):
        geometry.concatenate_area_defs(top, narrow)
    with pytest.raises(geometry.IncompatibleAreas):
        geometry.combine_area_extents_vertical(top, far)
    import types
    nested = geometry.StackedAreaDefinition(st, types.SimpleNamespace(height=0))  # empty members are dropped
    assert len(nested.defs) == 2 and nested.height == 90


def test_neighbour_info_file_round_trip(tmp_path):
    """save/load keeps masks, dtypes and shapes - also the empty-info form (int32 indices, distances 1.0)."""
    from pyresample_b200 import neighbour_cache as nc
    rng = np.random.default_rng(3)
    vii = rng.random(1003) < 0.7
    voi = rng.random((17, 13)) < 0.9
    ia = rng.integers(0, vii.sum() + 1, size=(int(voi.sum()), 4)).astype(np.uint32)
    da = rng.random((int(voi.sum()), 4)).astype(np.float32)
    path = tmp_path / "info.npz"
    nc.save_neighbour_info(path, vii, voi, ia, da)
    v1, v2, i2, d2 = nc.load_neighbour_info(path)
    assert v1.dtype == bool and np.array_equal(v1, vii) and v2.shape == (17, 13) and np.array_equal(v2, voi)
    assert i2.dtype == np.uint32 and np.array_equal(i2, ia) and d2.dtype == np.float32 and np.array_equal(d2, da)
    nc.save_neighbour_info(path, np.zeros(5, bool), np.ones(6, bool), np.full(6, 5, np.int32), None)
    v1, v2, i2, d2 = nc.load_neighbour_info(path)
    assert not v1.any() and v2.all() and i2.dtype == np.int32 and d2 is None


def test_registry_and_small_helpers_without_gpu():
    """The resampler registry, the weight-function probe and the index downcast are pure host logic."""
    from pyresample_b200 import future_nearest, kd_tree, registry, utils
    assert registry.list_resamplers() == ["nearest"]
    assert registry.RESAMPLER_REGISTRY["nearest"] is future_nearest.KDTreeNearestXarrayResampler
    with pytest.raises(ValueError, match="already registered"):
        registry.register_resampler("nearest", object)
    assert kd_tree._weights_finite_at_one(lambda d: 1 - d / 1000.0)
    assert kd_tree._weights_finite_at_one([lambda d: np.exp(-d), lambda d: d * 0 + 2])
    assert not kd_tree._weights_finite_at_one(lambda d: 1.0 / (d - 1.0))
    assert not kd_tree._weights_finite_at_one(lambda d: d["no such key"])
    idx = np.array([[0, 5, -1], [7, 9, 3]], dtype=np.int32)
    down = utils._downcast_index_array(idx.copy(), 8)
    assert down.dtype == np.uint16 and np.array_equal(down, [[0, 5, 8], [7, 8, 3]])
    big = utils._downcast_index_array(idx.copy(), 70000)
    assert big.dtype == np.int32 and np.array_equal(big, idx)
    assert utils.fwhm2sigma(2 * np.sqrt(np.log(2))) == pytest.approx(1.0)
    area = geometry.AreaDefinition("a", "a", "a", "+proj=eqc +datum=WGS84", 10, 10, (0, 0, 1e6, 1e6))
    with pytest.raises(ValueError, match="2 dimensions"):
        future_nearest.KDTreeNearestXarrayResampler(area, geometry.SwathDefinition(np.zeros(3), np.zeros(3)))
    r = registry.create_resampler(area, area)
    assert r.version == "0.1" and r._get_src_geo_dims() == ('y', 'x')
    with pytest.raises(ValueError, match="not the same shape"):
        r.precompute(mask=np.zeros((3, 3), dtype=bool))


def test_projection_parameters_are_whitelisted():
    """Definitions the device code would only half understand are refused: unknown keys, named prime
    meridians, datum shifts, +lon_0 on a geographic CRS; +proj=ups forces PROJ's constants."""
    from pyresample_b200.proj import ProjSpec
    for bad in ("+proj=laea +lat_0=50 +lon_0=10 +lat_1=30 +ellps=WGS84", "+proj=stere +lat_0=90 +axis=neu +ellps=WGS84",
                "+proj=longlat +pm=lisbon +ellps=WGS84", "+proj=merc +towgs84=1,2,3 +ellps=WGS84",
                "+proj=longlat +lon_0=10 +ellps=WGS84"):
        with pytest.raises(NotImplementedError):
            ProjSpec(bad)
    ProjSpec("+proj=merc +towgs84=0,0,0,0,0,0,0 +ellps=WGS84 +no_defs +type=crs")
    ups = ProjSpec("+proj=ups +k_0=1 +x_0=5 +y_0=7 +ellps=WGS84").cstruct
    assert (ups.x0, ups.y0) == (2000000.0, 2000000.0) and ups.phi0 > 1.5
    ref = ProjSpec("+proj=stere +lat_0=90 +k_0=0.994 +x_0=2000000 +y_0=2000000 +ellps=WGS84").cstruct
    assert ups.c[0] == ref.c[0]
