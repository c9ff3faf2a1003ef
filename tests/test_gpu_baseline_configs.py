"""Every BASELINE.json configuration (C1-C5, `workloads.make_config`) through the public API, as configured,
against the CPU oracle on the same seeded inputs.

C1 runs at full size; C2-C5 are shrunk (``scale``) so that the oracle finishes in seconds - the full sizes are
covered by size-independent properties in tests/test_gpu_fullsize.py.  The bar (BASELINE.json north_star): indices
and valid masks bit-exact; float32 distances / weighted outputs within 1e-4 relative.

float32 trees: the reference computes the target (and, for an area source, the source) lon/lat in float64 with PROJ
and casts to float32; the CUDA inverse projection and the oracle's numpy restatement agree to ~1e-12 degrees, so the
cast lands on the other side of a float32 rounding boundary for ~1e-7 of the pixels, i.e. one coordinate differs by
1 ulp (0.4 m).  Every index that differs from the oracle is therefore EXPLAINED, never tolerated: it must be an
exact distance tie or disappear when the CUDA search is re-run on the oracle's own lon/lat given as explicit arrays.
The counts are printed and written to gpurun_out/baseline_parity_counts.json.
"""
import json
import os
import warnings

import numpy as np
import pytest

from oracle import kd_tree_ref, proj_ref
from oracle.knn_ref import KDTreeRef

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_COUNTS = {}


def _record(name, **counts):
    _COUNTS[name] = {k: (int(v) if isinstance(v, (int, np.integer)) else v) for k, v in counts.items()}
    print("[baseline parity] %s: %s" % (name, json.dumps(_COUNTS[name])))
    out_dir = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out_dir, exist_ok=True)
        with open(os.path.join(out_dir, "baseline_parity_counts.json"), "w") as fh:
            json.dump(_COUNTS, fh, indent=1, sort_keys=True)
    except OSError:
        pass


@pytest.fixture(scope="module")
def kd(cuda_lib):
    from pyresample_b200 import kd_tree
    return kd_tree


@pytest.fixture(scope="module")
def geometry():
    from pyresample_b200 import geometry
    return geometry


def _oracle_lonlats(area, dtype):
    return proj_ref.area_lonlats(area.proj_dict, area.area_extent, area.width, area.height, dtype=dtype)


def _source_geo(cfg, geometry):
    if "source_area" in cfg:
        return cfg["source_area"]
    return geometry.SwathDefinition(cfg["lon"], cfg["lat"])


def _source_lonlats(cfg):
    """The oracle's lon/lat of the source (explicit arrays for a swath, restated PROJ inverse for an area)."""
    if "source_area" in cfg:
        sa = cfg["source_area"]
        return _oracle_lonlats(sa, sa.dtype)
    return cfg["lon"], cfg["lat"]


def _explained_info(kd, geometry, cfg, got, want, k, rtol, name):
    """Compare neighbour info with the oracle's; returns the boolean mask (over valid output pixels) of rows whose
    indices differ.  Every differing row must be a distance tie or vanish on the oracle's own coordinates."""
    vii, voi, ia, da = got
    ovii, ovoi, oia, oda = want
    assert np.array_equal(vii, ovii), "valid_input_index differs"
    assert np.array_equal(voi, ovoi), "valid_output_index differs"
    assert ia.dtype == oia.dtype == np.uint32 and ia.shape == oia.shape
    assert da.dtype == oda.dtype and da.shape == oda.shape
    ia2, oia2 = ia.reshape(len(ia), -1), oia.reshape(len(oia), -1)
    da2, oda2 = da.reshape(ia2.shape), oda.reshape(oia2.shape)
    diff = ia2 != oia2
    rows = np.flatnonzero(diff.any(axis=1))
    ties = int((diff & (da2 == oda2)).sum())
    n_entries = int(diff.sum())
    flips = 0
    if rows.size:
        assert rows.size <= max(8, 2e-4 * len(ia2)), "%d of %d rows differ" % (rows.size, len(ia2))
        # re-run the CUDA search on the oracle's coordinates (explicit arrays on both sides): must be exact there
        slon, slat = _source_lonlats(cfg)
        tlon, tlat = _oracle_lonlats(cfg["area"], np.asarray(slon).dtype)
        sel = np.flatnonzero(ovoi)[rows]
        src = geometry.SwathDefinition(np.asarray(slon).ravel()[ovii], np.asarray(slat).ravel()[ovii])
        tgt = geometry.SwathDefinition(tlon.ravel()[sel], tlat.ravel()[sel])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            v2, o2, ia_r, da_r = kd.get_neighbour_info(src, tgt, cfg["radius"], neighbours=k, reduce_data=False)
        assert v2.all() and o2.all()
        ia_r, da_r = ia_r.reshape(len(rows), -1), da_r.reshape(len(rows), -1)
        still = ia_r != oia2[rows]
        # what is left after the re-run may only be exact ties (lowest index first is this build's rule)
        assert not np.any(still & (da_r != oda2[rows])), "index mismatch that is neither a tie nor a coordinate flip"
        flips = int((diff[rows] & ~still).sum())
    same = ~diff
    fin = np.isfinite(oda2) & same
    assert np.array_equal(np.isinf(da2[same]), np.isinf(oda2[same]))
    np.testing.assert_allclose(da2[fin], oda2[fin], rtol=rtol, atol=0)
    _record(name, target_pixels=len(ia2), neighbours=k, index_entries_differing=n_entries,
            exact_distance_ties=ties, vanish_on_oracle_coordinates=flips, rows_differing=int(rows.size))
    bad = np.zeros(len(ia2), dtype=bool)
    bad[rows] = True
    return bad


def _pixels_of(bad_rows, voi, shape):
    full = np.zeros(voi.size, dtype=bool)
    full[np.flatnonzero(voi)[bad_rows]] = True
    return full.reshape(shape)


def _assert_same_except(got, want, allowed, rtol=0.0, atol=0.0):
    """got == want (within rtol) on every pixel outside ``allowed`` (2-D mask, broadcast over channels)."""
    assert got.shape == want.shape and got.dtype == want.dtype
    if rtol:
        ok = np.isclose(got, want, rtol=rtol, atol=atol, equal_nan=True)
    else:
        ok = (got == want) | (np.isnan(got) & np.isnan(want)) if got.dtype.kind == 'f' else (got == want)
    if ok.ndim > allowed.ndim:
        ok = ok.all(axis=-1)
    assert not np.any(~ok & ~allowed), "%d pixels differ outside the explained ones" % int((~ok & ~allowed).sum())
    return int((~ok).sum())


def _nearest_case(kd, geometry, cid, scale, name):
    import workloads
    cfg = workloads.make_config(cid, scale=scale)
    src, area, radius = _source_geo(cfg, geometry), cfg["area"], cfg["radius"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got_info = kd.get_neighbour_info(src, area, radius, neighbours=1)
        want_info = kd_tree_ref.get_neighbour_info(src, area, radius, neighbours=1)
        bad = _explained_info(kd, geometry, cfg, got_info, want_info, 1, 1e-4, name + " info")
        got = kd.resample_nearest(src, cfg["data"], area, radius, fill_value=-999)
        want = kd_tree_ref.resample_nearest(src, cfg["data"], area, radius, fill_value=-999)
    n = _assert_same_except(got, want, _pixels_of(bad, want_info[1], area.shape))
    found = (want_info[2] < want_info[0].sum()).mean()
    assert found > 0.005, "the configuration must actually produce neighbours"
    _record(name + " resample_nearest", pixels_differing=n, fraction_with_neighbour=float(found))
    # the two-step path on the product's own info
    two = kd.get_sample_from_neighbour_info('nn', area.shape, cfg["data"], *got_info[:3], fill_value=-999)
    assert np.array_equal(two, got, equal_nan=True)


def test_c1_full_size_nearest(kd, geometry):
    """C1 as configured: 1000x1000 f32 swath -> 500x500 laea, r = 50 km, k = 1, reduce_data on (full size)."""
    _nearest_case(kd, geometry, 1, 1.0, "C1 full")


def test_c2_nearest_5ch(kd, geometry):
    """C2 at scale 0.1: AVHRR-like swath, 5 channels -> global eqc grid, nearest."""
    _nearest_case(kd, geometry, 2, 0.1, "C2 x0.1")


def test_c5_geos_source_to_longlat(kd, geometry):
    """C5 at scale 0.1: geos (sweep=x) AREA source with off-disk pixels -> longlat grid, nearest."""
    _nearest_case(kd, geometry, 5, 0.1, "C5 x0.1")


def test_c3_gauss_k8_16ch(kd, geometry):
    """C3 at scale 0.1: f32 swath, 16 channels -> polar stere (lat_ts), resample_gauss k = 8, sigma = 25 km."""
    import workloads
    cfg = workloads.make_config(3, scale=0.1)
    src, area, radius, k = _source_geo(cfg, geometry), cfg["area"], cfg["radius"], cfg["k"]
    sigmas = [cfg["sigma"]] * cfg["channels"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got_info = kd.get_neighbour_info(src, area, radius, neighbours=k)
        want_info = kd_tree_ref.get_neighbour_info(src, area, radius, neighbours=k)
        bad = _explained_info(kd, geometry, cfg, got_info, want_info, k, 1e-4, "C3 x0.1 info")
        got = kd.resample_gauss(src, cfg["data"], area, radius, sigmas, neighbours=k, fill_value=-1)
        want = kd_tree_ref.resample_gauss(src, cfg["data"], area, radius, sigmas, neighbours=k, fill_value=-1)
    assert got.dtype == want.dtype == np.float64  # multi-channel weights promote to float64 (kd_tree.py:783)
    n = _assert_same_except(got, want, _pixels_of(bad, want_info[1], area.shape), rtol=1e-4, atol=1e-9)
    assert (want != -1).mean() > 0.05
    _record("C3 x0.1 resample_gauss", pixels_outside_tolerance=n)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        g, gs, gc = kd.resample_gauss(src, cfg["data"][:, :2], area, radius, sigmas[:2], neighbours=k,
                                      fill_value=None, with_uncert=True)
        w, ws, wc = kd_tree_ref.resample_gauss(src, cfg["data"][:, :2], area, radius, sigmas[:2], neighbours=k,
                                               fill_value=None, with_uncert=True)
    allowed = _pixels_of(bad, want_info[1], area.shape)
    assert np.array_equal(np.ma.getmaskarray(g) & ~allowed[..., None], np.ma.getmaskarray(w) & ~allowed[..., None])
    _assert_same_except(np.ma.filled(g, 0), np.ma.filled(w, 0), allowed, rtol=1e-4, atol=1e-9)
    _assert_same_except(np.ma.filled(gc, -1), np.ma.filled(wc, -1), allowed)
    _assert_same_except(np.ma.filled(gs, -1), np.ma.filled(ws, -1), allowed, rtol=1e-3, atol=1e-6)


def _wf(d):
    return 1 - d / 1e5


def test_c4_custom_k16_and_xarray_nn(kd, geometry, monkeypatch):
    """C4 at scale 0.05, as configured: four stacked granules -> EASE grid, resample_custom with the reference tests'
    weight function (test_kd_tree.py:74-75) x 16 channels, k = 16; and XArrayResamplerNN(neighbours=16) indices +
    neighbours=1 sampling (the reference class cannot sample k > 1, kd_tree.py:1082-1084)."""
    import workloads
    cfg = workloads.make_config(4, scale=0.05)
    src, area, radius, k = _source_geo(cfg, geometry), cfg["area"], cfg["radius"], cfg["k"]
    assert k == 16 and cfg["channels"] == 16
    funcs = [_wf] * 16
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got_info = kd.get_neighbour_info(src, area, radius, neighbours=k)
        want_info = kd_tree_ref.get_neighbour_info(src, area, radius, neighbours=k)
        bad = _explained_info(kd, geometry, cfg, got_info, want_info, k, 1e-4, "C4 x0.05 info")
        got = kd.resample_custom(src, cfg["data"], area, radius, funcs, neighbours=k, fill_value=-1)
        want = kd_tree_ref.resample_custom(src, cfg["data"], area, radius, funcs, neighbours=k, fill_value=-1)
    assert got.dtype == want.dtype == np.float64
    allowed = _pixels_of(bad, want_info[1], area.shape)
    n = _assert_same_except(got, want, allowed, rtol=1e-4, atol=1e-9)
    assert (want != -1).mean() > 0.05
    _record("C4 x0.05 resample_custom", pixels_outside_tolerance=n)

    # ---- XArrayResamplerNN: float64 tree over the float32 coordinates, no reduce_data, -1 sentinels
    from test_gpu_xarray_nn import FakeDataArray
    if kd.DataArray is None:
        monkeypatch.setattr(kd, "DataArray", FakeDataArray)
    lon, lat = cfg["lon"], cfg["lat"]
    swath = geometry.SwathDefinition(FakeDataArray(lon, dims=('y', 'x')), FakeDataArray(lat, dims=('y', 'x')))
    tlon, tlat = _oracle_lonlats(area, np.float64)
    with np.errstate(invalid="ignore"):
        vii = (lon >= -180) & (lon <= 180) & (lat <= 90) & (lat >= -90)
    otree = KDTreeRef(kd_tree_ref.lonlat2xyz(lon, lat).reshape(-1, 3)[vii.ravel()].astype(np.float64))
    voi = np.ones(tlon.shape, dtype=bool)

    def values(x):
        x = x.data if hasattr(x, "dims") else x
        return np.asarray(x.compute() if hasattr(x, "compute") else x)

    for nb in (16, 1):
        want_ia = kd_tree_ref.query_no_distance(tlon, tlat, voi, valid_input_index=vii, neighbours=nb, epsilon=0,
                                                radius=radius, kdtree=otree)
        r = kd.XArrayResamplerNN(swath, area, radius_of_influence=radius, neighbours=nb)
        g_vii, g_voi, g_ia, g_da = r.get_neighbour_info()
        assert g_da is None and np.array_equal(values(g_vii), vii) and values(g_voi).all()
        g_ia = values(g_ia)
        assert g_ia.dtype == np.int64 and g_ia.shape == want_ia.shape == area.shape + (nb,)
        # float64 tree: CUDA and libm sin/cos differ by a few ulp (<= 1e-8 m), so only exact-tie-scale flips may occur
        d = g_ia != want_ia
        _record("C4 x0.05 XArrayResamplerNN k=%d" % nb, index_entries=int(g_ia.size), differing=int(d.sum()))
        assert d.sum() <= 2e-6 * g_ia.size
        if nb == 1:
            data2d = cfg["data"][:, 0].reshape(lon.shape)
            res = r.get_sample_from_neighbour_info(FakeDataArray(data2d, dims=('y', 'x')))
            want_res = kd_tree_ref._my_index(want_ia[:, :, 0], vii, data2d.reshape(-1), vii_slices=[None],
                                             ia_slices=[None], fill_value=np.nan)
            assert res.dims == ('y', 'x')
            _assert_same_except(values(res), want_res, d[:, :, 0])
