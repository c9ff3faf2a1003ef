"""TEST INFRASTRUCTURE ONLY - CPU restatement of ``KDTreeNearestXarrayResampler``
(pyresample/future/resamplers/nearest.py:115-546) on plain numpy arrays: the numerics of ``precompute`` /
``resample`` / ``compute_data_mask`` / ``_compute_radius_of_influence`` without xarray and dask.

Parity: PINNED on the reference's own literals for this class (test_resamplers/test_nearest.py:55-228):
1-D swath -> 2-D grid arrays, 167913.0, 303048.0, 115591.0, 952386.0 (default radius from the geocentric
resolutions), 20666.0 (10 km fallback), 909144.0 - see tests/test_oracle_golden.py.
Nothing under pyresample_b200/ may import this module.
"""
import numpy as np

from . import kd_tree_ref, proj_ref
from .knn_ref import KDTreeRef

_WGS84 = (6378137.0, 298.257223563)


def _geocentric(lons, lats, radius=None):
    """``+proj=cart`` with altitude 0 on WGS84 (geometry.py:733-757, 2719-2724)."""
    if radius:
        a, es = float(radius), 0.0
    else:
        a, rf = _WGS84
        f = 1.0 / rf
        es = 2 * f - f * f
    lam, phi = np.deg2rad(np.asarray(lons, dtype=np.float64)), np.deg2rad(np.asarray(lats, dtype=np.float64))
    n = a / np.sqrt(1 - es * np.sin(phi) ** 2)
    return np.stack((n * np.cos(phi) * np.cos(lam), n * np.cos(phi) * np.sin(lam), n * (1 - es) * np.sin(phi)), axis=1)


def swath_geocentric_resolution(lons, lats):
    """SwathDefinition.geocentric_resolution, geometry.py:702-762 (no ``resolution`` attribute)."""
    lons, lats = np.asarray(lons), np.asarray(lats)
    if lons.ndim == 1:
        raise RuntimeError("Can't confidently determine geocentric resolution for 1D swath.")
    r = lons.shape[0] // 2
    xyz = _geocentric(lons[r:r + 1, :2].ravel(), lats[r:r + 1, :2].ravel())
    dist = np.linalg.norm(xyz[1] - xyz[0])
    if not np.isfinite(dist):
        raise RuntimeError("Could not calculate geocentric resolution")
    return dist


def area_geocentric_resolution(area):
    """AreaDefinition.geocentric_resolution, geometry.py:2691-2764: middle row and middle column, mean of the first
    two edges of a 10-bin histogram of the geocentric steps, maximum of both."""
    rows, cols = area.height, area.width
    mid_row = proj_ref.area_lonlats(area.proj_dict, area.area_extent, cols, rows, dtype=np.float64,
                                    data_slice=(rows // 2, slice(None)))
    mid_col = proj_ref.area_lonlats(area.proj_dict, area.area_extent, cols, rows, dtype=np.float64,
                                    data_slice=(slice(None), cols // 2))
    res = []
    for lo, la in (mid_row, mid_col):
        with np.errstate(invalid="ignore"):
            xyz = _geocentric(np.ravel(lo), np.ravel(la))
            dist = np.linalg.norm(np.diff(xyz, axis=0), axis=1)
        dist = dist[np.isfinite(dist)]
        try:
            edges = np.histogram_bin_edges(dist, bins=10)[:2]
        except ValueError:
            edges = dist[:2]
        res.append(np.mean(edges) if dist.size else 0)
    out = max(res)
    if not out:
        raise RuntimeError("Could not calculate geocentric resolution")
    return out


def geocentric_resolution(geo_def):
    if kd_tree_ref.is_area(geo_def) and getattr(geo_def, "lons", None) is None:
        return area_geocentric_resolution(geo_def)
    lons, lats = kd_tree_ref.get_lonlats(geo_def)
    return swath_geocentric_resolution(lons, lats)


def compute_radius_of_influence(source_geo_def, target_geo_def, fail_source=False, fail_target=False):
    """nearest.py:146-169; ``fail_*`` stands in for a geometry whose resolution cannot be determined."""
    res = []
    for g, fail in ((source_geo_def, fail_source), (target_geo_def, fail_target)):
        try:
            if fail:
                raise RuntimeError
            res.append(geocentric_resolution(g))
        except RuntimeError:
            res.append(np.nan)
    src_res, dst_res = res
    if np.isnan(src_res):
        radius = dst_res
    elif np.isnan(dst_res):
        radius = src_res
    else:
        radius = max(src_res, dst_res)
    if np.isnan(radius):
        radius = 10000
    return radius


def compute_data_mask(data, geo_axes):
    """nearest.py:527-543: invalid = NaN (floats) or the dtype maximum (integers), on every non-geo position."""
    data = np.asarray(data)
    if np.issubdtype(data.dtype, np.integer):
        mask = data == np.iinfo(data.dtype.type).max
    else:
        mask = np.isnan(data)
    other = tuple(i for i in range(data.ndim) if i not in geo_axes)
    return mask.all(axis=other) if other else mask


def precompute(source_geo_def, target_geo_def, mask=None, radius_of_influence=None, epsilon=0):
    """nearest.py:176-243 (+ the kd-tree of :171-184): ``(valid_input_index, index_array (rows, cols, 1))``."""
    if radius_of_influence is None:
        radius_of_influence = compute_radius_of_influence(source_geo_def, target_geo_def)
    slons, slats = kd_tree_ref.get_lonlats(source_geo_def)
    slons, slats = np.asarray(slons), np.asarray(slats)
    with np.errstate(invalid="ignore"):
        vii = (slons >= -180) & (slons <= 180) & (slats <= 90) & (slats >= -90)
    coords = kd_tree_ref.lonlat2xyz(slons, slats)[vii.ravel(), :].astype(np.float64)
    tree = KDTreeRef(coords)
    tlons, tlats = kd_tree_ref.get_lonlats(target_geo_def)
    tlons, tlats = np.asarray(tlons), np.asarray(tlats)
    with np.errstate(invalid="ignore"):
        voi = (tlons >= -180) & (tlons <= 180) & (tlats <= 90) & (tlats >= -90)
    if mask is not None and np.asarray(mask).shape != slons.shape:
        raise ValueError("'mask' must be the same shape as the source geo definition")
    ia = kd_tree_ref.query_no_distance(tlons, tlats, voi, mask=None if mask is None else np.asarray(mask),
                                       valid_input_index=vii, neighbours=1, epsilon=epsilon,
                                       radius=radius_of_influence, kdtree=tree)
    return vii, ia


def resample(source_geo_def, target_geo_def, data, geo_axes, mask=None, fill_value=np.nan,
             radius_of_influence=None, epsilon=0):
    """nearest.py:423-470 on a numpy array whose axes ``geo_axes`` (consecutive) are the source geometry's."""
    data = np.asarray(data)
    if fill_value is not None and np.isnan(fill_value) and np.issubdtype(data.dtype, np.integer):
        fill_value = np.iinfo(data.dtype.type).max
    vii, ia = precompute(source_geo_def, target_geo_def, mask, radius_of_influence, epsilon)
    first = geo_axes[0]
    flat_shape = data.shape[:first] + (-1,) + data.shape[first + len(geo_axes):]
    slices = [slice(None)] * len(flat_shape)
    slices[first] = None
    return kd_tree_ref._my_index(ia[:, :, 0], vii, data.reshape(flat_shape), vii_slices=list(slices),
                                 ia_slices=list(slices), fill_value=fill_value)
