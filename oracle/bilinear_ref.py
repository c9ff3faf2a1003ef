"""CPU restatement of pyresample's bilinear neighbour search and sampling (TEST INFRASTRUCTURE ONLY).

Follows pyresample/bilinear/_base.py (``BilinearBase.get_bil_info`` :99-119 and what it calls: valid input index
and tree :120-131, 228-249, 597-613; index array :149-162; four closest corners :391-411, 541-594; fractional
distances :414-538; slices :193-211; sampling :251-262, 653-660) and pyresample/bilinear/_numpy_resampler.py
(``resample_bilinear`` :37-99, ``get_sample_from_bil_info`` :102-158, ``get_bil_info`` :161-225,
``NumpyBilinearResampler`` :228-273).  pykdtree is restated by oracle/knn_ref.py, ``pyproj.Proj(...)(lon, lat)`` by
``proj_ref.forward`` and ``get_proj_coords`` by ``proj_ref.proj_vectors``.

Parity pins (tests/test_oracle_golden.py): the literals of pyresample/test/test_bilinear.py:139-383.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
import numpy as np

from . import kd_tree_ref, proj_ref
from .knn_ref import KDTreeRef


def _outside(data, lo, hi):
    """find_indices_outside_min_and_max (_base.py:382-384): NaN is NOT outside."""
    with np.errstate(invalid="ignore"):
        return (data < lo) | (data > hi)


# ------------------------------------------------------------------------------------------ fractional distances
def calc_abc(corner_points, out_y, out_x):
    """_calc_abc (_base.py:470-497)."""
    pt_1, pt_2, pt_3, pt_4 = corner_points
    x_21 = pt_2[:, 0] - pt_1[:, 0]
    x_31 = pt_3[:, 0] - pt_1[:, 0]
    x_42 = pt_4[:, 0] - pt_2[:, 0]
    y_21 = pt_2[:, 1] - pt_1[:, 1]
    y_31 = pt_3[:, 1] - pt_1[:, 1]
    y_42 = pt_4[:, 1] - pt_2[:, 1]
    a = x_31 * y_42 - y_31 * x_42
    b = out_y * (x_42 - x_31) - out_x * (y_42 - y_31) + x_31 * pt_2[:, 1] - y_31 * pt_2[:, 0] + \
        y_42 * pt_1[:, 0] - x_42 * pt_1[:, 1]
    c = out_y * x_21 - out_x * y_21 + pt_1[:, 0] * pt_2[:, 1] - pt_2[:, 0] * pt_1[:, 1]
    return a, b, c


def solve_quadratic(a, b, c, min_val=0.0, max_val=1.0):
    """_solve_quadratic (_base.py:431-460): the root inside [min_val, max_val], else the linear solution, else NaN."""
    a, b, c = (np.array([v]) if np.isscalar(v) else v for v in (a, b, c))
    with np.errstate(all="ignore"):
        disc = b * b - 4 * a * c
        x_1 = (-b + np.sqrt(disc)) / (2 * a)
        x_2 = (-b - np.sqrt(disc)) / (2 * a)
        x_3 = -c / b
        x = np.where(_outside(x_1, min_val, max_val) | np.isnan(x_1), x_2, x_1)
        x = np.where(_outside(x, min_val, max_val) | np.isnan(x), x_3, x)
        x = np.where(_outside(x, min_val, max_val), np.nan, x)
    return x


def _other_distance(f, y_corners, out_y):
    """_solve_another_fractional_distance (_base.py:500-517)."""
    y_1, y_2, y_3, y_4 = y_corners
    y_21 = y_2 - y_1
    y_43 = y_4 - y_3
    with np.errstate(all="ignore"):
        g = (out_y - y_1 - y_21 * f) / (y_3 + y_43 * f - y_1 - y_21 * f)
    return np.where(_outside(g, 0, 1), np.nan, g)


def _invalid_to_nan(t, s):
    bad = _outside(t, 0, 1) | _outside(s, 0, 1)
    return np.where(bad, np.nan, t), np.where(bad, np.nan, s)


def distances_irregular(corner_points, out_y, out_x):
    """_get_fractional_distances_irregular (_base.py:414-428)."""
    t = solve_quadratic(*calc_abc(corner_points, out_y, out_x))
    pt_1, pt_2, pt_3, pt_4 = corner_points
    s = _other_distance(t, (pt_1[:, 1], pt_3[:, 1], pt_2[:, 1], pt_4[:, 1]), out_y)
    return t, s


def distances_uprights_parallel(corner_points, out_y, out_x):
    """_get_fractional_distances_uprights_parallel (_base.py:529-543)."""
    pt_1, pt_2, pt_3, pt_4 = corner_points
    s = solve_quadratic(*calc_abc((pt_1, pt_3, pt_2, pt_4), out_y, out_x))
    t = _other_distance(s, (pt_1[:, 1], pt_2[:, 1], pt_3[:, 1], pt_4[:, 1]), out_y)
    return t, s


def distances_parallelogram(points, out_y, out_x):
    """_get_fractional_distances_parallellogram (_base.py:546-570)."""
    pt_1, pt_2, pt_3 = points
    x_21 = pt_2[:, 0] - pt_1[:, 0]
    x_31 = pt_3[:, 0] - pt_1[:, 0]
    y_21 = pt_2[:, 1] - pt_1[:, 1]
    y_31 = pt_3[:, 1] - pt_1[:, 1]
    with np.errstate(all="ignore"):
        t = (x_21 * (out_y - pt_1[:, 1]) - y_21 * (out_x - pt_1[:, 0])) / (x_21 * y_31 - y_21 * x_31)
        t = np.where(_outside(t, 0., 1.), np.nan, t)
        s = (out_x - pt_1[:, 0] + x_31 * t) / x_21
        s = np.where(_outside(s, 0., 1.), np.nan, s)
    return t, s


def fractional_distances(corner_points, out_x, out_y):
    """_get_fractional_distances (_base.py:395-411): irregular, then uprights parallel, then parallelogram."""
    t, s = _invalid_to_nan(*distances_irregular(corner_points, out_y, out_x))
    for func, pts in ((distances_uprights_parallel, corner_points), (distances_parallelogram, corner_points[:3])):
        idxs = np.ravel(np.isnan(t) | np.isnan(s))
        if np.any(idxs):
            nt, ns = func(pts, out_y, out_x)
            t, s = _invalid_to_nan(np.where(idxs, nt, t), np.where(idxs, ns, s))
    return t, s


# ------------------------------------------------------------------------------------------ corners
def four_closest_corners(in_x, in_y, out_x, out_y, neighbours, index_array):
    """_get_four_closest_corners (_base.py:391-411, 573-611): per output location the closest neighbour in each
    quadrant (upper left, upper right, lower left, lower right); NaN coordinates where a quadrant is empty (the index
    is then the first neighbour's)."""
    x_diff = out_x[:, None] - in_x
    y_diff = out_y[:, None] - in_y
    stride = np.arange(x_diff.shape[0])
    quadrants = ((x_diff > 0) & (y_diff < 0), (x_diff < 0) & (y_diff < 0), (x_diff > 0) & (y_diff > 0),
                 (x_diff < 0) & (y_diff > 0))
    points, indices = [], []
    for valid in quadrants:
        first = np.argmax(valid, axis=1)
        none = ~np.max(valid, axis=1)
        x = np.where(none, np.nan, in_x[stride, first])
        y = np.where(none, np.nan, in_y[stride, first])
        points.append(np.transpose(np.vstack((x, y))))
        indices.append(index_array[stride, first])
    return points, np.transpose(np.vstack(indices))


# ------------------------------------------------------------------------------------------ info
def _is_gridlike(g):
    return kd_tree_ref.is_area(g) or kd_tree_ref.is_grid(g)


def bil_info(source_geo_def, target_geo_def, radius=50e3, neighbours=32, reduce_data=True, epsilon=0):
    """BilinearBase.get_bil_info (_base.py:99-119) -> dict with everything the resampler object holds."""
    slons, slats = kd_tree_ref.get_lonlats(source_geo_def)
    slons, slats = np.ravel(slons), np.ravel(slats)
    valid_input = ~(_outside(slons, -180., 180.) | _outside(slats, -90., 90.))
    if reduce_data and ((kd_tree_ref.is_coord(source_geo_def) or _is_gridlike(source_geo_def)) and _is_gridlike(target_geo_def)):
        blons, blats = kd_tree_ref.get_boundary_lonlats(target_geo_def)
        valid_input &= kd_tree_ref.get_valid_index_from_lonlat_boundaries(blons, blats, slons, slats, radius)
    out = {"valid_input_index": valid_input}
    coords = kd_tree_ref.lonlat2xyz(slons, slats)[valid_input, :].astype(np.float64)
    tsize = int(np.prod(target_geo_def.shape))
    if coords.size == 0:  # _create_empty_bil_info (_base.py:133-141)
        out.update(valid_input_index=np.ones(slons.size, dtype=bool), index_array=np.ones((tsize, 4), dtype=np.int32),
                   bilinear_s=np.nan * np.zeros(tsize), bilinear_t=np.nan * np.zeros(tsize),
                   slices_x=np.zeros((tsize, 4), dtype=np.int32), slices_y=np.zeros((tsize, 4), dtype=np.int32))
        out["mask_slices"] = out["index_array"] >= slons.size
        return out
    tree = KDTreeRef(coords)
    tlons, tlats = kd_tree_ref.get_lonlats(target_geo_def)
    with np.errstate(invalid="ignore"):
        valid_output = np.ravel((tlons >= -180) & (tlons <= 180) & (tlats <= 90) & (tlats >= -90))
    q = kd_tree_ref.lonlat2xyz(np.ravel(tlons)[valid_output], np.ravel(tlats)[valid_output])
    _, index_array = tree.query(q, k=neighbours, eps=epsilon, distance_upper_bound=radius)
    index_array = np.where(index_array == valid_input.sum(), 0, index_array)
    tx, ty = proj_ref.proj_vectors(target_geo_def.area_extent, target_geo_def.width, target_geo_def.height, np.float64)
    gx, gy = np.meshgrid(tx, ty)
    out_x, out_y = np.ravel(gx)[valid_output], np.ravel(gy)[valid_output]
    lon_m = np.where(_outside(slons, -180., 180.) | _outside(slats, -90., 90.), np.nan, slons)[valid_input]
    lat_m = np.where(_outside(slons, -180., 180.) | _outside(slats, -90., 90.), np.nan, slats)[valid_input]
    in_x, in_y = proj_ref.forward(target_geo_def.proj_dict, lon_m[index_array], lat_m[index_array])
    corners, index_array = four_closest_corners(in_x, in_y, out_x, out_y, neighbours, index_array)
    t, s = fractional_distances(corners, out_x, out_y)
    shp = source_geo_def.shape
    if len(shp) > 1:
        cols, lines = np.meshgrid(np.arange(shp[1]), np.arange(shp[0]))
        lines, cols = np.ravel(lines), np.ravel(cols)
    else:
        lines, cols = np.zeros(shp[0], dtype=np.uint32), np.arange(shp[0])
    out.update(index_array=index_array, bilinear_t=t, bilinear_s=s, valid_output_index=valid_output,
               slices_y=lines[valid_input][index_array], slices_x=cols[valid_input][index_array],
               mask_slices=index_array >= slons.size, corner_points=corners, out_coords_x=tx, out_coords_y=ty)
    return out


def get_bil_info(source_geo_def, target_area_def, radius=50e3, neighbours=32, nprocs=1, masked=False, reduce_data=True,
                 segments=None, epsilon=0):
    """_numpy_resampler.get_bil_info (:161-225) -> (t, s, valid_input_index, index_array)."""
    info = bil_info(source_geo_def, target_area_def, radius, neighbours, reduce_data, epsilon)
    return info["bilinear_t"], info["bilinear_s"], info["valid_input_index"], info["index_array"]


# ------------------------------------------------------------------------------------------ sampling
def _bilinear(corner_values, s, t):
    """_resample (_base.py:653-660)."""
    p_1, p_2, p_3, p_4 = corner_values
    return p_1 * (1 - s) * (1 - t) + p_2 * s * (1 - t) + p_3 * (1 - s) * t + p_4 * s * t


def get_sample_from_bil_info(data, t, s, input_idxs, idx_arr, output_shape=None):
    """_numpy_resampler.get_sample_from_bil_info (:102-158)."""
    new_data = data[input_idxs]
    data_min = np.nanmin(new_data) - 1e-6
    data_max = np.nanmax(new_data) + 1e-6
    new_data = new_data[idx_arr]
    result = _bilinear((new_data[:, 0], new_data[:, 1], new_data[:, 2], new_data[:, 3]), s, t)
    if hasattr(result, 'mask'):
        mask = result.mask
        result = result.data
        result[mask] = np.nan
    result[_outside(result, data_min, data_max)] = np.nan
    if output_shape is not None:
        result = result.reshape(output_shape)
    return result


def resample_bilinear(data, source_geo_def, target_area_def, radius=50e3, neighbours=32, nprocs=1, fill_value=0,
                      reduce_data=True, segments=None, epsilon=0):
    """NumpyBilinearResampler.resample (_numpy_resampler.py:231-273) as called by resample_bilinear (:37-99)."""
    info = bil_info(source_geo_def, target_area_def, radius, neighbours, reduce_data, epsilon)
    if data.ndim > 2 and data.shape[0] * data.shape[1] == info["valid_input_index"].shape[0]:
        data = np.moveaxis(data, -1, 0)
    if data.ndim == 1:
        data = np.expand_dims(data, 1)
    sl_x, sl_y, mask = info["slices_x"], info["slices_y"], info["mask_slices"]
    if data.ndim == 2:
        arr = data[(sl_y, sl_x)]
        arr[(mask,)] = fill_value
        corner = (arr[:, 0], arr[:, 1], arr[:, 2], arr[:, 3])
    else:
        arr = data[(slice(None), sl_y, sl_x)]
        arr[(slice(None), mask)] = fill_value
        corner = (arr[:, :, 0], arr[:, :, 1], arr[:, :, 2], arr[:, :, 3])
    res = _bilinear(corner, info["bilinear_s"], info["bilinear_t"])
    shp = tuple(target_area_def.shape)
    if data.ndim == 3:
        res = np.moveaxis(np.reshape(res, (res.shape[0],) + shp), 0, -1)
    else:
        res = np.reshape(res, shp)
    res = np.squeeze(res)
    if fill_value is None:
        res = np.ma.masked_invalid(res)
    else:
        res[np.isnan(res)] = fill_value
    return np.squeeze(res)
