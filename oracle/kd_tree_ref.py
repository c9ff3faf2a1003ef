"""Oracle: numpy restatement of pyresample/kd_tree.py (numpy API) and its helpers.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Geometry objects are read
duck-typed (``lons``/``lats`` for coordinate definitions; ``proj_dict``,
``area_extent``, ``width``, ``height`` for area definitions), so the product's
host-side geometry classes can be passed in without this module touching the GPU.

Each function names the reference lines it follows.
"""
import warnings

import numpy as np

from . import proj_ref
from .knn_ref import KDTreeRef

R = 6370997.0  # pyresample/_spatial_mp.py:39


class EmptyResult(ValueError):
    """kd_tree.py:60."""


# ------------------------------------------------------------------ geometry access
def _mro_names(obj):
    return {c.__name__ for c in type(obj).__mro__}


def is_area(g):
    return "AreaDefinition" in _mro_names(g) or hasattr(g, "area_extent")


def is_grid(g):
    return "GridDefinition" in _mro_names(g)


def is_coord(g):
    return "CoordinateDefinition" in _mro_names(g) or (hasattr(g, "lons") and not is_area(g))


def geo_dtype(g):
    return np.dtype(g.dtype)


def get_lonlats(g, data_slice=None, dtype=None):
    """BaseDefinition.get_lonlats geometry.py:242-274 / AreaDefinition.get_lonlats geometry.py:2558-2645."""
    if is_area(g) and getattr(g, "lons", None) is None:
        if dtype is None:
            dtype = g.dtype
        return proj_ref.area_lonlats(g.proj_dict, g.area_extent, g.width, g.height, dtype=dtype, data_slice=data_slice)
    lons, lats = g.lons, g.lats
    if data_slice is not None:
        lons, lats = lons[data_slice], lats[data_slice]
    return lons, lats


def get_boundary_lonlats(g):
    """geometry.py:284-291: four sides clockwise from upper-left, each squeezed."""
    s1 = get_lonlats(g, data_slice=(0, slice(None)))
    s2 = get_lonlats(g, data_slice=(slice(None), -1))
    s3 = get_lonlats(g, data_slice=(-1, slice(None, None, -1)))
    s4 = get_lonlats(g, data_slice=(slice(None, None, -1), 0))
    lons = [np.squeeze(s[0]) for s in (s1, s2, s3, s4)]
    lats = [np.squeeze(s[1]) for s in (s1, s2, s3, s4)]
    return lons, lats


def get_slice(segments, shape):
    """geometry._get_slice geometry.py:3052-3068."""
    if not (1 <= len(shape) <= 2):
        raise ValueError('Cannot segment array of shape: %s' % str(shape))
    size = shape[0]
    slice_length = int(np.ceil(float(size) / segments))
    start_idx = 0
    end_idx = slice_length
    while start_idx < size:
        if len(shape) == 1:
            yield slice(start_idx, end_idx)
        else:
            yield slice(start_idx, end_idx), slice(None)
        start_idx = end_idx
        end_idx = min(start_idx + slice_length, size)


# ------------------------------------------------------------------ ECEF
def transform_lonlats(lons, lats):
    """Cartesian.transform_lonlats _spatial_mp.py:155-169 (numpy branch; numexpr is optional and absent)."""
    if np.issubdtype(lons.dtype, np.integer):
        lons = lons.astype(np.float64)
    coords = np.zeros((lons.size, 3), dtype=lons.dtype)
    coords[:, 0] = R * np.cos(np.deg2rad(lats)) * np.cos(np.deg2rad(lons))
    coords[:, 1] = R * np.cos(np.deg2rad(lats)) * np.sin(np.deg2rad(lons))
    coords[:, 2] = R * np.sin(np.deg2rad(lats))
    return coords


def lonlat2xyz(lons, lats):
    """future/resamplers/_transform_utils.py:22-33."""
    lats = np.deg2rad(lats)
    r_cos_lats = R * np.cos(lats)
    lons = np.deg2rad(lons)
    x_coords = r_cos_lats * np.cos(lons)
    y_coords = r_cos_lats * np.sin(lons)
    z_coords = R * np.sin(lats)
    return np.stack((x_coords.ravel(), y_coords.ravel(), z_coords.ravel()), axis=-1)


def get_cartesian_coords(g, data_slice=None):
    """BaseDefinition.get_cartesian_coords geometry.py:465-516."""
    lons, lats = get_lonlats(g, data_slice=data_slice)
    cc = transform_lonlats(np.ravel(lons), np.ravel(lats))
    if isinstance(lons, np.ndarray) and lons.ndim > 1:
        cc = cc.reshape(lons.shape[0], lons.shape[1], 3)
    return cc


# ------------------------------------------------------------------ data_reduce
def get_valid_index_from_lonlat_boundaries(boundary_lons, boundary_lats, lons, lats, radius_of_influence):
    """data_reduce.py:213-307 (``_get_valid_index``), statement for statement."""
    lons_side1, lons_side2, lons_side3, lons_side4 = boundary_lons
    lats_side1, lats_side2, lats_side3, lats_side4 = boundary_lats
    illegal_lons = (((lons_side1 < -180) | (lons_side1 > 180)).any() or
                    ((lons_side2 < -180) | (lons_side2 > 180)).any() or
                    ((lons_side3 < -180) | (lons_side3 > 180)).any() or
                    ((lons_side4 < -180) | (lons_side4 > 180)).any())
    illegal_lats = (((lats_side1 < -90) | (lats_side1 > 90)).any() or
                    ((lats_side2 < -90) | (lats_side2 > 90)).any() or
                    ((lats_side3 < -90) | (lats_side3 > 90)).any() or
                    ((lats_side4 < -90) | (lats_side4 > 90)).any())
    if illegal_lons or illegal_lats:
        return np.ones(lons.size, dtype=bool)

    angle_sum = 0
    for side in (lons_side1, lons_side2, lons_side3, lons_side4):
        prev = None
        for lon in side:
            if prev:  # data_reduce.py:250 - falsy for None AND for an exact 0.0 longitude
                delta = lon - prev
                if abs(delta) > 180:
                    delta = (abs(delta) - 360) * (delta // abs(delta))
                angle_sum += delta
            prev = lon

    lat_min = min(lats_side1.min(), lats_side2.min(), lats_side3.min(), lats_side4.min())
    lat_min_buffered = lat_min - np.degrees(float(radius_of_influence) / R)
    lat_max = max(lats_side1.max(), lats_side2.max(), lats_side3.max(), lats_side4.max())
    lat_max_buffered = lat_max + np.degrees(float(radius_of_influence) / R)

    max_angle_s2 = max(abs(lats_side2.max()), abs(lats_side2.min()))
    max_angle_s4 = max(abs(lats_side4.max()), abs(lats_side4.min()))
    with np.errstate(divide="ignore"):
        lon_min_buffered = (lons_side4.min() -
                            np.degrees(float(radius_of_influence) / (np.sin(np.radians(max_angle_s4)) * R)))
        lon_max_buffered = (lons_side2.max() +
                            np.degrees(float(radius_of_influence) / (np.sin(np.radians(max_angle_s2)) * R)))

    if round(angle_sum) == -360:
        valid_index = (lats >= lat_min_buffered)
    elif round(angle_sum) == 360:
        valid_index = (lats <= lat_max_buffered)
    elif round(angle_sum) == 0:
        valid_lats = (lats >= lat_min_buffered) * (lats <= lat_max_buffered)
        if lons_side2.min() > lons_side4.max():
            valid_lons = (lons >= lon_min_buffered) * (lons <= lon_max_buffered)
        else:
            seg1 = (lons >= lon_min_buffered) * (lons <= 180)
            seg2 = (lons <= lon_max_buffered) * (lons >= -180)
            valid_lons = seg1 + seg2
        valid_index = valid_lats * valid_lons
    else:
        valid_index = np.ones(lons.size, dtype=bool)
    return valid_index


# ------------------------------------------------------------------ neighbour info
def _get_valid_input_index(source_geo_def, target_geo_def, reduce_data, radius_of_influence):
    """kd_tree.py:389-430."""
    source_lons, source_lats = get_lonlats(source_geo_def)
    source_lons = np.asanyarray(source_lons).ravel()
    source_lats = np.asanyarray(source_lats).ravel()
    if source_lons.size == 0 or source_lats.size == 0:
        raise ValueError('Cannot resample empty data set')
    elif source_lons.size != source_lats.size or source_lons.shape != source_lats.shape:
        raise ValueError('Mismatch between lons and lats')
    with np.errstate(invalid="ignore"):
        valid_input_index = ((source_lons >= -180) & (source_lons <= 180) & (source_lats <= 90) & (source_lats >= -90))
    if reduce_data:
        source_is_griddish = is_grid(source_geo_def) or is_area(source_geo_def)
        target_is_griddish = is_grid(target_geo_def) or is_area(target_geo_def)
        source_is_coord = is_coord(source_geo_def)
        if (source_is_coord or source_is_griddish) and target_is_griddish:
            blons, blats = get_boundary_lonlats(target_geo_def)
            valid_input_index &= get_valid_index_from_lonlat_boundaries(blons, blats, source_lons, source_lats,
                                                                       radius_of_influence)
    if isinstance(valid_input_index, np.ma.core.MaskedArray):
        valid_input_index = valid_input_index.filled(False)
    return valid_input_index, source_lons, source_lats


def _get_valid_output_index(source_geo_def, target_geo_def, target_lons, target_lats, reduce_data, radius_of_influence):
    """kd_tree.py:433-461."""
    valid_output_index = np.ones(target_lons.size, dtype=bool)
    if reduce_data:
        if (is_grid(source_geo_def) or is_area(source_geo_def)) and is_coord(target_geo_def):
            blons, blats = get_boundary_lonlats(source_geo_def)
            valid_output_index = get_valid_index_from_lonlat_boundaries(blons, blats, target_lons, target_lats,
                                                                        radius_of_influence)
            valid_output_index = valid_output_index.astype(bool)
    with np.errstate(invalid="ignore"):
        valid_out = ((target_lons >= -180) & (target_lons <= 180) & (target_lats <= 90) & (target_lats >= -90))
    valid_output_index = (valid_output_index & valid_out)
    if isinstance(valid_output_index, np.ma.MaskedArray):
        valid_output_index = valid_output_index.filled(False)
    return valid_output_index


def _create_resample_kdtree(source_lons, source_lats, valid_input_index):
    """kd_tree.py:464-489."""
    source_lons_valid = source_lons[valid_input_index]
    source_lats_valid = source_lats[valid_input_index]
    input_coords = transform_lonlats(np.asarray(source_lons_valid), np.asarray(source_lats_valid))
    if input_coords.size == 0:
        raise EmptyResult('No valid data points in input data')
    return KDTreeRef(input_coords)


def _query_resample_kdtree(resample_kdtree, source_geo_def, target_geo_def, radius_of_influence, data_slice,
                           neighbours=8, epsilon=0, reduce_data=True):
    """kd_tree.py:492-550."""
    if not isinstance(radius_of_influence, (int, float)):
        raise TypeError('radius_of_influence must be number')
    elif not isinstance(neighbours, int):
        raise TypeError('neighbours must be integer')
    elif not isinstance(epsilon, (int, float)):
        raise TypeError('epsilon must be number')
    target_lons, target_lats = get_lonlats(target_geo_def, data_slice=data_slice, dtype=geo_dtype(source_geo_def))
    target_lons = np.asanyarray(target_lons)
    target_lats = np.asanyarray(target_lats)
    valid_output_index = _get_valid_output_index(source_geo_def, target_geo_def, target_lons.ravel(),
                                                 target_lats.ravel(), reduce_data, radius_of_influence)
    target_lons_valid = np.asarray(target_lons.ravel()[valid_output_index])
    target_lats_valid = np.asarray(target_lats.ravel()[valid_output_index])
    output_coords = transform_lonlats(target_lons_valid, target_lats_valid)
    output_coords = np.asarray(output_coords, dtype=resample_kdtree.data.dtype)
    distance_array, index_array = resample_kdtree.query(output_coords, k=neighbours, eps=epsilon,
                                                        distance_upper_bound=radius_of_influence)
    return valid_output_index, index_array, distance_array


def _create_empty_info(source_geo_def, target_geo_def, neighbours):
    """kd_tree.py:553-563."""
    valid_output_index = np.ones(target_geo_def.size, dtype=bool)
    if neighbours > 1:
        index_array = (np.ones((target_geo_def.size, neighbours), dtype=np.int32) * source_geo_def.size)
        distance_array = np.ones((target_geo_def.size, neighbours))
    else:
        index_array = (np.ones(target_geo_def.size, dtype=np.int32) * source_geo_def.size)
        distance_array = np.ones(target_geo_def.size)
    return valid_output_index, index_array, distance_array


def get_neighbour_info(source_geo_def, target_geo_def, radius_of_influence, neighbours=8, epsilon=0,
                       reduce_data=True, nprocs=1, segments=None):
    """kd_tree.py:281-386."""
    if source_geo_def.size < neighbours:
        warnings.warn('Searching for %s neighbours in %s data points' % (neighbours, source_geo_def.size), stacklevel=2)
    if segments is None:
        cut_off = 3000000
        if target_geo_def.size > cut_off:
            segments = int(target_geo_def.size / cut_off)
        else:
            segments = 1
    valid_input_index, source_lons, source_lats = _get_valid_input_index(source_geo_def, target_geo_def, reduce_data,
                                                                         radius_of_influence)
    try:
        resample_kdtree = _create_resample_kdtree(source_lons, source_lats, valid_input_index)
    except EmptyResult:
        valid_output_index, index_array, distance_array = _create_empty_info(source_geo_def, target_geo_def,
                                                                             neighbours)
        return (valid_input_index, valid_output_index, index_array, distance_array)
    if segments > 1:
        voi, ia, da = [], [], []
        for target_slice in get_slice(segments, target_geo_def.shape):
            next_voi, next_ia, next_da = _query_resample_kdtree(resample_kdtree, source_geo_def, target_geo_def,
                                                                radius_of_influence, target_slice,
                                                                neighbours=neighbours, epsilon=epsilon,
                                                                reduce_data=reduce_data)
            voi.append(next_voi)
            ia.append(next_ia)
            da.append(next_da)
        valid_output_index = np.concatenate(voi)
        index_array = np.concatenate(ia)
        distance_array = np.concatenate(da)
    else:
        valid_output_index, index_array, distance_array = _query_resample_kdtree(
            resample_kdtree, source_geo_def, target_geo_def, radius_of_influence, slice(None),
            neighbours=neighbours, epsilon=epsilon, reduce_data=reduce_data)
    if neighbours > 1:
        if not np.all(np.isinf(distance_array[:, -1])):
            warnings.warn(('Possible more than %s neighbours within %s m for some data points') %
                          (neighbours, radius_of_influence), stacklevel=2)
    return valid_input_index, valid_output_index, index_array, distance_array


# ------------------------------------------------------------------ sampling
def _get_fill_mask_value(data_dtype):
    """kd_tree.py:1186-1195."""
    if issubclass(data_dtype.type, np.floating):
        return np.finfo(data_dtype.type).max
    elif issubclass(data_dtype.type, np.integer):
        return np.iinfo(data_dtype.type).max
    raise TypeError('Type %s is unsupported for masked fill values' % data_dtype.type)


def _remask_data(data, is_to_be_masked=True):
    """kd_tree.py:1198-1211."""
    channels = data.shape[-1]
    if is_to_be_masked:
        mask = data[..., (channels // 2):]
        mask = (mask != 0)
        data = np.ma.array(data[..., :(channels // 2)], mask=mask)
    else:
        data = data[..., :(channels // 2)]
    if data.shape[-1] == 1:
        data = data.reshape(data.shape[:-1])
    return data


def _get_empty_sample(data, output_shape, is_multi_channel, fill_value):
    """kd_tree.py:655-671."""
    if is_multi_channel:
        output_shape = list(output_shape)
        output_shape.append(data.shape[1])
    if fill_value is None:
        return np.ma.array(np.zeros(output_shape, data.dtype), mask=np.ones(output_shape, dtype=bool))
    return np.ones(output_shape, dtype=data.dtype) * fill_value


def _resample_with_weights(new_data, index_array, distance_array, neighbours, input_size, valid_output_index,
                           weight_funcs, with_uncert, fill_value):
    """kd_tree.py:741-818."""
    ch_neighbour_list = []
    index_mask_list = []
    for i in range(neighbours):
        index_ni = index_array[:, i].copy()
        index_mask_ni = (index_ni == input_size)
        index_ni[index_mask_ni] = 0
        ch_neighbour_list.append(new_data[index_ni])
        index_mask_list.append(index_mask_ni)
    weight_list = []
    for i in range(neighbours):
        distance = distance_array[:, i].copy()
        distance[index_mask_list[i]] = 1
        if new_data.ndim == 1:
            weight_list.append(weight_funcs(distance))
            continue
        weights = []
        num_weights = valid_output_index.sum()
        for j in range(new_data.shape[1]):
            calc_weight = weight_funcs[j](distance)
            weights.append(np.ones(num_weights) * calc_weight)
        weight_list.append(np.column_stack(weights))
    result = 0
    norm = 0
    for i in range(neighbours):
        if new_data.ndim > 1:
            inv_index_mask = np.expand_dims(np.invert(index_mask_list[i]), axis=1)
        else:
            inv_index_mask = np.invert(index_mask_list[i])
        weights_tmp = inv_index_mask * weight_list[i]
        result += weights_tmp * ch_neighbour_list[i]
        norm += weights_tmp
    result_valid_index = (norm > 0)
    result[result_valid_index] /= norm[result_valid_index]
    result[np.invert(result_valid_index)] = fill_value
    if with_uncert:
        stddev, count = _calculate_uncertainty(neighbours, new_data, index_mask_list, weight_list,
                                               ch_neighbour_list, result, norm)
        return result, stddev, count
    return result, None, None


def _calculate_uncertainty(neighbours, new_data, index_mask_list, weight_list, ch_neighbour_list, result, norm):
    """kd_tree.py:821-859."""
    count = 0
    norm_sqr = 0
    stddev = 0
    for i in range(neighbours):
        if new_data.ndim > 1:
            inv_index_mask = np.expand_dims(np.invert(index_mask_list[i]), axis=1)
        else:
            inv_index_mask = np.invert(index_mask_list[i])
        weights_tmp = inv_index_mask * weight_list[i]
        count += inv_index_mask
        norm_sqr += weights_tmp ** 2
        values = inv_index_mask * ch_neighbour_list[i]
        stddev += weights_tmp * (values - result) ** 2
    new_valid_index = (count > 1)
    with np.errstate(invalid="ignore", divide="ignore"):
        if stddev.ndim >= 2:
            new_valid_index = new_valid_index[:, 0]
            for i in range(stddev.shape[-1]):
                v1 = norm[new_valid_index, i]
                v2 = norm_sqr[new_valid_index, i]
                stddev[new_valid_index, i] = np.sqrt((v1 / (v1 ** 2 - v2)) * stddev[new_valid_index, i])
                stddev[~new_valid_index, i] = np.nan
        else:
            v1 = norm[new_valid_index]
            v2 = norm_sqr[new_valid_index]
            stddev[new_valid_index] = np.sqrt((v1 / (v1 ** 2 - v2)) * stddev[new_valid_index])
            stddev[~new_valid_index] = np.nan
    return stddev, count


def _prepare_and_fill_uncertainty_result(stddev, count, output_raw_shape, output_shape, valid_output_index,
                                         is_masked_data):
    """kd_tree.py:862-881."""
    full_stddev = np.ones(output_raw_shape) * np.nan
    full_count = np.zeros(output_raw_shape)
    full_stddev[valid_output_index] = stddev
    full_count[valid_output_index] = count
    stddev = full_stddev.reshape(output_shape)
    count = full_count.reshape(output_shape)
    if is_masked_data:
        stddev = _remask_data(stddev, is_to_be_masked=False)
        count = _remask_data(count, is_to_be_masked=False)
    stddev = np.ma.array(stddev, mask=np.isnan(stddev))
    return stddev, count


def _prepare_result(result, output_shape, is_masked_data, use_masked_fill_value, fill_value, input_data_type):
    """kd_tree.py:884-899."""
    result = result.reshape(output_shape)
    if is_masked_data:
        result = _remask_data(result)
    if use_masked_fill_value:
        result = np.ma.masked_equal(result, fill_value)
    if input_data_type is not None:
        result = result.astype(input_data_type)
    return result


def get_sample_from_neighbour_info(resample_type, output_shape, data, valid_input_index, valid_output_index,
                                   index_array, distance_array=None, weight_funcs=None, fill_value=0,
                                   with_uncert=False):
    """kd_tree.py:566-652 and _extract_resample_result :693-738."""
    if data.ndim > 2 and data.shape[0] * data.shape[1] == valid_input_index.size:
        data = data.reshape(data.shape[0] * data.shape[1], data.shape[2])
    elif data.shape[0] != valid_input_index.size:
        data = data.ravel()
    if valid_input_index.size != data.shape[0]:
        raise ValueError('Mismatch between geometry and dataset')
    is_multi_channel = (data.ndim > 1)
    valid_input_size = valid_input_index.sum()
    valid_output_size = valid_output_index.sum()
    if valid_input_size == 0 or valid_output_size == 0:
        return _get_empty_sample(data, output_shape, is_multi_channel, fill_value)
    if not isinstance(data, np.ndarray):
        raise TypeError('data must be numpy array')
    elif valid_input_index.ndim != 1:
        raise TypeError('valid_index must be one dimensional array')
    if resample_type not in ("nn", "custom"):
        raise TypeError(f"Invalid resampling type: {resample_type}")
    if resample_type == "custom" and weight_funcs is None:
        raise ValueError("weight_funcs must be supplied when using custom resampling")
    if index_array.ndim == 1:
        neighbours = 1
    else:
        neighbours = index_array.shape[1]
        if resample_type == 'nn':
            raise ValueError('index_array contains more neighbours than just the nearest')
    new_data = data[valid_input_index]
    input_data_type = new_data.dtype if resample_type == 'nn' else None
    is_masked_data = False
    if np.ma.is_masked(new_data):
        is_masked_data = True
        new_data = np.column_stack((new_data.data, new_data.mask))
    output_shape = tuple(output_shape)
    # _extract_resample_result
    if weight_funcs is not None and is_masked_data:
        weight_funcs = weight_funcs * 2 if is_multi_channel else (weight_funcs,) * 2
    use_masked_fill_value = False
    if fill_value is None:
        use_masked_fill_value = True
        fill_value = _get_fill_mask_value(new_data.dtype)
    if resample_type == 'nn' or neighbours == 1:
        index_mask = (index_array == valid_input_size)
        new_index_array = np.where(index_mask, 0, index_array)
        result = new_data[new_index_array].copy()
        result[index_mask] = fill_value
        stddev = count = None
    else:
        result, stddev, count = _resample_with_weights(new_data, index_array, distance_array, neighbours,
                                                       valid_input_size, valid_output_index, weight_funcs,
                                                       with_uncert, fill_value)
    output_size = output_shape[0] * output_shape[1] if len(output_shape) > 1 else output_shape[0]
    output_raw_shape = (output_size, new_data.shape[1]) if new_data.ndim > 1 else output_size
    full_result = np.full(output_raw_shape, fill_value, dtype=result.dtype)
    full_result[valid_output_index] = result
    result = full_result
    final_output_shape = output_shape + (new_data.shape[1],) if new_data.ndim > 1 else output_shape
    result = _prepare_result(result, final_output_shape, is_masked_data, use_masked_fill_value, fill_value,
                             input_data_type)
    if with_uncert:
        stddev, count = _prepare_and_fill_uncertainty_result(stddev, count, output_raw_shape, final_output_shape,
                                                             valid_output_index, is_masked_data)
        if np.ma.isMA(result):
            stddev = np.ma.array(stddev, mask=(result.mask | stddev.mask))
            count = np.ma.array(count, mask=result.mask)
        return result, stddev, count
    return result


def _resample(source_geo_def, data, target_geo_def, resample_type, radius_of_influence, neighbours=8, epsilon=0,
              weight_funcs=None, fill_value=0, reduce_data=True, nprocs=1, segments=None, with_uncert=False):
    """kd_tree.py:256-278."""
    vii, voi, ia, da = get_neighbour_info(source_geo_def, target_geo_def, radius_of_influence, neighbours=neighbours,
                                          epsilon=epsilon, reduce_data=reduce_data, nprocs=nprocs, segments=segments)
    return get_sample_from_neighbour_info(resample_type, target_geo_def.shape, data, vii, voi, ia, distance_array=da,
                                          weight_funcs=weight_funcs, fill_value=fill_value, with_uncert=with_uncert)


def resample_nearest(source_geo_def, data, target_geo_def, radius_of_influence, epsilon=0, fill_value=0,
                     reduce_data=True, nprocs=1, segments=None):
    """kd_tree.py:64-110."""
    return _resample(source_geo_def, data, target_geo_def, 'nn', radius_of_influence, neighbours=1, epsilon=epsilon,
                     fill_value=fill_value, reduce_data=reduce_data, nprocs=nprocs, segments=segments)


def resample_gauss(source_geo_def, data, target_geo_def, radius_of_influence, sigmas, neighbours=8, epsilon=0,
                   fill_value=0, reduce_data=True, nprocs=1, segments=None, with_uncert=False):
    """kd_tree.py:113-189."""
    def gauss(sigma):
        return lambda r: np.exp(-r ** 2 / float(sigma) ** 2)
    is_multi_channel = False
    try:
        sigmas.__iter__()
        sigma_list = sigmas
        is_multi_channel = True
    except AttributeError:
        sigma_list = [sigmas]
    for sigma in sigma_list:
        if not isinstance(sigma, (int, float)):
            raise TypeError('sigma must be number')
    if is_multi_channel:
        weight_funcs = list(map(gauss, sigma_list))
    else:
        weight_funcs = gauss(sigmas)
    return _resample(source_geo_def, data, target_geo_def, 'custom', radius_of_influence, neighbours=neighbours,
                     epsilon=epsilon, weight_funcs=weight_funcs, fill_value=fill_value, reduce_data=reduce_data,
                     nprocs=nprocs, segments=segments, with_uncert=with_uncert)


def resample_custom(source_geo_def, data, target_geo_def, radius_of_influence, weight_funcs, neighbours=8, epsilon=0,
                    fill_value=0, reduce_data=True, nprocs=1, segments=None, with_uncert=False):
    """kd_tree.py:192-253."""
    import types
    if not isinstance(weight_funcs, (list, tuple)):
        if not isinstance(weight_funcs, types.FunctionType):
            raise TypeError('weight_func must be function object')
    else:
        for weight_func in weight_funcs:
            if not isinstance(weight_func, types.FunctionType):
                raise TypeError('weight_func must be function object')
    return _resample(source_geo_def, data, target_geo_def, 'custom', radius_of_influence, neighbours=neighbours,
                     epsilon=epsilon, weight_funcs=weight_funcs, fill_value=fill_value, reduce_data=reduce_data,
                     nprocs=nprocs, segments=segments, with_uncert=with_uncert)


# ------------------------------------------------------------------ xarray-path block kernels
def query_no_distance(target_lons, target_lats, valid_output_index, mask=None, valid_input_index=None,
                      neighbours=None, epsilon=None, radius=None, kdtree=None):
    """future/resamplers/nearest.py:44-94."""
    voi = valid_output_index
    shape = voi.shape + (neighbours,)
    voir = voi.ravel()
    if mask is not None:
        mask = mask.ravel()[valid_input_index.ravel()]
    target_lons_valid = target_lons.ravel()[voir]
    target_lats_valid = target_lats.ravel()[voir]
    coords = lonlat2xyz(target_lons_valid, target_lats_valid)
    distance_array, index_array = kdtree.query(coords, k=neighbours, eps=epsilon, distance_upper_bound=radius,
                                               mask=mask)
    if neighbours == 1:
        index_array = np.asarray(index_array, dtype=np.int64)
        index_array[index_array >= kdtree.n] = -1
        out = np.full(voi.size, -1, dtype=np.int64)
        out[voir] = index_array
        return out.reshape(voi.shape + (1,))
    if index_array.ndim == 1:
        index_array = index_array[:, None]
    out = np.full(shape, -1, dtype=np.int64)
    out[voi, :] = index_array
    out[out >= kdtree.n] = -1
    return out


def _my_index(index_arr, vii, data_arr, vii_slices=None, ia_slices=None, fill_value=np.nan):
    """future/resamplers/nearest.py:97-108."""
    vii_slices = tuple(x if x is not None else vii.ravel() for x in vii_slices)
    mask_slices = tuple(x if x is not None else (index_arr == -1) for x in ia_slices)
    ia_slices = tuple(x if x is not None else index_arr for x in ia_slices)
    res = data_arr[vii_slices][ia_slices]
    res[mask_slices] = fill_value
    return res
