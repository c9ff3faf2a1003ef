"""Caller conveniences around the kd-tree path (pyresample/utils/__init__.py:102-183), SURVEY 8 f4."""
from __future__ import annotations

import numpy as np


def fwhm2sigma(fwhm):
    """Sigma of the gauss weight function from the FWHM (3 dB level) (utils/__init__.py:161-174)."""
    return fwhm / (2 * np.sqrt(np.log(2)))


def _downcast_index_array(index_array, size):
    """Try to downcast array to uint16; out-of-range entries become ``size`` (utils/__init__.py:177-183)."""
    if size <= np.iinfo(np.uint16).max:
        mask = (index_array < 0) | (index_array >= size)
        index_array[mask] = size
        index_array = index_array.astype(np.uint16)
    return index_array


def generate_nearest_neighbour_linesample_arrays(source_area_def, target_area_def, radius_of_influence, nprocs=1):
    """Row / column index arrays for nearest-neighbour grid resampling (utils/__init__.py:102-158).

    The neighbour search is ``kd_tree.get_neighbour_info`` on the GPU; pixels without a neighbour hold -1 (or the
    source dimension size after the uint16 downcast).
    """
    from .kd_tree import get_neighbour_info
    valid_input_index, valid_output_index, index_array, distance_array = get_neighbour_info(
        source_area_def, target_area_def, radius_of_influence, neighbours=1, nprocs=nprocs)
    rows = np.fromfunction(lambda i, j: i, source_area_def.shape, dtype=np.int32).ravel()
    cols = np.fromfunction(lambda i, j: j, source_area_def.shape, dtype=np.int32).ravel()
    rows_valid = rows[valid_input_index]
    cols_valid = cols[valid_input_index]
    number_of_valid_points = valid_input_index.sum()
    full = index_array
    if not np.all(valid_output_index):  # get_neighbour_info returns rows for valid target points only
        full = np.full(valid_output_index.size, number_of_valid_points, dtype=index_array.dtype)
        full[valid_output_index] = index_array
    index_mask = (full == number_of_valid_points)
    full = np.where(index_mask, 0, full)
    row_sample = rows_valid[full] if number_of_valid_points else np.zeros(full.shape, np.int32)
    col_sample = cols_valid[full] if number_of_valid_points else np.zeros(full.shape, np.int32)
    row_sample[index_mask] = -1
    col_sample[index_mask] = -1
    row_indices = row_sample.reshape(target_area_def.shape)
    col_indices = col_sample.reshape(target_area_def.shape)
    row_indices = _downcast_index_array(row_indices, source_area_def.shape[0])
    col_indices = _downcast_index_array(col_indices, source_area_def.shape[1])
    return row_indices, col_indices
