"""Caller conveniences around the kd-tree path: the row / column "linesample" arrays that
``pyresample.utils.generate_nearest_neighbour_linesample_arrays`` (utils/__init__.py:102-158) derives from
neighbour info, produced here straight from the device search result."""
from __future__ import annotations

import math

import numpy as np

_FWHM_TO_SIGMA = 1.0 / (2.0 * math.sqrt(math.log(2.0)))


def fwhm2sigma(fwhm):
    """``sigma`` of ``exp(-r^2 / sigma^2)`` for a full width at half maximum ``fwhm`` (utils/__init__.py:161-174)."""
    return fwhm * _FWHM_TO_SIGMA


def _downcast_index_array(index_array, size):
    """uint16 form of an index array along an axis of length ``size`` (utils/__init__.py:177-183): when every
    in-range index fits, entries outside ``[0, size)`` (the -1 "no neighbour" marks) are stored as ``size``.
    Works in place on numpy arrays and on torch tensors."""
    if size > np.iinfo(np.uint16).max:
        return index_array
    outside = (index_array < 0) | (index_array >= size)
    index_array[outside] = size
    if isinstance(index_array, np.ndarray):
        return index_array.astype(np.uint16)
    import torch
    return index_array.to(torch.uint16)


def generate_nearest_neighbour_linesample_arrays(source_area_def, target_area_def, radius_of_influence, nprocs=1):
    """(row_indices, col_indices) of the nearest source pixel for every target pixel, -1 where there is none
    (``size`` after the uint16 downcast) - same values as utils/__init__.py:102-158.

    The reference enumerates the source rows / columns, compacts them by ``valid_input_index`` and gathers them by
    ``index_array`` on the host.  Here the search result never leaves the device as neighbour info: the index of
    the nearest *source pixel* is split into ``// width`` and ``% width`` on the GPU and only the two result planes
    are read back.
    """
    from . import geometry, kd_tree
    torch = kd_tree._torch()
    del nprocs  # one GPU pass whatever the value
    source = kd_tree._Source(source_area_def, target_area_def, True, radius_of_influence, 1, need_ranks=True)
    tgt = geometry.as_geo_def(target_area_def)
    src_rows, src_cols = (int(v) for v in source_area_def.shape)
    if source.index is None:  # every source point reduced away: no pixel has a neighbour
        none = np.full(tuple(tgt.shape), -1, dtype=np.int32)
        return _downcast_index_array(none.copy(), src_rows), _downcast_index_array(none, src_cols)
    kd_tree._check_query_args(tgt, radius_of_influence, 1, 0)
    rank, _, valid_out, _, _ = kd_tree._query_device(source, tgt, radius_of_influence, 1, True, want_dist=False)
    rank = rank.reshape(-1).to(torch.int64) & 0xFFFFFFFF
    found = (rank < int(source.n_valid)) & valid_out.bool()
    rank_to_source = kd_tree._rank_to_source(source)
    pixel = rank.clamp_(max=max(int(source.n_valid) - 1, 0))
    if rank_to_source is not None:
        pixel = rank_to_source.to(torch.int64)[pixel] & 0xFFFFFFFF
    minus_one = torch.full_like(pixel, -1)
    row = torch.where(found, torch.div(pixel, src_cols, rounding_mode="floor"), minus_one).to(torch.int32)
    col = torch.where(found, pixel % src_cols, minus_one).to(torch.int32)
    row = _downcast_index_array(row.reshape(tuple(tgt.shape)), src_rows)
    col = _downcast_index_array(col.reshape(tuple(tgt.shape)), src_cols)
    return kd_tree._host(row.contiguous()), kd_tree._host(col.contiguous())
