"""``XArrayResamplerNN`` and its block kernels on the GPU (pyresample/kd_tree.py:902-1183,
pyresample/future/resamplers/nearest.py:44-108).

The reference builds a lazy dask graph whose blocks call ``query_no_distance`` (pykdtree) and ``_my_index``
(numpy fancy indexing).  Here the neighbour search runs eagerly in one ``prs_query`` launch over the whole target
(the GPU needs no chunking) and the sampling is one ``prs_gather_nearest`` launch; results are handed back as
``xarray.DataArray`` (wrapped in dask arrays when dask is importable, so callers can keep calling ``.compute()``).
xarray is required exactly like in the reference (``ImportError`` otherwise, kd_tree.py:931-932).

Contract kept from the reference: the tree is float64 (kd_tree.py:979), ``valid_input_index`` is the 2-D range mask
(no reduce_data), ``index_array`` is int64 ``(rows, cols, neighbours)`` with ``-1`` for "no neighbour", no distances
are returned (kd_tree.py:1004), sampling supports ``neighbours == 1`` only (kd_tree.py:1082-1084).
"""
from __future__ import annotations

import ctypes
import warnings
from copy import deepcopy
from logging import getLogger

import numpy as np

from . import CHUNK_SIZE, _cabi, geometry

logger = getLogger(__name__)

try:
    from xarray import DataArray
except ImportError:  # pragma: no cover
    DataArray = None
try:
    import dask.array as da
except ImportError:  # pragma: no cover
    da = None


def _kd():
    from . import kd_tree
    return kd_tree


def _materialise(arr):
    """xarray / dask / numpy / torch -> numpy or torch (device arrays stay where they are)."""
    if hasattr(arr, "data") and hasattr(arr, "dims"):
        arr = arr.data
    if hasattr(arr, "compute"):
        arr = arr.compute()
    return arr


class _GpuTree(object):
    """What ``query_no_distance`` receives as ``kdtree``: the device index over the valid source points."""

    def __init__(self, source):
        self.source = source
        self.n = int(source.n_valid)


def build_tree(source_lons, source_lats, mask=None):
    """Range-valid mask + float64 device index (kd_tree.py:970-981).  ``mask`` (True = exclude) is pykdtree's
    query ``mask=`` (nearest.py:56-67); it is applied when the index is built."""
    kd = _kd()
    lons, lats = _materialise(source_lons), _materialise(source_lats)
    swath = geometry.SwathDefinition(lons, lats)
    excl = None
    if mask is not None:
        excl = np.ascontiguousarray(np.asarray(_materialise(mask), dtype=bool)).ravel()
    source = kd._Source(swath, None, False, 0.0, 1, exclude_mask=excl, tree_dtype=np.float64)
    valid = source.valid.cpu().numpy().astype(bool).reshape(swath.shape)
    return valid, _GpuTree(source)


def query_no_distance(target_lons, target_lats, valid_output_index, mask=None, valid_input_index=None,
                      neighbours=None, epsilon=None, radius=None, kdtree=None):
    """Query the index; no distances are returned (nearest.py:44-94).

    ``kdtree`` is the object returned by :func:`build_tree`.  A source ``mask`` must be given to
    :func:`build_tree` (the exclusion is part of the device index); passing a different one here raises.
    """
    kd = _kd()
    torch = kd._torch()
    if mask is not None and getattr(kdtree, "mask_applied", None) is not True:
        # rebuild with the mask (pykdtree applies it per query; the device index applies it at build time)
        src = kdtree.source
        excl = np.ascontiguousarray(np.asarray(_materialise(mask), dtype=bool)).ravel()
        source = kd._Source(src.geo_def, None, False, 0.0, 1, exclude_mask=excl, tree_dtype=np.float64)
        kdtree = _GpuTree(source)
    voi = np.asarray(_materialise(valid_output_index), dtype=bool)
    tlons = np.asarray(_materialise(target_lons))
    tlats = np.asarray(_materialise(target_lats))
    shape = voi.shape + (neighbours,)
    target = geometry.SwathDefinition(tlons.ravel(), tlats.ravel())
    source = kdtree.source
    out = np.full((voi.size, neighbours), -1, dtype=np.int64)
    if source.index is not None:
        idx, _, valid, _, _ = kd._query_device(source, target, float(radius), int(neighbours), False, want_dist=False)
        idx = idx.cpu().numpy().view(np.uint32).astype(np.int64)
        idx[idx >= kdtree.n] = -1
        idx[~(voi.ravel() & valid.cpu().numpy().astype(bool))] = -1
        out = idx
    return out.reshape(shape)


def _my_index(index_arr, vii, data_arr, vii_slices=None, ia_slices=None, fill_value=np.nan):
    """``data_arr[vii][index_arr]`` with ``-1 -> fill_value`` along the flattened geo dimension (nearest.py:97-108),
    as one device gather."""
    kd = _kd()
    torch = kd._torch()
    L = _cabi.lib()
    axis = [i for i, s in enumerate(vii_slices) if s is None][0]
    data = np.moveaxis(np.asarray(data_arr), axis, 0)
    lead = data.shape[0]
    rest = data.shape[1:]
    rows = np.ascontiguousarray(data.reshape(lead, -1))
    vii = np.asarray(vii, dtype=bool).ravel()
    ia = np.asarray(index_arr)
    n_valid = int(vii.sum())
    out_dtype = rows.dtype
    fill_row = np.full((1, rows.shape[1]), fill_value, dtype=out_dtype)
    idx = np.where(ia.ravel() < 0, n_valid, ia.ravel()).astype(np.uint32)
    dev_rows = kd._dev(kd._byte_rows(rows))
    row_bytes = dev_rows.shape[1]
    r2s = None
    if n_valid != vii.size:
        r2s = torch.empty(n_valid, dtype=torch.int32, device="cuda")
        vt = kd._dev(vii)
        _cabi.check(L.prs_rank_to_source(kd._ptr(vt), vt.numel(), kd._ptr(r2s), None, kd._stream()))
    idx_t = kd._dev(idx.view(np.int32))
    fill_t = kd._dev(kd._byte_rows(fill_row))
    out = torch.empty((idx.size, row_bytes), dtype=torch.uint8, device="cuda")
    _cabi.check(L.prs_gather_nearest(kd._ptr(idx_t), idx.size, n_valid, kd._ptr(r2s), kd._ptr(dev_rows), row_bytes,
                                     kd._ptr(fill_t), kd._ptr(out), kd._stream()))
    res = kd._host(out).view(out_dtype).reshape(ia.shape + rest)
    # put the (now possibly multi-dimensional) geo axes back where the flattened axis was
    nd = ia.ndim
    order = list(range(nd, nd + axis)) + list(range(nd)) + list(range(nd + axis, res.ndim))
    return np.transpose(res, order)


class XArrayResamplerNN(object):
    """Resampler for xarray DataArrays with the nearest-neighbour algorithm (pyresample/kd_tree.py:902-1183)."""

    def __init__(self, source_geo_def, target_geo_def, radius_of_influence=None, neighbours=1, epsilon=0):
        if _kd().DataArray is None:
            raise ImportError("Missing 'xarray' and 'dask' dependencies")
        self.valid_input_index = None
        self.valid_output_index = None
        self.index_array = None
        self.distance_array = None
        self.delayed_kdtree = None
        self.neighbours = neighbours
        self.epsilon = epsilon
        self.source_geo_def = source_geo_def
        self.target_geo_def = target_geo_def
        if radius_of_influence is None:
            radius_of_influence = self._compute_radius_of_influence()
        self.radius_of_influence = radius_of_influence
        if self.target_geo_def.ndim != 2:
            raise ValueError("Target area definition must be 2 dimensions")

    def _compute_radius_of_influence(self):
        """Estimate a good default radius_of_influence (kd_tree.py:949-968)."""
        try:
            src_res = self.source_geo_def.geocentric_resolution()
        except RuntimeError:
            logger.warning("Could not calculate source definition resolution")
            src_res = np.nan
        try:
            dst_res = self.target_geo_def.geocentric_resolution()
        except RuntimeError:
            logger.warning("Could not calculate destination definition resolution")
            dst_res = np.nan
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            radius_of_influence = np.nanmax([src_res, dst_res])
        if np.isnan(radius_of_influence):
            logger.warning("Could not calculate radius_of_influence, falling back to 10000 meters. This may produce "
                           "lower quality results than expected.")
            radius_of_influence = 10000
        return radius_of_influence

    def _wrap(self, arr):
        if da is not None:
            return da.from_array(arr, chunks=CHUNK_SIZE)
        return arr

    def get_neighbour_info(self, mask=None):
        """Return neighbour info (kd_tree.py:1006-1043)."""
        kd = _kd()
        if self.source_geo_def.size < self.neighbours:
            warnings.warn('Searching for %s neighbours in %s data points' %
                          (self.neighbours, self.source_geo_def.size), stacklevel=2)
        if mask is not None:
            if tuple(mask.shape) != tuple(self.source_geo_def.shape):
                raise ValueError("'mask' must be the same shape as the source geo definition")
        src = geometry.as_geo_def(self.source_geo_def)
        excl = None
        if mask is not None:
            excl = np.ascontiguousarray(np.asarray(_materialise(mask), dtype=bool)).ravel()
        if geometry.is_area(src):
            source = kd._Source(src, None, False, 0.0, self.neighbours, exclude_mask=excl, tree_dtype=np.float64)
        else:
            lons, lats = src.get_lonlats()
            source = kd._Source(geometry.SwathDefinition(_materialise(lons), _materialise(lats)), None, False, 0.0,
                                self.neighbours, exclude_mask=excl, tree_dtype=np.float64)
        self.delayed_kdtree = _GpuTree(source)
        valid_input = source.valid.cpu().numpy().astype(bool).reshape(tuple(src.shape))
        tgt = geometry.as_geo_def(self.target_geo_def)
        rows, cols = tgt.shape
        k = int(self.neighbours)
        if source.index is not None:
            idx, _, valid, _, _ = kd._query_device(source, tgt, float(self.radius_of_influence), k, False,
                                                   want_dist=False)
            valid_out = valid.cpu().numpy().astype(bool).reshape(rows, cols)
            ia = kd._host(idx).view(np.uint32).astype(np.int64)
            ia[ia >= source.n_valid] = -1
            ia[~valid_out.ravel()] = -1
            ia = ia.reshape(rows, cols, k)
        else:
            tl, tt = tgt.get_lonlats()
            tl, tt = np.asarray(_materialise(tl)), np.asarray(_materialise(tt))
            with np.errstate(invalid="ignore"):
                valid_out = ((tl >= -180) & (tl <= 180) & (tt <= 90) & (tt >= -90))
            ia = np.full((rows, cols, k), -1, dtype=np.int64)
        self.valid_input_index = self._wrap(valid_input)
        self.valid_output_index = self._wrap(valid_out)
        self.index_array = self._wrap(ia)
        self.distance_array = None
        return (self.valid_input_index, self.valid_output_index, self.index_array, self.distance_array)

    def _get_valid_dims(self, data):
        """kd_tree.py:1166-1183."""
        if geometry.is_swath(self.source_geo_def) and hasattr(self.source_geo_def.lons, "dims"):
            src_geo_dims = tuple(self.source_geo_def.lons.dims)
        else:
            src_geo_dims = ('y', 'x')
        dst_geo_dims = ('y', 'x')
        data_geo_dims = tuple(d for d in data.dims if d in src_geo_dims)
        if data_geo_dims != src_geo_dims:
            raise ValueError("Data dimensions do not match source area dimensions")
        first_dim_idx = data.dims.index(src_geo_dims[0])
        num_dims = len(src_geo_dims)
        if tuple(data.dims[first_dim_idx:first_dim_idx + num_dims]) != data_geo_dims:
            raise ValueError("Data's geolocation dimensions are not consecutive.")
        return src_geo_dims, dst_geo_dims

    def get_sample_from_neighbour_info(self, data, fill_value=np.nan):
        """Get the pixels matching the target area (kd_tree.py:1045-1164)."""
        kd = _kd()
        if fill_value is not None and np.isnan(fill_value) and np.issubdtype(data.dtype, np.integer):
            fill_value = kd._get_fill_mask_value(np.dtype(data.dtype))
            logger.warning("Fill value incompatible with integer data using {:d} instead.".format(fill_value))
        if self.neighbours > 1:
            raise NotImplementedError("Nearest neighbor resampling can not handle more than 1 neighbor yet.")
        ia = np.asarray(_materialise(self.index_array))[:, :, 0]
        vii = np.asarray(_materialise(self.valid_input_index)).ravel()
        src_geo_dims, dst_geo_dims = self._get_valid_dims(data)

        def contain_coords(var, coord_list):
            return bool(set(coord_list).intersection(set(getattr(var, "dims", ()))))

        coords = {c: c_var for c, c_var in data.coords.items()
                  if not contain_coords(c_var, src_geo_dims + dst_geo_dims)}
        try:
            coord_x, coord_y = self.target_geo_def.get_proj_vectors()
            coords['y'] = coord_y
            coords['x'] = coord_x
        except AttributeError:
            logger.debug("No geo coordinates created")
        flat_src_shape, vii_slices, ia_slices, dst_dims = [], [], [], []
        geo_handled = False
        for dim in data.dims:
            if dim in src_geo_dims and not geo_handled:
                flat_src_shape.append(-1)
                vii_slices.append(None)
                ia_slices.append(None)
                dst_dims.extend(dst_geo_dims)
                geo_handled = True
            elif dim not in src_geo_dims:
                flat_src_shape.append(data.sizes[dim])
                vii_slices.append(slice(None))
                ia_slices.append(slice(None))
                dst_dims.append(dim)
        new_data = np.asarray(_materialise(data.data)).reshape(flat_src_shape)
        res = _my_index(ia, vii, new_data, vii_slices=vii_slices, ia_slices=ia_slices, fill_value=fill_value)
        res = self._wrap(res)
        return kd.DataArray(res, dims=dst_dims, coords=coords, attrs=deepcopy(data.attrs))
