"""``KDTreeNearestXarrayResampler`` on the GPU (pyresample/future/resamplers/nearest.py:115-546) - the 2.0 resampler
interface over the same device index and gather kernels as ``kd_tree``.

The reference builds a lazy dask graph of ``query_no_distance`` / ``_my_index`` blocks.  Here ``precompute`` runs one
eager ``prs_query`` over the whole target and ``resample`` one ``prs_gather_nearest``; what is kept:

* the tree is float64 over the range-valid source pixels, a ``mask`` (True = invalid data pixel) excludes pixels
  from the search, ``index_array`` is int64 ``(rows, cols, 1)`` with -1 for "no neighbour" (nearest.py:44-94);
* ``resample`` accepts what the reference accepts - 2-D numpy arrays, dask arrays, DataArrays with any extra
  dimensions - and returns the same kind of object (nearest.py:472-504); DataArray results carry the non-geo
  coordinates, ``y``/``x`` projection coordinates and ``attrs['area']`` (resampler.py:173-201; the ``crs``
  coordinate needs pyproj and is left out);
* the default ``radius_of_influence`` (geocentric resolutions, 10 km fallback), the default data mask for swath
  sources, the integer fill value rule, the internal cache keyed by (mask, neighbours, radius, epsilon), and the
  error messages of the reference's input checks.

No dask graph is built, so the reference's ``PerformanceWarning`` for numpy input does not apply.
"""
from __future__ import annotations

import hashlib
import warnings
from copy import deepcopy
from logging import getLogger

import numpy as np

from . import geometry
from . import xarray_nn
from .xarray_nn import _materialise, _my_index

logger = getLogger(__name__)

try:
    from xarray import DataArray
except ImportError:  # pragma: no cover - xarray is optional here: numpy / duck-typed DataArrays still work
    DataArray = None


def _is_data_array(obj):
    return hasattr(obj, "dims") and hasattr(obj, "data") and not isinstance(obj, np.ndarray)


class Resampler(object):
    """Base resampler class (pyresample/future/resamplers/resampler.py:204-270)."""

    def __init__(self, source_geo_def, target_geo_def, cache=None):
        self.source_geo_def = source_geo_def
        self.target_geo_def = target_geo_def
        self.cache = cache

    @property
    def version(self) -> str:
        raise NotImplementedError()

    def precompute(self, **kwargs):
        return None

    def resample(self, data):
        raise NotImplementedError()


class KDTreeNearestXarrayResampler(Resampler):
    """Resampler using the basic nearest neighbor algorithm (nearest.py:115-546)."""

    def __init__(self, source_geo_def, target_geo_def, cache=None):
        super().__init__(source_geo_def, target_geo_def, cache=cache)
        self._internal_cache: dict[tuple, dict] = {}
        if self.target_geo_def.ndim != 2:
            raise ValueError("Target area definition must be 2 dimensions")

    @property
    def version(self) -> str:
        return "0.1"

    def _compute_radius_of_influence(self):
        """Estimate a good default radius_of_influence (nearest.py:146-169)."""
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
        if np.isnan(src_res):
            radius_of_influence = dst_res
        elif np.isnan(dst_res):
            radius_of_influence = src_res
        else:
            radius_of_influence = max(src_res, dst_res)
        if np.isnan(radius_of_influence):
            logger.warning("Could not calculate radius_of_influence, falling back to 10000 meters. This may produce "
                           "lower quality results than expected.")
            radius_of_influence = 10000
        return radius_of_influence

    def _get_neighbor_info(self, mask, neighbors, radius_of_influence, epsilon):
        """``(valid_input_index, index_array)`` (nearest.py:214-243)."""
        if self.source_geo_def.size < neighbors:
            warnings.warn('Searching for %s neighbors in %s data points' % (neighbors, self.source_geo_def.size),
                          stacklevel=3)
        if mask is not None and tuple(mask.shape) != tuple(self.source_geo_def.shape):
            raise ValueError("'mask' must be the same shape as the source geo definition")
        nn = xarray_nn.XArrayResamplerNN.__new__(xarray_nn.XArrayResamplerNN)
        nn.source_geo_def, nn.target_geo_def = self.source_geo_def, self.target_geo_def
        nn.neighbours, nn.epsilon, nn.radius_of_influence = neighbors, epsilon, radius_of_influence
        valid_input_index, _, index_array, _ = nn.get_neighbour_info(mask=mask)
        return np.asarray(_materialise(valid_input_index)), np.asarray(_materialise(index_array))

    def _get_src_geo_dims(self):
        if geometry.is_swath(self.source_geo_def) and hasattr(self.source_geo_def.lons, "dims"):
            return tuple(self.source_geo_def.lons.dims)  # could be 1D or 2D
        if geometry.is_swath(self.source_geo_def) and self.source_geo_def.ndim == 1:
            return ('y',)
        return ('y', 'x')

    def _verify_data_geo_dims(self, data, src_geo_dims):
        """nearest.py:364-384."""
        data_geo_dims = tuple(d for d in data.dims if d in src_geo_dims)
        if data_geo_dims != tuple(src_geo_dims):
            raise ValueError("Data dimensions do not match source area dimensions.")
        first_dim_idx = tuple(data.dims).index(src_geo_dims[0])
        num_dims = len(src_geo_dims)
        if tuple(data.dims[first_dim_idx:first_dim_idx + num_dims]) != data_geo_dims:
            raise ValueError("Data's geolocation dimensions are not consecutive.")
        for dim_offset, dim_name in enumerate(data_geo_dims):
            geom_size = self.source_geo_def.shape[dim_offset]
            data_size = data.shape[first_dim_idx + dim_offset]
            if geom_size != data_size:
                raise ValueError("Input data shape is not equal to the shape of the source geometry: "
                                 f"{dim_name}={data_size} versus {geom_size}")

    @staticmethod
    def _mask_key(mask):
        if mask is None:
            return None
        m = np.ascontiguousarray(np.asarray(_materialise(mask), dtype=bool))
        return hashlib.sha1(m.tobytes()).hexdigest() + str(m.shape)

    def precompute(self, mask=None, radius_of_influence=None, epsilon=0):
        """Generate neighbor indexes from the geolocation and an optional data mask (nearest.py:395-429)."""
        neighbors = 1
        if mask is not None and tuple(mask.shape) != tuple(self.source_geo_def.shape):
            raise ValueError("'mask' provided to 'precompute' is not the same shape as the source geometry.")
        if radius_of_influence is None:
            radius_of_influence = self._compute_radius_of_influence()
        key = (self._mask_key(mask), neighbors, radius_of_influence, epsilon)
        if key not in self._internal_cache:
            valid_input_index, index_arr = self._get_neighbor_info(mask, neighbors, radius_of_influence, epsilon)
            self._internal_cache[key] = {"valid_input_index": valid_input_index, "index_array": index_arr}
        return key

    def get_sample_from_neighbor_info(self, data, valid_input_index, index_array, neighbors=1, fill_value=np.nan):
        """Get the pixels matching the target area (nearest.py:245-362); ``data`` is DataArray-like."""
        from . import kd_tree
        if fill_value is not None and np.isnan(fill_value) and np.issubdtype(data.dtype, np.integer):
            fill_value = kd_tree._get_fill_mask_value(np.dtype(data.dtype))
            logger.warning("Fill value incompatible with integer data using {:d} instead.".format(fill_value))
        if neighbors > 1:
            raise NotImplementedError("Nearest neighbor resampling can not handle more than 1 neighbor yet.")
        ia = np.asarray(_materialise(index_array))[:, :, 0]
        vii = np.asarray(_materialise(valid_input_index)).ravel()
        src_geo_dims = self._get_src_geo_dims()
        dst_geo_dims = ('y', 'x')
        self._verify_data_geo_dims(data, src_geo_dims)
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
                flat_src_shape.append(dict(zip(data.dims, data.shape))[dim])
                vii_slices.append(slice(None))
                ia_slices.append(slice(None))
                dst_dims.append(dim)
        new_data = np.asarray(_materialise(data.data)).reshape(flat_src_shape)
        res = _my_index(ia, vii, new_data, vii_slices=vii_slices, ia_slices=ia_slices, fill_value=fill_value)
        return self._new_data_array(data, res, dst_dims)

    def _new_data_array(self, old, values, dims):
        """New DataArray with the coordinates of resampler.py:173-201 (without the pyproj ``crs`` coordinate)."""
        ignore = ('y', 'x', 'crs')
        coords = {}
        for cname, cval in dict(getattr(old, "coords", {}) or {}).items():
            cdims = getattr(cval, "dims", ())
            if cname in ignore or any(d in cdims for d in ignore):
                continue
            coords[cname] = cval
        try:
            coord_x, coord_y = self.target_geo_def.get_proj_vectors()
            coords['y'], coords['x'] = coord_y, coord_x
        except AttributeError:
            logger.debug("No geo coordinates created")
        attrs = deepcopy(dict(getattr(old, "attrs", {}) or {}))
        attrs['area'] = self.target_geo_def
        cls = DataArray if DataArray is not None and isinstance(old, DataArray) else type(old)
        return cls(values, dims=tuple(dims), coords=coords, attrs=attrs)

    def resample(self, data, mask_area=None, fill_value=np.nan, radius_of_influence=None, epsilon=0):
        """Resample ``data`` from the source geometry to the target geometry (nearest.py:431-470)."""
        new_data = self._verify_input_object_type(data)
        src_geo_dims = self._get_src_geo_dims()
        self._verify_data_geo_dims(new_data, src_geo_dims)
        mask = self._get_area_mask(mask_area, new_data)
        if radius_of_influence is None:
            radius_of_influence = self._compute_radius_of_influence()
        key = self.precompute(mask=mask, radius_of_influence=radius_of_influence, epsilon=epsilon)
        info = self._internal_cache[key]
        result = self.get_sample_from_neighbor_info(new_data, info["valid_input_index"], info["index_array"],
                                                    fill_value=fill_value)
        return self._verify_result_object_type(result, data)

    def _verify_input_object_type(self, data):
        """nearest.py:472-494: plain arrays must be 2-D and get the source geometry's dimension names."""
        if _is_data_array(data):
            return data
        if data.ndim != 2:
            raise ValueError(f"{self.__class__.__name__} requires DataArrays with a dask array. Input array is not "
                             "2D and can't be automatically converted.")
        return _PlainDataArray(data, dims=self._get_src_geo_dims())

    def _verify_result_object_type(self, data, orig_data):
        """nearest.py:496-504: the result is the kind of object that came in."""
        if _is_data_array(orig_data):
            return data
        values = data.data
        if hasattr(orig_data, "compute") and xarray_nn.da is not None:  # a dask array came in
            return xarray_nn.da.from_array(values, chunks=getattr(orig_data, "chunksize", "auto"))
        return np.asarray(values)

    def _get_area_mask(self, mask_area, data):
        """nearest.py:506-516."""
        if isinstance(mask_area, np.ndarray) or _is_data_array(mask_area) or hasattr(mask_area, "compute"):
            return mask_area
        if mask_area is None and geometry.is_swath(self.source_geo_def):
            mask_area = True
        if mask_area:
            return self.compute_data_mask(data)
        return None

    def compute_data_mask(self, data):
        """Mask array where the data are invalid on every non-geo position (nearest.py:518-534)."""
        geo_dims = self._get_src_geo_dims()
        values = np.asarray(_materialise(data.data))
        if np.issubdtype(values.dtype, np.integer):
            fill = dict(getattr(data, "attrs", {}) or {}).get('_FillValue', np.iinfo(values.dtype.type).max)
            mask = values == fill
        else:
            mask = np.isnan(values)
        flat_axes = tuple(i for i, dim in enumerate(data.dims) if dim not in geo_dims)
        if flat_axes:
            mask = mask.all(axis=flat_axes)
        return mask


class _PlainDataArray(object):
    """Dimension names around a plain array (what ``xr.DataArray(data, dims=...)`` gives the reference)."""

    def __init__(self, data, dims=None, coords=None, attrs=None):
        self.data = data
        self.dims = tuple(dims)
        self.coords = dict(coords or {})
        self.attrs = dict(attrs or {})

    @property
    def shape(self):
        return tuple(self.data.shape)

    @property
    def ndim(self):
        return self.data.ndim

    @property
    def dtype(self):
        return self.data.dtype

    @property
    def values(self):
        return np.asarray(_materialise(self.data))
