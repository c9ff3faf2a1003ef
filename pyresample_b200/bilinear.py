"""Bilinear resampling for irregular grids on the GPU neighbour index - drop-in for ``pyresample.bilinear``'s numpy API
(pyresample/bilinear/_numpy_resampler.py, pyresample/bilinear/_base.py).

Same names, arguments and attributes as the reference: ``NumpyBilinearResampler`` (``get_bil_info``,
``get_sample_from_bil_info``, ``resample``; ``bilinear_t``, ``bilinear_s``, ``slices_x``, ``slices_y``, ``mask_slices``,
``out_coords_x``, ``out_coords_y``), the deprecated ``NumpyResamplerBilinear`` and the legacy functions ``get_bil_info``,
``get_sample_from_bil_info`` and ``resample_bilinear``.  Underneath:

* the k = 32 neighbour search is ``prs_query`` on the same cube-face index as the kd_tree path, with the float64 tree
  the reference uses here (``_base.py:240-249``);
* the source points are projected into the target CRS once (``prs_project_forward`` instead of
  ``Proj(target.proj_str)(lon, lat)`` on the 32 x expanded arrays of ``_base.py:299-307``);
* corner selection and the fractional distances run per pixel in ``prs_bilinear_info``; sampling in
  ``prs_bilinear_sample``.

Differences from the reference: source coordinates that are NaN count as invalid (the reference's range test lets NaN
through and hands NaN coordinates to pykdtree, ``_base.py:382-384, 597-608``); ``nprocs`` / ``segments`` are accepted
and ignored.
"""
from __future__ import annotations

import ctypes
import warnings

import numpy as np

from . import _cabi, geometry, kd_tree


class BilinearBase(object):
    """Base class for bilinear interpolation (pyresample/bilinear/_base.py:44-262)."""

    def __init__(self, source_geo_def, target_geo_def, radius_of_influence, neighbours=32, epsilon=0, reduce_data=True):
        self.bilinear_t = None
        self.bilinear_s = None
        self.slices_x = None
        self.slices_y = None
        self.mask_slices = None
        self.out_coords_x = None
        self.out_coords_y = None
        self._out_coords = {'x': self.out_coords_x, 'y': self.out_coords_y}
        self._valid_input_index = None
        self._index_array = None
        self._distance_array = None
        self._neighbours = neighbours
        self._epsilon = epsilon
        self._reduce_data = reduce_data
        self._source_geo_def = source_geo_def
        self._target_geo_def = target_geo_def
        self._radius_of_influence = radius_of_influence
        self._resample_kdtree = None
        self._target_lons = None
        self._target_lats = None
        self._device = None  # (idx4, t, s, rank_to_source) kept on the GPU for sampling

    def resample(self, data, fill_value=None, nprocs=1):
        raise NotImplementedError

    # ------------------------------------------------------------------ info
    def get_bil_info(self, kdtree_class=None, nprocs=1):
        """Calculate bilinear neighbour info (_base.py:99-119)."""
        torch = kd_tree._torch()
        L = _cabi.lib()
        src, tgt = geometry.as_geo_def(self._source_geo_def), geometry.as_geo_def(self._target_geo_def)
        if src.size < self._neighbours:
            warnings.warn('Searching for %s neighbours in %s data points' % (self._neighbours, src.size), stacklevel=2)
        if not (geometry.is_area(tgt) and getattr(tgt, "lons", None) is None):
            raise NotImplementedError("bilinear resampling needs an AreaDefinition target (projection coordinates)")
        k = int(self._neighbours)
        # float64 tree from coordinates computed in the source dtype (lonlat2xyz(...).astype(float64), _base.py:240-249)
        source = kd_tree._Source(src, tgt, self._reduce_data, self._radius_of_influence, k, need_ranks=True,
                                 tree_dtype=np.float64)
        self._valid_input_index = source.valid.cpu().numpy().astype(bool)
        tsize = int(np.prod(tgt.shape))
        self.out_coords_x, self.out_coords_y = tgt.get_proj_vectors()
        if source.index is None:
            self._create_empty_bil_info()
            return
        self._resample_kdtree = source
        # neighbour ranks of every target pixel (no distances, _base.py:644-656)
        idx, _, valid_out, _, n_invalid = kd_tree._query_device(source, tgt, self._radius_of_influence, k, False,
                                                                want_dist=False)
        # source points in the target CRS, once per point
        n_src = int(source.lon_t.numel())
        xy = torch.empty((n_src, 2), dtype=torch.float64, device="cuda")
        _cabi.check(L.prs_project_forward(kd_tree._ptr(source.lon_t), kd_tree._ptr(source.lat_t), n_src,
                                          kd_tree._tree_code(source.dtype), ctypes.byref(tgt.proj_spec.cstruct),
                                          kd_tree._ptr(xy), kd_tree._stream()))
        r2s = kd_tree._rank_to_source(source)
        t = torch.empty(tsize, dtype=torch.float64, device="cuda")
        s = torch.empty(tsize, dtype=torch.float64, device="cuda")
        idx4 = torch.empty((tsize, 4), dtype=torch.int32, device="cuda")
        grid = kd_tree._area_grid(tgt)
        _cabi.check(L.prs_bilinear_info(kd_tree._ptr(idx), tsize, k, int(source.n_valid), kd_tree._ptr(r2s),
                                        kd_tree._ptr(xy), ctypes.byref(grid), kd_tree._ptr(t), kd_tree._ptr(s),
                                        kd_tree._ptr(idx4), kd_tree._stream()))
        if n_invalid:  # the reference's arrays only cover the valid target pixels (_base.py:143-162, 182-186)
            keep = valid_out.bool()
            t, s, idx4 = t[keep], s[keep], idx4[keep].contiguous()
        self._device = (idx4, t, s, r2s)
        self.bilinear_t, self.bilinear_s = kd_tree._host(t), kd_tree._host(s)
        self._index_array = kd_tree._host(idx4).view(np.uint32)
        self._get_slices()

    def _create_empty_bil_info(self):
        """_base.py:133-141."""
        tsize = int(np.prod(self._target_geo_def.shape))
        ssize = int(self._source_geo_def.size)
        self._valid_input_index = np.ones(ssize, dtype=bool)
        self._index_array = np.ones((tsize, 4), dtype=np.int32)
        self.bilinear_s = np.nan * np.zeros(tsize)
        self.bilinear_t = np.nan * np.zeros(tsize)
        self.slices_x = np.zeros((tsize, 4), dtype=np.int32)
        self.slices_y = np.zeros((tsize, 4), dtype=np.int32)
        self.out_coords_x, self.out_coords_y = self._target_geo_def.get_proj_vectors()
        self.mask_slices = self._index_array >= ssize
        self._device = None

    def _get_slices(self):
        """Line / column of each corner in the source array (_base.py:193-211)."""
        shp = tuple(self._source_geo_def.shape)
        flat = np.flatnonzero(self._valid_input_index)[self._index_array]
        if len(shp) > 1:
            self.slices_y, self.slices_x = np.divmod(flat, shp[1])
        else:
            self.slices_y, self.slices_x = np.zeros_like(flat, dtype=np.uint32), flat
        self.mask_slices = self._index_array >= int(self._source_geo_def.size)

    # ------------------------------------------------------------------ sampling
    def _sample_device(self, rows, fill):
        """rows: (n_source, C) host array -> (n_out, C) float64 host array."""
        torch = kd_tree._torch()
        idx4, t, s, r2s = self._device
        ddt = np.dtype(np.float32) if rows.dtype == np.float32 else np.dtype(np.float64)
        dev = kd_tree._dev(np.ascontiguousarray(rows, dtype=ddt))
        n_out, C = int(idx4.shape[0]), int(rows.shape[1])
        out = torch.empty((n_out, C), dtype=torch.float64, device="cuda")
        _cabi.check(_cabi.lib().prs_bilinear_sample(kd_tree._ptr(idx4), kd_tree._ptr(t), kd_tree._ptr(s), n_out,
                                                    kd_tree._ptr(r2s), kd_tree._ptr(dev), kd_tree._tree_code(ddt), C,
                                                    float(fill), kd_tree._ptr(out), kd_tree._stream()))
        return kd_tree._host(out)

    def get_sample_from_bil_info(self, data, fill_value=None, output_shape=None):
        raise NotImplementedError

    def save_resampling_info(self, filename):
        raise NotImplementedError

    def load_resampling_info(self, filename):
        raise NotImplementedError


class NumpyBilinearResampler(BilinearBase):
    """Bilinear interpolation for numpy arrays (pyresample/bilinear/_numpy_resampler.py:228-273)."""

    def resample(self, data, fill_value=0, nprocs=1):
        """Resample the given data."""
        self.get_bil_info(nprocs=nprocs)
        return self.get_sample_from_bil_info(data, fill_value=fill_value, output_shape=None)

    def get_sample_from_bil_info(self, data, fill_value=None, output_shape=None):
        """Resample using the pre-computed look-up tables (:240-251)."""
        del output_shape
        data = np.ma.getdata(np.asarray(data)) if not isinstance(data, np.ma.MaskedArray) else np.ma.filled(data, np.nan)
        n_src = self._valid_input_index.shape[0]
        multi = data.ndim > 2 and data.shape[0] * data.shape[1] == n_src
        rows = data.reshape(n_src, -1) if (multi or data.ndim <= 2) else None
        if rows is None:
            raise ValueError("Only 2D and 3D arrays are supported")
        shp = tuple(self._target_geo_def.shape)
        if self._device is None:  # empty info: every corner is "outside" (mask_slices) and t / s are NaN
            res = np.full((int(np.prod(shp)), rows.shape[1]), np.nan)
        else:
            res = self._sample_device(rows, np.nan)
        res = res.reshape(shp + (rows.shape[1],)) if multi else res.reshape(shp)
        res = np.squeeze(res)
        if fill_value is None:
            res = np.ma.masked_invalid(res)
        else:
            res[np.isnan(res)] = fill_value
        return np.squeeze(res)


class NumpyResamplerBilinear(NumpyBilinearResampler):
    """Wrapper for the old resampler class (:305-318)."""

    def __init__(self, source_geo_def, target_geo_def, radius_of_influence, **kwargs):
        warnings.warn("Use of NumpyResamplerBilinear is deprecated, use NumpyBilinearResampler instead", stacklevel=2)
        super().__init__(source_geo_def, target_geo_def, radius_of_influence, **kwargs)


def resample_bilinear(data, source_geo_def, target_area_def, radius=50e3, neighbours=32, nprocs=1, fill_value=0,
                      reduce_data=True, segments=None, epsilon=0):
    """Resample using bilinear interpolation (:37-99)."""
    warnings.warn("Usage of resample_bilinear() is deprecated, please use NumpyResamplerBilinear class instead",
                  FutureWarning, stacklevel=2)
    resampler = NumpyBilinearResampler(source_geo_def, target_area_def, radius, neighbours=neighbours, epsilon=epsilon,
                                       reduce_data=reduce_data)
    resampler.get_bil_info(nprocs=nprocs)
    return resampler.get_sample_from_bil_info(data, fill_value=fill_value, output_shape=None)


def get_bil_info(source_geo_def, target_area_def, radius=50e3, neighbours=32, nprocs=1, masked=False, reduce_data=True,
                 segments=None, epsilon=0):
    """Calculate information needed for bilinear resampling (:161-225) -> (t, s, valid_input_index, index_array)."""
    warnings.warn("Usage of get_bil_info() is deprecated, please use NumpyResamplerBilinear class instead",
                  FutureWarning, stacklevel=2)
    resampler = NumpyBilinearResampler(source_geo_def, target_area_def, radius, neighbours=neighbours, epsilon=epsilon,
                                       reduce_data=reduce_data)
    resampler.get_bil_info(nprocs=nprocs)
    return resampler.bilinear_t, resampler.bilinear_s, resampler._valid_input_index, resampler._index_array


def get_sample_from_bil_info(data, t__, s__, input_idxs, idx_arr, output_shape=None):
    """Resample 1-D data with look-up tables the caller holds (:102-158): ``data[input_idxs][idx_arr]`` weighted
    by (s, t); results outside the data's own range (+- 1e-6) become NaN."""
    warnings.warn("Usage of get_sample_from_bil_info() is deprecated, please use NumpyResamplerBilinear class instead",
                  FutureWarning, stacklevel=2)
    torch = kd_tree._torch()
    L = _cabi.lib()
    masked = isinstance(data, np.ma.MaskedArray)
    values = np.ma.filled(data, np.nan) if masked else np.asarray(data)
    input_idxs = np.asarray(input_idxs, dtype=bool)
    new_data = values[input_idxs]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        data_min, data_max = np.nanmin(new_data) - 1e-6, np.nanmax(new_data) + 1e-6
    ddt = np.dtype(np.float32) if new_data.dtype == np.float32 else np.dtype(np.float64)
    dev = kd_tree._dev(np.ascontiguousarray(new_data, dtype=ddt).reshape(-1, 1))
    idx4 = kd_tree._dev(np.ascontiguousarray(idx_arr).astype(np.uint32).view(np.int32))
    t_t, s_t = kd_tree._dev(np.asarray(t__, dtype=np.float64)), kd_tree._dev(np.asarray(s__, dtype=np.float64))
    n_out = int(idx4.shape[0])
    out = torch.empty((n_out, 1), dtype=torch.float64, device="cuda")
    _cabi.check(L.prs_bilinear_sample(kd_tree._ptr(idx4), kd_tree._ptr(t_t), kd_tree._ptr(s_t), n_out, None,
                                      kd_tree._ptr(dev), kd_tree._tree_code(ddt), 1, float("nan"), kd_tree._ptr(out),
                                      kd_tree._stream()))
    result = kd_tree._host(out).reshape(-1)
    with np.errstate(invalid='ignore'):
        result[(result < data_min) | (result > data_max)] = np.nan
    if output_shape is not None:
        result = result.reshape(output_shape)
    return result
