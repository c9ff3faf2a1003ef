"""Geometry definitions consumed by the B200 kd-tree resampling path.

These are the *host-side mirror* of the parts of ``pyresample.geometry`` that the
hot path reads (SURVEY 8b "duck-type surface"): ``size``, ``shape``, ``ndim``,
``dtype``, ``get_lonlats(data_slice=, dtype=)``, ``get_boundary_lonlats()``,
``get_cartesian_coords()`` plus, for areas, the projection / extent / pixel-size
attributes of ``AreaDefinition.__init__`` (pyresample/geometry.py:1577-1619).
They are not a re-implementation of pyresample's geometry algebra.  Genuine
pyresample geometry objects are accepted too - :func:`as_geo_def` adapts them by
attribute, using ``crs.to_dict()`` for the projection.

Coordinate generation for areas runs on the GPU (``prs_area_lonlats`` /
``prs_area_xyz``): there is no CPU implementation in this package.
"""
from __future__ import annotations

import math

import numpy as np

from .proj import ProjSpec, proj_to_dict


def _np_dtype(dtype):
    """numpy dtype of a numpy or torch dtype object."""
    text = str(dtype)
    if text.startswith("torch."):
        return np.dtype(text[len("torch."):])
    return np.dtype(dtype)


class DimensionError(ValueError):
    """Wrong dimensionality (pyresample/geometry.py:75)."""


class SimpleBoundary(object):
    """Container for the four sides of a geometry (pyresample/boundary/legacy_boundary.py:135-145)."""

    def __init__(self, side1, side2, side3, side4):
        self.side1 = side1
        self.side2 = side2
        self.side3 = side3
        self.side4 = side4


class BaseDefinition(object):
    """Base class (pyresample/geometry.py:96-140)."""

    def __init__(self, lons=None, lats=None, nprocs=1):
        if type(lons) is not type(lats):
            raise TypeError('lons and lats must be of same type')
        elif lons is not None:
            if not isinstance(lons, (np.ndarray,)) and not hasattr(lons, "shape"):
                lons = np.asanyarray(lons)
                lats = np.asanyarray(lats)
            if lons.shape != lats.shape:
                raise ValueError('lons and lats must have same shape')
        self.nprocs = nprocs
        self.lons = lons
        self.lats = lats
        self.ndim = None
        self.cartesian_coords = None

    def get_lonlats(self, data_slice=None, chunks=None, **kwargs):
        """Stored lon/lat arrays, optionally sliced (pyresample/geometry.py:242-274)."""
        lons, lats = self.lons, self.lats
        if lons is None or lats is None:
            raise ValueError('lon/lat values are not defined')
        if hasattr(lons, "dims") and hasattr(lons, "data"):
            # xarray DataArray: use the array underneath (pyresample/geometry.py:257-260)
            lons, lats = lons.data, lats.data
        if hasattr(lons, "compute"):  # dask arrays are materialised: the GPU path is eager
            lons, lats = lons.compute(), lats.compute()
        if chunks is not None:
            raise NotImplementedError("dask chunking is not part of the B200 path; arrays are device tensors")
        if data_slice is not None:
            lons, lats = lons[data_slice], lats[data_slice]
        return lons, lats

    def get_boundary_lonlats(self):
        """Four boundary sides, clockwise from the upper left (pyresample/geometry.py:284-291)."""
        s1_lon, s1_lat = self.get_lonlats(data_slice=(0, slice(None)))
        s2_lon, s2_lat = self.get_lonlats(data_slice=(slice(None), -1))
        s3_lon, s3_lat = self.get_lonlats(data_slice=(-1, slice(None, None, -1)))
        s4_lon, s4_lat = self.get_lonlats(data_slice=(slice(None, None, -1), 0))
        return (SimpleBoundary(s1_lon.squeeze(), s2_lon.squeeze(), s3_lon.squeeze(), s4_lon.squeeze()),
                SimpleBoundary(s1_lat.squeeze(), s2_lat.squeeze(), s3_lat.squeeze(), s4_lat.squeeze()))

    def get_cartesian_coords(self, nprocs=None, data_slice=None, cache=False):
        """ECEF coordinates on the R=6370997 m sphere, computed on the GPU (pyresample/geometry.py:465-516)."""
        from . import kd_tree  # late import: needs the CUDA library
        return kd_tree._cartesian_coords_numpy(self, data_slice)


class CoordinateDefinition(BaseDefinition):
    """Geometry defined by lons and lats only (pyresample/geometry.py:653-700)."""

    def __init__(self, lons, lats, nprocs=1):
        if not hasattr(lons, "shape") or not hasattr(lons, "dtype"):
            lons = np.asanyarray(lons)
            lats = np.asanyarray(lats)
        super().__init__(lons, lats, nprocs)
        if tuple(lons.shape) == tuple(lats.shape) and lons.dtype == lats.dtype:
            self.shape = tuple(lons.shape)
            self.size = int(np.prod(self.shape)) if len(self.shape) else 1
            self.ndim = len(self.shape)
            self.dtype = _np_dtype(lons.dtype)
        else:
            raise ValueError(('%s must be created with either '
                              'lon/lats of the same shape with same dtype') % self.__class__.__name__)

    def geocentric_resolution(self, ellps='WGS84', radius=None, nadir_factor=2):
        """Distance between the first two pixels of the middle row in geocentric metres
        (pyresample/geometry.py:702-762).  Host-side, O(1)."""
        if hasattr(self.lons, 'attrs') and self.lons.attrs.get('resolution') is not None:
            return self.lons.attrs['resolution'] * nadir_factor
        if self.ndim == 1:
            raise RuntimeError("Can't confidently determine geocentric resolution for 1D swath.")
        start_row = self.shape[0] // 2
        lons = _to_numpy(self.lons[start_row: start_row + 1, :2]).ravel().astype(np.float64)
        lats = _to_numpy(self.lats[start_row: start_row + 1, :2]).ravel().astype(np.float64)
        xyz = _geodetic_to_geocentric(lons, lats, ellps, radius)
        dist = np.linalg.norm(xyz[1] - xyz[0])
        if not np.isfinite(dist):
            raise RuntimeError("Could not calculate geocentric resolution")
        return dist


_GEOCENTRIC_ELLPS = {"WGS84": (6378137.0, 298.257223563), "GRS80": (6378137.0, 298.257222101)}


def _to_numpy(a):
    if hasattr(a, "detach"):
        return a.detach().cpu().numpy()
    if hasattr(a, "values") and not isinstance(a, np.ndarray):
        return np.asarray(a.values)
    return np.asarray(a)


def _geodetic_to_geocentric(lons, lats, ellps='WGS84', radius=None):
    """``+proj=cart`` on an ellipsoid (or a sphere of ``radius``), altitude 0 - used only for resolution estimates."""
    if radius:
        a, es = float(radius), 0.0
    else:
        a, rf = _GEOCENTRIC_ELLPS[ellps]
        f = 1.0 / rf
        es = 2 * f - f * f
    lam, phi = np.deg2rad(lons), np.deg2rad(lats)
    n = a / np.sqrt(1 - es * np.sin(phi) ** 2)
    return np.stack((n * np.cos(phi) * np.cos(lam), n * np.cos(phi) * np.sin(lam), n * (1 - es) * np.sin(phi)), axis=1)


class GridDefinition(CoordinateDefinition):
    """Grid defined by 2-D lons and lats (pyresample/geometry.py:765-810)."""

    def __init__(self, lons, lats, nprocs=1):
        super().__init__(lons, lats, nprocs)
        if lons.shape != lats.shape:
            raise ValueError('lon and lat grid must have same shape')
        elif lons.ndim != 2:
            raise ValueError('2 dimensional lon lat grid expected')


class SwathDefinition(CoordinateDefinition):
    """Swath defined by lons and lats (pyresample/geometry.py:813-870)."""

    def __init__(self, lons, lats, nprocs=1, crs=None):
        super().__init__(lons, lats, nprocs)
        if lons.shape != lats.shape:
            raise ValueError('lon and lat arrays must have same shape')
        elif lons.ndim > 3:
            raise ValueError('Only 1, 2 or 3 dimensional swaths are allowed')
        self.crs = crs


class AreaDefinition(BaseDefinition):
    """Uniform grid in a cartographic projection (pyresample/geometry.py:1403-1619).

    ``projection`` is a PROJ.4 dict or string (or a pyproj CRS when pyproj exists).
    """

    def __init__(self, area_id, description, proj_id, projection, width, height, area_extent, nprocs=1,
                 lons=None, lats=None, dtype=np.float64):
        super().__init__(lons, lats, nprocs)
        self.area_id = area_id
        self.description = description
        self.proj_id = proj_id
        self.width = int(width)
        self.height = int(height)
        self.crop_offset = (0, 0)
        self.shape = (self.height, self.width)
        self.size = self.height * self.width
        if lons is not None and tuple(lons.shape) != self.shape:
            raise ValueError('Shape of lon lat grid must match area definition')
        self.ndim = 2
        self.pixel_size_x = (area_extent[2] - area_extent[0]) / float(width)
        self.pixel_size_y = (area_extent[3] - area_extent[1]) / float(height)
        self._area_extent = tuple(area_extent)
        self.proj_dict = proj_to_dict(projection)
        self.proj_spec = ProjSpec(self.proj_dict)
        self.upper_left_extent = (float(area_extent[0]), float(area_extent[3]))
        self.pixel_upper_left = (float(area_extent[0]) + float(self.pixel_size_x) / 2,
                                 float(area_extent[3]) - float(self.pixel_size_y) / 2)
        self.pixel_offset_x = -self.area_extent[0] / self.pixel_size_x
        self.pixel_offset_y = self.area_extent[3] / self.pixel_size_y
        self.dtype = np.dtype(dtype)

    @property
    def area_extent(self):
        return self._area_extent

    @property
    def proj_str(self):
        return " ".join("+%s=%s" % (k, v) if v is not True else "+%s" % k for k, v in sorted(self.proj_dict.items()))

    def __getitem__(self, key):
        """Sub-area of a (row slice, column slice) with unit steps - what a slicer's result is applied with
        (AreaDefinition.__getitem__, pyresample/geometry.py)."""
        ys, xs = key
        r0, r1, rstep = ys.indices(self.height)
        c0, c1, cstep = xs.indices(self.width)
        if rstep != 1 or cstep != 1 or r1 <= r0 or c1 <= c0:
            raise NotImplementedError("only non-empty slices with step 1 select a sub-area")
        x0, y0, x1, y1 = (float(v) for v in self.area_extent)
        extent = (x0 + c0 * self.pixel_size_x, y1 - r1 * self.pixel_size_y, x0 + c1 * self.pixel_size_x,
                  y1 - r0 * self.pixel_size_y)
        sub = AreaDefinition(self.area_id, self.description, self.proj_id, self.proj_dict, c1 - c0, r1 - r0, extent,
                             dtype=self.dtype)
        sub.crop_offset = (self.crop_offset[0] + r0, self.crop_offset[1] + c0)
        return sub

    def get_proj_vectors(self, dtype=None, chunks=None):
        """1-D projection coordinates of the pixel centres (pyresample/geometry.py:1374-1380, 2407-2437).

        Evaluated as ``arange(n, dtype) * pixel_size + offset`` in ``dtype`` with Python-float
        (weak) scalars, exactly like the reference - SURVEY 9.3.
        """
        if dtype is None:
            dtype = self.dtype
        x = np.arange(0, self.width, dtype=dtype) * float(self.pixel_size_x) + self.pixel_upper_left[0]
        y = np.arange(0, self.height, dtype=dtype) * -float(self.pixel_size_y) + self.pixel_upper_left[1]
        return x, y

    def get_proj_coords(self, data_slice=None, dtype=None, chunks=None):
        """2-D projection coordinates (pyresample/geometry.py:2449-2488)."""
        x, y = self.get_proj_vectors(dtype=dtype)
        ys, xs = _yx_slice(data_slice)
        if ys is not None:
            y = y[ys]
        if xs is not None:
            x = x[xs]
        return tuple(np.meshgrid(x, y))

    def geocentric_resolution(self, ellps='WGS84', radius=None):
        """Best estimate of the overall geocentric pixel size (pyresample/geometry.py:2691-2764): distances along
        the middle row and middle column, mean of the first two of 10 histogram bin edges, maximum of both."""
        def _safe_bin_edges(arr):
            try:
                return np.histogram_bin_edges(arr, bins=10)[:2]
            except ValueError:
                return arr[:2]
        rows, cols = self.shape
        hor_lon, hor_lat = self.get_lonlats(data_slice=(rows // 2, slice(None)), dtype=np.float64)
        ver_lon, ver_lat = self.get_lonlats(data_slice=(slice(None), cols // 2), dtype=np.float64)
        res = []
        with np.errstate(invalid="ignore"):
            for lo, la in ((hor_lon, hor_lat), (ver_lon, ver_lat)):
                xyz = _geodetic_to_geocentric(np.asarray(lo).ravel(), np.asarray(la).ravel(), ellps, radius)
                dist = np.linalg.norm(np.diff(xyz, axis=0), axis=1)
                dist = dist[np.isfinite(dist)]
                res.append(np.mean(_safe_bin_edges(dist)) if dist.size else 0)
        out = max(res)
        if not out:
            raise RuntimeError("Could not calculate geocentric resolution")
        return out

    def get_lonlats(self, nprocs=None, data_slice=None, cache=False, dtype=None, chunks=None):
        """Lon/lat of the pixel centres, inverse-projected on the GPU (pyresample/geometry.py:2558-2645)."""
        if dtype is None:
            dtype = self.dtype
        if self.lons is not None:
            lons, lats = self.lons, self.lats
            if data_slice is not None:
                lons, lats = lons[data_slice], lats[data_slice]
            return lons, lats
        from . import kd_tree  # late import: needs the CUDA library
        lons, lats = kd_tree._area_lonlats_numpy(self, data_slice, np.dtype(dtype))
        if cache and data_slice is None:
            self.lons, self.lats = lons, lats
        return lons, lats


class IncompatibleAreas(ValueError):
    """Error when the areas to combine are not compatible (pyresample/geometry.py:66-67)."""


def combine_area_extents_vertical(area1, area2):
    """Extent of two areas stacked vertically (pyresample/geometry.py:2878-2896)."""
    e1, e2 = area1.area_extent, area2.area_extent
    if not (e1[0] == e2[0] and e1[2] == e2[2]):
        raise IncompatibleAreas("Can't concatenate area definitions with incompatible area extents: "
                                "{0} and {1}".format(area1, area2))
    current_extent = list(e1)
    if np.isclose(e1[1], e2[3]):
        current_extent[1] = e2[1]
    elif np.isclose(e1[3], e2[1]):
        current_extent[3] = e2[3]
    else:
        raise IncompatibleAreas("Can't concatenate non-contiguous area definitions: "
                                "{0} and {1}".format(area1, area2))
    return current_extent


def concatenate_area_defs(area1, area2, axis=0):
    """Append *area2* to *area1* (pyresample/geometry.py:2899-2920); vertical only, like the reference."""
    if axis != 0:
        raise NotImplementedError('Only vertical contatenation is supported.')
    if area1.proj_dict != area2.proj_dict or area1.width != area2.width:
        raise IncompatibleAreas("Can't concatenate area definitions with different projections: "
                                "{0} and {1}".format(area1, area2))
    area_extent = combine_area_extents_vertical(area1, area2)
    return AreaDefinition(area1.area_id, area1.description, area1.proj_id, area1.proj_dict, int(area1.width),
                          int(area1.height + area2.height), area_extent, dtype=area1.dtype)


class StackedAreaDefinition(BaseDefinition):
    """Several vertically stacked AreaDefinitions (pyresample/geometry.py:2923-3049).

    Contiguous members are merged on ``append`` like in the reference; ``get_lonlats`` stacks the members'
    (GPU inverse-projected) coordinates.  For kd_tree it is a plain 2-D geometry: no reduce_data box (it is neither
    Coordinate-, Grid- nor AreaDefinition, kd_tree.py:410-414).
    """

    def __init__(self, *definitions, **kwargs):
        super().__init__(None, None, kwargs.get('nprocs', 1))
        self.dtype = np.dtype(kwargs.get('dtype', np.float64))
        self.defs = []
        self.ndim = 2
        for definition in definitions:
            self.append(definition)

    @property
    def width(self):
        return self.defs[0].width

    @property
    def height(self):
        return sum(definition.height for definition in self.defs)

    @property
    def shape(self):
        return (self.height, self.width)

    @shape.setter
    def shape(self, value):  # BaseDefinition.__init__ assigns it
        pass

    @property
    def size(self):
        return self.height * self.width

    @size.setter
    def size(self, value):
        pass

    @property
    def proj_str(self):
        return self.defs[0].proj_str

    def append(self, definition):
        if isinstance(definition, StackedAreaDefinition):
            for area in definition.defs:
                self.append(area)
            return
        if definition.height == 0:
            return
        if self.defs and self.defs[0].proj_dict != definition.proj_dict:
            raise NotImplementedError('Cannot append areas: CRS mismatch')
        try:
            self.defs[-1] = concatenate_area_defs(self.defs[-1], definition)
        except (IncompatibleAreas, IndexError):
            self.defs.append(definition)

    def get_lonlats(self, nprocs=None, data_slice=None, cache=False, dtype=None, chunks=None):
        """Stack of the members' lon/lat arrays (pyresample/geometry.py:2966-2998)."""
        llons, llats = [], []
        try:
            row_slice, col_slice = data_slice
        except TypeError:
            row_slice, col_slice = slice(0, self.height), slice(0, self.width)
        if not isinstance(row_slice, slice):
            raise TypeError("StackedAreaDefinition.get_lonlats takes row slices")
        start = 0 if row_slice.start is None else row_slice.start
        stop = self.height if row_slice.stop is None else row_slice.stop
        offset = 0
        for areadef in self.defs:
            local = slice(max(start - offset, 0), min(max(stop - offset, 0), areadef.height), row_slice.step)
            lons, lats = areadef.get_lonlats(nprocs=nprocs, data_slice=(local, col_slice), cache=cache, dtype=dtype)
            llons.append(lons)
            llats.append(lats)
            offset += areadef.height
        self.lons, self.lats = np.vstack(llons), np.vstack(llats)
        return self.lons, self.lats

    def squeeze(self):
        """A single AreaDefinition if possible."""
        return self.defs[0] if len(self.defs) == 1 else self


def _yx_slice(data_slice):
    """AreaDefinition._get_yx_data_slice (pyresample/geometry.py:2491-2496)."""
    if data_slice is not None and isinstance(data_slice, slice):
        return data_slice, slice(None, None, None)
    elif data_slice is not None:
        return data_slice[0], data_slice[1]
    return None, None


def _get_slice(segments, shape):
    """Row blocks of ``ceil(rows / segments)`` rows as slices: the segmentation rule of geometry.py:3052-3068
    (identical block boundaries; for 2-D shapes every item is ``(row_slice, slice(None))``)."""
    if len(shape) not in (1, 2):
        raise ValueError('Cannot segment array of shape: %s' % str(shape))
    rows = int(shape[0])
    block = int(math.ceil(float(rows) / segments))
    for start in range(0, rows, max(block, 1)):
        row_slice = slice(start, min(start + block, rows))
        yield row_slice if len(shape) == 1 else (row_slice, slice(None))


# --------------------------------------------------------------------------- duck typing
def _mro_names(obj):
    return {c.__name__ for c in type(obj).__mro__}


def is_base_definition(obj):
    return isinstance(obj, BaseDefinition) or "BaseDefinition" in _mro_names(obj)


def is_area(obj):
    return isinstance(obj, AreaDefinition) or "AreaDefinition" in _mro_names(obj)


def is_griddish(obj):
    """``isinstance(obj, (GridDefinition, AreaDefinition))`` (pyresample/kd_tree.py:410-412)."""
    return is_area(obj) or isinstance(obj, GridDefinition) or "GridDefinition" in _mro_names(obj)


def is_coord(obj):
    """``isinstance(obj, CoordinateDefinition)`` (pyresample/kd_tree.py:413)."""
    return isinstance(obj, CoordinateDefinition) or "CoordinateDefinition" in _mro_names(obj)


def is_swath(obj):
    return isinstance(obj, SwathDefinition) or "SwathDefinition" in _mro_names(obj)


def as_geo_def(obj):
    """Adapt a genuine pyresample geometry object; objects of this module pass through."""
    if isinstance(obj, BaseDefinition):
        return obj
    names = _mro_names(obj)
    if "AreaDefinition" in names:
        crs = getattr(obj, "crs", None)
        projection = crs.to_dict() if crs is not None else obj.proj_dict
        area = AreaDefinition(getattr(obj, "area_id", ""), getattr(obj, "description", ""),
                              getattr(obj, "proj_id", ""), projection, obj.width, obj.height, obj.area_extent,
                              dtype=getattr(obj, "dtype", np.float64))
        return area
    if "GridDefinition" in names:
        return GridDefinition(np.asanyarray(obj.lons), np.asanyarray(obj.lats))
    if "SwathDefinition" in names:
        return SwathDefinition(np.asanyarray(obj.lons), np.asanyarray(obj.lats))
    if "CoordinateDefinition" in names:
        return CoordinateDefinition(np.asanyarray(obj.lons), np.asanyarray(obj.lats))
    return obj
