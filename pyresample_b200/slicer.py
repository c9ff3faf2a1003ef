"""Source cropping before the resampling path: the slices of a source geometry that enclose a target
area - what ``pyresample.slicer.create_slicer(area_to_crop, area_to_contain).get_slices()`` (slicer.py:38-49, 93-96) and
``AreaDefinition.get_area_slices`` are used for.

The reference intersects shapely polygons of the two boundaries.  Here every source pixel is projected into the
target CRS on the device (``prs_project_forward``) and tested against the target's extent grown by a margin; the
slices are the bounding rows / columns of the pixels that fall inside.  The result encloses everything a resampling
with ``radius_of_influence <= margin`` can touch (resampling the cropped source gives the same grid, see
tests/test_gpu_slicer.py); it is not promised to be pixel-identical to the polygon method's slices.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _cabi, geometry, kd_tree


class IncompatibleAreas(ValueError):
    """The two geometries do not overlap (pyresample/geometry.py ``IncompatibleAreas``)."""


def expand_slice(small_slice):
    """Expand a slice by one element on both sides (slicer.py:165-167)."""
    return slice(max(small_slice.start - 1, 0), small_slice.stop + 1, small_slice.step)


class Slicer(object):
    """Slices of ``area_to_crop`` that enclose ``area_to_contain`` (slicer.py:52-104)."""

    def __init__(self, area_to_crop, area_to_contain, margin=None):
        self.area_to_crop = geometry.as_geo_def(area_to_crop)
        self.area_to_contain = geometry.as_geo_def(area_to_contain)
        if not (geometry.is_area(self.area_to_contain) and getattr(self.area_to_contain, "lons", None) is None):
            raise NotImplementedError("the area to contain must be an AreaDefinition")
        if self.area_to_crop.ndim != 2:
            raise NotImplementedError("only 2-D source geometries can be sliced")
        self.margin = margin

    def _margin(self):
        """Margin in the target CRS' units: given, else two target pixels (geographic CRS: degrees)."""
        if self.margin is not None:
            m = float(self.margin)
            if self.area_to_contain.proj_spec.is_geographic:
                m = float(np.degrees(m / kd_tree.R)) * 1.5  # metres -> degrees, generous away from the equator
            return m
        return 2.0 * max(abs(self.area_to_contain.pixel_size_x), abs(self.area_to_contain.pixel_size_y))

    def get_slices(self):
        """(column slice, row slice) of ``area_to_crop``, like the reference's slicers."""
        torch = kd_tree._torch()
        src, tgt = self.area_to_crop, self.area_to_contain
        lon_t, lat_t, _ = kd_tree._geo_lonlats_device(src)
        n = int(lon_t.numel())
        xy = torch.empty((n, 2), dtype=torch.float64, device="cuda")
        _cabi.check(_cabi.lib().prs_project_forward(kd_tree._ptr(lon_t), kd_tree._ptr(lat_t), n,
                                                    kd_tree._tree_code(kd_tree._np_dtype_of(lon_t)),
                                                    ctypes.byref(tgt.proj_spec.cstruct), kd_tree._ptr(xy), kd_tree._stream()))
        x0, y0, x1, y1 = (float(v) for v in tgt.area_extent)
        m = self._margin()
        inside = ((xy[:, 0] >= min(x0, x1) - m) & (xy[:, 0] <= max(x0, x1) + m) &
                  (xy[:, 1] >= min(y0, y1) - m) & (xy[:, 1] <= max(y0, y1) + m)).reshape(tuple(src.shape))
        rows = torch.nonzero(inside.any(dim=1)).reshape(-1)
        cols = torch.nonzero(inside.any(dim=0)).reshape(-1)
        if rows.numel() == 0 or cols.numel() == 0:
            raise IncompatibleAreas("Areas not overlapping.")
        row_slice = expand_slice(slice(int(rows[0]), int(rows[-1]) + 1))
        col_slice = expand_slice(slice(int(cols[0]), int(cols[-1]) + 1))
        row_slice = slice(row_slice.start, min(row_slice.stop, src.shape[0]))
        col_slice = slice(col_slice.start, min(col_slice.stop, src.shape[1]))
        return col_slice, row_slice


class AreaSlicer(Slicer):
    """A Slicer for cropping AreaDefinitions (slicer.py:170-251)."""


class SwathSlicer(Slicer):
    """A Slicer for cropping SwathDefinitions (slicer.py:107-162)."""


def create_slicer(area_to_crop, area_to_contain, margin=None):
    """An ``AreaSlicer`` or a ``SwathSlicer`` depending on the geometry to crop (slicer.py:38-49).  ``margin``: how far
    (metres) beyond the target's edges source pixels are still needed - the resampling's ``radius_of_influence``."""
    g = geometry.as_geo_def(area_to_crop)
    if geometry.is_swath(g):
        return SwathSlicer(g, area_to_contain, margin)
    if geometry.is_area(g):
        return AreaSlicer(g, area_to_contain, margin)
    raise NotImplementedError("Don't know how to slice a " + str(type(area_to_crop)))


def get_area_slices(area_to_crop, area_to_contain, margin=None):
    """``AreaDefinition.get_area_slices``: (column slice, row slice) of ``area_to_crop`` covering ``area_to_contain``."""
    return create_slicer(area_to_crop, area_to_contain, margin).get_slices()
