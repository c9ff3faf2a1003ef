"""Image containers over the kd-tree path (pyresample/image.py:28-128, 203-293), SURVEY 8 f4.

``ImageContainerNearest.resample`` is ``kd_tree.resample_nearest`` on the GPU; ``get_array_from_linesample`` is one
device gather (``prs_gather_nearest``) with the semantics of ``grid.get_image_from_linesample`` (grid.py:28-88).
The quick (projection-coordinate) and bilinear containers are other resamplers and out of scope here.
"""
from __future__ import annotations

import ctypes
import warnings

import numpy as np

from . import _cabi


def get_image_from_linesample(row_indices, col_indices, source_image, fill_value=0):
    """Sample ``source_image[rows, cols]``; indices outside the image give ``fill_value`` or a masked pixel
    (pyresample/grid.py:28-88).  ``source_image``: ``(H, W)`` or ``(H, W, C)``."""
    from . import kd_tree as kd
    torch = kd._torch()
    L = _cabi.lib()
    row_indices, col_indices = np.asarray(row_indices), np.asarray(col_indices)
    H, W = source_image.shape[:2]
    rows, cols = row_indices.astype(np.int64), col_indices.astype(np.int64)
    valid = (rows >= 0) & (rows < H) & (cols >= 0) & (cols < W)
    n = H * W
    lin = np.where(valid, rows * W + cols, n).astype(np.uint32).ravel()
    data = np.ma.getdata(source_image)
    flat = np.ascontiguousarray(data.reshape(n, -1))
    masked_result = fill_value is None
    in_mask = np.ma.getmaskarray(source_image) if np.ma.is_masked(source_image) else None
    fill_row = np.full((1, flat.shape[1]), 0 if masked_result else fill_value).astype(flat.dtype)
    dev_rows = kd._dev(kd._byte_rows(flat))
    row_bytes = dev_rows.shape[1]
    idx_t = kd._dev(lin.view(np.int32))
    fill_t = kd._dev(kd._byte_rows(fill_row))
    out = torch.empty((lin.size, row_bytes), dtype=torch.uint8, device="cuda")
    _cabi.check(L.prs_gather_nearest(kd._ptr(idx_t), lin.size, n, None, kd._ptr(dev_rows), row_bytes,
                                     kd._ptr(fill_t), kd._ptr(out), kd._stream()))
    res = kd._host(out).view(flat.dtype).reshape(row_indices.shape + data.shape[2:])
    if not masked_result:
        return res
    mask = ~valid
    if res.ndim > mask.ndim:
        mask = np.broadcast_to(mask[..., None], res.shape).copy()
    if in_mask is not None:
        safe_r, safe_c = np.where(valid, rows, 0), np.where(valid, cols, 0)
        mask = mask | in_mask[safe_r, safe_c]
    return np.ma.array(res, mask=mask)


class ImageContainer(object):
    """Holds image with geometry definition.  Allows indexing with linesample arrays (image.py:28-128)."""

    def __init__(self, image_data, geo_def, fill_value=0, nprocs=1):
        warnings.warn("Usage of ImageContainer is deprecated, please use NumpyResamplerBilinear class instead",
                      FutureWarning, stacklevel=2)
        if not isinstance(image_data, (np.ndarray, np.ma.core.MaskedArray)):
            raise TypeError('image_data must be either an ndarray or a masked array')
        elif (image_data.ndim > geo_def.ndim + 1) or (image_data.ndim < geo_def.ndim):
            raise ValueError('Unexpected number of dimensions for image_data: %s' % image_data.ndim)
        for i, size in enumerate(geo_def.shape):
            if image_data.shape[i] != size:
                raise ValueError('Size mismatch for image_data. Expected size %s for dimension %s and got %s' %
                                 (size, i, image_data.shape[i]))
        self.shape = geo_def.shape
        self.size = geo_def.size
        self.ndim = geo_def.ndim
        self.image_data = image_data
        self.channels = image_data.shape[-1] if image_data.ndim > geo_def.ndim else 1
        self.geo_def = geo_def
        self.fill_value = fill_value
        self.nprocs = nprocs

    def __str__(self):
        return 'Image:\n %s' % self.image_data.__str__()

    def __repr__(self):
        return self.image_data.__repr__()

    def resample(self, target_geo_def):
        raise NotImplementedError('Method "resample" is not implemented in class %s' % self.__class__.__name__)

    def get_array_from_linesample(self, row_indices, col_indices):
        """Array sampled from the image by index arrays (image.py:100-123)."""
        if self.geo_def.ndim != 2:
            raise TypeError('Resampling from linesamples only makes sense on 2D data')
        return get_image_from_linesample(row_indices, col_indices, self.image_data, self.fill_value)

    def get_array_from_neighbour_info(self, *args, **kwargs):
        raise NotImplementedError('Method "get_array_from_neighbour_info" is not implemented in class %s' %
                                  self.__class__.__name__)


class ImageContainerNearest(ImageContainer):
    """Holds image with geometry definition; nearest-neighbour resampling to a new geometry (image.py:203-293)."""

    def __init__(self, image_data, geo_def, radius_of_influence, epsilon=0, fill_value=0, reduce_data=True, nprocs=1,
                 segments=None):
        super(ImageContainerNearest, self).__init__(image_data, geo_def, fill_value=fill_value, nprocs=nprocs)
        self.radius_of_influence = radius_of_influence
        self.epsilon = epsilon
        self.reduce_data = reduce_data
        self.segments = segments

    def resample(self, target_geo_def):
        """Resample the image with ``kd_tree.resample_nearest`` (image.py:262-293)."""
        from . import kd_tree
        if self.image_data.ndim > 2 and self.ndim > 1:
            image_data = self.image_data.reshape(self.image_data.shape[0] * self.image_data.shape[1],
                                                 self.image_data.shape[2])
        else:
            image_data = self.image_data.ravel()
        resampled_image = kd_tree.resample_nearest(self.geo_def, image_data, target_geo_def, self.radius_of_influence,
                                                   epsilon=self.epsilon, fill_value=self.fill_value,
                                                   nprocs=self.nprocs, reduce_data=self.reduce_data,
                                                   segments=self.segments)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", FutureWarning)
            return ImageContainerNearest(resampled_image, target_geo_def, self.radius_of_influence,
                                         epsilon=self.epsilon, fill_value=self.fill_value,
                                         reduce_data=self.reduce_data, nprocs=self.nprocs, segments=self.segments)
