"""Persistence and device residency of neighbour info.

Users of the reference cache ``(valid_input_index, valid_output_index, index_array, distance_array)`` from
``get_neighbour_info`` and feed it to ``get_sample_from_neighbour_info`` for every new image of the same geometry
(docs/source/howtos/swath.rst:207-243).  Two helpers for that pattern:

* ``save_neighbour_info`` / ``load_neighbour_info``: one ``.npz`` file, masks bit-packed, arrays exactly as
  ``get_neighbour_info`` returned them (dtypes and shapes included, also the int32 / distance-1.0 "empty" form);
* ``DeviceNeighbourInfo``: the same four arrays uploaded once; ``sample_nearest(data)`` is then the upload of the
  image, one ``prs_gather_nearest`` (+ ``prs_scatter_rows``) and the read-back - what
  ``get_sample_from_neighbour_info('nn', ...)`` does minus the re-upload of the index arrays on every call.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _cabi

FORMAT_VERSION = 1


def save_neighbour_info(filename, valid_input_index, valid_output_index, index_array, distance_array=None):
    """Write neighbour info to ``filename`` (``.npz``)."""
    vii = np.asarray(valid_input_index, dtype=bool)
    voi = np.asarray(valid_output_index, dtype=bool)
    arrays = dict(version=np.int64(FORMAT_VERSION),
                  vii_bits=np.packbits(vii.ravel()), vii_shape=np.asarray(vii.shape, dtype=np.int64),
                  voi_bits=np.packbits(voi.ravel()), voi_shape=np.asarray(voi.shape, dtype=np.int64),
                  index_array=np.asarray(index_array))
    if distance_array is not None:
        arrays["distance_array"] = np.asarray(distance_array)
    with open(filename, "wb") as fh:
        np.savez(fh, **arrays)


def load_neighbour_info(filename):
    """``(valid_input_index, valid_output_index, index_array, distance_array)`` as written by
    :func:`save_neighbour_info` (``distance_array`` is ``None`` when it was not stored)."""
    with np.load(filename) as f:
        if int(f["version"]) != FORMAT_VERSION:
            raise ValueError("unknown neighbour info format version %d" % int(f["version"]))

        def unpack(bits, shape):
            n = int(np.prod(shape)) if len(shape) else 1
            return np.unpackbits(bits, count=n).astype(bool).reshape(tuple(int(s) for s in shape))

        vii = unpack(f["vii_bits"], f["vii_shape"])
        voi = unpack(f["voi_bits"], f["voi_shape"])
        index_array = f["index_array"]
        distance_array = f["distance_array"] if "distance_array" in f.files else None
    return vii, voi, index_array, distance_array


class DeviceNeighbourInfo(object):
    """Neighbour info held in HBM for repeated nearest-neighbour sampling of the same geometry pair."""

    def __init__(self, valid_input_index, valid_output_index, index_array, distance_array=None):
        from . import kd_tree as kd
        torch = kd._torch()
        L = _cabi.lib()
        vii = np.asarray(valid_input_index, dtype=bool).ravel()
        voi = np.asarray(valid_output_index, dtype=bool).ravel()
        index_array = np.asarray(index_array)
        if index_array.ndim != 1:
            raise ValueError('index_array contains more neighbours than just the nearest')
        self.n_source, self.n_valid = vii.size, int(vii.sum())
        self.n_target, self.n_out = voi.size, int(voi.sum())
        if index_array.shape[0] != self.n_out:
            raise ValueError("index_array does not match valid_output_index")
        self.empty = self.n_valid == 0 or self.n_out == 0
        self.distance_array = distance_array
        self._idx = self._r2s = self._voi = None
        if self.empty:
            return
        self._idx = kd._dev(np.ascontiguousarray(index_array).astype(np.uint32, copy=False).view(np.int32))
        if self.n_valid != self.n_source:
            vt = kd._dev(vii)
            self._r2s = torch.empty(self.n_valid, dtype=torch.int32, device="cuda")
            _cabi.check(L.prs_rank_to_source(kd._ptr(vt), vt.numel(), kd._ptr(self._r2s), None, kd._stream()))
        if self.n_out != self.n_target:
            self._voi = kd._dev(voi)

    @classmethod
    def load(cls, filename):
        return cls(*load_neighbour_info(filename))

    def sample_nearest(self, data, output_shape, fill_value=0):
        """``get_sample_from_neighbour_info('nn', output_shape, data, ...)`` (kd_tree.py:566-652, 705-723) for
        unmasked ``data`` of shape ``(n_source,)`` / ``(n_source, C)`` / ``(rows, cols[, C])``."""
        from . import kd_tree as kd
        torch = kd._torch()
        L = _cabi.lib()
        if isinstance(data, np.ma.MaskedArray) and np.ma.is_masked(data):
            raise TypeError("masked data: use kd_tree.get_sample_from_neighbour_info")
        data = kd._flatten_data(np.ma.getdata(data), self.n_source)
        channels = data.shape[1] if data.ndim > 1 else 1
        output_shape = tuple(output_shape)
        final_shape = output_shape + (channels,) if data.ndim > 1 else output_shape
        masked = fill_value is None
        if masked:
            fill_value = kd._get_fill_mask_value(data.dtype)
        if self.empty:
            res = np.full(final_shape, fill_value, dtype=data.dtype)
        else:
            dev_data = kd._dev(kd._byte_rows(data))
            row_bytes = dev_data.shape[1]
            fill_row = kd._dev(kd._byte_rows(np.full((1, channels), fill_value, dtype=data.dtype)))
            res_t = torch.empty((self.n_out, row_bytes), dtype=torch.uint8, device="cuda")
            _cabi.check(L.prs_gather_nearest(kd._ptr(self._idx), self.n_out, self.n_valid, kd._ptr(self._r2s),
                                             kd._ptr(dev_data), row_bytes, kd._ptr(fill_row), kd._ptr(res_t),
                                             kd._stream()))
            res_t = kd._scatter_full(res_t, self._voi, self.n_target, fill_row)
            res = kd._host(res_t).view(data.dtype).reshape(final_shape)
        if masked:
            res = np.ma.masked_equal(res, fill_value)
        return res
