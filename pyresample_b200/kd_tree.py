"""Drop-in for ``pyresample.kd_tree`` (numpy API) running on hand-written sm_100a CUDA kernels.

Same names, argument order, defaults, warnings and exceptions as pyresample/kd_tree.py:64-899.
What changes is everything underneath:

* lon/lat -> ECEF, validity masks and (for ``AreaDefinition`` targets) the inverse projection run in
  CUDA kernels (``prs_valid_mask``, ``prs_area_lonlats``, in-kernel for query targets);
* pykdtree is replaced by a GPU-built cube-face cell index (``prs_index_build``) queried by an exact,
  radius-bounded k-NN kernel (``prs_query``) with ties broken to the lowest source index;
* ``resample_nearest`` / ``resample_gauss`` fuse the sample gather / Gaussian weighting into the query
  kernel's epilogue (``prs_resample_nearest`` / ``prs_resample_gauss``) - the neighbour arrays are never
  materialised; ``get_sample_from_neighbour_info`` runs ``prs_gather_nearest`` / ``prs_gather_weighted``.

numpy arrays in -> numpy (or masked) arrays out; CUDA torch tensors in -> CUDA torch tensors out for the
fused entry points.  There is no CPU fallback: without the CUDA library and a GPU every call raises.
"""
from __future__ import annotations

import ctypes
import threading
import time
import types
import warnings
from logging import getLogger

import numpy as np

from . import _cabi, geometry
from ._cabi import (PRS_F32, PRS_F64, PRS_REDUCE_ALL, PRS_REDUCE_BOX, PRS_REDUCE_BOX_DATELINE,
                    PRS_REDUCE_LAT_GE, PRS_REDUCE_LAT_LE, PRS_TARGET_AREA, PRS_TARGET_LONLAT,
                    prs_area_grid, prs_reduce_box, prs_target)

logger = getLogger(__name__)

R = 6370997.0  # pyresample/_spatial_mp.py:39

try:  # pyresample/kd_tree.py:43-54
    from xarray import DataArray
except ImportError:  # pragma: no cover - xarray is optional
    DataArray = None


class EmptyResult(ValueError):
    """No valid data is produced (pyresample/kd_tree.py:60)."""


# =========================================================================================== device plumbing
def _torch():
    import torch
    return torch


def _stream():
    return ctypes.c_void_p(_torch().cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _is_tensor(x):
    return type(x).__module__.startswith("torch") and hasattr(x, "data_ptr")


def _np_dtype_of(x):
    if _is_tensor(x):
        return np.dtype(str(x.dtype).replace("torch.", ""))
    return np.dtype(x.dtype)


# Ordinary (pageable) numpy arrays of at least _STAGED_MIN_BYTES go up in chunks through a few pinned staging
# buffers that a small thread pool fills (numpy releases the GIL while copying) while earlier chunks are already on
# the bus: ~3x the rate of a plain pageable cudaMemcpy, which is bound by one thread's memcpy.
_STAGED_MIN_BYTES = 16 << 20
_STAGE_CHUNK = 8 << 20
_STAGE_BUFFERS = 6
_STAGE_THREADS = 4
_stage = {"lock": threading.Lock(), "bufs": None, "pool": None}


def _staged_upload(a):
    """Contiguous pageable numpy array -> CUDA tensor of the same dtype / shape on the current stream."""
    torch = _torch()
    flat = a.reshape(-1).view(np.uint8)
    n = flat.size
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream()
    with _stage["lock"]:
        if _stage["bufs"] is None:
            from concurrent.futures import ThreadPoolExecutor
            _stage["bufs"] = [[torch.empty(_STAGE_CHUNK, dtype=torch.uint8, pin_memory=True), None]
                              for _ in range(_STAGE_BUFFERS)]
            for b in _stage["bufs"]:
                b.append(b[0].numpy())
            _stage["pool"] = ThreadPoolExecutor(max_workers=_STAGE_THREADS)
        bufs, pool = _stage["bufs"], _stage["pool"]
        n_chunks = (n + _STAGE_CHUNK - 1) // _STAGE_CHUNK

        def fill(i):
            b = bufs[i % _STAGE_BUFFERS]
            if b[1] is not None:
                b[1].synchronize()  # the previous transfer out of this buffer has finished
            off = i * _STAGE_CHUNK
            m = min(_STAGE_CHUNK, n - off)
            return pool.submit(np.copyto, b[2][:m], flat[off:off + m])

        pending = [(i, fill(i)) for i in range(min(_STAGE_BUFFERS, n_chunks))]
        while pending:
            i, fut = pending.pop(0)
            fut.result()
            b = bufs[i % _STAGE_BUFFERS]
            off = i * _STAGE_CHUNK
            m = min(_STAGE_CHUNK, n - off)
            dev[off:off + m].copy_(b[0][:m], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(stream)
            b[1] = ev
            if i + _STAGE_BUFFERS < n_chunks:
                pending.append((i + _STAGE_BUFFERS, fill(i + _STAGE_BUFFERS)))
    return dev.view(getattr(torch, a.dtype.name)).reshape(a.shape)


def _dev(x, dtype=None):
    """numpy array / torch tensor -> contiguous CUDA tensor (optionally cast)."""
    torch = _torch()
    if _is_tensor(x):
        t = x if x.is_cuda else x.cuda(non_blocking=True)
    else:
        a = np.ascontiguousarray(np.ma.getdata(x))
        if a.dtype == np.bool_:
            a = a.view(np.uint8)
        if not a.flags.writeable:
            a = a.copy()
        h = torch.from_numpy(a)
        if a.nbytes >= _STAGED_MIN_BYTES and not h.is_pinned():
            t = _staged_upload(a)
        else:
            t = h.cuda(non_blocking=True)
    if dtype is not None:
        td = getattr(torch, np.dtype(dtype).name)
        if t.dtype != td:
            t = t.to(td)
    return t.contiguous()


def _host(t):
    """CUDA tensor -> numpy array through a pinned staging buffer (torch's caching host allocator)."""
    torch = _torch()
    h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    h.copy_(t, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return h.numpy()


# Outputs at least this large take the pipelined path of resample_nearest (numpy in, area target, numpy out).
_PIPELINE_MIN_BYTES = 32 << 20
_PIPELINE_MIN_RUN_BYTES = 4 << 20   # shorter runs of finished bands are not worth a copy of their own
_side = {}
_TIMELINE = None  # set to a list to collect host-side time stamps of the pipelined path (tools/timeline_e2e.py)


def _side_streams():
    """(upload, read-back) streams of the current device for the pipelined path."""
    torch = _torch()
    dev = torch.cuda.current_device()
    if dev not in _side:
        _side[dev] = (torch.cuda.Stream(), torch.cuda.Stream())
    return _side[dev]


def _band_runs(live, band_rows, nrows, row_bytes_total, min_run_bytes):
    """Split the tile-row flags into runs [(row0, row1, is_live)]; short finished runs are merged into live ones."""
    live = np.asarray(live, dtype=bool).copy()
    n = live.size
    edges = np.flatnonzero(np.diff(live.astype(np.int8))) + 1
    starts = np.concatenate([[0], edges])
    ends = np.concatenate([edges, [n]])
    for b0, b1 in zip(starts, ends):
        if not live[b0] and (min(b1 * band_rows, nrows) - b0 * band_rows) * row_bytes_total < min_run_bytes:
            live[b0:b1] = True
    edges = np.flatnonzero(np.diff(live.astype(np.int8))) + 1
    starts = np.concatenate([[0], edges])
    ends = np.concatenate([edges, [n]])
    return [(int(b0) * band_rows, int(min(b1 * band_rows, nrows)), bool(live[b0])) for b0, b1 in zip(starts, ends)]


class _HostUpload(object):
    """Arrival of the source rows on the device for _nearest_pipelined: upload from host memory on a side stream."""

    def __init__(self, host_rows):
        self.host_rows = host_rows
        self.row_bytes = host_rows.shape[1]
        # pinned memory goes up asynchronously: start right away.  Pageable memory keeps this thread busy while it
        # is staged (_staged_upload): start once the early read-back has been queued.
        self.start_early = bool(_torch().from_numpy(np.ascontiguousarray(host_rows)).is_pinned())

    def start(self, main, up):
        torch = _torch()
        up.wait_stream(main)
        with torch.cuda.stream(up):
            self.dev = _dev(self.host_rows)
            self.ready = torch.cuda.Event()
            self.ready.record(up)

    def finish(self, main):
        """Make ``main`` wait for the rows; returns the device tensor."""
        main.wait_event(self.ready)
        return self.dev


def _nearest_pipelined(L, source, tg, n_t, arrival, fill_row_host, radius):
    """resample_nearest for host arrays with the PCIe transfers overlapped (both directions at once).

    The result of every tile that cannot have a neighbour is the fill value whatever the data are, so: start the
    arrival of the data (``arrival.start``: an upload on a side stream; pyresample_b200/distributed.py adds a
    broadcast); meanwhile classify + fill (prs_resample_nearest_fill) and start reading back the bands of rows
    that are already final; when the data have arrived (``arrival.finish``), search the remaining tiles
    (prs_resample_nearest_search) and read back the other bands.  Output identical to prs_resample_nearest.
    """
    torch = _torch()
    main = torch.cuda.current_stream()
    up, down = _side_streams()
    row_bytes = arrival.row_bytes
    width, nrows = int(tg.grid.width), int(tg.nrows)
    stamps = [("enter", time.perf_counter())] if _TIMELINE is not None else None
    early = getattr(arrival, "start_early", True)
    if early:
        arrival.start(main, up)
    fill_row = _dev(fill_row_host)
    out = torch.empty((n_t, row_bytes), dtype=torch.uint8, device="cuda")
    scratch = torch.empty(int(L.prs_resample_scratch_bytes(ctypes.byref(tg))), dtype=torch.uint8, device="cuda")
    n_bands = (nrows + _cabi.PRS_TILE_ROWS - 1) // _cabi.PRS_TILE_ROWS
    band_live = torch.empty(n_bands, dtype=torch.uint8, device="cuda")
    handle = source.index.handle
    _cabi.check(L.prs_resample_nearest_fill(handle, ctypes.byref(tg), float(radius), row_bytes, _ptr(fill_row),
                                            _ptr(out), _ptr(scratch), _ptr(band_live), _stream()))
    host_out = torch.empty((n_t, row_bytes), dtype=torch.uint8, pin_memory=True)
    runs = _band_runs(_host(band_live), _cabi.PRS_TILE_ROWS, nrows, width * row_bytes, _PIPELINE_MIN_RUN_BYTES)
    if stamps is not None:
        stamps.append(("fill stage done", time.perf_counter()))
    down.wait_stream(main)
    with torch.cuda.stream(down):
        for r0, r1, is_live in runs:
            if not is_live:
                host_out[r0 * width:r1 * width].copy_(out[r0 * width:r1 * width], non_blocking=True)
    if not early:
        arrival.start(main, up)
    dev_data = arrival.finish(main)
    _cabi.check(L.prs_resample_nearest_search(handle, ctypes.byref(tg), float(radius), _ptr(dev_data), row_bytes,
                                              _ptr(fill_row), _ptr(out), _ptr(scratch), _stream()))
    down.wait_stream(main)
    with torch.cuda.stream(down):
        for r0, r1, is_live in runs:
            if is_live:
                host_out[r0 * width:r1 * width].copy_(out[r0 * width:r1 * width], non_blocking=True)
    if stamps is not None:
        stamps.append(("all work queued", time.perf_counter()))
    down.synchronize()
    main.synchronize()
    if stamps is not None:
        stamps.append(("done", time.perf_counter()))
        _TIMELINE.append([(n, (t - stamps[0][1]) * 1e3) for n, t in stamps] +
                         [("finished rows", sum(r1 - r0 for r0, r1, lv in runs if not lv)), ("rows", nrows)])
    return host_out.numpy()


def _tree_code(np_dtype):
    np_dtype = np.dtype(np_dtype)
    if np_dtype == np.float32:
        return PRS_F32
    if np_dtype == np.float64:
        return PRS_F64
    raise TypeError("coordinate dtype must be float32 or float64, got %s" % np_dtype)


def _coord_dtype(np_dtype):
    """dtype the ECEF tree is built in: float32 stays float32, everything else is float64
    (Cartesian.transform_lonlats, _spatial_mp.py:157-159; pykdtree converts other dtypes to float64)."""
    np_dtype = np.dtype(np_dtype)
    return np.dtype(np.float32) if np_dtype == np.float32 else np.dtype(np.float64)


class _Index(object):
    """Owner of a ``prs_index`` handle (stands in for the pykdtree ``KDTree`` object).

    The build is queued without waiting for the device; the numbers that describe the result (how many points are
    valid, the cell size, ...) live in device memory and are fetched the first time one of them is read."""

    def __init__(self, handle, lon_t, lat_t):
        self.handle = handle
        self._keep = (lon_t, lat_t)
        self._info = None

    def _fetch(self):
        if self._info is None:
            info = (ctypes.c_int64 * 8)()
            _cabi.check(_cabi.lib().prs_index_info(self.handle, info))
            self._info = tuple(int(v) for v in info)
        return self._info

    n_source = property(lambda self: self._fetch()[0])
    n = property(lambda self: self._fetch()[1])
    n_indexed = property(lambda self: self._fetch()[2])
    cells_per_side = property(lambda self: self._fetch()[3])
    bytes = property(lambda self: self._fetch()[4])
    tree_dtype = property(lambda self: self._fetch()[5])
    coarse_per_side = property(lambda self: self._fetch()[6])
    all_valid = property(lambda self: bool(self._fetch()[7]))

    def __del__(self):
        try:
            if self.handle is not None:
                _cabi.load_library().prs_index_destroy(self.handle)
                self.handle = None
        except Exception:  # pragma: no cover - interpreter shutdown
            pass


# =========================================================================================== geometry -> device
def _area_grid(area):
    return prs_area_grid(int(area.width), int(area.height), float(area.pixel_size_x), float(area.pixel_size_y),
                         float(area.pixel_upper_left[0]), float(area.pixel_upper_left[1]))


def _slice_params(sl, size):
    """(start, step, count, is_scalar) of an int or slice index along an axis of length ``size``."""
    if isinstance(sl, (int, np.integer)):
        i = int(sl)
        if i < 0:
            i += size
        if not 0 <= i < size:
            raise IndexError("index out of range")
        return i, 1, 1, True
    start, stop, step = sl.indices(size)
    return start, step, len(range(start, stop, step)), False


def _area_lonlats_device(area, data_slice, dtype, want_xyz=False):
    """AreaDefinition.get_lonlats on the GPU -> (lon, lat[, xyz]) CUDA tensors shaped like numpy's result."""
    torch = _torch()
    L = _cabi.lib()
    ys, xs = geometry._yx_slice(data_slice)
    r0, rs, nr, rscalar = _slice_params(ys if ys is not None else slice(None), area.height)
    c0, cs, nc, cscalar = _slice_params(xs if xs is not None else slice(None), area.width)
    td = getattr(torch, np.dtype(dtype).name)
    lon = torch.empty((nr, nc), dtype=td, device="cuda")
    lat = torch.empty((nr, nc), dtype=td, device="cuda")
    xyz = torch.empty((nr, nc, 3), dtype=td, device="cuda") if want_xyz else None
    grid = _area_grid(area)
    _cabi.check(L.prs_area_lonlats(ctypes.byref(area.proj_spec.cstruct), ctypes.byref(grid), _tree_code(dtype),
                                   r0, rs, nr, c0, cs, nc, _ptr(lon), _ptr(lat), _ptr(xyz), _stream()))
    if rscalar and cscalar:
        lon, lat = lon[0, 0], lat[0, 0]
    elif rscalar:
        lon, lat = lon[0], lat[0]
    elif cscalar:
        lon, lat = lon[:, 0], lat[:, 0]
    if want_xyz:
        return lon, lat, xyz
    return lon, lat


def _area_lonlats_covered(area, cover):
    """Raveled lon/lat of a whole area as one rank of a row-sharded run needs them (prs_area_lonlats_covered)."""
    torch = _torch()
    td = getattr(torch, np.dtype(area.dtype).name)
    lon = torch.empty(area.size, dtype=td, device="cuda")
    lat = torch.empty(area.size, dtype=td, device="cuda")
    grid = _area_grid(area)
    _cabi.check(_cabi.lib().prs_area_lonlats_covered(ctypes.byref(area.proj_spec.cstruct), ctypes.byref(grid),
                                                     _tree_code(area.dtype), _ptr(cover), _ptr(lon), _ptr(lat), _stream()))
    return lon, lat, None


def _area_lonlats_numpy(area, data_slice, dtype):
    lon, lat = _area_lonlats_device(area, data_slice, dtype)
    return lon.cpu().numpy(), lat.cpu().numpy()


def _cartesian_coords_numpy(geo_def, data_slice):
    """BaseDefinition.get_cartesian_coords (pyresample/geometry.py:465-516) on the GPU."""
    torch = _torch()
    L = _cabi.lib()
    if geometry.is_area(geo_def) and geo_def.lons is None:
        _, _, xyz = _area_lonlats_device(geo_def, data_slice, geo_def.dtype, want_xyz=True)
        return xyz.cpu().numpy()
    lons, lats = geo_def.get_lonlats(data_slice=data_slice)
    dt = _coord_dtype(_np_dtype_of(lons))
    lon_t, lat_t = _dev(lons, dt), _dev(lats, dt)
    n = lon_t.numel()
    xyz = torch.empty((n, 3), dtype=lon_t.dtype, device="cuda")
    _cabi.check(L.prs_lonlat_to_xyz(_ptr(lon_t), _ptr(lat_t), n, _tree_code(dt), _ptr(xyz), _stream()))
    out = xyz.cpu().numpy()
    if lon_t.dim() > 1:
        out = out.reshape(lon_t.shape[0], lon_t.shape[1], 3)
    return out


def _geo_lonlats_device(geo_def, dtype=None):
    """Full lon/lat of a geometry as raveled CUDA tensors + optional np.ma premask (uint8 tensor, 1 = masked)."""
    if geometry.is_area(geo_def) and getattr(geo_def, "lons", None) is None:
        dt = np.dtype(dtype if dtype is not None else geo_def.dtype)
        lon, lat = _area_lonlats_device(geo_def, None, dt)
        return lon.reshape(-1), lat.reshape(-1), None
    lons, lats = geo_def.get_lonlats()
    premask = None
    if isinstance(lons, np.ma.MaskedArray) or isinstance(lats, np.ma.MaskedArray):
        m = np.ma.getmaskarray(lons) | np.ma.getmaskarray(lats)
        if m.any():
            premask = _dev(m.ravel())
    dt = _coord_dtype(_np_dtype_of(lons))
    return _dev(lons, dt).reshape(-1), _dev(lats, dt).reshape(-1), premask


# =========================================================================================== reduce_data
def _to_host(a):
    if _is_tensor(a):
        return a.detach().cpu().numpy()
    return np.asanyarray(a)


def _boundary_sides(geo_def):
    """Host copies of the four boundary sides (geometry.py:284-291), lon then lat."""
    if geometry.is_area(geo_def) and getattr(geo_def, "lons", None) is None:
        cached = getattr(geo_def, "_prs_boundary_sides", None)
        if cached is not None:  # an AreaDefinition is immutable: its boundary is computed once
            return cached
        H, W = geo_def.height, geo_def.width
        slices = ((0, slice(None)), (slice(None), W - 1), (H - 1, slice(None, None, -1)), (slice(None, None, -1), 0))
        lons, lats = [], []
        for sl in slices:
            lo, la = _area_lonlats_device(geo_def, sl, geo_def.dtype)
            lons.append(lo.cpu().numpy().squeeze())
            lats.append(la.cpu().numpy().squeeze())
        try:
            geo_def._prs_boundary_sides = (lons, lats)
        except AttributeError:  # pragma: no cover - foreign object without a __dict__
            pass
        return lons, lats
    blons, blats = geo_def.get_boundary_lonlats()
    lons = [_to_host(s) for s in (blons.side1, blons.side2, blons.side3, blons.side4)]
    lats = [_to_host(s) for s in (blats.side1, blats.side2, blats.side3, blats.side4)]
    return lons, lats


def _reduce_box(boundary_geo_def, radius_of_influence):
    """Boundary analysis of data_reduce._get_valid_index (pyresample/data_reduce.py:224-305).

    This is the O(boundary) host part; the O(N) comparisons run in ``prs_valid_mask``.  The arithmetic is
    kept in the boundary arrays' own dtype because the buffered limits are part of the bit-exact
    ``valid_input_index`` contract.
    """
    is_cacheable = geometry.is_area(boundary_geo_def) and getattr(boundary_geo_def, "lons", None) is None
    if is_cacheable:  # an AreaDefinition is immutable: its box depends on the radius only
        cache = getattr(boundary_geo_def, "_prs_reduce_boxes", None)
        if cache is not None and float(radius_of_influence) in cache:
            return cache[float(radius_of_influence)]
    box = _reduce_box_uncached(boundary_geo_def, radius_of_influence)
    if is_cacheable:
        try:
            if getattr(boundary_geo_def, "_prs_reduce_boxes", None) is None:
                boundary_geo_def._prs_reduce_boxes = {}
            boundary_geo_def._prs_reduce_boxes[float(radius_of_influence)] = box
        except AttributeError:  # pragma: no cover - foreign object without a __dict__
            pass
    return box


def _reduce_box_uncached(boundary_geo_def, radius_of_influence):
    lon_sides, lat_sides = _boundary_sides(boundary_geo_def)
    box = prs_reduce_box()
    box.kind = PRS_REDUCE_ALL
    with np.errstate(invalid="ignore"):
        for s in lon_sides:
            if ((s < -180) | (s > 180)).any():
                return box
        for s in lat_sides:
            if ((s < -90) | (s > 90)).any():
                return box
    # winding sum of longitude steps; a step is skipped when the previous longitude is falsy (exactly 0.0)
    angle_sum = 0.0
    for side in lon_sides:
        s = np.atleast_1d(side).astype(np.float64)
        if s.size < 2:
            continue
        prev, cur = s[:-1], s[1:]
        delta = cur - prev
        with np.errstate(invalid="ignore", divide="ignore"):
            big = np.abs(delta) > 180
            wrapped = (np.abs(delta) - 360) * np.floor_divide(delta, np.abs(delta))
        delta = np.where(big, wrapped, delta)
        angle_sum += float(delta[prev != 0].sum())
    lat_min = min(s.min() for s in lat_sides)
    lat_max = max(s.max() for s in lat_sides)
    lat_min_buffered = lat_min - np.degrees(float(radius_of_influence) / R)
    lat_max_buffered = lat_max + np.degrees(float(radius_of_influence) / R)
    max_angle_s2 = max(abs(lat_sides[1].max()), abs(lat_sides[1].min()))
    max_angle_s4 = max(abs(lat_sides[3].max()), abs(lat_sides[3].min()))
    with np.errstate(divide="ignore", invalid="ignore"):
        lon_min_buffered = (lon_sides[3].min() -
                            np.degrees(float(radius_of_influence) / (np.sin(np.radians(max_angle_s4)) * R)))
        lon_max_buffered = (lon_sides[1].max() +
                            np.degrees(float(radius_of_influence) / (np.sin(np.radians(max_angle_s2)) * R)))
    box.lat_min, box.lat_max = float(lat_min_buffered), float(lat_max_buffered)
    box.lon_min, box.lon_max = float(lon_min_buffered), float(lon_max_buffered)
    if np.isnan(angle_sum):
        return box
    rounded = round(angle_sum)
    if rounded == -360:
        box.kind = PRS_REDUCE_LAT_GE
    elif rounded == 360:
        box.kind = PRS_REDUCE_LAT_LE
    elif rounded == 0:
        box.kind = PRS_REDUCE_BOX if lon_sides[1].min() > lon_sides[3].max() else PRS_REDUCE_BOX_DATELINE
    return box


def _valid_mask_device(lon_t, lat_t, dtype, box, premask, count=True):
    torch = _torch()
    n = lon_t.numel()
    valid = torch.empty(n, dtype=torch.uint8, device="cuda")
    n_valid = ctypes.c_int64(-1)
    _cabi.check(_cabi.lib().prs_valid_mask(_ptr(lon_t), _ptr(lat_t), n, _tree_code(dtype),
                                           ctypes.byref(box) if box is not None else None, _ptr(premask), _ptr(valid),
                                           ctypes.byref(n_valid) if count else None, _stream()))
    return valid, n_valid.value


# =========================================================================================== neighbour search
def _check_query_args(target_geo_def, radius_of_influence, neighbours, epsilon):
    """Type checks of _query_resample_kdtree (pyresample/kd_tree.py:503-510)."""
    if not geometry.is_base_definition(target_geo_def):
        raise TypeError('target_geo_def must be of geometry type')
    elif not isinstance(radius_of_influence, (int, float)):
        raise TypeError('radius_of_influence must be number')
    elif not isinstance(neighbours, int):
        raise TypeError('neighbours must be integer')
    elif not isinstance(epsilon, (int, float)):
        raise TypeError('epsilon must be number')


class _Source(object):
    """Device-side state of the source geometry: lon/lat, valid_input_index and the neighbour index."""

    def __init__(self, source_geo_def, target_geo_def, reduce_data, radius_of_influence, neighbours,
                 exclude_mask=None, tree_dtype=None, need_ranks=True, cover_rows=None, crop_to_cover=False):
        """``need_ranks``: keep the rank table (index_array is in rank space; the fused resample paths gather by
        source index and do not need it).  ``cover_rows``: this rank's rows of an area target - source points that
        cannot be within the radius of any of them are left out of the index (multi-GPU row sharding).
        ``crop_to_cover``: an AREA source is not even transformed where the rank cannot need it
        (prs_area_lonlats_covered): same index, but ``valid`` / ``n_valid`` then describe the cropped source - only
        for the fused nearest path, which reads neither."""
        _cabi.lib()  # fail loudly when the CUDA library or the GPU is missing: there is no CPU fallback
        source_geo_def = geometry.as_geo_def(source_geo_def)
        target_geo_def = geometry.as_geo_def(target_geo_def)
        self.geo_def = source_geo_def
        cover = None
        lon_t = None
        if cover_rows is not None and target_geo_def is not None:
            src_is_area = geometry.is_area(source_geo_def) and getattr(source_geo_def, "lons", None) is None
            tdt0 = np.dtype(tree_dtype) if tree_dtype is not None else (
                np.dtype(source_geo_def.dtype) if src_is_area else None)
            if tdt0 is not None:
                cover = _target_cover(target_geo_def, cover_rows, radius_of_influence, tdt0)
            if crop_to_cover and cover is not None and src_is_area and not need_ranks:
                lon_t, lat_t, premask = _area_lonlats_covered(source_geo_def, cover)
        if lon_t is None:
            lon_t, lat_t, premask = _geo_lonlats_device(source_geo_def)
        if lon_t.numel() == 0 or lat_t.numel() == 0:
            raise ValueError('Cannot resample empty data set')
        elif lon_t.numel() != lat_t.numel():
            raise ValueError('Mismatch between lons and lats')
        self.dtype = _np_dtype_of(lon_t)  # coordinate dtype of the chain (kd_tree.py:513-514)
        box = None
        if reduce_data and target_geo_def is not None:
            # kd_tree.py:410-424: swath/grid/area source onto a grid-like target
            if (geometry.is_coord(source_geo_def) or geometry.is_griddish(source_geo_def)) and \
                    geometry.is_griddish(target_geo_def):
                box = _reduce_box(target_geo_def, radius_of_influence)
        # argument types (kd_tree.py:503-510); the reference checks them when it queries, i.e. after the
        # reduce_data analysis above has already used the radius
        if not isinstance(radius_of_influence, (int, float)):
            raise TypeError('radius_of_influence must be number')
        elif not isinstance(neighbours, int):
            raise TypeError('neighbours must be integer')
        # validity (range test, reduce_data box, np.ma mask) is evaluated inside the index build; the number of
        # valid points comes back with the build's single host round trip
        self.valid = _torch().empty(lon_t.numel(), dtype=_torch().uint8, device="cuda")
        self.lon_t, self.lat_t = lon_t, lat_t
        tdt = np.dtype(tree_dtype) if tree_dtype is not None else self.dtype
        self.tree_dtype = tdt
        # float32 lon/lat into a float64 tree (XArrayResamplerNN): the reference transforms in float32 and widens the
        # ECEF coordinates afterwards (kd_tree.py:970-981) - the build does the same, so the arrays go in as they are
        if tdt != self.dtype and not (self.dtype == np.float32 and tdt == np.float64):
            lon_t, lat_t = _dev(lon_t, tdt), _dev(lat_t, tdt)
        in_dtype = _np_dtype_of(lon_t)
        handle = ctypes.c_void_p()
        excl = _dev(exclude_mask).reshape(-1) if exclude_mask is not None else None
        if cover is None and cover_rows is not None and target_geo_def is not None:
            cover = _target_cover(target_geo_def, cover_rows, radius_of_influence, tdt)
        flags = _cabi.PRS_BUILD_COMPUTE_VALID | (_cabi.PRS_BUILD_RANKS if need_ranks else 0)
        _cabi.check(_cabi.lib().prs_index_build(_ptr(lon_t), _ptr(lat_t), lon_t.numel(), _tree_code(in_dtype),
                                               _ptr(self.valid), ctypes.byref(box) if box is not None else None,
                                               _ptr(premask), _ptr(excl), _tree_code(tdt), int(neighbours),
                                               float(radius_of_influence), _ptr(cover), flags, ctypes.byref(handle),
                                               _stream()))
        self._index = _Index(handle, lon_t, lat_t)
        self.handle = handle  # for launches that need not know yet whether any source point is valid

    @property
    def n_valid(self):
        """Number of valid source points (pykdtree's ``.n``); the first read waits for the build's plan kernel."""
        return int(self._index.n)

    @property
    def index(self):
        """The index, or None when no source point is valid (EmptyResult of _create_resample_kdtree,
        kd_tree.py:480-481)."""
        return self._index if self.n_valid > 0 else None


class _TreeSpec(object):
    """What _make_target needs to know about the source before the index exists."""

    def __init__(self, tree_dtype, geo_def=None):
        self.tree_dtype, self.geo_def = np.dtype(tree_dtype), geo_def


def _rows_tuple(rows):
    """(row0, nrows, row_block, row_stride) of a row specification: None, (row0, nrows) or a 4-tuple."""
    if rows is None:
        return 0, None, 0, 0
    rows = tuple(int(v) for v in rows)
    return rows + (0, 0) if len(rows) == 2 else rows


def _target_cover(target_geo_def, rows, radius_of_influence, tree_dtype):
    """Coverage mask (prs_target_coverage) of one rank's rows of an area target; cached on the area object."""
    target_geo_def = geometry.as_geo_def(target_geo_def)
    if not (geometry.is_area(target_geo_def) and getattr(target_geo_def, "lons", None) is None):
        return None
    torch = _torch()
    key = (_rows_tuple(rows), float(radius_of_influence), np.dtype(tree_dtype).name, torch.cuda.current_device())
    cache = getattr(target_geo_def, "_prs_cover_cache", None)
    if cache is None:
        cache = target_geo_def._prs_cover_cache = {}
    if key not in cache:
        tg, n, keep = _make_target(_TreeSpec(tree_dtype), target_geo_def, False, radius_of_influence, rows=rows)
        cover = torch.empty(6 * _cabi.PRS_COVER_CELLS ** 2, dtype=torch.uint8, device="cuda")
        _cabi.check(_cabi.lib().prs_target_coverage(ctypes.byref(tg), float(radius_of_influence), _ptr(cover), _stream()))
        cache[key] = cover
    return cache[key]


def _make_target(source, target_geo_def, reduce_data, radius_of_influence, row0=0, nrows=None, rows=None):
    """Describe the query points for the C-ABI.  Returns (prs_target, n, keepalive)."""
    target_geo_def = geometry.as_geo_def(target_geo_def)
    row_block = row_stride = 0
    if rows is not None:
        row0, nrows, row_block, row_stride = _rows_tuple(rows)
    tg = prs_target()
    keep = []
    tree_dt = source.tree_dtype
    valid_extra = None
    grid_to_swath = reduce_data and geometry.is_griddish(source.geo_def) and geometry.is_coord(target_geo_def)
    if geometry.is_area(target_geo_def) and getattr(target_geo_def, "lons", None) is None:
        if nrows is None:
            nrows = target_geo_def.height
        tg.kind = PRS_TARGET_AREA
        tg.dtype = _tree_code(tree_dt)
        tg.proj = target_geo_def.proj_spec.cstruct
        tg.grid = _area_grid(target_geo_def)
        tg.row0, tg.nrows = int(row0), int(nrows)
        tg.row_block, tg.row_stride = int(row_block), int(row_stride)
        tg.n = int(nrows) * int(target_geo_def.width)
        return tg, tg.n, keep
    if row_block:
        raise NotImplementedError("block-cyclic row ownership is implemented for AreaDefinition targets")
    lons, lats = target_geo_def.get_lonlats()
    premask = None
    if isinstance(lons, np.ma.MaskedArray) or isinstance(lats, np.ma.MaskedArray):
        m = np.ma.getmaskarray(lons) | np.ma.getmaskarray(lats)
        if m.any():
            premask = m.ravel()
    lon_t = _dev(lons, tree_dt).reshape(-1)
    lat_t = _dev(lats, tree_dt).reshape(-1)
    if row0 != 0 or nrows is not None:
        ncols = int(np.prod(target_geo_def.shape[1:])) if len(target_geo_def.shape) > 1 else 1
        nrows = target_geo_def.shape[0] - row0 if nrows is None else nrows
        lon_t = lon_t[row0 * ncols:(row0 + nrows) * ncols]
        lat_t = lat_t[row0 * ncols:(row0 + nrows) * ncols]
        if premask is not None:
            premask = premask[row0 * ncols:(row0 + nrows) * ncols]
    if grid_to_swath or premask is not None:
        # kd_tree.py:438-451: reduce the *target* points against the source grid's boundary
        box = _reduce_box(source.geo_def, radius_of_influence) if grid_to_swath else None
        valid_extra, _ = _valid_mask_device(lon_t, lat_t, tree_dt, box, _dev(premask) if premask is not None else None)
        keep.append(valid_extra)
    tg.kind = PRS_TARGET_LONLAT
    tg.dtype = _tree_code(tree_dt)
    tg.lon, tg.lat = lon_t.data_ptr(), lat_t.data_ptr()
    tg.n = lon_t.numel()
    tg.valid_extra = valid_extra.data_ptr() if valid_extra is not None else None
    keep += [lon_t, lat_t]
    return tg, tg.n, keep


def _query_device(source, target_geo_def, radius_of_influence, neighbours, reduce_data, want_dist=True,
                  row0=0, nrows=None, rows=None):
    """Run ``prs_query``; returns device tensors covering ALL target points + (kth_found, n_invalid)."""
    torch = _torch()
    tg, n, keep = _make_target(source, target_geo_def, reduce_data, radius_of_influence, row0, nrows, rows=rows)
    k = int(neighbours)
    idx = torch.empty((n, k), dtype=torch.int32, device="cuda")
    td = getattr(torch, source.tree_dtype.name)
    dist = torch.empty((n, k), dtype=td, device="cuda") if want_dist else None
    valid = torch.empty(n, dtype=torch.uint8, device="cuda")
    flags = (ctypes.c_int64 * 2)()
    _cabi.check(_cabi.lib().prs_query(source.index.handle, ctypes.byref(tg), k, float(radius_of_influence),
                                      _ptr(idx), _ptr(dist), _ptr(valid), flags, _stream()))
    del keep
    return idx, dist, valid, bool(flags[0]), int(flags[1])


def _compact_rows(mask_t, rows_t, kept=None):
    """``rows[mask]`` on the device (prs_compact_rows); ``kept`` = the number of set mask bytes when already known."""
    torch = _torch()
    n = mask_t.numel()
    flat = rows_t.reshape(n, -1)
    row_bytes = flat.shape[1] * flat.element_size()
    L = _cabi.lib()
    if kept is None:
        counted = ctypes.c_int64(0)
        _cabi.check(L.prs_compact_rows(_ptr(mask_t), n, row_bytes, None, None, ctypes.byref(counted), _stream()))
        kept = counted.value
    kept = int(kept)
    out = torch.empty((kept, flat.shape[1]), dtype=flat.dtype, device="cuda")
    if kept:
        _cabi.check(L.prs_compact_rows(_ptr(mask_t), n, row_bytes, _ptr(flat), _ptr(out), None, _stream()))
    return out.reshape((kept,) + tuple(rows_t.shape[1:]))


def _create_empty_info(source_geo_def, target_geo_def, neighbours):
    """Dummy info for an empty result set (pyresample/kd_tree.py:553-563)."""
    valid_output_index = np.ones(target_geo_def.size, dtype=bool)
    if neighbours > 1:
        index_array = (np.ones((target_geo_def.size, neighbours), dtype=np.int32) * source_geo_def.size)
        distance_array = np.ones((target_geo_def.size, neighbours))
    else:
        index_array = (np.ones(target_geo_def.size, dtype=np.int32) * source_geo_def.size)
        distance_array = np.ones(target_geo_def.size)
    return valid_output_index, index_array, distance_array


def get_neighbour_info(source_geo_def, target_geo_def, radius_of_influence,
                       neighbours=8, epsilon=0, reduce_data=True,
                       nprocs=1, segments=None):
    """Return neighbour info (pyresample/kd_tree.py:281-386).

    ``nprocs`` and ``segments`` are accepted for compatibility; the GPU processes the whole target at
    once and the result is identical for any value.  ``epsilon`` > 0 permits an approximate answer in
    the reference; the exact answer returned here always satisfies it.

    Returns
    -------
    (valid_input_index, valid_output_index, index_array, distance_array) : tuple of numpy arrays
    """
    if source_geo_def.size < neighbours:
        warnings.warn('Searching for %s neighbours in %s data points' %
                      (neighbours, source_geo_def.size), stacklevel=2)
    source = _Source(source_geo_def, target_geo_def, reduce_data, radius_of_influence, neighbours)
    valid_input_index = source.valid.cpu().numpy().astype(bool)
    if source.index is None:
        # all input data reduced away (kd_tree.py:336-341)
        valid_output_index, index_array, distance_array = _create_empty_info(source_geo_def, target_geo_def,
                                                                             neighbours)
        return valid_input_index, valid_output_index, index_array, distance_array
    _check_query_args(geometry.as_geo_def(target_geo_def), radius_of_influence, neighbours, epsilon)
    idx, dist, valid, kth_found, n_invalid = _query_device(source, target_geo_def, radius_of_influence, neighbours,
                                                           reduce_data)
    if n_invalid:
        idx = _compact_rows(valid, idx)
        dist = _compact_rows(valid, dist)
    valid_output_index = valid.cpu().numpy().astype(bool)
    index_array = _host(idx).view(np.uint32)
    distance_array = _host(dist)
    if neighbours == 1:
        index_array = index_array.reshape(-1)
        distance_array = distance_array.reshape(-1)
    elif kth_found:
        # kd_tree.py:380-384
        warnings.warn(('Possible more than %s neighbours within %s m for some data points') %
                      (neighbours, radius_of_influence), stacklevel=2)
    return valid_input_index, valid_output_index, index_array, distance_array


# =========================================================================================== sampling helpers
def _get_fill_mask_value(data_dtype):
    """Maximum value of dtype (pyresample/kd_tree.py:1186-1195)."""
    if issubclass(data_dtype.type, np.floating):
        fill_value = np.finfo(data_dtype.type).max
    elif issubclass(data_dtype.type, np.integer):
        fill_value = np.iinfo(data_dtype.type).max
    else:
        raise TypeError('Type %s is unsupported for masked fill values' % data_dtype.type)
    return fill_value


def _remask_data(data, is_to_be_masked=True):
    """Interpret half the channels as the mask of the other half (pyresample/kd_tree.py:1198-1211)."""
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


def _prepare_result(result, output_shape, is_masked_data, use_masked_fill_value, fill_value, input_data_type):
    """pyresample/kd_tree.py:884-899."""
    result = result.reshape(output_shape)
    if is_masked_data:
        result = _remask_data(result)
    if use_masked_fill_value:
        result = np.ma.masked_equal(result, fill_value)
    if input_data_type is not None:
        # the reference's astype copies; the array here is already a fresh buffer owned by the caller
        result = result.astype(input_data_type, copy=False)
    return result


def _prepare_uncertainty(stddev, count, output_shape, is_masked_data, result):
    """_prepare_and_fill_uncertainty_result + masking (pyresample/kd_tree.py:862-881, 734-736)."""
    stddev = np.asarray(stddev, dtype=np.float64).reshape(output_shape)
    count = np.asarray(count, dtype=np.float64).reshape(output_shape)
    if is_masked_data:
        stddev = _remask_data(stddev, is_to_be_masked=False)
        count = _remask_data(count, is_to_be_masked=False)
    stddev = np.ma.array(stddev, mask=np.isnan(stddev))
    if np.ma.isMA(result):
        stddev = np.ma.array(stddev, mask=(result.mask | stddev.mask))
        count = np.ma.array(count, mask=result.mask)
    return stddev, count


def _get_empty_sample(data, output_shape, is_multi_channel, fill_value):
    """pyresample/kd_tree.py:655-671."""
    if is_multi_channel:
        output_shape = list(output_shape)
        output_shape.append(data.shape[1])
    if fill_value is None:
        return np.ma.array(np.zeros(output_shape, data.dtype), mask=np.ones(output_shape, dtype=bool))
    return np.ones(output_shape, dtype=data.dtype) * fill_value


def _flatten_data(data, source_size):
    """Reshape data to (N_src[, C]) (pyresample/kd_tree.py:607-613)."""
    if data.ndim > 2 and data.shape[0] * data.shape[1] == source_size:
        data = data.reshape(data.shape[0] * data.shape[1], data.shape[2])
    elif data.shape[0] != source_size:
        data = data.ravel()
    if source_size != data.shape[0]:
        raise ValueError('Mismatch between geometry and dataset')
    return data


def _byte_rows(arr):
    """(N[, C]) array of any dtype -> contiguous (N, row_bytes) uint8 view: the gather kernels move opaque rows."""
    arr = np.ascontiguousarray(arr)
    return arr.reshape(arr.shape[0], -1).view(np.uint8)


def _stack_mask_channels(data):
    """Masked input: append the mask as extra channels (pyresample/kd_tree.py:640-645).

    The reference tests ``np.ma.is_masked(data[valid_input_index])`` - at least one masked *valid* element.
    """
    return np.column_stack((np.ma.getdata(data), np.ma.getmaskarray(data)))


def _weight_funcs_for_masked(weight_funcs, is_multi_channel):
    """pyresample/kd_tree.py:697-698."""
    return weight_funcs * 2 if is_multi_channel else (weight_funcs,) * 2


class _GaussWeight(object):
    """Marker carried by the closures built in :func:`resample_gauss` so the Gaussian can be fused on the GPU."""

    registry = {}


def _gauss_sigma_of(func):
    return getattr(func, "_prs_gauss_sigma", None)


def _weighted_dtypes(dist_dtype, data_dtype, is_multi):
    """(device data dtype, result dtype) following numpy promotion in _resample_with_weights."""
    wdt = np.dtype(np.float64) if is_multi else np.dtype(dist_dtype)
    if data_dtype.kind == 'c':
        raise NotImplementedError("complex data is only supported for nearest neighbour resampling")
    rt = np.result_type(wdt, data_dtype)
    if rt not in (np.dtype(np.float32), np.dtype(np.float64)):
        rt = np.dtype(np.float64)
    if rt == np.float32 or data_dtype == np.float32:
        return np.dtype(np.float32), rt
    return np.dtype(np.float64), rt


# =========================================================================================== public sampling
def get_sample_from_neighbour_info(resample_type, output_shape, data,
                                   valid_input_index, valid_output_index,
                                   index_array, distance_array=None,
                                   weight_funcs=None, fill_value=0,
                                   with_uncert=False):
    """Resample swath based on precomputed neighbour info (pyresample/kd_tree.py:566-652)."""
    data = _flatten_data(data, valid_input_index.size)
    is_multi_channel = (data.ndim > 1)
    valid_input_size = valid_input_index.sum()
    valid_output_size = valid_output_index.sum()
    if valid_input_size == 0 or valid_output_size == 0:
        return _get_empty_sample(data, output_shape, is_multi_channel, fill_value)
    if not isinstance(data, np.ndarray):
        raise TypeError('data must be numpy array')
    elif valid_input_index.ndim != 1:
        raise TypeError('valid_index must be one dimensional array')
    elif data.shape[0] != valid_input_index.size:
        raise TypeError('Not the same number of datapoints in valid_input_index and data')
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

    torch = _torch()
    L = _cabi.lib()
    output_shape = tuple(output_shape)
    input_data_type = data.dtype if resample_type == 'nn' else None
    # masked input -> mask channels (decided on the valid subset, kd_tree.py:642)
    is_masked_data = bool(np.ma.is_masked(data[valid_input_index])) if isinstance(data, np.ma.MaskedArray) else False
    if is_masked_data:
        new_data = _stack_mask_channels(data)
        if weight_funcs is not None:
            weight_funcs = _weight_funcs_for_masked(weight_funcs, is_multi_channel)
    else:
        new_data = np.ma.getdata(data)
    channels = new_data.shape[1] if new_data.ndim > 1 else 1
    use_masked_fill_value = False
    if fill_value is None:
        use_masked_fill_value = True
        fill_value = _get_fill_mask_value(new_data.dtype)

    valid_in_t = _dev(np.asarray(valid_input_index, dtype=bool))
    all_valid_in = int(valid_input_size) == valid_input_index.size
    rank_to_source = None
    if not all_valid_in:
        rank_to_source = torch.empty(int(valid_input_size), dtype=torch.int32, device="cuda")
        _cabi.check(L.prs_rank_to_source(_ptr(valid_in_t), valid_in_t.numel(), _ptr(rank_to_source), None, _stream()))
    n_out = int(valid_output_size)
    idx_np = np.ascontiguousarray(index_array).astype(np.uint32, copy=False)
    idx_t = _dev(idx_np.view(np.int32).reshape(n_out, neighbours))
    output_size = int(np.prod(output_shape))
    all_valid_out = n_out == output_size
    voi_t = None if all_valid_out else _dev(np.asarray(valid_output_index, dtype=bool))
    final_output_shape = output_shape + (channels,) if new_data.ndim > 1 else output_shape
    stddev = count = None

    if resample_type == 'nn' or neighbours == 1:
        # kd_tree.py:705-711 (note: 'custom' with one neighbour is a plain gather too)
        dev_data = _dev(_byte_rows(new_data))
        row_bytes = dev_data.shape[1]
        fill_row = _dev(_byte_rows(np.full((1, channels), fill_value, dtype=new_data.dtype)))
        res_t = torch.empty((n_out, row_bytes), dtype=torch.uint8, device="cuda")
        _cabi.check(L.prs_gather_nearest(_ptr(idx_t), n_out, int(valid_input_size), _ptr(rank_to_source),
                                         _ptr(dev_data), row_bytes, _ptr(fill_row), _ptr(res_t), _stream()))
        res_t = _scatter_full(res_t, voi_t, output_size, fill_row)
        result = _host(res_t).view(new_data.dtype)
        if with_uncert:
            stddev = count = None
    else:
        if distance_array is None:
            raise ValueError("distance_array is required for custom resampling")
        dist_np = np.ascontiguousarray(distance_array)
        if dist_np.dtype not in (np.float32, np.float64):
            dist_np = dist_np.astype(np.float64)
        ddt, rt = _weighted_dtypes(dist_np.dtype, new_data.dtype, new_data.ndim > 1)
        dev_data = _dev(new_data.astype(ddt, copy=False)).reshape(new_data.shape[0], -1)
        dist_t = _dev(dist_np.reshape(n_out, neighbours))
        res_t, sd_t, cnt_t = _weighted_from_info(idx_t, dist_t, n_out, neighbours, int(valid_input_size), rank_to_source,
                                                 dev_data, ddt, channels, weight_funcs, new_data.ndim > 1, fill_value,
                                                 with_uncert, dist_np)
        fill_row = torch.full((channels,), float(fill_value), dtype=res_t.dtype, device="cuda")
        res_t = _scatter_full(res_t, voi_t, output_size, fill_row)
        result = _host(res_t).astype(rt, copy=False)
        if with_uncert:
            nan_row = torch.full((channels,), float("nan"), dtype=sd_t.dtype, device="cuda")
            zero_row = torch.zeros((channels,), dtype=cnt_t.dtype, device="cuda")
            stddev = _scatter_full(sd_t, voi_t, output_size, nan_row).cpu().numpy()
            count = _scatter_full(cnt_t, voi_t, output_size, zero_row).cpu().numpy()
    result = _prepare_result(result, final_output_shape, is_masked_data, use_masked_fill_value, fill_value,
                             input_data_type)
    if with_uncert:
        if stddev is None:  # neighbours == 1: the reference returns (result, None, None) prepared values
            raise TypeError("uncertainty estimates need more than one neighbour")
        stddev, count = _prepare_uncertainty(stddev, count, final_output_shape, is_masked_data, result)
        return result, stddev, count
    return result


def _scatter_full(rows_t, voi_t, output_size, fill_row_t):
    """full_result[valid_output_index] = result (pyresample/kd_tree.py:719-723) on the device."""
    if voi_t is None:
        return rows_t
    torch = _torch()
    flat = rows_t.reshape(rows_t.shape[0], -1)
    row_bytes = flat.shape[1] * flat.element_size()
    out = torch.empty((output_size, flat.shape[1]), dtype=flat.dtype, device="cuda")
    _cabi.check(_cabi.lib().prs_scatter_rows(_ptr(voi_t), output_size, row_bytes, _ptr(flat), _ptr(fill_row_t),
                                             _ptr(out), _stream()))
    return out


def _weights_finite_at_one(weight_funcs):
    """True when every weight function maps the distance 1 (what missing neighbours are given, kd_tree.py:767-768)
    to a finite value, so that ``0 * w`` is 0 for them."""
    funcs = weight_funcs if isinstance(weight_funcs, (list, tuple)) else [weight_funcs]
    try:
        with np.errstate(all="ignore"):
            return all(bool(np.all(np.isfinite(np.asarray(f(np.ones(1)), dtype=np.float64)))) for f in funcs)
    except Exception:  # a function that does not take a length-1 array: no shortcut
        return False


def _eval_weight_planes(weight_funcs, channels, is_multi, dist_np, index_missing):
    """Evaluate arbitrary Python weight functions on the host distances (kd_tree.py:764-787).

    Returns (planes ndarray [n_planes, n_out, k], plane_of_channel list).  Identical function objects share
    a plane.  Distances of missing neighbours are set to 1 before the call (kd_tree.py:767-768).
    """
    funcs = list(weight_funcs) if is_multi else [weight_funcs]
    if is_multi and len(funcs) < channels:
        raise IndexError("list index out of range")
    dist = dist_np.copy()
    dist[index_missing] = 1
    planes, plane_of_channel, seen = [], [], {}
    n_out, k = dist.shape
    for c in range(channels):
        f = funcs[c] if is_multi else funcs[0]
        key = id(f)
        if key not in seen:
            cols = []
            for i in range(k):
                w = f(dist[:, i])
                if is_multi:
                    w = np.ones(n_out) * w  # kd_tree.py:783: always float64
                else:
                    w = np.asarray(w)
                    if w.ndim == 0:
                        w = np.full(n_out, w)
                cols.append(w)
            plane = np.stack(cols, axis=1)
            seen[key] = len(planes)
            planes.append(plane)
        plane_of_channel.append(seen[key])
    wdt = np.result_type(*[p.dtype for p in planes])
    if wdt not in (np.dtype(np.float32), np.dtype(np.float64)):
        wdt = np.dtype(np.float64)
    planes = np.stack([p.astype(wdt, copy=False) for p in planes], axis=0)
    return np.ascontiguousarray(planes), plane_of_channel


_DEVICE_WEIGHT_CHUNK = 1 << 29  # (row, neighbour) entries evaluated per call of a weight function on the device (4 GiB of float64)


def _missing_and_dist(idx_t, dist_t, n_valid):
    """Device copies of what the reference feeds to a weight function: the missing-neighbour mask and the distances
    with ``1`` in place of missing neighbours (kd_tree.py:751-768), evaluated in float64."""
    torch = _torch()
    idx_t, dist_t = idx_t.contiguous(), dist_t.contiguous()
    d = torch.empty(dist_t.shape, dtype=torch.float64, device="cuda")
    _cabi.check(_cabi.lib().prs_weight_distances(_ptr(idx_t), _ptr(dist_t), _tree_code(_np_dtype_of(dist_t)),
                                                 idx_t.numel(), int(n_valid), _ptr(d), _stream()))
    return d


def _eval_weight_planes_device(weight_funcs, channels, is_multi, idx_t, dist_t, n_valid):
    """Evaluate the weight functions (kd_tree.py:764-787) on a CUDA view of the ``(n_out, k)`` distances.

    Works for anything written with arithmetic operators or ``torch`` functions (``1 - d / 1e5``, constants, ...).
    The function sees float64 distances; the plane is then cast to the dtype numpy would have produced (probed on a
    two-element host array), so the result is the correctly rounded value of the exact weight - within 1 ulp of the
    reference's own float32 evaluation.  Returns None when some function does not accept a tensor (``np.cos`` ...):
    the caller then takes the host route.  Identical function objects share a plane.
    """
    torch = _torch()
    funcs = list(weight_funcs) if is_multi else [weight_funcs]
    if is_multi and len(funcs) < channels:
        raise IndexError("list index out of range")
    dist_dtype = _np_dtype_of(dist_t)
    n_out, k = dist_t.shape
    planes, plane_of_channel, seen, dtypes = [], [], {}, []
    for c in range(channels):
        f = funcs[c] if is_multi else funcs[0]
        if id(f) in seen:
            plane_of_channel.append(seen[id(f)])
            continue
        try:
            with np.errstate(all="ignore"):
                probe = f(np.ones(2, dtype=dist_dtype))
            if is_multi:
                wdt = np.dtype(np.float64)  # np.ones(n) * w (kd_tree.py:783)
            else:
                wdt = np.asarray(probe).dtype
                if wdt not in (np.dtype(np.float32), np.dtype(np.float64)):
                    wdt = np.dtype(np.float64)
            pdt = getattr(torch, wdt.name)
            flat_i, flat_d = idx_t.reshape(-1), dist_t.reshape(-1)
            if n_out * k <= _DEVICE_WEIGHT_CHUNK:
                # one call: a tensor result IS the plane (no copy)
                w = f(_missing_and_dist(flat_i, flat_d, n_valid))
                if _is_tensor(w) and w.is_cuda and w.numel() == n_out * k and w.dim() == 1:
                    plane = w.to(pdt).contiguous().reshape(n_out, k)
                    seen[id(f)] = len(planes)
                    planes.append(plane)
                    dtypes.append(wdt)
                    plane_of_channel.append(seen[id(f)])
                    continue
            plane = torch.empty((n_out, k), dtype=pdt, device="cuda")
            flat_p = plane.reshape(-1)
            for off in range(0, n_out * k, _DEVICE_WEIGHT_CHUNK):
                sl = slice(off, min(off + _DEVICE_WEIGHT_CHUNK, n_out * k))
                d = _missing_and_dist(flat_i[sl], flat_d[sl], n_valid)
                w = f(d)
                if _is_tensor(w):
                    if not w.is_cuda or tuple(w.shape) != tuple(d.shape):
                        return None
                    flat_p[sl] = w
                elif isinstance(w, (int, float, np.integer, np.floating)):
                    flat_p[sl] = float(w)
                else:
                    return None
        except Exception:  # the function needs a numpy array: host route
            return None
        seen[id(f)] = len(planes)
        planes.append(plane)
        dtypes.append(wdt)
        plane_of_channel.append(seen[id(f)])
    wdt = np.result_type(*dtypes)
    td = getattr(torch, np.dtype(wdt).name)
    if len(planes) == 1:
        return planes[0].to(td).reshape(1, n_out, k), plane_of_channel, np.dtype(wdt)
    return torch.stack([p.to(td) for p in planes], dim=0).contiguous(), plane_of_channel, np.dtype(wdt)


def _weighted_from_info(idx_t, dist_t, n_out, k, n_valid, rank_to_source, dev_data, ddt, channels, weight_funcs,
                        is_multi, fill_value, with_uncert, dist_np=None):
    """prs_gather_weighted with Gaussian weights (computed in the kernel) or arbitrary weight functions: evaluated on
    the device when they accept a tensor, otherwise on the host."""
    torch = _torch()
    L = _cabi.lib()
    funcs = list(weight_funcs) if is_multi else [weight_funcs] * channels
    if is_multi and len(funcs) < channels:
        raise IndexError("list index out of range")
    sigmas = [_gauss_sigma_of(f) for f in funcs[:channels]]
    dist_dtype = _np_dtype_of(dist_t)
    if all(s is not None for s in sigmas):
        acc64 = channels > 1 or dist_dtype == np.float64 or ddt == np.float64
        weights_t, wcode, n_planes, planes_c, sig_c = None, _tree_code(dist_dtype), 0, None, \
            (ctypes.c_double * channels)(*[float(s) for s in sigmas])
    else:
        dev_planes = _eval_weight_planes_device(weight_funcs, channels, is_multi, idx_t, dist_t, n_valid)
        if dev_planes is not None:
            weights_t, plane_of_channel, wdt = dev_planes
        else:
            if dist_np is None:
                dist_np = dist_t.cpu().numpy()
            idx_np = idx_t.cpu().numpy().view(np.uint32)
            planes, plane_of_channel = _eval_weight_planes(weight_funcs, channels, is_multi, dist_np, idx_np >= n_valid)
            weights_t = _dev(planes)
            wdt = planes.dtype
        wcode = _tree_code(wdt)
        n_planes = int(weights_t.shape[0])
        planes_c = (ctypes.c_int32 * channels)(*plane_of_channel)
        sig_c = None
        acc64 = channels > 1 or wdt == np.float64 or ddt == np.float64
    odt = torch.float64 if acc64 else torch.float32
    out = torch.empty((n_out, channels), dtype=odt, device="cuda")
    sd = torch.empty((n_out, channels), dtype=odt, device="cuda") if with_uncert else None
    cnt = torch.empty((n_out, channels), dtype=odt, device="cuda") if with_uncert else None
    _cabi.check(L.prs_gather_weighted(_ptr(idx_t), _ptr(dist_t), n_out, k, n_valid, _ptr(rank_to_source),
                                      _ptr(dev_data), _tree_code(ddt), channels, _ptr(weights_t), wcode, n_planes,
                                      planes_c, sig_c, float(fill_value), _ptr(out), _ptr(sd), _ptr(cnt), _stream()))
    return out, sd, cnt


# =========================================================================================== fused resampling
def _resample(source_geo_def, data, target_geo_def, resample_type,
              radius_of_influence, neighbours=8, epsilon=0, weight_funcs=None,
              fill_value=0, reduce_data=True, nprocs=1, segments=None, with_uncert=False, _rows=None):
    """Resample using the kd-tree approach (pyresample/kd_tree.py:256-278), fused on the GPU when possible.

    ``_rows=(row0, nrows)`` restricts the output to one row block of the target (multi-GPU sharding,
    pyresample_b200/distributed.py); the source mask and index are those of the whole target.
    """
    if source_geo_def.size < neighbours:
        warnings.warn('Searching for %s neighbours in %s data points' %
                      (neighbours, source_geo_def.size), stacklevel=3)
    tensor_io = _is_tensor(data)
    # the fused paths gather by source index; only arbitrary Python weight functions go through index_array
    fused = resample_type == 'nn' or neighbours == 1 or weight_funcs is None or all(
        _gauss_sigma_of(f) is not None for f in (weight_funcs if isinstance(weight_funcs, (list, tuple))
                                                 else [weight_funcs]))
    source = _Source(source_geo_def, target_geo_def, reduce_data, radius_of_influence, neighbours,
                     need_ranks=not fused, cover_rows=_rows,
                     crop_to_cover=resample_type == 'nn' and _rows is not None and _is_tensor(data))
    tgt = geometry.as_geo_def(target_geo_def)
    output_shape = tuple(tgt.shape)
    n_src = int(source.lon_t.numel())
    if tensor_io:
        return _resample_tensor(source, data, tgt, resample_type, radius_of_influence, neighbours, weight_funcs,
                                fill_value, reduce_data, with_uncert, _rows)
    if _rows is not None:
        raise NotImplementedError("row-block resampling is implemented for CUDA tensor inputs")
    data = _flatten_data(data, n_src)
    is_multi_channel = data.ndim > 1
    if source.index is None:
        # empty neighbour info -> empty sample (kd_tree.py:336-341, 620-621)
        return _get_empty_sample(data, output_shape, is_multi_channel, fill_value)
    _check_query_args(tgt, radius_of_influence, neighbours, epsilon)
    if not isinstance(data, np.ndarray):
        raise TypeError('data must be numpy array')
    if resample_type not in ("nn", "custom"):
        raise TypeError(f"Invalid resampling type: {resample_type}")
    if resample_type == "custom" and weight_funcs is None:
        raise ValueError("weight_funcs must be supplied when using custom resampling")

    torch = _torch()
    L = _cabi.lib()
    valid_input_index_host = None
    is_masked_data = False
    if isinstance(data, np.ma.MaskedArray) and np.ma.is_masked(data):
        valid_input_index_host = source.valid.cpu().numpy().astype(bool)
        is_masked_data = bool(np.ma.is_masked(data[valid_input_index_host]))
    if is_masked_data:
        new_data = _stack_mask_channels(data)
        if weight_funcs is not None:
            weight_funcs = _weight_funcs_for_masked(weight_funcs, is_multi_channel)
    else:
        new_data = np.ma.getdata(data)
    channels = new_data.shape[1] if new_data.ndim > 1 else 1
    input_data_type = data.dtype if resample_type == 'nn' else None
    use_masked_fill_value = False
    if fill_value is None:
        use_masked_fill_value = True
        fill_value = _get_fill_mask_value(new_data.dtype)
    final_output_shape = output_shape + (channels,) if new_data.ndim > 1 else output_shape
    tg, n_t, keep = _make_target(source, tgt, reduce_data, radius_of_influence)
    stddev = count = None

    if resample_type == 'nn' or neighbours == 1:
        if neighbours != 1:
            raise ValueError('index_array contains more neighbours than just the nearest')
        host_rows = _byte_rows(new_data)
        row_bytes = host_rows.shape[1]
        fill_row_host = _byte_rows(np.full((1, channels), fill_value, dtype=new_data.dtype))
        if tg.kind == PRS_TARGET_AREA and n_t * row_bytes >= _PIPELINE_MIN_BYTES:
            result = _nearest_pipelined(L, source, tg, n_t, _HostUpload(host_rows), fill_row_host,
                                        radius_of_influence).view(new_data.dtype)
        else:
            dev_data = _dev(host_rows)
            fill_row = _dev(fill_row_host)
            out = torch.empty((n_t, row_bytes), dtype=torch.uint8, device="cuda")
            _cabi.check(L.prs_resample_nearest(source.index.handle, ctypes.byref(tg), float(radius_of_influence),
                                               _ptr(dev_data), row_bytes, _ptr(fill_row), _ptr(out), _stream()))
            result = _host(out).view(new_data.dtype)
        if with_uncert:
            raise TypeError("uncertainty estimates need more than one neighbour")
    else:
        funcs = list(weight_funcs) if new_data.ndim > 1 else [weight_funcs] * channels
        if new_data.ndim > 1 and len(funcs) < channels:
            raise IndexError("list index out of range")
        sigmas = [_gauss_sigma_of(f) for f in funcs[:channels]]
        ddt, rt = _weighted_dtypes(source.tree_dtype, new_data.dtype, new_data.ndim > 1)
        dev_data = _dev(new_data.astype(ddt, copy=False)).reshape(n_src, -1)
        if all(s is not None for s in sigmas):
            # fused query + Gaussian weighting
            acc64 = channels > 1 or source.tree_dtype == np.float64 or ddt == np.float64
            odt = torch.float64 if acc64 else torch.float32
            out = torch.empty((n_t, channels), dtype=odt, device="cuda")
            sd_t = torch.empty((n_t, channels), dtype=odt, device="cuda") if with_uncert else None
            cnt_t = torch.empty((n_t, channels), dtype=odt, device="cuda") if with_uncert else None
            flags = (ctypes.c_int64 * 2)()
            sig_c = (ctypes.c_double * channels)(*[float(s) for s in sigmas])
            _cabi.check(L.prs_resample_gauss(source.index.handle, ctypes.byref(tg), int(neighbours),
                                             float(radius_of_influence), _ptr(dev_data), _tree_code(ddt), channels,
                                             sig_c, float(fill_value), _ptr(out), _ptr(sd_t), _ptr(cnt_t), flags,
                                             _stream()))
            kth_found = bool(flags[0])
        else:
            # arbitrary Python weight functions: device query -> f(dist) on the device (host when the function
            # needs numpy) -> device weighted sum
            res = _custom_on_device(source, tgt, radius_of_influence, neighbours, reduce_data, dev_data, ddt,
                                    channels, weight_funcs, new_data.ndim > 1, fill_value, with_uncert)
            if res is None:
                return _get_empty_sample(data, output_shape, is_multi_channel,
                                         None if use_masked_fill_value else fill_value)
            out, sd_t, cnt_t, kth_found = res
        if kth_found:
            warnings.warn(('Possible more than %s neighbours within %s m for some data points') %
                          (neighbours, radius_of_influence), stacklevel=3)
        result = _host(out).astype(rt, copy=False)
        if with_uncert:
            stddev, count = sd_t.cpu().numpy(), cnt_t.cpu().numpy()
    del keep
    result = _prepare_result(result, final_output_shape, is_masked_data, use_masked_fill_value, fill_value,
                             input_data_type)
    if with_uncert:
        stddev, count = _prepare_uncertainty(stddev, count, final_output_shape, is_masked_data, result)
        return result, stddev, count
    return result


def _custom_on_device(source, tgt, radius_of_influence, neighbours, reduce_data, dev_data, ddt, channels,
                      weight_funcs, is_multi, fill_value, with_uncert, rows=None):
    """``resample_custom`` with arbitrary weight functions: ``prs_query`` -> weights -> ``prs_gather_weighted``, all
    on device tensors.  Returns (out, stddev, count, kth_found) over every target point, or None when no target
    point is valid."""
    torch = _torch()
    idx, dist, valid, kth_found, n_invalid = _query_device(source, tgt, radius_of_influence, neighbours,
                                                           reduce_data, rows=rows)
    n_t = idx.shape[0]
    voi_t = None
    if _weights_finite_at_one(weight_funcs):
        # pixels without any neighbour end as fill / NaN / 0 whatever the weights are (norm == 0,
        # kd_tree.py:807-809, 842-858) as long as f(1) - the value the reference feeds for missing
        # neighbours - is finite: only rows with a neighbour are weighted
        # the nearest neighbour's column: ranks below n_valid, the count itself where there is none (int32 holds both:
        # a rank is an index into float arrays of at most 2**31 - 1 points)
        has = idx[:, 0] < int(source.n_valid) if int(source.n_valid) < 2 ** 31 else \
            (idx[:, 0].to(torch.int64) & 0xFFFFFFFF) < int(source.n_valid)
        if n_invalid:
            has &= valid.view(torch.bool)
        n_has = int(has.sum())
        if 0 < n_has < n_t // 2:
            valid = has.view(torch.uint8)
            n_invalid = n_t - n_has
    n_kept = n_t - n_invalid
    if n_invalid:
        voi_t = valid
        idx = _compact_rows(valid, idx, kept=n_kept)
        dist = _compact_rows(valid, dist, kept=n_kept)
    n_out = idx.shape[0]
    if n_out == 0:
        return None
    out, sd_t, cnt_t = _weighted_from_info(idx, dist, n_out, int(neighbours), int(source.n_valid),
                                           _rank_to_source(source), dev_data, ddt, channels, weight_funcs,
                                           is_multi, fill_value, with_uncert)
    fill_row = torch.full((channels,), float(fill_value), dtype=out.dtype, device="cuda")
    out = _scatter_full(out, voi_t, n_t, fill_row)
    if with_uncert:
        nan_row = torch.full((channels,), float("nan"), dtype=sd_t.dtype, device="cuda")
        zero_row = torch.zeros((channels,), dtype=cnt_t.dtype, device="cuda")
        sd_t = _scatter_full(sd_t, voi_t, n_t, nan_row)
        cnt_t = _scatter_full(cnt_t, voi_t, n_t, zero_row)
    return out, sd_t, cnt_t, kth_found


def _rank_to_source(source):
    if source.index.all_valid:
        return None
    torch = _torch()
    out = torch.empty(int(source.n_valid), dtype=torch.int32, device="cuda")
    _cabi.check(_cabi.lib().prs_rank_to_source(_ptr(source.valid), source.valid.numel(), _ptr(out), None, _stream()))
    return out


def _resample_tensor(source, data, tgt, resample_type, radius_of_influence, neighbours, weight_funcs, fill_value,
                     reduce_data, with_uncert, rows=None):
    """Device-resident variant: CUDA tensor in -> CUDA tensor out (no host copies, no masked arrays)."""
    torch = _torch()
    L = _cabi.lib()
    n_src = int(source.lon_t.numel())
    data = data if data.is_cuda else data.cuda()
    if data.dim() > 2 and data.shape[0] * data.shape[1] == n_src:
        data = data.reshape(n_src, data.shape[2])
    elif data.shape[0] != n_src:
        data = data.reshape(-1)
    if data.shape[0] != n_src:
        raise ValueError('Mismatch between geometry and dataset')
    multi = data.dim() > 1
    channels = data.shape[1] if multi else 1
    row0, nrows, _, _ = _rows_tuple(rows)
    lead = tuple(tgt.shape) if rows is None else (nrows,) + tuple(tgt.shape[1:])
    output_shape = lead + ((channels,) if multi else ())
    if fill_value is None:
        raise ValueError("fill_value=None (masked output) needs numpy input")
    # no check for an empty index here: the kernels give every pixel the fill value then (kd_tree.py:336-341, 620-621),
    # and asking would make the host wait for the device
    tg, n_t, keep = _make_target(source, tgt, reduce_data, radius_of_influence, rows=rows)
    if n_t == 0:  # a rank without rows
        return torch.empty(output_shape, dtype=data.dtype, device="cuda")
    dev_data = data.contiguous().reshape(n_src, -1)
    if resample_type == 'nn':
        row_bytes = dev_data.shape[1] * dev_data.element_size()
        fill_row = torch.full((channels,), fill_value, dtype=data.dtype, device="cuda")
        out = torch.empty((n_t, channels), dtype=data.dtype, device="cuda")
        _cabi.check(L.prs_resample_nearest(source.handle, ctypes.byref(tg), float(radius_of_influence),
                                           _ptr(dev_data), row_bytes, _ptr(fill_row), _ptr(out), _stream()))
        return out.reshape(output_shape)
    funcs = list(weight_funcs) if multi else [weight_funcs] * channels
    sigmas = [_gauss_sigma_of(f) for f in funcs[:channels]]
    ddt = np.dtype(np.float32) if data.dtype == torch.float32 else np.dtype(np.float64)
    dev_data = dev_data.to(getattr(torch, ddt.name))
    if not all(s is not None for s in sigmas):
        res = _custom_on_device(source, tgt, radius_of_influence, neighbours, reduce_data, dev_data, ddt, channels,
                                weight_funcs, multi, fill_value, with_uncert, rows=rows)
        if res is None:
            return torch.full(output_shape, fill_value, dtype=data.dtype, device="cuda")
        out, sd_t, cnt_t, _ = res
        if with_uncert:
            return out.reshape(output_shape), sd_t.reshape(output_shape), cnt_t.reshape(output_shape)
        return out.reshape(output_shape)
    acc64 = channels > 1 or source.tree_dtype == np.float64 or ddt == np.float64
    odt = torch.float64 if acc64 else torch.float32
    out = torch.empty((n_t, channels), dtype=odt, device="cuda")
    sd_t = torch.empty((n_t, channels), dtype=odt, device="cuda") if with_uncert else None
    cnt_t = torch.empty((n_t, channels), dtype=odt, device="cuda") if with_uncert else None
    sig_c = (ctypes.c_double * channels)(*[float(s) for s in sigmas])
    _cabi.check(L.prs_resample_gauss(source.handle, ctypes.byref(tg), int(neighbours), float(radius_of_influence),
                                     _ptr(dev_data), _tree_code(ddt), channels, sig_c, float(fill_value), _ptr(out),
                                     _ptr(sd_t), _ptr(cnt_t), None, _stream()))
    del keep
    if with_uncert:
        return out.reshape(output_shape), sd_t.reshape(output_shape), cnt_t.reshape(output_shape)
    return out.reshape(output_shape)


def resample_nearest(source_geo_def,
                     data,
                     target_geo_def,
                     radius_of_influence,
                     epsilon=0,
                     fill_value=0,
                     reduce_data=True,
                     nprocs=1,
                     segments=None):
    """Resample data using the nearest neighbour approach (pyresample/kd_tree.py:64-110)."""
    return _resample(source_geo_def, data, target_geo_def, 'nn',
                     radius_of_influence, neighbours=1,
                     epsilon=epsilon, fill_value=fill_value,
                     reduce_data=reduce_data, nprocs=nprocs, segments=segments)


def _make_gauss(sigma):
    """Gaussian weight function ``w = exp(-r^2 / sigma^2)`` (pyresample/kd_tree.py:163-165), tagged for fusion."""
    def gauss(r):
        return np.exp(-r ** 2 / float(sigma) ** 2)
    gauss._prs_gauss_sigma = float(sigma)
    return gauss


def resample_gauss(source_geo_def, data, target_geo_def,
                   radius_of_influence, sigmas, neighbours=8, epsilon=0,
                   fill_value=0, reduce_data=True, nprocs=1, segments=None,
                   with_uncert=False):
    """Resample data using Gaussian-weighted neighbours (pyresample/kd_tree.py:113-189)."""
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
        weight_funcs = list(map(_make_gauss, sigma_list))
    else:
        weight_funcs = _make_gauss(sigmas)
    return _resample(source_geo_def, data, target_geo_def, 'custom',
                     radius_of_influence, neighbours=neighbours,
                     epsilon=epsilon, weight_funcs=weight_funcs, fill_value=fill_value,
                     reduce_data=reduce_data, nprocs=nprocs, segments=segments, with_uncert=with_uncert)


def resample_custom(source_geo_def, data, target_geo_def,
                    radius_of_influence, weight_funcs, neighbours=8,
                    epsilon=0, fill_value=0, reduce_data=True, nprocs=1,
                    segments=None, with_uncert=False):
    """Resample data using custom radial weight functions (pyresample/kd_tree.py:192-253)."""
    if not isinstance(weight_funcs, (list, tuple)):
        if not isinstance(weight_funcs, types.FunctionType):
            raise TypeError('weight_func must be function object')
    else:
        for weight_func in weight_funcs:
            if not isinstance(weight_func, types.FunctionType):
                raise TypeError('weight_func must be function object')
    return _resample(source_geo_def, data, target_geo_def, 'custom',
                     radius_of_influence, neighbours=neighbours,
                     epsilon=epsilon, weight_funcs=weight_funcs,
                     fill_value=fill_value, reduce_data=reduce_data,
                     nprocs=nprocs, segments=segments, with_uncert=with_uncert)


# the xarray API of this module (pyresample/kd_tree.py:902-1183) and the block kernels it imports at kd_tree.py:38
from .xarray_nn import XArrayResamplerNN, _my_index, query_no_distance  # noqa: E402,F401
