"""ctypes binding of ``libprs_b200.so`` (C-ABI declared in ``include/prs_b200.h``).

The library is built in-tree by :func:`build` (``nvcc -gencode arch=compute_100a,code=sm_100a``)
and loaded lazily.  There is NO fallback: if the library is missing or no CUDA device is present,
:func:`lib` raises and every compute entry point of the package fails loudly.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import threading

from .proj import prs_proj_params

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC_DIR = os.path.join(_HERE, "csrc")
_SO_PATH = os.environ.get("PRS_B200_LIBRARY") or os.path.join(_HERE, "libprs_b200.so")  # override: kernel experiments
_SOURCES = ["prs_b200.cu", "prs_query_inst.cu", "prs_classify_inst.cu"]
_HEADERS = ["prs_common.cuh", "prs_scan.cuh", "prs_proj.cuh", "prs_index.cuh", "prs_query.cuh", "prs_launch.cuh",
            "prs_bilinear.cuh"]
_INCLUDE = os.path.join(os.path.dirname(_HERE), "include", "prs_b200.h")

PRS_F32, PRS_F64 = 0, 1
PRS_TARGET_LONLAT, PRS_TARGET_AREA = 0, 1
PRS_BUILD_RANKS = 1
PRS_BUILD_COMPUTE_VALID = 2
PRS_COVER_CELLS = 512
PRS_TILE_ROWS = 8
PRS_REDUCE_ALL, PRS_REDUCE_LAT_GE, PRS_REDUCE_LAT_LE, PRS_REDUCE_BOX, PRS_REDUCE_BOX_DATELINE = 0, 1, 2, 3, 4

EXPORTED_SYMBOLS = (
    "prs_abi_version", "prs_last_error", "prs_launch_count", "prs_release_cached_memory", "prs_lonlat_to_xyz", "prs_area_lonlats", "prs_valid_mask",
    "prs_index_build", "prs_target_coverage", "prs_index_destroy", "prs_index_info", "prs_query", "prs_resample_nearest",
    "prs_resample_scratch_bytes", "prs_resample_nearest_fill", "prs_resample_nearest_search",
    "prs_resample_gauss", "prs_gather_nearest", "prs_gather_weighted", "prs_compact_rows",
    "prs_rank_to_source", "prs_scatter_rows", "prs_project_forward", "prs_bilinear_info", "prs_bilinear_sample",
    "prs_weight_distances", "prs_area_lonlats_covered",
)


class prs_area_grid(ctypes.Structure):
    _fields_ = [("width", ctypes.c_int64), ("height", ctypes.c_int64),
                ("pixel_size_x", ctypes.c_double), ("pixel_size_y", ctypes.c_double),
                ("pixel_upper_left_x", ctypes.c_double), ("pixel_upper_left_y", ctypes.c_double)]


class prs_reduce_box(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("_pad", ctypes.c_int32),
                ("lat_min", ctypes.c_double), ("lat_max", ctypes.c_double),
                ("lon_min", ctypes.c_double), ("lon_max", ctypes.c_double)]


class prs_target(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("dtype", ctypes.c_int32),
                ("lon", ctypes.c_void_p), ("lat", ctypes.c_void_p), ("n", ctypes.c_int64),
                ("proj", prs_proj_params), ("grid", prs_area_grid),
                ("row0", ctypes.c_int64), ("nrows", ctypes.c_int64),
                ("valid_extra", ctypes.c_void_p),
                ("row_block", ctypes.c_int64), ("row_stride", ctypes.c_int64)]


class PrsError(RuntimeError):
    """An entry point of libprs_b200 returned a non-zero status."""


def needs_rebuild():
    if not os.path.exists(_SO_PATH):
        return True
    so_time = os.path.getmtime(_SO_PATH)
    deps = [os.path.join(_SRC_DIR, f) for f in _SOURCES + _HEADERS] + [_INCLUDE]
    return any(os.path.getmtime(d) > so_time for d in deps if os.path.exists(d))


# search kernels: (tree dtype, top-k width K, epilogue) with epilogue 0 = neighbour info, 1 = nearest gather (K = 1),
# 2 / 3 / 4 = Gaussian weighting with float32 / float64-over-float32-data / float64 accumulation
_QUERY_INSTANCES = [(t, k, e) for t in ("float", "double") for k, e in
                    [(1, 0), (4, 0), (8, 0), (16, 0), (32, 0), (1, 1)] + [(k, e) for e in (2, 3, 4) for k in (4, 8, 16, 32)]]


def build(force=False, verbose=False, jobs=None):
    """Compile the CUDA library for sm_100a into ``pyresample_b200/libprs_b200.so``.

    ``prs_b200.cu`` holds the C-ABI and the small kernels; the classification kernels are instantiated per tree
    dtype in ``prs_classify_inst.cu`` and the search kernels per (tree dtype, K, epilogue) in
    ``prs_query_inst.cu``; the translation units are compiled in parallel and linked.
    """
    if not force and not needs_rebuild():
        return _SO_PATH
    from concurrent.futures import ThreadPoolExecutor
    nvcc = os.environ.get("NVCC", "nvcc")
    if not any(os.path.exists(os.path.join(p, nvcc)) for p in os.environ.get("PATH", "").split(os.pathsep)) \
            and os.path.exists("/usr/local/cuda/bin/nvcc"):
        nvcc = "/usr/local/cuda/bin/nvcc"
    obj_dir = os.path.join(_HERE, "build")
    os.makedirs(obj_dir, exist_ok=True)
    flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
             "-Xcompiler", "-fPIC"] + os.environ.get("PRS_NVCC_FLAGS", "").split()
    if verbose:
        flags.append("-Xptxas=-v")
    units = [(os.path.join(_SRC_DIR, "prs_b200.cu"), os.path.join(obj_dir, "prs_b200.o"), [])]
    for t in ("float", "double"):
        units.append((os.path.join(_SRC_DIR, "prs_classify_inst.cu"), os.path.join(obj_dir, "prs_classify_%s.o" % t),
                      ["-DPRS_INST_T=%s" % t]))
    for t, k, e in _QUERY_INSTANCES:
        units.append((os.path.join(_SRC_DIR, "prs_query_inst.cu"),
                      os.path.join(obj_dir, "prs_query_%s_%d_%d.o" % (t, k, e)),
                      ["-DPRS_INST_T=%s" % t, "-DPRS_INST_K=%d" % k, "-DPRS_INST_EPI=%d" % e]))
    units.sort(key=lambda u: 0 if "_32_" in u[1] else 1 if "_16_" in u[1] else 2)  # longest first

    # kernel experiments: PRS_BUILD_ONLY="float_1,prs_b200" recompiles just those units and relinks with the
    # existing objects of the others (only sound while the structs shared between units are unchanged)
    only = [s for s in os.environ.get("PRS_BUILD_ONLY", "").split(",") if s]

    def compile_unit(unit):
        src, obj, defs = unit
        if only and os.path.exists(obj) and not any(os.path.basename(obj) in (o + ".o", "prs_query_" + o + ".o")
                                                    for o in only):
            return obj
        subprocess.check_call([nvcc] + flags + defs + ["-c", src, "-o", obj])
        return obj

    with ThreadPoolExecutor(max_workers=jobs or min(len(units), os.cpu_count() or 4)) as pool:
        objs = list(pool.map(compile_unit, units))
    subprocess.check_call([nvcc, "-shared", "-o", _SO_PATH + ".tmp"] + objs)
    os.replace(_SO_PATH + ".tmp", _SO_PATH)
    return _SO_PATH


_lib = None
_lock = threading.Lock()


def load_library():
    """dlopen the library and declare prototypes.  Does not need a GPU (used by the CPU-side symbol test)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(_SO_PATH):
            raise PrsError("libprs_b200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                           "pyresample_b200 has no CPU fallback")
        L = ctypes.CDLL(_SO_PATH)
        vp, i64, i32, dbl = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_double
        pi64 = ctypes.POINTER(ctypes.c_int64)
        L.prs_abi_version.restype = i32
        L.prs_abi_version.argtypes = []
        L.prs_last_error.restype = ctypes.c_char_p
        L.prs_last_error.argtypes = []
        L.prs_launch_count.restype = ctypes.c_int64
        L.prs_launch_count.argtypes = []
        L.prs_release_cached_memory.argtypes = []
        L.prs_lonlat_to_xyz.argtypes = [vp, vp, i64, i32, vp, vp]
        L.prs_area_lonlats.argtypes = [ctypes.POINTER(prs_proj_params), ctypes.POINTER(prs_area_grid), i32,
                                       i64, i64, i64, i64, i64, i64, vp, vp, vp, vp]
        L.prs_valid_mask.argtypes = [vp, vp, i64, i32, ctypes.POINTER(prs_reduce_box), vp, vp, pi64, vp]
        L.prs_index_build.argtypes = [vp, vp, i64, i32, vp, ctypes.POINTER(prs_reduce_box), vp, vp, i32, i32, dbl, vp,
                                      i32, ctypes.POINTER(vp), vp]
        L.prs_target_coverage.argtypes = [ctypes.POINTER(prs_target), dbl, vp, vp]
        L.prs_index_destroy.argtypes = [vp]
        L.prs_index_info.argtypes = [vp, pi64]
        L.prs_query.argtypes = [vp, ctypes.POINTER(prs_target), i32, dbl, vp, vp, vp, pi64, vp]
        L.prs_resample_nearest.argtypes = [vp, ctypes.POINTER(prs_target), dbl, vp, i64, vp, vp, vp]
        L.prs_resample_scratch_bytes.restype = i64
        L.prs_resample_scratch_bytes.argtypes = [ctypes.POINTER(prs_target)]
        L.prs_resample_nearest_fill.argtypes = [vp, ctypes.POINTER(prs_target), dbl, i64, vp, vp, vp, vp, vp]
        L.prs_resample_nearest_search.argtypes = [vp, ctypes.POINTER(prs_target), dbl, vp, i64, vp, vp, vp, vp]
        L.prs_resample_gauss.argtypes = [vp, ctypes.POINTER(prs_target), i32, dbl, vp, i32, i32,
                                         ctypes.POINTER(ctypes.c_double), dbl, vp, vp, vp, pi64, vp]
        L.prs_gather_nearest.argtypes = [vp, i64, ctypes.c_uint32, vp, vp, i64, vp, vp, vp]
        L.prs_gather_weighted.argtypes = [vp, vp, i64, i32, ctypes.c_uint32, vp, vp, i32, i32, vp, i32, i32,
                                          ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_double), dbl,
                                          vp, vp, vp, vp]
        L.prs_compact_rows.argtypes = [vp, i64, i64, vp, vp, pi64, vp]
        L.prs_rank_to_source.argtypes = [vp, i64, vp, pi64, vp]
        L.prs_area_lonlats_covered.argtypes = [ctypes.POINTER(prs_proj_params), ctypes.POINTER(prs_area_grid), ctypes.c_int, vp, vp,
                                               vp, vp]
        L.prs_weight_distances.argtypes = [vp, vp, ctypes.c_int, i64, ctypes.c_uint32, vp, vp]
        L.prs_scatter_rows.argtypes = [vp, i64, i64, vp, vp, vp, vp]
        L.prs_project_forward.argtypes = [vp, vp, i64, i32, ctypes.POINTER(prs_proj_params), vp, vp]
        L.prs_bilinear_info.argtypes = [vp, i64, i32, ctypes.c_uint32, vp, vp, ctypes.POINTER(prs_area_grid), vp, vp, vp, vp]
        L.prs_bilinear_sample.argtypes = [vp, vp, vp, i64, vp, vp, i32, i32, dbl, vp, vp]
        for name in EXPORTED_SYMBOLS:
            if name not in ("prs_abi_version", "prs_last_error", "prs_launch_count", "prs_resample_scratch_bytes"):
                getattr(L, name).restype = i32
        _lib = L
        return L


def lib():
    """The loaded library, for compute calls: additionally requires a CUDA device."""
    import torch
    if not torch.cuda.is_available():
        raise PrsError("pyresample_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return load_library()


def check(status):
    if status != 0:
        msg = load_library().prs_last_error()
        raise PrsError("libprs_b200 error %d: %s" % (status, msg.decode() if msg else "?"))
