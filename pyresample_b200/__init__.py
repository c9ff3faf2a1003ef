"""B200-native drop-in for the ``pyresample.kd_tree`` neighbour resampling path.

``pyresample_b200.kd_tree`` mirrors ``pyresample.kd_tree`` (pyresample/kd_tree.py) and
``pyresample_b200.geometry`` the geometry surface that path reads.  All computation
runs in hand-written sm_100a CUDA kernels behind the C-ABI of ``include/prs_b200.h``;
there is no CPU fallback - importing :mod:`pyresample_b200.kd_tree` works without a GPU
(for argument validation), any computation without the CUDA library/GPU raises.
"""
import os

CHUNK_SIZE = int(os.getenv('PYTROLL_CHUNK_SIZE', 4096))  # pyresample/__init__.py:27

from . import geometry  # noqa: E402,F401
from .geometry import (AreaDefinition, CoordinateDefinition, GridDefinition,  # noqa: E402,F401
                       SwathDefinition)

__all__ = ["geometry", "kd_tree", "AreaDefinition", "CoordinateDefinition", "GridDefinition", "SwathDefinition"]


def __getattr__(name):
    if name == "kd_tree":
        import importlib
        return importlib.import_module(".kd_tree", __name__)
    raise AttributeError(name)
