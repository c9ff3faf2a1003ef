"""Registry of named resamplers (pyresample/future/resamplers/registry.py:33-151).

The reference fills its registry from the ``pyresample.resamplers`` entry-point group (setup.py:81-85, where
``nearest`` is the only built-in); here the GPU nearest-neighbour resampler is registered under the same name.
"""
from __future__ import annotations

from .future_nearest import KDTreeNearestXarrayResampler, Resampler

RESAMPLER_REGISTRY: dict = {}


def register_resampler(resampler_name, resampler_cls) -> None:
    """Register a :class:`Resampler` subclass under ``resampler_name`` (registry.py:33-64)."""
    if resampler_name in RESAMPLER_REGISTRY:
        raise ValueError(f"Resampler with name '{resampler_name} is already registered. "
                         "Use 'unregister_resampler' to make the name available.")
    RESAMPLER_REGISTRY[resampler_name] = resampler_cls


def unregister_resampler(resampler_name) -> None:
    """Remove a previously registered resampler (registry.py:67-69)."""
    del RESAMPLER_REGISTRY[resampler_name]


def list_resamplers() -> list:
    """Sorted names of the registered resamplers (registry.py:108-112)."""
    return sorted(RESAMPLER_REGISTRY.keys())


def create_resampler(src_geom, dst_geom, resampler=None, cache=None, **kwargs) -> Resampler:
    """Instance of a registered resampler class; ``nearest`` when no name is given (registry.py:115-151)."""
    if resampler is None:
        resampler = "nearest"
    resampler_cls = RESAMPLER_REGISTRY[resampler]
    return resampler_cls(src_geom, dst_geom, cache=cache, **kwargs)


register_resampler("nearest", KDTreeNearestXarrayResampler)
