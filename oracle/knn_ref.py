"""Oracle: exact bounded k-NN with the pykdtree call contract.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

pykdtree (``pykdtree>=1.3.1``, setup.py:28, unpinned, NOT in the reference tree) is
what the reference calls at pyresample/kd_tree.py:487 (``KDTree(input_coords)``),
:545-548 (``.query(coords, k=, eps=, distance_upper_bound=)``), :538 (``.data``)
and future/resamplers/nearest.py:62-67,73 (``mask=``, ``.n``).  Its published
algorithm (kdtree.pyx + _kdtree_core.c) is restated here as a *specification*:

* squared Euclidean distance accumulated in the tree dtype, in x, y, z order,
  without fused multiply-add:  d2 = ((dx*dx) + dy*dy) + dz*dz ;
* the bound is ``dub = dtype(distance_upper_bound ** 2)`` and a candidate is
  accepted iff ``d2 < dub`` (strict);
* results sorted ascending; missing -> index ``n`` and distance ``inf``;
  distance = sqrt(d2) in the tree dtype; index dtype uint32; 1-D outputs for k=1;
* ``mask[i] == True`` removes tree point ``i`` from consideration (numbering unchanged).

Ties (equal d2) are broken to the LOWEST index - the deterministic rule named by
BASELINE.json's north_star.  pykdtree itself breaks ties by traversal order; no
reference test pins that, so exact ties are "parity unpinned" (DESIGN.md).

Candidates come from scipy.spatial.cKDTree (the reference's own alternative
backend, _spatial_mp.py:215-216) and are re-ranked in the tree dtype; the
selection is proven complete per row or the row is redone with more candidates.
``brute_query`` (C, oracle/knn_brute.c) is the independent check of this file.
"""
import ctypes
import os

import numpy as np
from scipy.spatial import cKDTree


def _d2(q, p, dtype):
    """Squared distance exactly as pykdtree accumulates it (no FMA: numpy ops are separate)."""
    dx = (q[..., 0] - p[..., 0]).astype(dtype, copy=False)
    dy = (q[..., 1] - p[..., 1]).astype(dtype, copy=False)
    dz = (q[..., 2] - p[..., 2]).astype(dtype, copy=False)
    return (dx * dx + dy * dy) + dz * dz


class KDTreeRef:
    """Drop-in for ``pykdtree.kdtree.KDTree`` as far as the reference path uses it."""

    def __init__(self, data_pts, leafsize=16):
        data_pts = np.asarray(data_pts)
        if data_pts.dtype not in (np.float32, np.float64):
            data_pts = data_pts.astype(np.float64)
        if data_pts.ndim == 1:
            data_pts = data_pts[:, None]
        self.data = np.ascontiguousarray(data_pts)
        self.n = self.data.shape[0]
        self.ndim = self.data.shape[1]
        self.leafsize = leafsize
        self._tree = None
        self._tree_mask_key = None
        self._tree_map = None

    def _get_tree(self, mask):
        key = None if mask is None else mask.tobytes()
        if self._tree is None or key != self._tree_mask_key:
            if mask is None:
                self._tree_map = None
                pts = self.data.astype(np.float64)
            else:
                self._tree_map = np.flatnonzero(~mask)
                pts = self.data[self._tree_map].astype(np.float64)
            self._tree = cKDTree(pts, leafsize=self.leafsize) if pts.shape[0] else None
            self._tree_mask_key = key
        return self._tree, self._tree_map

    def query(self, query_pts, k=1, eps=0, distance_upper_bound=None, sqr_dists=False, mask=None, workers=-1):
        dtype = self.data.dtype
        q = np.asarray(query_pts)
        if q.ndim == 1:
            q = q[:, None]
        if q.dtype != dtype:
            if dtype == np.float32:
                raise TypeError("Type mismatch. query points must be of type float32 when data points are of type float32")
            q = q.astype(dtype)
        m = q.shape[0]
        if mask is not None:
            mask = np.asarray(mask, dtype=bool).ravel()
            if mask.size != self.n:
                raise ValueError("Mask must have the same size as data points")
        if distance_upper_bound is None:
            dub = dtype.type(np.finfo(dtype).max)
            r_search = np.inf
        else:
            dub = dtype.type(float(distance_upper_bound) * float(distance_upper_bound))
            r_search = float(distance_upper_bound) * (1 + 1e-5) + 1e-3
        idx_out = np.full((m, k), self.n, dtype=np.uint32)
        d2_out = np.full((m, k), np.inf, dtype=dtype)
        tree, tmap = self._get_tree(mask)
        if tree is not None and m > 0:
            rows = np.arange(m)
            extra = 4
            n_tree = tree.n
            while rows.size:
                kk = min(k + extra, n_tree)
                qq = q[rows].astype(np.float64)
                _, ci = tree.query(qq, k=kk, distance_upper_bound=r_search, workers=workers)
                if kk == 1:
                    ci = ci[:, None]
                missing = ci >= n_tree
                ci_safe = np.where(missing, 0, ci)
                if tmap is not None:
                    ci_glob = tmap[ci_safe]
                else:
                    ci_glob = ci_safe
                cand = self.data[ci_glob]  # (rows, kk, 3)
                d2 = _d2(q[rows][:, None, :], cand, dtype).astype(dtype)
                bad = missing | ~(d2 < dub)
                d2s = np.where(bad, np.inf, d2)
                idxs = np.where(bad, self.n, ci_glob).astype(np.int64)
                order = np.lexsort((idxs, d2s), axis=1)
                d2s = np.take_along_axis(d2s, order, axis=1)
                idxs = np.take_along_axis(idxs, order, axis=1)
                # completeness: a candidate beyond the kk-th (in true distance) can only matter if the
                # kk-th true-nearest exists and its tree-dtype d2 is not clearly above the selected k-th.
                last_present = ~missing[:, -1]
                d2_last = d2[:, -1]  # cKDTree returns ascending true distance: last column = farthest fetched
                kth = d2s[:, min(k, kk) - 1]
                tol = 1e-6 if dtype == np.float32 else 1e-12
                ambiguous = last_present & (kk < n_tree) & ~(d2_last > kth * (1 + tol)) & np.isfinite(kth) | \
                    (last_present & (kk < n_tree) & ~np.isfinite(kth) & (d2_last < dub * (1 + tol)))
                good = ~ambiguous
                take = min(k, kk)
                idx_out[rows[good], :take] = idxs[good, :take]
                d2_out[rows[good], :take] = d2s[good, :take]
                rows = rows[ambiguous]
                extra = extra * 4 + 4
        if sqr_dists:
            dist = d2_out
        else:
            dist = np.sqrt(d2_out)
        if k == 1:
            return dist[:, 0], idx_out[:, 0]
        return dist, idx_out


# ---------------------------------------------------------------- C brute-force checker
_HERE = os.path.dirname(os.path.abspath(__file__))
_BRUTE_SO = os.path.join(_HERE, "_build", "libknn_brute.so")
_brute = None


def build_brute(force=False):
    """Compile oracle/knn_brute.c -> oracle/_build/libknn_brute.so (gcc, no FMA contraction)."""
    import subprocess
    src = os.path.join(_HERE, "knn_brute.c")
    os.makedirs(os.path.dirname(_BRUTE_SO), exist_ok=True)
    if force or not os.path.exists(_BRUTE_SO) or os.path.getmtime(_BRUTE_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-shared", "-fPIC",
                               "-o", _BRUTE_SO, src, "-lm"])
    return _BRUTE_SO


def brute_query(data, query, k, radius, mask=None):
    """Exhaustive O(N*M) k-NN in C with the same arithmetic and (d2, index) ordering."""
    global _brute
    if _brute is None:
        _brute = ctypes.CDLL(build_brute())
    data = np.ascontiguousarray(data)
    dtype = data.dtype
    query = np.ascontiguousarray(query, dtype=dtype)
    n, m = data.shape[0], query.shape[0]
    idx = np.empty((m, k), dtype=np.uint32)
    d2 = np.empty((m, k), dtype=dtype)
    mk = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
    fn = _brute.knn_brute_f32 if dtype == np.float32 else _brute.knn_brute_f64
    fn.restype = None
    fn(data.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(n), query.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(m),
       ctypes.c_int(k), ctypes.c_double(float(radius) * float(radius)),
       None if mk is None else mk.ctypes.data_as(ctypes.c_void_p),
       idx.ctypes.data_as(ctypes.c_void_p), d2.ctypes.data_as(ctypes.c_void_p))
    dist = np.sqrt(d2)
    if k == 1:
        return dist[:, 0], idx[:, 0]
    return dist, idx
