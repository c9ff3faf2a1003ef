"""Shared synthetic inputs for the parity tests (seeded, jittered so exact distance ties do not occur)."""
import numpy as np

from pyresample_b200 import geometry

STERE_AREA_ARGS = ('areaD', 'Europe (3km, HRV, VTC)', 'areaD',
                   {'a': '6378144.0', 'b': '6356759.0', 'lat_0': '50.00', 'lat_ts': '50.00', 'lon_0': '8.00',
                    'proj': 'stere'}, 800, 800,
                   [-1370912.72, -909968.64000000001, 1029087.28, 1490031.3600000001])


def stere_area():
    """The 800x800 oblique-stereographic area of pyresample/test/test_kd_tree.py:35-49."""
    return geometry.AreaDefinition(*STERE_AREA_ARGS)


def jittered_swath(rows, cols, dtype=np.float64, seed=0, lon0=0.0, lat0=40.0, dlon=30.0, dlat=25.0, rot_deg=10.0,
                   jitter=0.2):
    """Smooth curvilinear swath + uniform jitter of +-`jitter` pixel (SURVEY 8d synthetic-input proposal)."""
    rng = np.random.default_rng(seed)
    i, j = np.meshgrid(np.arange(rows, dtype=np.float64), np.arange(cols, dtype=np.float64), indexing="ij")
    i = i + rng.uniform(-jitter, jitter, i.shape)
    j = j + rng.uniform(-jitter, jitter, j.shape)
    u = j / max(cols - 1, 1) - 0.5
    v = i / max(rows - 1, 1) - 0.5
    a = np.deg2rad(rot_deg)
    lon = lon0 + dlon * (0.5 + u * np.cos(a) - v * np.sin(a))
    lat = lat0 + dlat * (0.5 + u * np.sin(a) + v * np.cos(a))
    return lon.astype(dtype), lat.astype(dtype)


def laea_area(width=500, height=500, half=1000000.0, lat_0=52, lon_0=15, dtype=np.float64):
    return geometry.AreaDefinition("c1", "C1 target", "laea",
                                   "+proj=laea +lat_0=%s +lon_0=%s +a=6371228 +units=m" % (lat_0, lon_0),
                                   width, height, (-half, -half, half, half), dtype=dtype)


def index_mismatch_report(idx_a, idx_b, dist_a, dist_b):
    """Rows where indices differ, split into exact-distance ties and real mismatches."""
    idx_a = np.asarray(idx_a).reshape(len(idx_a), -1)
    idx_b = np.asarray(idx_b).reshape(len(idx_b), -1)
    dist_a = np.asarray(dist_a).reshape(idx_a.shape)
    dist_b = np.asarray(dist_b).reshape(idx_b.shape)
    diff = idx_a != idx_b
    ties = diff & (dist_a == dist_b)
    return int(diff.sum()), int(ties.sum())
