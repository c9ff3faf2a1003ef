"""Synthetic workloads of BASELINE.json `configs` (shapes named there; geometry per SURVEY 8d).

Shared by bench.py and the tests.  Everything is seeded; swath positions carry a +-0.2 pixel jitter so that
exact distance ties (which no reference test pins) do not occur.
"""
import numpy as np

from pyresample_b200 import geometry

R_EARTH = 6371000.0


def polar_orbiter_swath(n_lines, n_pixels, scan_half_angle_deg, sat_height_m, line_spacing_m, seed,
                        inclination_deg=98.7, start_arc_deg=-20.0, node_lon_deg=10.0, dtype=np.float32, jitter=0.2):
    """Lon/lat (n_lines, n_pixels) of a cross-track scanner on an inclined circular orbit (no Earth rotation)."""
    rng = np.random.default_rng(seed)
    inc = np.deg2rad(inclination_deg)
    node = np.deg2rad(node_lon_deg)
    # orthonormal basis of the orbital plane: a = ascending node direction, b = 90 deg further along the orbit
    a = np.array([np.cos(node), np.sin(node), 0.0])
    b = np.array([-np.sin(node) * np.cos(inc), np.cos(node) * np.cos(inc), np.sin(inc)])
    w = np.cross(a, b)
    li = np.arange(n_lines, dtype=np.float64)[:, None] + rng.uniform(-jitter, jitter, (n_lines, n_pixels))
    pj = np.arange(n_pixels, dtype=np.float64)[None, :] + rng.uniform(-jitter, jitter, (n_lines, n_pixels))
    t = np.deg2rad(start_arc_deg) + li * (line_spacing_m / R_EARTH)
    alpha = np.deg2rad(scan_half_angle_deg) * (2.0 * pj / (n_pixels - 1) - 1.0)
    gamma = np.arcsin(np.clip((R_EARTH + sat_height_m) / R_EARTH * np.sin(alpha), -1, 1)) - alpha
    ct, st = np.cos(t), np.sin(t)
    cg, sg = np.cos(gamma), np.sin(gamma)
    x = cg * (ct * a[0] + st * b[0]) + sg * w[0]
    y = cg * (ct * a[1] + st * b[1]) + sg * w[1]
    z = cg * (ct * a[2] + st * b[2]) + sg * w[2]
    lon = np.rad2deg(np.arctan2(y, x)).astype(dtype)
    lat = np.rad2deg(np.arcsin(np.clip(z, -1, 1))).astype(dtype)
    return lon, lat


def brightness_data(n, channels, seed, nan_frac=0.01):
    rng = np.random.default_rng(seed)
    data = (rng.standard_normal((n, channels), dtype=np.float32) * 10 + 280).astype(np.float32)
    if nan_frac:
        data[rng.random((n, channels)) < nan_frac] = np.nan
    return data


def make_config(cid, scale=1.0, dtype=np.float32):
    """Return dict(source lon/lat, data, target AreaDefinition, radius, k, mode, channels, name).

    ``scale`` < 1 shrinks every dimension proportionally (tests); 1.0 is the BASELINE.json shape.
    """
    seed = 20261019 + cid
    s = float(scale)

    def sz(v):
        return max(8, int(round(v * s)))

    if cid == 1:
        rows, cols = sz(1000), sz(1000)
        rng = np.random.default_rng(seed)
        i, j = np.meshgrid(np.arange(rows, dtype=np.float64), np.arange(cols, dtype=np.float64), indexing="ij")
        i += rng.uniform(-0.2, 0.2, i.shape)
        j += rng.uniform(-0.2, 0.2, j.shape)
        u, v = j / (cols - 1) - 0.5, i / (rows - 1) - 0.5
        ang = np.deg2rad(10.0)
        lon = (0.0 + 30.0 * (0.5 + u * np.cos(ang) - v * np.sin(ang))).astype(dtype)
        lat = (40.0 + 25.0 * (0.5 + u * np.sin(ang) + v * np.cos(ang))).astype(dtype)
        area = geometry.AreaDefinition("c1", "laea 4 km", "laea", "+proj=laea +lat_0=52 +lon_0=15 +a=6371228 +units=m",
                                       sz(500), sz(500), (-1000000.0, -1000000.0, 1000000.0, 1000000.0))
        return dict(name="C1 1000x1000 f32 swath -> 500x500 laea, nearest r=50km", lon=lon, lat=lat,
                    data=brightness_data(rows * cols, 1, seed + 100)[:, 0], area=area, radius=50000, k=1, mode="nearest",
                    channels=1)
    if cid == 2:
        lines, pix = sz(5400), sz(2048)
        lon, lat = polar_orbiter_swath(lines, pix, 55.4, 833000.0, 1100.0 / s, seed, dtype=dtype)
        area = geometry.AreaDefinition("pc_world", "global 0.05 deg eqc", "eqc", "+proj=eqc +datum=WGS84",
                                       sz(7200), sz(3600),
                                       (-20037508.342789244, -10018754.171394622, 20037508.342789244, 10018754.171394622))
        return dict(name="C2 AVHRR-like 2048x5400 swath, 5 ch -> global 0.05deg eqc 7200x3600, nearest r=5km",
                    lon=lon, lat=lat, data=brightness_data(lines * pix, 5, seed + 100), area=area, radius=5000.0 / s,
                    k=1, mode="nearest", channels=5)
    if cid == 3:
        lines, pix = sz(6400), sz(3200)
        lon, lat = polar_orbiter_swath(lines, pix, 56.28, 829000.0, 750.0 / s, seed, start_arc_deg=55.0,
                                       node_lon_deg=-60.0, dtype=dtype)
        area = geometry.AreaDefinition("nsidc_stere", "polar stere 2 km", "stere",
                                       "+proj=stere +lat_0=90 +lat_ts=70 +lon_0=-45 +a=6378273 +b=6356889.449",
                                       sz(4000), sz(4000), (-4000000.0, -4000000.0, 4000000.0, 4000000.0))
        return dict(name="C3 VIIRS-like 3200x6400 swath, 16 ch -> polar stere 4000x4000, gauss k=8 sigma=25km",
                    lon=lon, lat=lat, data=brightness_data(lines * pix, 16, seed + 100), area=area, radius=50000.0,
                    k=8, mode="gauss", sigma=25000.0, channels=16)
    if cid == 4:
        lons, lats = [], []
        for g in range(4):
            lo, la = polar_orbiter_swath(sz(6400), sz(3200), 56.28, 829000.0, 750.0 / s, seed + g,
                                         start_arc_deg=40.0 + 10.0 * g, node_lon_deg=-120.0 + 60.0 * g, dtype=dtype)
            lons.append(lo)
            lats.append(la)
        lon, lat = np.concatenate(lons), np.concatenate(lats)
        area = geometry.AreaDefinition("ease_nh", "EASE north", "laea",
                                       "+proj=laea +lat_0=90 +lon_0=0 +a=6371228.0 +units=m",
                                       sz(8000), sz(8000), (-5326849.0625, -5326849.0625, 5326849.0625, 5326849.0625))
        return dict(name="C4 4x VIIRS granules -> 8000x8000 EASE, custom weights k=16", lon=lon, lat=lat,
                    data=brightness_data(lon.size, 16, seed + 100), area=area, radius=50000.0, k=16, mode="custom",
                    channels=16)
    if cid == 5:
        src = geometry.AreaDefinition("abi_fd", "ABI full disk like", "geos",
                                      "+proj=geos +h=35786023 +lon_0=-75 +sweep=x +ellps=GRS80", sz(5424), sz(5424),
                                      (-5434894.885056, -5434894.885056, 5434894.885056, 5434894.885056), dtype=dtype)
        area = geometry.AreaDefinition("latlon", "lat/lon grid", "longlat", "+proj=longlat +datum=WGS84",
                                       sz(10000), sz(10000), (-156.3, -81.3, 6.3, 81.3))
        return dict(name="C5 ABI full-disk 5424x5424 geos -> 10000x10000 lat/lon, nearest r=10km", source_area=src,
                    data=brightness_data(src.size, 1, seed + 100)[:, 0], area=area, radius=10000.0 / s, k=1,
                    mode="nearest", channels=1)
    raise ValueError("unknown config %r" % (cid,))
