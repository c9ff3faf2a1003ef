"""Oracle: numpy restatement of the PROJ inverse projections the path needs.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference obtains target lon/lats with
``pyproj.Transformer.from_crs(geodetic_crs_no_pm_shift, crs, always_xy=True)
.transform(x, y, direction=INVERSE)`` (pyresample/geometry.py:2630-2636, dask
variant ``_invproj`` geometry.py:1347-1354).  pyproj/PROJ is a third-party
dependency (``pyproj>=3.0``, setup.py:27, unpinned, not vendored), so the
formulas below restate PROJ's published algorithms
(src/projections/{latlong,eqc,laea,stere,geos,merc}.cpp and pj_inv):

    x' = (x * to_meter - x_0) / a ;  y' likewise
    (lam, phi) = inverse(x', y')
    lam = adjlon(lam + lon_0)  ;  degrees out ;  failure -> (inf, inf)

Pinned by the reference's tests test_geometry/test_area.py:293-299 (oblique
ellipsoidal stere), :748-761 (polar spherical laea), :589-600 (latlong) and by
every cross-sum of test_kd_tree.py through oracle/kd_tree_ref.py.
"""
import math

import numpy as np

_ELLPS = {  # name: (a, rf or None, b or None)
    "WGS84": (6378137.0, 298.257223563, None),
    "GRS80": (6378137.0, 298.257222101, None),
    "WGS72": (6378135.0, 298.26, None),
    "bessel": (6377397.155, 299.1528128, None),
    "intl": (6378388.0, 297.0, None),
    "clrk66": (6378206.4, None, 6356583.8),
    "krass": (6378245.0, 298.3, None),
    "airy": (6377563.396, None, 6356256.910),
    "sphere": (6370997.0, None, 6370997.0),
}
_DATUM = {"WGS84": "WGS84", "NAD83": "GRS80", "NAD27": "clrk66"}
_UNITS = {"m": 1.0, "km": 1000.0, "dm": 0.1, "cm": 0.01, "mm": 0.001,
          "ft": 0.3048, "us-ft": 0.304800609601219, "kmi": 1852.0}

EPS10 = 1e-10
HALFPI = math.pi / 2
RAD_TO_DEG = 57.295779513082321
DEG_TO_RAD = .017453292519943296


def parse_proj(projection):
    """PROJ.4 dict or '+k=v' string -> plain dict of str values (flags -> True)."""
    if isinstance(projection, dict):
        out = {}
        for k, v in projection.items():
            out[str(k).lstrip("+")] = v
        return out
    out = {}
    s = str(projection).strip()
    if s.upper().startswith("EPSG:"):
        code = int(s.split(":")[1])
        if code == 4326:
            return {"proj": "longlat", "datum": "WGS84"}
        raise ValueError("oracle supports only EPSG:4326 among EPSG codes")
    for tok in s.split():
        tok = tok.lstrip("+")
        if not tok:
            continue
        if "=" in tok:
            k, v = tok.split("=", 1)
            out[k] = v
        else:
            out[tok] = True
    return out


def _f(d, key, default=None):
    v = d.get(key, default)
    if v is None:
        return None
    return float(v)


def ellipsoid(d):
    """(a, es) following PROJ's pj_ellipsoid precedence: R > (a with b|rf|f|es|e) > ellps > datum > WGS84/GRS80 default."""
    if d.get("R") is not None:
        return float(d["R"]), 0.0
    a = _f(d, "a")
    name = d.get("ellps")
    if name is None and d.get("datum") is not None:
        name = _DATUM[str(d["datum"])]
    b = rf = None
    if a is None:
        if name is None:
            name = "WGS84" if d.get("proj") in ("longlat", "latlong", "lonlat", "latlon") else "GRS80"
        a, rf, b = _ELLPS[str(name)]
    if d.get("b") is not None:
        b = _f(d, "b")
        rf = None
    elif d.get("rf") is not None:
        rf = _f(d, "rf")
        b = None
    elif d.get("f") is not None:
        f = _f(d, "f")
        rf = (1.0 / f) if f != 0 else None
        b = a if f == 0 else None
    elif d.get("es") is not None:
        return a, _f(d, "es")
    elif d.get("e") is not None:
        return a, _f(d, "e") ** 2
    elif _f(d, "a") is not None and name is None:
        return a, 0.0  # only +a given: sphere
    elif _f(d, "a") is not None and name is not None:
        _, rf, b = _ELLPS[str(name)]
    if b is not None:
        if b == a:
            return a, 0.0
        return a, 1.0 - (b * b) / (a * a)
    if rf is not None:
        f = 1.0 / rf
        return a, 2 * f - f * f
    return a, 0.0


def _adjlon(lam):
    """PROJ adjlon: wrap into [-pi, pi] only when outside (tolerance 1e-12)."""
    lam = np.asarray(lam, dtype=np.float64)
    out = lam.copy()
    big = np.abs(lam) >= math.pi + 1e-12
    if big.any():
        t = lam[big] + math.pi
        t = t - 2 * math.pi * np.floor(t / (2 * math.pi))
        out[big] = t - math.pi
    return out


def _qsfn(sinphi, e, one_es):
    if e >= 1e-7:
        con = e * sinphi
        return one_es * (sinphi / (1 - con * con) - (.5 / e) * math.log((1 - con) / (1 + con)))
    return sinphi + sinphi


def _authset(es):
    t = es * es
    apa0 = es * (1 / 3.) + t * (31 / 180.)
    apa1 = t * (23 / 360.)
    t *= es
    apa0 += t * (517 / 5040.)
    apa1 += t * (251 / 3780.)
    apa2 = t * (761 / 45360.)
    return apa0, apa1, apa2


def _tsfn(phi, sinphi, e):
    sinphi = sinphi * e
    return math.tan(.5 * (HALFPI - phi)) / math.pow((1 - sinphi) / (1 + sinphi), .5 * e)


# --------------------------------------------------------------------------- inverses
def _inv_longlat(d, a, es, x, y):
    return x * 1.0, y * 1.0, None


def _inv_eqc(d, a, es, x, y):
    # eqc.cpp: spherical only, rc = cos(lat_ts); lam = x/rc; phi = y + phi0
    rc = math.cos(_f(d, "lat_ts", 0.0) * DEG_TO_RAD)
    phi0 = _f(d, "lat_0", 0.0) * DEG_TO_RAD
    return x / rc, y + phi0, None


def _inv_merc(d, a, es, x, y):
    # merc.cpp, spherical and ellipsoidal (pj_phi2, 15 iterations, tol 1e-10)
    k0 = _f(d, "k_0", _f(d, "k", 1.0))
    if d.get("lat_ts") is not None:
        phits = abs(_f(d, "lat_ts") * DEG_TO_RAD)
        k0 = math.cos(phits) / math.sqrt(1 - es * math.sin(phits) ** 2) if es else math.cos(phits)
    lam = x / k0
    if es == 0:
        return lam, np.arctan(np.sinh(y / k0)), None
    e = math.sqrt(es)
    ts = np.exp(-y / k0)
    eccnth = .5 * e
    phi = HALFPI - 2 * np.arctan(ts)
    for _ in range(15):
        con = e * np.sin(phi)
        dphi = HALFPI - 2 * np.arctan(ts * np.power((1 - con) / (1 + con), eccnth)) - phi
        phi = phi + dphi
        if np.all(np.abs(dphi) <= 1e-10):
            break
    return lam, phi, None


def _inv_laea(d, a, es, x, y):
    phi0 = _f(d, "lat_0", 0.0) * DEG_TO_RAD
    t = abs(phi0)
    if abs(t - HALFPI) < EPS10:
        mode = "S" if phi0 < 0 else "N"
    elif abs(t) < EPS10:
        mode = "E"
    else:
        mode = "O"
    x = x.copy()
    y = y.copy()
    bad = np.zeros(x.shape, dtype=bool)
    if es == 0:
        sinb1, cosb1 = math.sin(phi0), math.cos(phi0)
        rh = np.hypot(x, y)
        half = rh * .5
        bad |= half > 1.
        half = np.where(bad, 0., half)
        c = 2. * np.arcsin(half)
        with np.errstate(invalid="ignore", divide="ignore"):
            if mode == "E":
                sinz, cosz = np.sin(c), np.cos(c)
                phi = np.where(np.abs(rh) <= EPS10, 0., np.arcsin(y * sinz / rh))
                xx = x * sinz
                yy = cosz * rh
                lam = np.where(yy == 0., 0., np.arctan2(xx, yy))
            elif mode == "O":
                sinz, cosz = np.sin(c), np.cos(c)
                phi = np.where(np.abs(rh) <= EPS10, phi0, np.arcsin(cosz * sinb1 + y * sinz * cosb1 / rh))
                xx = x * (sinz * cosb1)
                yy = (cosz - np.sin(phi) * sinb1) * rh
                lam = np.where(yy == 0., 0., np.arctan2(xx, yy))
            elif mode == "N":
                phi = HALFPI - c
                lam = np.arctan2(x, -y)
            else:
                phi = c - HALFPI
                lam = np.arctan2(x, y)
        return lam, phi, bad
    # ellipsoid (classic 3-term authalic series)
    e = math.sqrt(es)
    one_es = 1 - es
    qp = _qsfn(1., e, one_es)
    apa = _authset(es)
    if mode in ("N", "S"):
        yy = -y if mode == "N" else y
        q = x * x + yy * yy
        ab = 1. - q / qp
        if mode == "S":
            ab = -ab
        lam = np.arctan2(x, yy)
        zero = q == 0
    else:
        rq = math.sqrt(.5 * qp)
        if mode == "E":
            dd = 1. / rq
            sinb1, cosb1 = 0., 1.
        else:
            sinphi = math.sin(phi0)
            sinb1 = _qsfn(sinphi, e, one_es) / qp
            cosb1 = math.sqrt(1. - sinb1 * sinb1)
            dd = math.cos(phi0) / (math.sqrt(1. - es * sinphi * sinphi) * rq * cosb1)
        xx = x / dd
        yy = y * dd
        rho = np.hypot(xx, yy)
        zero = rho < EPS10
        rho_s = np.where(zero, 1., rho)
        arg = .5 * rho_s / rq
        bad |= arg > 1.
        sce = 2. * np.arcsin(np.where(bad, 0., arg))
        cce = np.cos(sce)
        sce = np.sin(sce)
        xx = xx * sce
        if mode == "O":
            ab = cce * sinb1 + yy * sce * cosb1 / rho_s
            yy = rho_s * cosb1 * cce - yy * sinb1 * sce
        else:
            ab = yy * sce / rho_s
            yy = rho_s * cce
        lam = np.arctan2(xx, yy)
    beta = np.arcsin(np.clip(ab, -1., 1.))
    t2 = beta + beta
    phi = beta + apa[0] * np.sin(t2) + apa[1] * np.sin(t2 + t2) + apa[2] * np.sin(t2 + t2 + t2)
    lam = np.where(zero, 0., lam)
    phi = np.where(zero, phi0, phi)
    return lam, phi, bad


def _inv_stere(d, a, es, x, y):
    phi0 = _f(d, "lat_0", 0.0) * DEG_TO_RAD
    k0 = _f(d, "k_0", _f(d, "k", 1.0))
    phits = abs(_f(d, "lat_ts") * DEG_TO_RAD) if d.get("lat_ts") is not None else HALFPI
    t = abs(phi0)
    if abs(t - HALFPI) < EPS10:
        mode = "S" if phi0 < 0 else "N"
    else:
        mode = "O" if t > EPS10 else "E"
    x = np.array(x, dtype=np.float64)
    y = np.array(y, dtype=np.float64)
    bad = np.zeros(x.shape, dtype=bool)
    if es != 0:
        e = math.sqrt(es)
        if mode in ("N", "S"):
            if abs(phits - HALFPI) < EPS10:
                akm1 = 2. * k0 / math.sqrt(math.pow(1 + e, 1 + e) * math.pow(1 - e, 1 - e))
            else:
                t = math.sin(phits)
                akm1 = math.cos(phits) / _tsfn(phits, t, e)
                t *= e
                akm1 /= math.sqrt(1. - t * t)
            sinX1 = cosX1 = 0.
        else:
            t = math.sin(phi0)
            ssfn = math.tan(.5 * (HALFPI + phi0)) * math.pow((1. - t * e) / (1. + t * e), .5 * e)
            X = 2. * math.atan(ssfn) - HALFPI
            t *= e
            akm1 = 2. * k0 * math.cos(phi0) / math.sqrt(1. - t * t)
            sinX1, cosX1 = math.sin(X), math.cos(X)
        rho = np.hypot(x, y)
        if mode in ("O", "E"):
            tp = 2. * np.arctan2(rho * cosX1, akm1)
            cosphi, sinphi = np.cos(tp), np.sin(tp)
            with np.errstate(invalid="ignore", divide="ignore"):
                phi_l = np.where(rho == 0., np.arcsin(cosphi * sinX1),
                                 np.arcsin(cosphi * sinX1 + (y * sinphi * cosX1 / rho)))
            tp = np.tan(.5 * (HALFPI + phi_l))
            xx = x * sinphi
            yy = rho * cosX1 * cosphi - y * sinX1 * sinphi
            halfpi, halfe = HALFPI, .5 * e
        else:
            yy = -y if mode == "N" else y
            xx = x
            tp = -rho / akm1
            phi_l = HALFPI - 2. * np.arctan(tp)
            halfpi, halfe = -HALFPI, -.5 * e
        phi = np.full(x.shape, np.nan)
        done = np.zeros(x.shape, dtype=bool)
        for _ in range(8):  # per-point early exit, as PROJ iterates per point
            sp = e * np.sin(phi_l)
            new = 2. * np.arctan(tp * np.power((1. + sp) / (1. - sp), halfe)) - halfpi
            conv = (~done) & (np.abs(phi_l - new) < EPS10)
            phi[conv] = new[conv]
            done |= conv
            phi_l = np.where(done, phi_l, new)
            if done.all():
                break
        bad |= ~done
        if mode == "S":
            phi = -phi
        lam = np.where((xx == 0.) & (yy == 0.), 0., np.arctan2(xx, yy))
        return lam, phi, bad
    # sphere
    if mode == "O":
        sinX1, cosX1 = math.sin(phi0), math.cos(phi0)
        akm1 = 2. * k0
    elif mode == "E":
        sinX1, cosX1 = 0., 1.
        akm1 = 2. * k0
    else:
        sinX1 = cosX1 = 0.
        akm1 = math.cos(phits) / math.tan(math.pi / 4 - .5 * phits) if abs(phits - HALFPI) >= EPS10 else 2. * k0
    rh = np.hypot(x, y)
    c = 2. * np.arctan(rh / akm1)
    sinc, cosc = np.sin(c), np.cos(c)
    lam = np.zeros(x.shape)
    with np.errstate(invalid="ignore", divide="ignore"):
        if mode == "E":
            phi = np.where(np.abs(rh) <= EPS10, 0., np.arcsin(y * sinc / rh))
            m = (cosc != 0.) | (x != 0.)
            lam = np.where(m, np.arctan2(x * sinc, cosc * rh), 0.)
        elif mode == "O":
            phi = np.where(np.abs(rh) <= EPS10, phi0, np.arcsin(cosc * sinX1 + y * sinc * cosX1 / rh))
            cc = cosc - sinX1 * np.sin(phi)
            m = (cc != 0.) | (x != 0.)
            lam = np.where(m, np.arctan2(x * sinc * cosX1, cc * rh), 0.)
        else:
            yy = -y if mode == "N" else y
            phi = np.where(np.abs(rh) <= EPS10, phi0, np.arcsin(-cosc if mode == "S" else cosc))
            lam = np.where((x == 0.) & (yy == 0.), 0., np.arctan2(x, yy))
    return lam, phi, bad


def _inv_geos(d, a, es, x, y):
    h = _f(d, "h")
    flip = str(d.get("sweep", "y")) == "x"
    rg1 = h / a
    rg = 1. + rg1
    C = rg * rg - 1.
    if es != 0:
        radius_p = math.sqrt(1 - es)
        radius_p_inv2 = 1. / (1 - es)
    else:
        radius_p = radius_p_inv2 = 1.
    if flip:
        Vz = np.tan(y / rg1)
        Vy = np.tan(x / rg1) * np.hypot(1.0, Vz)
    else:
        Vy = np.tan(x / rg1)
        Vz = np.tan(y / rg1) * np.hypot(1.0, Vy)
    Vx = -1.0
    aa = Vz / radius_p
    aa = Vy * Vy + aa * aa + Vx * Vx
    b = 2 * rg * Vx
    det = b * b - 4 * aa * C
    bad = det < 0
    k = (-b - np.sqrt(np.where(bad, 0., det))) / (2. * aa)
    Vxx = rg + k * Vx
    Vy = Vy * k
    Vz = Vz * k
    lam = np.arctan2(Vy, Vxx)
    phi = np.arctan(Vz * np.cos(lam) / Vxx)
    if es != 0:
        phi = np.arctan(radius_p_inv2 * np.tan(phi))
    return lam, phi, bad


_INVERSES = {"longlat": _inv_longlat, "latlong": _inv_longlat, "lonlat": _inv_longlat, "latlon": _inv_longlat,
             "eqc": _inv_eqc, "merc": _inv_merc, "laea": _inv_laea, "stere": _inv_stere, "geos": _inv_geos}


def inverse(projection, x, y):
    """Projected (x, y) -> (lon, lat) degrees in float64; failures -> inf (pyproj errcheck=False)."""
    d = parse_proj(projection)
    name = str(d["proj"])
    if name == "ups":
        d = dict(d)
        d["proj"] = name = "stere"
        d["lat_0"] = -90.0 if d.get("south") else 90.0
        d.setdefault("k_0", 0.994)
        d.setdefault("x_0", 2000000.0)
        d.setdefault("y_0", 2000000.0)
    a, es = ellipsoid(d)
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    pm = _f(d, "pm", 0.0) or 0.0
    if name in ("longlat", "latlong", "lonlat", "latlon"):
        # geographic CRS -> its own geodetic CRS: identity (plus prime-meridian shift back to Greenwich, wrapped
        # into [-180, 180] like PROJ's adjlon; pinned by test_resamplers/test_nearest.py:130-143, sum 115591.0)
        lon = x + pm
        if pm != 0.0 and not d.get("over"):
            lon = np.where(lon > 180.0, lon - 360.0, np.where(lon < -180.0, lon + 360.0, lon))
        return lon, y * 1.0
    to_meter = _f(d, "to_meter", None)
    if to_meter is None:
        to_meter = _UNITS[str(d.get("units", "m"))]
    x0 = _f(d, "x_0", 0.0)
    y0 = _f(d, "y_0", 0.0)
    ra = 1. / a
    xs = (x * to_meter - x0) * ra
    ys = (y * to_meter - y0) * ra
    lam, phi, bad = _INVERSES[name](d, a, es, xs, ys)
    lam0 = _f(d, "lon_0", 0.0) * DEG_TO_RAD
    lam = lam + lam0
    if not d.get("over"):
        lam = _adjlon(lam)
    lon = lam * RAD_TO_DEG + pm
    lat = phi * RAD_TO_DEG
    if bad is not None and np.any(bad):
        lon = np.where(bad, np.inf, lon)
        lat = np.where(bad, np.inf, lat)
    nonfinite = ~np.isfinite(x) | ~np.isfinite(y)
    if nonfinite.any():
        lon = np.where(nonfinite, np.inf, lon)
        lat = np.where(nonfinite, np.inf, lat)
    return lon, lat


# ---------------------------------------------------------------- area pixel centres
def proj_vectors(area_extent, width, height, dtype, pixel_size=None):
    """1-D projection coordinates of pixel centres (geometry.py:1374-1380, 1593-1609).

    ``np.arange(n, dtype) * pixel_size + offset`` is evaluated in ``dtype`` because
    pixel sizes/offsets are Python floats (NEP 50 weak scalars) - SURVEY 9.3.
    """
    psx = (area_extent[2] - area_extent[0]) / float(width)
    psy = (area_extent[3] - area_extent[1]) / float(height)
    if pixel_size is not None:
        psx, psy = pixel_size
    ulx = float(area_extent[0]) + float(psx) / 2
    uly = float(area_extent[3]) - float(psy) / 2
    x = np.arange(0, width, dtype=dtype) * float(psx) + ulx
    y = np.arange(0, height, dtype=dtype) * -float(psy) + uly
    return x, y


def area_lonlats(projection, area_extent, width, height, dtype=np.float64, data_slice=None):
    """AreaDefinition.get_lonlats (geometry.py:2558-2645) without caching/dask."""
    x, y = proj_vectors(area_extent, width, height, dtype)
    if data_slice is not None:
        if isinstance(data_slice, slice):
            ys, xs = data_slice, slice(None)
        else:
            ys, xs = data_slice
        y = y[ys]
        x = x[xs]
    scalar_y = np.ndim(y) == 0
    scalar_x = np.ndim(x) == 0
    tx, ty = np.meshgrid(x, y)
    lon, lat = inverse(projection, tx, ty)
    lon = np.asanyarray(lon, dtype=dtype)
    lat = np.asanyarray(lat, dtype=dtype)
    if scalar_y and not scalar_x:
        lon, lat = lon[0], lat[0]
    elif scalar_x and not scalar_y:
        lon, lat = lon[:, 0], lat[:, 0]
    elif scalar_x and scalar_y:
        lon, lat = lon[0, 0], lat[0, 0]
    return lon, lat


# --------------------------------------------------------------------------- forwards
# PROJ's forward projections (src/projections/{eqc,merc,laea,stere,geos}.cpp, pj_fwd), needed by the bilinear
# resampler, which projects the source lon/lat into the TARGET CRS (bilinear/_base.py:188-191, 299-307).
def _modes(d, laea):
    phi0 = _f(d, "lat_0", 0.0) * DEG_TO_RAD
    t = abs(phi0)
    if abs(t - HALFPI) < EPS10:
        return phi0, ("S" if phi0 < 0 else "N")
    if laea:
        return phi0, ("E" if abs(t) < EPS10 else "O")
    return phi0, ("O" if t > EPS10 else "E")


def _fwd_eqc(d, a, es, lam, phi):
    rc = math.cos(_f(d, "lat_ts", 0.0) * DEG_TO_RAD)
    return rc * lam, phi - _f(d, "lat_0", 0.0) * DEG_TO_RAD, None


def _fwd_merc(d, a, es, lam, phi):
    k0 = _f(d, "k_0", _f(d, "k", 1.0))
    if d.get("lat_ts") is not None:
        phits = abs(_f(d, "lat_ts") * DEG_TO_RAD)
        k0 = math.cos(phits) / math.sqrt(1 - es * math.sin(phits) ** 2) if es else math.cos(phits)
    bad = np.abs(np.abs(phi) - HALFPI) <= EPS10
    with np.errstate(all="ignore"):
        if es == 0:
            y = k0 * np.arcsinh(np.tan(phi))
        else:
            e = math.sqrt(es)
            y = k0 * (np.arcsinh(np.tan(phi)) - e * np.arctanh(e * np.sin(phi)))
    return k0 * lam, y, bad


def _fwd_laea(d, a, es, lam, phi):
    phi0, mode = _modes(d, True)
    coslam, sinlam = np.cos(lam), np.sin(lam)
    sinphi = np.sin(phi)
    bad = np.zeros(lam.shape, dtype=bool)
    with np.errstate(all="ignore"):
        if es == 0:
            cosphi = np.cos(phi)
            sinb1, cosb1 = math.sin(phi0), math.cos(phi0)
            if mode in ("E", "O"):
                y = 1. + cosphi * coslam if mode == "E" else 1. + sinb1 * sinphi + cosb1 * cosphi * coslam
                bad |= y <= EPS10
                y = np.sqrt(2. / np.where(bad, 1., y))
                x = y * cosphi * sinlam
                y = y * (sinphi if mode == "E" else cosb1 * sinphi - sinb1 * cosphi * coslam)
            else:
                if mode == "N":
                    coslam = -coslam
                bad |= np.abs(phi + phi0) < EPS10
                y = math.pi / 4 - phi * .5
                y = 2. * (np.cos(y) if mode == "S" else np.sin(y))
                x = y * sinlam
                y = y * coslam
            return x, y, bad
        e = math.sqrt(es)
        one_es = 1 - es
        qp = _qsfn(1., e, one_es)
        rq = math.sqrt(.5 * qp)
        con = e * sinphi
        q = one_es * (sinphi / (1 - con * con) - (.5 / e) * np.log((1 - con) / (1 + con)))
        if mode in ("E", "O"):
            if mode == "E":
                dd, sinb1, cosb1 = 1. / rq, 0., 1.
                xmf, ymf = 1., .5 * qp
            else:
                sp0 = math.sin(phi0)
                sinb1 = _qsfn(sp0, e, one_es) / qp
                cosb1 = math.sqrt(1. - sinb1 * sinb1)
                dd = math.cos(phi0) / (math.sqrt(1. - es * sp0 * sp0) * rq * cosb1)
                xmf, ymf = rq * dd, rq / dd
            sinb = q / qp
            cosb2 = 1. - sinb * sinb
            cosb = np.where(cosb2 > 0, np.sqrt(np.where(cosb2 > 0, cosb2, 0.)), 0.)
            b = 1. + sinb1 * sinb + cosb1 * cosb * coslam if mode == "O" else 1. + cosb * coslam
            bad |= np.abs(b) < EPS10
            b = np.sqrt(2. / np.where(bad, 1., b))
            if mode == "O":
                y = ymf * b * (cosb1 * sinb - sinb1 * cosb * coslam)
            else:
                y = b * sinb * ymf
            x = xmf * b * cosb * sinlam
            return x, y, bad
        if mode == "N":
            b = HALFPI + phi
            q = qp - q
        else:
            b = phi - HALFPI
            q = qp + q
        bad |= np.abs(b) < EPS10
        ok = q >= 1e-15
        bb = np.sqrt(np.where(ok, q, 0.))
        x = np.where(ok, bb * sinlam, 0.)
        y = np.where(ok, coslam * (bb if mode == "S" else -bb), 0.)
        return x, y, bad


def _fwd_stere(d, a, es, lam, phi):
    phi0, mode = _modes(d, False)
    k0 = _f(d, "k_0", _f(d, "k", 1.0))
    phits = abs(_f(d, "lat_ts") * DEG_TO_RAD) if d.get("lat_ts") is not None else HALFPI
    coslam, sinlam = np.cos(lam), np.sin(lam)
    sinphi = np.sin(phi)
    bad = np.zeros(lam.shape, dtype=bool)
    with np.errstate(all="ignore"):
        if es != 0:
            e = math.sqrt(es)
            if mode in ("N", "S"):
                if abs(phits - HALFPI) < EPS10:
                    akm1 = 2. * k0 / math.sqrt(math.pow(1 + e, 1 + e) * math.pow(1 - e, 1 - e))
                else:
                    t = math.sin(phits)
                    akm1 = math.cos(phits) / _tsfn(phits, t, e)
                    t *= e
                    akm1 /= math.sqrt(1. - t * t)
                if mode == "S":
                    phi, coslam, sinphi = -phi, -coslam, -sinphi
                sp = sinphi * e
                ts = np.tan(.5 * (HALFPI - phi)) / np.power((1 - sp) / (1 + sp), .5 * e)
                x = np.where(np.abs(phi - HALFPI) < 1e-15, 0., akm1 * ts)
                y = -x * coslam
                return x * sinlam, y, bad
            t = math.sin(phi0)
            X1 = 2. * math.atan(math.tan(.5 * (HALFPI + phi0)) * math.pow((1. - t * e) / (1. + t * e), .5 * e)) - HALFPI
            t *= e
            akm1 = 2. * k0 * math.cos(phi0) / math.sqrt(1. - t * t)
            sinX1, cosX1 = math.sin(X1), math.cos(X1)
            sp = sinphi * e
            X = 2. * np.arctan(np.tan(.5 * (HALFPI + phi)) * np.power((1. - sp) / (1. + sp), .5 * e)) - HALFPI
            sinX, cosX = np.sin(X), np.cos(X)
            if mode == "O":
                denom = cosX1 * (1. + sinX1 * sinX + cosX1 * cosX * coslam)
                bad |= denom == 0
                A = akm1 / np.where(bad, 1., denom)
                y = A * (cosX1 * sinX - sinX1 * cosX * coslam)
            else:
                denom = 1. + cosX * coslam
                bad |= denom == 0
                A = akm1 / np.where(bad, 1., denom)
                y = A * sinX
            return A * cosX * sinlam, y, bad
        cosphi = np.cos(phi)
        if mode in ("O", "E"):
            sinph0, cosph0 = (math.sin(phi0), math.cos(phi0)) if mode == "O" else (0., 1.)
            akm1 = 2. * k0
            y = 1. + cosphi * coslam if mode == "E" else 1. + sinph0 * sinphi + cosph0 * cosphi * coslam
            bad |= y <= EPS10
            y = akm1 / np.where(bad, 1., y)
            x = y * cosphi * sinlam
            y = y * (sinphi if mode == "E" else cosph0 * sinphi - sinph0 * cosphi * coslam)
            return x, y, bad
        akm1 = math.cos(phits) / math.tan(math.pi / 4 - .5 * phits) if abs(phits - HALFPI) >= EPS10 else 2. * k0
        if mode == "N":
            coslam, phi = -coslam, -phi
        bad |= np.abs(phi - HALFPI) < 1e-8
        y = akm1 * np.tan(math.pi / 4 + .5 * phi)
        return sinlam * y, y * coslam, bad


def _fwd_geos(d, a, es, lam, phi):
    h = _f(d, "h")
    flip = str(d.get("sweep", "y")) == "x"
    rg1 = h / a
    rg = 1. + rg1
    with np.errstate(all="ignore"):
        if es != 0:
            radius_p, radius_p2, radius_p_inv2 = math.sqrt(1 - es), 1 - es, 1. / (1 - es)
            phi = np.arctan(radius_p2 * np.tan(phi))
            r = radius_p / np.hypot(radius_p * np.cos(phi), np.sin(phi))
        else:
            radius_p_inv2 = 1.
            r = 1.
        Vx = r * np.cos(lam) * np.cos(phi)
        Vy = r * np.sin(lam) * np.cos(phi)
        Vz = r * np.sin(phi)
        bad = ((rg - Vx) * Vx - Vy * Vy - Vz * Vz * radius_p_inv2) < 0.
        tmp = rg - Vx
        if flip:
            x = rg1 * np.arctan(Vy / np.hypot(Vz, tmp))
            y = rg1 * np.arctan(Vz / tmp)
        else:
            x = rg1 * np.arctan(Vy / tmp)
            y = rg1 * np.arctan(Vz / np.hypot(Vy, tmp))
    return x, y, bad


_FORWARDS = {"eqc": _fwd_eqc, "merc": _fwd_merc, "laea": _fwd_laea, "stere": _fwd_stere, "geos": _fwd_geos}


def forward(projection, lon, lat):
    """(lon, lat) degrees -> projected (x, y) in float64, like ``pyproj.Proj(projection)(lon, lat)``; failures and
    non-finite input -> inf."""
    d = parse_proj(projection)
    name = str(d["proj"])
    if name == "ups":
        d = dict(d)
        d["proj"] = name = "stere"
        d["lat_0"] = -90.0 if d.get("south") else 90.0
        d["k_0"], d["x_0"], d["y_0"], d["lon_0"] = 0.994, 2000000.0, 2000000.0, 0.0
    a, es = ellipsoid(d)
    lon = np.asarray(lon, dtype=np.float64)
    lat = np.asarray(lat, dtype=np.float64)
    pm = _f(d, "pm", 0.0) or 0.0
    if name in ("longlat", "latlong", "lonlat", "latlon"):
        return lon - pm, lat * 1.0
    nonfinite = ~np.isfinite(lon) | ~np.isfinite(lat)
    lam = (np.where(nonfinite, 0., lon) - pm) * DEG_TO_RAD - _f(d, "lon_0", 0.0) * DEG_TO_RAD
    phi = np.where(nonfinite, 0., lat) * DEG_TO_RAD
    over_range = np.abs(phi) - HALFPI > 1e-12  # pj_fwd: latitude outside [-90, 90] is an error
    if not d.get("over"):
        lam = _adjlon(lam)
    x, y, bad = _FORWARDS[name](d, a, es, lam, phi)
    to_meter = _f(d, "to_meter", None)
    if to_meter is None:
        to_meter = _UNITS[str(d.get("units", "m"))]
    x = (a * x + _f(d, "x_0", 0.0)) / to_meter
    y = (a * y + _f(d, "y_0", 0.0)) / to_meter
    fail = nonfinite | over_range
    if bad is not None:
        fail = fail | bad
    return np.where(fail, np.inf, x), np.where(fail, np.inf, y)
