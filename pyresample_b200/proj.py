"""PROJ-string handling for the device-side inverse projections (host side, no GPU needed).

The reference hands every ``AreaDefinition`` to pyproj (pyresample/geometry.py:1596-1602,
2630-2636).  pyproj/PROJ is not available on the B200 boxes, and the per-pixel inverse
projection is part of the hot path, so the projection is described
by a small plain-old-data record (``ProjSpec`` -> ``prs_proj_params`` in
include/prs_b200.h) that the CUDA kernels evaluate themselves.

Supported ``+proj=``: longlat/latlong, eqc, merc, laea (sphere + ellipsoid),
stere/ups (sphere + ellipsoid; polar, oblique, equatorial), geos (sweep x/y).
Anything else raises ``NotImplementedError`` - there is no CPU fallback.
"""
from __future__ import annotations

import ctypes
import math

PROJ_LONGLAT = 0
PROJ_EQC = 1
PROJ_MERC = 2
PROJ_LAEA = 3
PROJ_STERE = 4
PROJ_GEOS = 5

MODE_N_POLE = 0
MODE_S_POLE = 1
MODE_OBLIQ = 2
MODE_EQUIT = 3

_ELLIPSOIDS = {
    # name: (a, b, rf)  -- exactly one of b / rf is given
    "WGS84": (6378137.0, None, 298.257223563),
    "GRS80": (6378137.0, None, 298.257222101),
    "WGS72": (6378135.0, None, 298.26),
    "bessel": (6377397.155, None, 299.1528128),
    "intl": (6378388.0, None, 297.0),
    "clrk66": (6378206.4, 6356583.8, None),
    "krass": (6378245.0, None, 298.3),
    "airy": (6377563.396, 6356256.910, None),
    "sphere": (6370997.0, 6370997.0, None),
}
_DATUMS = {"WGS84": "WGS84", "NAD83": "GRS80", "NAD27": "clrk66"}
_UNIT_TO_METER = {"m": 1.0, "km": 1000.0, "dm": 0.1, "cm": 0.01, "mm": 0.001,
                  "ft": 0.3048, "us-ft": 0.304800609601219, "kmi": 1852.0}
_GEOGRAPHIC = ("longlat", "latlong", "lonlat", "latlon")

_HALFPI = math.pi / 2
_EPS10 = 1e-10
_D2R = .017453292519943296


class prs_proj_params(ctypes.Structure):
    """Mirror of ``struct prs_proj_params`` (include/prs_b200.h)."""

    _fields_ = [
        ("kind", ctypes.c_int32),
        ("mode", ctypes.c_int32),
        ("ellipsoidal", ctypes.c_int32),
        ("flip_axis", ctypes.c_int32),
        ("a", ctypes.c_double),
        ("ra", ctypes.c_double),
        ("es", ctypes.c_double),
        ("e", ctypes.c_double),
        ("one_es", ctypes.c_double),
        ("lam0", ctypes.c_double),
        ("phi0", ctypes.c_double),
        ("x0", ctypes.c_double),
        ("y0", ctypes.c_double),
        ("to_meter", ctypes.c_double),
        ("pm_deg", ctypes.c_double),
        ("over", ctypes.c_int32),
        ("_pad", ctypes.c_int32),
        ("c", ctypes.c_double * 12),
    ]


def proj_to_dict(projection) -> dict:
    """Normalise a PROJ.4 dict / string / pyproj CRS into a plain ``{key: value}`` dict."""
    if hasattr(projection, "to_dict") and not isinstance(projection, dict):  # pyproj.CRS duck type
        projection = projection.to_dict()
    if isinstance(projection, dict):
        return {str(k).lstrip("+"): v for k, v in projection.items()}
    text = str(projection).strip()
    if text.upper().startswith("EPSG:"):
        code = int(text.split(":", 1)[1])
        if code == 4326:
            return {"proj": "longlat", "datum": "WGS84"}
        raise NotImplementedError("EPSG:%d needs the PROJ database; give the PROJ.4 parameters instead" % code)
    out = {}
    for token in text.split():
        token = token.lstrip("+")
        if not token:
            continue
        if "=" in token:
            key, value = token.split("=", 1)
            out[key] = value
        else:
            out[token] = True
    return out


def _num(params, key, default=None):
    value = params.get(key)
    if value is None:
        return default
    return float(value)


def _resolve_ellipsoid(params):
    """Return (a, es).  Precedence follows PROJ's pj_ellipsoid: R, then a with a shape parameter, then ellps/datum."""
    if params.get("R") is not None:
        return float(params["R"]), 0.0
    name = params.get("ellps")
    if name is None and params.get("datum") is not None:
        name = _DATUMS.get(str(params["datum"]))
        if name is None:
            raise NotImplementedError("datum %r" % (params["datum"],))
    a = _num(params, "a")
    b = rf = None
    if name is not None or a is None:
        if name is None:
            name = "WGS84" if params.get("proj") in _GEOGRAPHIC else "GRS80"
        if str(name) not in _ELLIPSOIDS:
            raise NotImplementedError("ellipsoid %r" % (name,))
        ea, b, rf = _ELLIPSOIDS[str(name)]
        if a is None:
            a = ea
    explicit = False
    for key in ("b", "rf", "f", "es", "e"):
        if params.get(key) is not None:
            explicit = True
    if explicit:
        if params.get("b") is not None:
            b, rf = _num(params, "b"), None
        elif params.get("rf") is not None:
            rf, b = _num(params, "rf"), None
        elif params.get("f") is not None:
            f = _num(params, "f")
            return a, 2 * f - f * f
        elif params.get("es") is not None:
            return a, _num(params, "es")
        else:
            return a, _num(params, "e") ** 2
    if b is not None:
        return a, (0.0 if b == a else 1.0 - (b * b) / (a * a))
    if rf is not None:
        f = 1.0 / rf
        return a, 2 * f - f * f
    return a, 0.0


# What each projection reads, beyond the ellipsoid / unit / false-origin parameters common to all.  Anything else in a
# definition would silently change the coordinates PROJ computes (axis order, datum shifts, a second standard
# parallel ...): such definitions are refused instead of half-understood.
_COMMON_KEYS = {"proj", "a", "b", "rf", "f", "es", "e", "R", "ellps", "datum", "units", "to_meter", "x_0", "y_0", "lon_0",
                "pm", "over", "no_defs", "type", "wktext", "towgs84", "nadgrids", "vunits", "geoidgrids"}
_PROJ_KEYS = {"longlat": set(), "latlong": set(), "latlon": set(), "lonlat": set(),
              "eqc": {"lat_ts", "lat_0"}, "merc": {"lat_ts", "k_0", "k", "lat_0"},
              "laea": {"lat_0"}, "stere": {"lat_0", "lat_ts", "k_0", "k"},
              "ups": {"south", "lat_0", "lat_ts", "k_0", "k"}, "geos": {"h", "sweep", "lat_0"}}
_NEUTRAL = {"towgs84": ("0,0,0", "0,0,0,0,0,0,0", "0", 0, 0.0), "nadgrids": ("@null",), "vunits": ("m",)}


def _check_known_parameters(p, name):
    if name not in _PROJ_KEYS:
        return  # reported as an unsupported projection by the caller
    for key, value in p.items():
        if key in _NEUTRAL:
            if value is True or str(value).replace(" ", "") in [str(v) for v in _NEUTRAL[key]]:
                continue
            raise NotImplementedError("+%s=%s changes the coordinates and is not implemented on the device" % (key, value))
        if key == "geoidgrids":
            raise NotImplementedError("+geoidgrids is not implemented on the device")
        if key not in _COMMON_KEYS and key not in _PROJ_KEYS[name]:
            raise NotImplementedError("+%s is not understood for +proj=%s (supported: %s)"
                                      % (key, name, ", ".join(sorted(_PROJ_KEYS[name])) or "none"))
    pm = p.get("pm")
    if pm is not None:
        try:
            float(pm)
        except (TypeError, ValueError):
            raise NotImplementedError("named prime meridian +pm=%s (give it in degrees)" % (pm,))


def _qsfn(sinphi, e, one_es):
    if e >= 1e-7:
        con = e * sinphi
        return one_es * (sinphi / (1 - con * con) - (.5 / e) * math.log((1 - con) / (1 + con)))
    return sinphi + sinphi


class ProjSpec:
    """Host-side description of one projection; ``.cstruct`` is what crosses the C-ABI."""

    def __init__(self, projection):
        p = proj_to_dict(projection)
        self.params = p
        name = str(p.get("proj"))
        _check_known_parameters(p, name)
        if name == "ups":
            # PROJ's ups fixes these whatever the string says (src/projections/stere.cpp, PJ_PROJECTION(ups))
            p = dict(p)
            name = "stere"
            p["lat_0"] = -90.0 if p.get("south") else 90.0
            p["k_0"], p["x_0"], p["y_0"], p["lon_0"] = 0.994, 2000000.0, 2000000.0, 0.0
            p.pop("lat_ts", None)
            p.pop("k", None)
        self.name = name
        a, es = _resolve_ellipsoid(p)
        s = prs_proj_params()
        s.a, s.ra, s.es, s.e, s.one_es = a, 1.0 / a, es, math.sqrt(es), 1.0 - es
        s.ellipsoidal = 1 if es != 0 else 0
        s.lam0 = _num(p, "lon_0", 0.0) * _D2R
        s.phi0 = _num(p, "lat_0", 0.0) * _D2R
        s.x0 = _num(p, "x_0", 0.0)
        s.y0 = _num(p, "y_0", 0.0)
        to_meter = _num(p, "to_meter")
        if to_meter is None:
            unit = str(p.get("units", "m"))
            if unit not in _UNIT_TO_METER:
                raise NotImplementedError("units=%s" % unit)
            to_meter = _UNIT_TO_METER[unit]
        s.to_meter = to_meter
        s.pm_deg = _num(p, "pm", 0.0) or 0.0
        s.over = 1 if p.get("over") else 0
        c = [0.0] * 12
        k0 = _num(p, "k_0", _num(p, "k", 1.0))
        if name in _GEOGRAPHIC:
            s.kind = PROJ_LONGLAT
            if s.lam0 != 0.0:  # PROJ's pj_inv would add it; nothing in the reference's tests or docs uses it
                raise NotImplementedError("+lon_0 on a geographic CRS is not implemented (use +pm for a shifted meridian)")
        elif name == "eqc":
            s.kind = PROJ_EQC
            rc = math.cos(_num(p, "lat_ts", 0.0) * _D2R)
            if rc <= 0:
                raise ValueError("eqc: lat_ts larger than 90 degrees")
            c[0] = rc
            s.es, s.e, s.one_es, s.ellipsoidal = 0.0, 0.0, 1.0, 0
        elif name == "merc":
            s.kind = PROJ_MERC
            if p.get("lat_ts") is not None:
                phits = abs(_num(p, "lat_ts") * _D2R)
                if es != 0:
                    k0 = math.cos(phits) / math.sqrt(1 - es * math.sin(phits) ** 2)
                else:
                    k0 = math.cos(phits)
            c[0] = k0
        elif name == "laea":
            s.kind = PROJ_LAEA
            s.mode = self._mode(s.phi0, laea=True)
            if es == 0:
                c[0], c[1] = math.sin(s.phi0), math.cos(s.phi0)
            else:
                e, one_es = s.e, s.one_es
                qp = _qsfn(1.0, e, one_es)
                t = es * es
                apa0 = es * (1 / 3.) + t * (31 / 180.)
                apa1 = t * (23 / 360.)
                t *= es
                apa0 += t * (517 / 5040.)
                apa1 += t * (251 / 3780.)
                apa2 = t * (761 / 45360.)
                rq = math.sqrt(.5 * qp)
                dd, sinb1, cosb1 = 1.0, 0.0, 1.0
                if s.mode == MODE_EQUIT:
                    dd = 1.0 / rq
                elif s.mode == MODE_OBLIQ:
                    sinphi = math.sin(s.phi0)
                    sinb1 = _qsfn(sinphi, e, one_es) / qp
                    cosb1 = math.sqrt(1.0 - sinb1 * sinb1)
                    dd = math.cos(s.phi0) / (math.sqrt(1.0 - es * sinphi * sinphi) * rq * cosb1)
                c[0], c[1], c[2], c[3], c[4] = sinb1, cosb1, qp, rq, dd
                c[5], c[6], c[7] = apa0, apa1, apa2
        elif name == "stere":
            s.kind = PROJ_STERE
            s.mode = self._mode(s.phi0, laea=False)
            phits = abs(_num(p, "lat_ts") * _D2R) if p.get("lat_ts") is not None else _HALFPI
            sinX1 = cosX1 = 0.0
            if es != 0:
                e = s.e
                if s.mode in (MODE_N_POLE, MODE_S_POLE):
                    if abs(phits - _HALFPI) < _EPS10:
                        akm1 = 2. * k0 / math.sqrt(math.pow(1 + e, 1 + e) * math.pow(1 - e, 1 - e))
                    else:
                        t = math.sin(phits)
                        tsfn = math.tan(.5 * (_HALFPI - phits)) / math.pow((1 - t * e) / (1 + t * e), .5 * e)
                        akm1 = math.cos(phits) / tsfn
                        t *= e
                        akm1 /= math.sqrt(1. - t * t)
                else:
                    t = math.sin(s.phi0)
                    ssfn = math.tan(.5 * (_HALFPI + s.phi0)) * math.pow((1. - t * e) / (1. + t * e), .5 * e)
                    X = 2. * math.atan(ssfn) - _HALFPI
                    t *= e
                    akm1 = 2. * k0 * math.cos(s.phi0) / math.sqrt(1. - t * t)
                    sinX1, cosX1 = math.sin(X), math.cos(X)
            else:
                if s.mode == MODE_OBLIQ:
                    sinX1, cosX1 = math.sin(s.phi0), math.cos(s.phi0)
                    akm1 = 2. * k0
                elif s.mode == MODE_EQUIT:
                    sinX1, cosX1 = 0.0, 1.0
                    akm1 = 2. * k0
                else:
                    if abs(phits - _HALFPI) >= _EPS10:
                        akm1 = math.cos(phits) / math.tan(math.pi / 4 - .5 * phits)
                    else:
                        akm1 = 2. * k0
            c[0], c[1], c[2] = akm1, sinX1, cosX1
        elif name == "geos":
            s.kind = PROJ_GEOS
            h = _num(p, "h")
            if h is None:
                raise ValueError("geos: +h is required")
            s.flip_axis = 1 if str(p.get("sweep", "y")) == "x" else 0
            rg1 = h / a
            rg = 1.0 + rg1
            c[0], c[1], c[2] = rg1, rg, rg * rg - 1.0
            if es != 0:
                c[3], c[4] = math.sqrt(s.one_es), 1.0 / s.one_es
            else:
                c[3], c[4] = 1.0, 1.0
        else:
            raise NotImplementedError("+proj=%s has no device-side inverse (supported: longlat, eqc, merc, laea, "
                                      "stere, ups, geos)" % name)
        for i, v in enumerate(c):
            s.c[i] = v
        self.cstruct = s

    @staticmethod
    def _mode(phi0, laea):
        t = abs(phi0)
        if abs(t - _HALFPI) < _EPS10:
            return MODE_S_POLE if phi0 < 0 else MODE_N_POLE
        if laea:
            return MODE_EQUIT if abs(t) < _EPS10 else MODE_OBLIQ
        return MODE_OBLIQ if t > _EPS10 else MODE_EQUIT

    @property
    def is_geographic(self):
        return self.cstruct.kind == PROJ_LONGLAT

    @property
    def is_separable(self):
        """True when lon depends only on x and lat only on y (row/column tables replace per-pixel maths)."""
        return self.cstruct.kind in (PROJ_LONGLAT, PROJ_EQC, PROJ_MERC)
