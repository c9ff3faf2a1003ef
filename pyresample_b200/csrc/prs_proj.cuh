// Device-side inverse cartographic projections (double precision), replacing the per-pixel
// pyproj.Transformer(...).transform(x, y, direction=INVERSE) call of AreaDefinition.get_lonlats
// (pyresample/geometry.py:2630-2636).  Formulas follow PROJ's published algorithms
// (src/projections/{latlong,eqc,merc,laea,stere,geos}.cpp, pj_inv); constants are prepared on the host
// in pyresample_b200/proj.py.  Failures return false -> the caller reports (inf, inf) like pyproj.
#pragma once
#include "prs_common.cuh"

namespace prs {

#define PRS_HALFPI 1.5707963267948966
#define PRS_PI 3.14159265358979323846
#define PRS_EPS10 1e-10
#define PRS_RAD_TO_DEG 57.295779513082321

__device__ __forceinline__ double adjlon(double lon)
{
    if (fabs(lon) < PRS_PI + 1e-12) return lon;
    lon += PRS_PI;
    lon -= 2.0 * PRS_PI * floor(lon / (2.0 * PRS_PI));
    lon -= PRS_PI;
    return lon;
}

// lam, phi in radians relative to lon_0; returns false on projection failure.
__device__ inline bool inv_laea(const prs_proj_params &P, double x, double y, double &lam, double &phi)
{
    if (!P.ellipsoidal) {
        const double sinb1 = P.c[0], cosb1 = P.c[1];
        double rh = hypot(x, y);
        double half = rh * .5;
        if (half > 1.) return false;
        double c = 2. * asin(half);
        switch (P.mode) {
        case PRS_MODE_EQUIT: {
            double sinz, cosz;
            sincos(c, &sinz, &cosz);
            phi = fabs(rh) <= PRS_EPS10 ? 0. : asin(y * sinz / rh);
            x *= sinz;
            y = cosz * rh;
            lam = (y == 0.) ? 0. : atan2(x, y);
            break;
        }
        case PRS_MODE_OBLIQ: {
            double sinz, cosz;
            sincos(c, &sinz, &cosz);
            phi = fabs(rh) <= PRS_EPS10 ? P.phi0 : asin(cosz * sinb1 + y * sinz * cosb1 / rh);
            x *= sinz * cosb1;
            y = (cosz - sin(phi) * sinb1) * rh;
            lam = (y == 0.) ? 0. : atan2(x, y);
            break;
        }
        case PRS_MODE_N_POLE:
            y = -y;
            phi = PRS_HALFPI - c;
            lam = atan2(x, y);
            break;
        default:
            phi = c - PRS_HALFPI;
            lam = atan2(x, y);
            break;
        }
        return true;
    }
    const double sinb1 = P.c[0], cosb1 = P.c[1], qp = P.c[2], rq = P.c[3], dd = P.c[4];
    double ab;
    if (P.mode == PRS_MODE_N_POLE || P.mode == PRS_MODE_S_POLE) {
        if (P.mode == PRS_MODE_N_POLE) y = -y;
        double q = x * x + y * y;
        if (q == 0.) {
            lam = 0.;
            phi = P.phi0;
            return true;
        }
        ab = 1. - q / qp;
        if (P.mode == PRS_MODE_S_POLE) ab = -ab;
    } else {
        x /= dd;
        y *= dd;
        double rho = hypot(x, y);
        if (rho < PRS_EPS10) {
            lam = 0.;
            phi = P.phi0;
            return true;
        }
        double arg = .5 * rho / rq;
        if (arg > 1.) return false;
        double sCe, cCe;
        sincos(2. * asin(arg), &sCe, &cCe);
        x *= sCe;
        if (P.mode == PRS_MODE_OBLIQ) {
            ab = cCe * sinb1 + y * sCe * cosb1 / rho;
            y = rho * cosb1 * cCe - y * sinb1 * sCe;
        } else {
            ab = y * sCe / rho;
            y = rho * cCe;
        }
    }
    lam = atan2(x, y);
    double beta = asin(fmin(fmax(ab, -1.), 1.));
    double t = beta + beta;
    phi = beta + P.c[5] * sin(t) + P.c[6] * sin(t + t) + P.c[7] * sin(t + t + t);
    return true;
}

// ((1 + s) / (1 - s)) ** halfe  with s = e * sin(phi), |s| <= e.  For the eccentricities of real ellipsoids
// (e < 0.12) this is exp(2 * halfe * atanh(s)) with the atanh series summed to s^17 (< 1e-17 relative): one exp
// instead of a divide and a double-precision pow, and accurate to the last bits of what PROJ's pow() returns.
__device__ __forceinline__ double pow_ratio(double s, double halfe, double e)
{
    if (e < 0.12) {
        const double s2 = s * s;
        double a = 1. / 17.;
        a = a * s2 + 1. / 15.;
        a = a * s2 + 1. / 13.;
        a = a * s2 + 1. / 11.;
        a = a * s2 + 1. / 9.;
        a = a * s2 + 1. / 7.;
        a = a * s2 + 1. / 5.;
        a = a * s2 + 1. / 3.;
        a = a * s2 + 1.;
        return exp(2. * halfe * (a * s));
    }
    return pow((1. + s) / (1. - s), halfe);
}

__device__ inline bool inv_stere(const prs_proj_params &P, double x, double y, double &lam, double &phi)
{
    const double akm1 = P.c[0], sinX1 = P.c[1], cosX1 = P.c[2];
    if (P.ellipsoidal) {
        double rho = hypot(x, y);
        double tp, phi_l, halfpi, halfe;
        if (P.mode == PRS_MODE_OBLIQ || P.mode == PRS_MODE_EQUIT) {
            tp = 2. * atan2(rho * cosX1, akm1);
            double sinphi, cosphi;
            sincos(tp, &sinphi, &cosphi);
            if (rho == 0.0)
                phi_l = asin(cosphi * sinX1);
            else
                phi_l = asin(cosphi * sinX1 + (y * sinphi * cosX1 / rho));
            tp = tan(.5 * (PRS_HALFPI + phi_l));
            x *= sinphi;
            y = rho * cosX1 * cosphi - y * sinX1 * sinphi;
            halfpi = PRS_HALFPI;
            halfe = .5 * P.e;
        } else {
            if (P.mode == PRS_MODE_N_POLE) y = -y;
            tp = -rho / akm1;
            phi_l = PRS_HALFPI - 2. * atan(tp);
            halfpi = -PRS_HALFPI;
            halfe = -.5 * P.e;
        }
        for (int i = 0; i < 8; ++i) {
            double sinphi = P.e * sin(phi_l);
            double p = 2. * atan(tp * pow_ratio(sinphi, halfe, P.e)) - halfpi;
            if (fabs(phi_l - p) < PRS_EPS10) {
                phi = (P.mode == PRS_MODE_S_POLE) ? -p : p;
                lam = (x == 0. && y == 0.) ? 0. : atan2(x, y);
                return true;
            }
            phi_l = p;
        }
        return false;
    }
    double rh = hypot(x, y);
    double c = 2. * atan(rh / akm1);
    double sinc, cosc;
    sincos(c, &sinc, &cosc);
    lam = 0.;
    switch (P.mode) {
    case PRS_MODE_EQUIT:
        phi = fabs(rh) <= PRS_EPS10 ? 0. : asin(y * sinc / rh);
        if (cosc != 0. || x != 0.) lam = atan2(x * sinc, cosc * rh);
        break;
    case PRS_MODE_OBLIQ: {
        phi = fabs(rh) <= PRS_EPS10 ? P.phi0 : asin(cosc * sinX1 + y * sinc * cosX1 / rh);
        double cc = cosc - sinX1 * sin(phi);
        if (cc != 0. || x != 0.) lam = atan2(x * sinc * cosX1, cc * rh);
        break;
    }
    case PRS_MODE_N_POLE:
        y = -y;
        phi = fabs(rh) <= PRS_EPS10 ? P.phi0 : asin(cosc);
        lam = (x == 0. && y == 0.) ? 0. : atan2(x, y);
        break;
    default:
        phi = fabs(rh) <= PRS_EPS10 ? P.phi0 : asin(-cosc);
        lam = (x == 0. && y == 0.) ? 0. : atan2(x, y);
        break;
    }
    return true;
}

__device__ inline bool inv_geos(const prs_proj_params &P, double x, double y, double &lam, double &phi)
{
    const double rg1 = P.c[0], rg = P.c[1], C = P.c[2], radius_p = P.c[3], radius_p_inv2 = P.c[4];
    double Vx = -1.0, Vy, Vz;
    if (P.flip_axis) {
        Vz = tan(y / rg1);
        Vy = tan(x / rg1) * hypot(1.0, Vz);
    } else {
        Vy = tan(x / rg1);
        Vz = tan(y / rg1) * hypot(1.0, Vy);
    }
    double a = Vz / radius_p;
    a = Vy * Vy + a * a + Vx * Vx;
    double b = 2 * rg * Vx;
    double det = b * b - 4 * a * C;
    if (det < 0.) return false;
    double k = (-b - sqrt(det)) / (2. * a);
    Vx = rg + k * Vx;
    Vy *= k;
    Vz *= k;
    lam = atan2(Vy, Vx);
    phi = atan(Vz * cos(lam) / Vx);
    if (P.ellipsoidal) phi = atan(radius_p_inv2 * tan(phi));
    return true;
}

__device__ inline bool inv_merc(const prs_proj_params &P, double x, double y, double &lam, double &phi)
{
    const double k0 = P.c[0];
    lam = x / k0;
    if (!P.ellipsoidal) {
        phi = atan(sinh(y / k0));
        return true;
    }
    double ts = exp(-y / k0);
    double eccnth = .5 * P.e;
    double p = PRS_HALFPI - 2. * atan(ts);
    for (int i = 0; i < 15; ++i) {
        double con = P.e * sin(p);
        double dphi = PRS_HALFPI - 2. * atan(ts * pow_ratio(-con, eccnth, P.e)) - p;
        p += dphi;
        if (fabs(dphi) <= 1e-10) break;
    }
    phi = p;
    return true;
}

// Projected (x, y) in the CRS's own units -> geodetic lon/lat in degrees (double).
// Off-projection points give +inf for both, like pyproj with errcheck=False.
static __device__ __noinline__ void proj_inverse(const prs_proj_params &P, double x, double y, double &lon, double &lat)
{
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    if (P.kind == PRS_PROJ_LONGLAT) {
        lon = x + P.pm_deg;
        // a prime-meridian shift is wrapped back into [-180, 180] (PROJ's adjlon; test_resamplers/test_nearest.py:130-143)
        if (P.pm_deg != 0.0 && !P.over) lon = lon > 180.0 ? lon - 360.0 : (lon < -180.0 ? lon + 360.0 : lon);
        lat = y;
        return;
    }
    if (!isfinite(x) || !isfinite(y)) {
        lon = lat = INF;
        return;
    }
    // pj_inv: descale, no contraction (PROJ is built for generic x86-64, which has no FMA)
    double xs = __dmul_rn(__dsub_rn(__dmul_rn(x, P.to_meter), P.x0), P.ra);
    double ys = __dmul_rn(__dsub_rn(__dmul_rn(y, P.to_meter), P.y0), P.ra);
    double lam = 0., phi = 0.;
    bool ok = true;
    switch (P.kind) {
    case PRS_PROJ_EQC:
        lam = xs / P.c[0];
        phi = ys + P.phi0;
        break;
    case PRS_PROJ_MERC:
        ok = inv_merc(P, xs, ys, lam, phi);
        break;
    case PRS_PROJ_LAEA:
        ok = inv_laea(P, xs, ys, lam, phi);
        break;
    case PRS_PROJ_STERE:
        ok = inv_stere(P, xs, ys, lam, phi);
        break;
    case PRS_PROJ_GEOS:
        ok = inv_geos(P, xs, ys, lam, phi);
        break;
    default:
        ok = false;
    }
    if (!ok) {
        lon = lat = INF;
        return;
    }
    lam += P.lam0;
    if (!P.over) lam = adjlon(lam);
    lon = __dadd_rn(__dmul_rn(lam, PRS_RAD_TO_DEG), P.pm_deg);
    lat = __dmul_rn(phi, PRS_RAD_TO_DEG);
}

// Pixel-centre projection coordinates, evaluated in the requested dtype exactly like
// ``np.arange(n, dtype) * pixel_size + offset`` with Python-float scalars (geometry.py:1378-1379).
template <typename T> __device__ __forceinline__ T pixel_x(const prs_area_grid &g, int64_t col)
{
    return Ar<T>::add(Ar<T>::mul((T)col, (T)g.pixel_size_x), (T)g.pixel_upper_left_x);
}
template <typename T> __device__ __forceinline__ T pixel_y(const prs_area_grid &g, int64_t row)
{
    return Ar<T>::add(Ar<T>::mul((T)row, (T)(-g.pixel_size_y)), (T)g.pixel_upper_left_y);
}

// lon/lat of one area pixel in dtype T (geometry.py:2637-2638: results are cast to the requested dtype).
template <typename T>
__device__ __forceinline__ void area_pixel_lonlat(const prs_proj_params &P, const prs_area_grid &g, int64_t row,
                                                  int64_t col, T &lon, T &lat)
{
    double dlon, dlat;
    proj_inverse(P, (double)pixel_x<T>(g, col), (double)pixel_y<T>(g, row), dlon, dlat);
    lon = (T)dlon;
    lat = (T)dlat;
}

}  // namespace prs
