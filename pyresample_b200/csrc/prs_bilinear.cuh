// Bilinear resampling on the neighbour index: the per-pixel epilogue of pyresample's
// BilinearBase.get_bil_info (bilinear/_base.py:99-119) behind a k = 32 search - forward projection of the source
// points into the target CRS (the `Proj(target.proj_str)(lon, lat)` of _base.py:188-191, 299-307), the closest
// neighbour in each quadrant (:391-411, 573-611), the fractional distances t and s (:414-570) - and the sampling
// p1 (1-s)(1-t) + p2 s (1-t) + p3 (1-s) t + p4 s t (:653-660).
#pragma once
#include "prs_common.cuh"
#include "prs_proj.cuh"

namespace prs {

// ------------------------------------------------------------------------------------------ forward projections
// PROJ's forward formulas (src/projections/{eqc,merc,laea,stere,geos}.cpp) with the constants prepared in
// pyresample_b200/proj.py for the inverses.  lam relative to lon_0 (radians), phi (radians) -> x, y in units of a.
__device__ inline bool fwd_laea(const prs_proj_params &P, double lam, double phi, double &x, double &y)
{
    double sinlam, coslam;
    sincos(lam, &sinlam, &coslam);
    const double sinphi = sin(phi);
    if (!P.ellipsoidal) {
        const double cosphi = cos(phi), sinb1 = P.c[0], cosb1 = P.c[1];
        if (P.mode == PRS_MODE_EQUIT || P.mode == PRS_MODE_OBLIQ) {
            y = P.mode == PRS_MODE_EQUIT ? 1. + cosphi * coslam : 1. + sinb1 * sinphi + cosb1 * cosphi * coslam;
            if (y <= PRS_EPS10) return false;
            y = sqrt(2. / y);
            x = y * cosphi * sinlam;
            y *= P.mode == PRS_MODE_EQUIT ? sinphi : cosb1 * sinphi - sinb1 * cosphi * coslam;
            return true;
        }
        if (P.mode == PRS_MODE_N_POLE) coslam = -coslam;
        if (fabs(phi + P.phi0) < PRS_EPS10) return false;
        y = 0.78539816339744833 - phi * .5;
        y = 2. * (P.mode == PRS_MODE_S_POLE ? cos(y) : sin(y));
        x = y * sinlam;
        y *= coslam;
        return true;
    }
    const double sinb1 = P.c[0], cosb1 = P.c[1], qp = P.c[2], rq = P.c[3], dd = P.c[4];
    const double con = P.e * sinphi;
    double q = P.one_es * (sinphi / (1. - con * con) - (.5 / P.e) * log((1. - con) / (1. + con)));
    if (P.mode == PRS_MODE_EQUIT || P.mode == PRS_MODE_OBLIQ) {
        const double sinb = q / qp, cosb2 = 1. - sinb * sinb, cosb = cosb2 > 0 ? sqrt(cosb2) : 0.;
        double b = P.mode == PRS_MODE_OBLIQ ? 1. + sinb1 * sinb + cosb1 * cosb * coslam : 1. + cosb * coslam;
        if (fabs(b) < PRS_EPS10) return false;
        b = sqrt(2. / b);
        const double xmf = P.mode == PRS_MODE_OBLIQ ? rq * dd : 1., ymf = P.mode == PRS_MODE_OBLIQ ? rq / dd : .5 * qp;
        y = P.mode == PRS_MODE_OBLIQ ? ymf * b * (cosb1 * sinb - sinb1 * cosb * coslam) : b * sinb * ymf;
        x = xmf * b * cosb * sinlam;
        return true;
    }
    double b;
    if (P.mode == PRS_MODE_N_POLE) {
        b = PRS_HALFPI + phi;
        q = qp - q;
    } else {
        b = phi - PRS_HALFPI;
        q = qp + q;
    }
    if (fabs(b) < PRS_EPS10) return false;
    if (q >= 1e-15) {
        b = sqrt(q);
        x = b * sinlam;
        y = coslam * (P.mode == PRS_MODE_S_POLE ? b : -b);
    } else {
        x = y = 0.;
    }
    return true;
}

__device__ inline bool fwd_stere(const prs_proj_params &P, double lam, double phi, double &x, double &y)
{
    double sinlam, coslam;
    sincos(lam, &sinlam, &coslam);
    double sinphi = sin(phi);
    const double akm1 = P.c[0], sinX1 = P.c[1], cosX1 = P.c[2];
    if (P.ellipsoidal) {
        if (P.mode == PRS_MODE_N_POLE || P.mode == PRS_MODE_S_POLE) {
            if (P.mode == PRS_MODE_S_POLE) {
                phi = -phi;
                coslam = -coslam;
                sinphi = -sinphi;
            }
            const double sp = sinphi * P.e;
            x = fabs(phi - PRS_HALFPI) < 1e-15 ? 0. : akm1 * (tan(.5 * (PRS_HALFPI - phi)) / pow((1. - sp) / (1. + sp), .5 * P.e));
            y = -x * coslam;
            x *= sinlam;
            return true;
        }
        const double sp = sinphi * P.e;
        const double X = 2. * atan(tan(.5 * (PRS_HALFPI + phi)) * pow((1. - sp) / (1. + sp), .5 * P.e)) - PRS_HALFPI;
        double sinX, cosX;
        sincos(X, &sinX, &cosX);
        double A;
        if (P.mode == PRS_MODE_OBLIQ) {
            const double denom = cosX1 * (1. + sinX1 * sinX + cosX1 * cosX * coslam);
            if (denom == 0.) return false;
            A = akm1 / denom;
            y = A * (cosX1 * sinX - sinX1 * cosX * coslam);
        } else {
            const double denom = 1. + cosX * coslam;
            if (denom == 0.) return false;
            A = akm1 / denom;
            y = A * sinX;
        }
        x = A * cosX * sinlam;
        return true;
    }
    const double cosphi = cos(phi);
    if (P.mode == PRS_MODE_OBLIQ || P.mode == PRS_MODE_EQUIT) {
        y = P.mode == PRS_MODE_EQUIT ? 1. + cosphi * coslam : 1. + sinX1 * sinphi + cosX1 * cosphi * coslam;
        if (y <= PRS_EPS10) return false;
        y = akm1 / y;
        x = y * cosphi * sinlam;
        y *= P.mode == PRS_MODE_EQUIT ? sinphi : cosX1 * sinphi - sinX1 * cosphi * coslam;
        return true;
    }
    if (P.mode == PRS_MODE_N_POLE) {
        coslam = -coslam;
        phi = -phi;
    }
    if (fabs(phi - PRS_HALFPI) < 1e-8) return false;
    y = akm1 * tan(0.78539816339744833 + .5 * phi);
    x = sinlam * y;
    y *= coslam;
    return true;
}

__device__ inline bool fwd_geos(const prs_proj_params &P, double lam, double phi, double &x, double &y)
{
    const double rg1 = P.c[0], rg = P.c[1];
    double r = 1., inv2 = 1.;
    if (P.ellipsoidal) {
        const double radius_p = P.c[3];
        inv2 = P.c[4];
        phi = atan(P.one_es * tan(phi));
        r = radius_p / hypot(radius_p * cos(phi), sin(phi));
    }
    const double Vx = r * cos(lam) * cos(phi), Vy = r * sin(lam) * cos(phi), Vz = r * sin(phi);
    if (((rg - Vx) * Vx - Vy * Vy - Vz * Vz * inv2) < 0.) return false;
    const double tmp = rg - Vx;
    if (P.flip_axis) {
        x = rg1 * atan(Vy / hypot(Vz, tmp));
        y = rg1 * atan(Vz / tmp);
    } else {
        x = rg1 * atan(Vy / tmp);
        y = rg1 * atan(Vz / hypot(Vy, tmp));
    }
    return true;
}

// pj_fwd: (lon, lat) degrees -> projected coordinates; (inf, inf) on failure or non-finite input, like
// pyproj.Proj(...)(lon, lat) with errcheck off.
static __device__ __noinline__ void proj_forward(const prs_proj_params &P, double lon, double lat, double &x, double &y)
{
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    if (P.kind == PRS_PROJ_LONGLAT) {
        x = lon - P.pm_deg;
        y = lat;
        return;
    }
    x = y = INF;
    if (!isfinite(lon) || !isfinite(lat)) return;
    double lam = (lon - P.pm_deg) * 0.017453292519943296 - P.lam0;
    const double phi = lat * 0.017453292519943296;
    if (fabs(phi) - PRS_HALFPI > 1e-12) return;
    if (!P.over) lam = adjlon(lam);
    double px = 0., py = 0.;
    bool ok = true;
    switch (P.kind) {
    case PRS_PROJ_EQC:
        px = P.c[0] * lam;
        py = phi - P.phi0;
        break;
    case PRS_PROJ_MERC:
        ok = !(fabs(fabs(phi) - PRS_HALFPI) <= PRS_EPS10);
        px = P.c[0] * lam;
        py = P.c[0] * (asinh(tan(phi)) - (P.ellipsoidal ? P.e * atanh(P.e * sin(phi)) : 0.));
        break;
    case PRS_PROJ_LAEA:
        ok = fwd_laea(P, lam, phi, px, py);
        break;
    case PRS_PROJ_STERE:
        ok = fwd_stere(P, lam, phi, px, py);
        break;
    case PRS_PROJ_GEOS:
        ok = fwd_geos(P, lam, phi, px, py);
        break;
    default:
        ok = false;
    }
    if (!ok) return;
    x = (P.a * px + P.x0) / P.to_meter;
    y = (P.a * py + P.y0) / P.to_meter;
}

template <typename T>
__global__ void project_forward_kernel(const T *__restrict__ lon, const T *__restrict__ lat, int64_t n,
                                       const __grid_constant__ prs_proj_params P, double2 *__restrict__ xy)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x, y;
    proj_forward(P, (double)lon[i], (double)lat[i], x, y);
    xy[i] = make_double2(x, y);
}

// ------------------------------------------------------------------------------------------ fractional distances
// The reference's numpy expressions, one pixel at a time, same operations in the same order (double, no contraction).
namespace bil {
__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dv(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ bool outside(double v, double lo, double hi) { return v < lo || v > hi; }  // NaN: false
__device__ __forceinline__ double nan_() { return __longlong_as_double(0x7ff8000000000000LL); }

struct Pt {
    double x, y;
};

// _calc_abc (_base.py:470-497)
__device__ __forceinline__ void calc_abc(Pt p1, Pt p2, Pt p3, Pt p4, double out_y, double out_x, double &a, double &b, double &c)
{
    const double x_21 = sub(p2.x, p1.x), x_31 = sub(p3.x, p1.x), x_42 = sub(p4.x, p2.x);
    const double y_21 = sub(p2.y, p1.y), y_31 = sub(p3.y, p1.y), y_42 = sub(p4.y, p2.y);
    a = sub(mul(x_31, y_42), mul(y_31, x_42));
    b = sub(add(sub(add(sub(mul(out_y, sub(x_42, x_31)), mul(out_x, sub(y_42, y_31))), mul(x_31, p2.y)), mul(y_31, p2.x)),
                mul(y_42, p1.x)),
            mul(x_42, p1.y));
    c = sub(add(sub(mul(out_y, x_21), mul(out_x, y_21)), mul(p1.x, p2.y)), mul(p2.x, p1.y));
}

// _solve_quadratic (_base.py:431-460) on [0, 1]
__device__ __forceinline__ double solve_quadratic(double a, double b, double c)
{
    const double disc = sub(mul(b, b), mul(mul(4., a), c));
    const double sq = sqrt(disc);  // NaN for a negative discriminant, like np.sqrt
    const double x_1 = dv(add(-b, sq), mul(2., a)), x_2 = dv(sub(-b, sq), mul(2., a)), x_3 = dv(-c, b);
    double x = (outside(x_1, 0., 1.) || isnan(x_1)) ? x_2 : x_1;
    x = (outside(x, 0., 1.) || isnan(x)) ? x_3 : x;
    return outside(x, 0., 1.) ? nan_() : x;
}

// _solve_another_fractional_distance (_base.py:500-517)
__device__ __forceinline__ double other_distance(double f, double y_1, double y_2, double y_3, double y_4, double out_y)
{
    const double y_21 = sub(y_2, y_1), y_43 = sub(y_4, y_3);
    const double g = dv(sub(sub(out_y, y_1), mul(y_21, f)), sub(sub(add(y_3, mul(y_43, f)), y_1), mul(y_21, f)));
    return outside(g, 0., 1.) ? nan_() : g;
}

__device__ __forceinline__ void invalid_to_nan(double &t, double &s)
{
    if (outside(t, 0., 1.) || outside(s, 0., 1.)) t = s = nan_();
}

// _get_fractional_distances (_base.py:395-411)
__device__ __forceinline__ void fractional_distances(Pt p1, Pt p2, Pt p3, Pt p4, double out_x, double out_y, double &t, double &s)
{
    double a, b, c;
    calc_abc(p1, p2, p3, p4, out_y, out_x, a, b, c);
    t = solve_quadratic(a, b, c);
    s = other_distance(t, p1.y, p3.y, p2.y, p4.y, out_y);
    invalid_to_nan(t, s);
    if (isnan(t) || isnan(s)) {  // uprights parallel (:529-543)
        calc_abc(p1, p3, p2, p4, out_y, out_x, a, b, c);
        s = solve_quadratic(a, b, c);
        t = other_distance(s, p1.y, p2.y, p3.y, p4.y, out_y);
        invalid_to_nan(t, s);
    }
    if (isnan(t) || isnan(s)) {  // parallelogram (:546-570)
        const double x_21 = sub(p2.x, p1.x), x_31 = sub(p3.x, p1.x), y_21 = sub(p2.y, p1.y), y_31 = sub(p3.y, p1.y);
        t = dv(sub(mul(x_21, sub(out_y, p1.y)), mul(y_21, sub(out_x, p1.x))), sub(mul(x_21, y_31), mul(y_21, x_31)));
        if (outside(t, 0., 1.)) t = nan_();
        s = dv(add(sub(out_x, p1.x), mul(x_31, t)), x_21);
        if (outside(s, 0., 1.)) s = nan_();
        invalid_to_nan(t, s);
    }
}
}  // namespace bil

struct BilInfoParams {
    const uint32_t *index;  // [n_t * k] ranks from prs_query, sentinel n_valid
    int64_t n_t;
    int k;
    uint32_t n_valid;
    const uint32_t *rank_to_source;  // NULL = identity
    const double2 *xy;               // projected source coordinates by raveled source index
    prs_area_grid grid;
    double *t_out, *s_out;  // [n_t]
    uint32_t *idx4_out;     // [n_t * 4] ranks of the four corners
};

// One thread per target pixel: closest neighbour per quadrant (neighbours arrive sorted by distance), then t and s.
static __global__ void bilinear_info_kernel(const __grid_constant__ BilInfoParams P)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.n_t) return;
    const int64_t W = P.grid.width;
    const int64_t row = p / W, col = p - row * W;
    const double out_x = pixel_x<double>(P.grid, col), out_y = pixel_y<double>(P.grid, row);
    const uint32_t *idx = P.index + (size_t)p * P.k;
    bil::Pt pt[4];
    uint32_t rk[4];
    bool found[4] = {false, false, false, false};
    const double NaN = bil::nan_();
#pragma unroll
    for (int q = 0; q < 4; ++q) pt[q].x = pt[q].y = NaN;
    uint32_t first_rank = 0;
    for (int i = 0; i < P.k; ++i) {
        uint32_t r = idx[i];
        if (r >= P.n_valid) r = 0;  // "no neighbour" entries point at element 0 (_base.py:158-162)
        if (i == 0) first_rank = r;
        const uint32_t src = P.rank_to_source ? __ldg(P.rank_to_source + r) : r;
        const double2 v = P.xy[src];
        const double xd = out_x - v.x, yd = out_y - v.y;
        // upper left, upper right, lower left, lower right (_base.py:573-586); NaN / inf never compare true
        const bool in_q[4] = {xd > 0 && yd < 0, xd < 0 && yd < 0, xd > 0 && yd > 0, xd < 0 && yd > 0};
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (in_q[q] && !found[q]) {
                found[q] = true;
                pt[q].x = v.x, pt[q].y = v.y;
                rk[q] = r;
            }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) P.idx4_out[(size_t)p * 4 + q] = found[q] ? rk[q] : first_rank;  // argmax of all-False is 0
    double t, s;
    bil::fractional_distances(pt[0], pt[1], pt[2], pt[3], out_x, out_y, t, s);
    P.t_out[p] = t;
    P.s_out[p] = s;
}

struct BilSampleParams {
    const uint32_t *idx4;  // [n_out * 4] ranks
    const double *t, *s;   // [n_out]
    int64_t n_out;
    const uint32_t *rank_to_source;
    const void *data;  // [n_source * channels]
    int channels;
    double fill;  // replaces NaN results (NaN: keep them)
    double *out;  // [n_out * channels]
};

// _resample (_base.py:653-660) in float64, then NaN -> fill (_numpy_resampler.py:276-282)
template <typename Td> __global__ void bilinear_sample_kernel(const __grid_constant__ BilSampleParams P)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int C = P.channels;
    const int64_t p = i / C;
    if (p >= P.n_out) return;
    const int c = (int)(i - p * C);
    const double t = P.t[p], s = P.s[p];
    double v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t r = P.idx4[(size_t)p * 4 + q];
        const uint32_t src = P.rank_to_source ? __ldg(P.rank_to_source + r) : r;
        v[q] = (double)((const Td *)P.data)[(size_t)src * C + c];
    }
    using namespace bil;
    const double one_s = sub(1., s), one_t = sub(1., t);
    double r = add(add(add(mul(mul(v[0], one_s), one_t), mul(mul(v[1], s), one_t)), mul(mul(v[2], one_s), t)), mul(mul(v[3], s), t));
    if (isnan(r) && !isnan(P.fill)) r = P.fill;
    P.out[(size_t)p * C + c] = r;
}

}  // namespace prs
