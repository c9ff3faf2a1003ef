// Common device/host helpers for the B200 kd-tree resampling kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>

#include "../../include/prs_b200.h"

#define PRS_R 6370997.0  // pyresample/_spatial_mp.py:39, future/resamplers/_transform_utils.py:24
#define PRS_NONE 0xFFFFFFFFu

namespace prs {

// ------------------------------------------------------------------------------------------ errors
inline std::string &err_slot()
{
    static thread_local std::string s;
    return s;
}
inline int fail(int code, const std::string &msg)
{
    err_slot() = msg;
    return code;
}
#define PRS_CK(call)                                                                                     \
    do {                                                                                                 \
        cudaError_t _e = (call);                                                                         \
        if (_e != cudaSuccess) {                                                                         \
            char _b[512];                                                                                \
            snprintf(_b, sizeof(_b), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return prs::fail(PRS_ECUDA, _b);                                                             \
        }                                                                                                \
    } while (0)
inline long long &launch_counter()
{
    static long long n = 0;  // kernels launched by this library (bench.py's gpu_launches); not thread-exact
    return n;
}
#define PRS_CKL()                      \
    do {                               \
        ++prs::launch_counter();       \
        PRS_CK(cudaGetLastError());    \
    } while (0)

// ------------------------------------------------------------------------------------------ arithmetic
// All distance / ECEF arithmetic uses explicit round-to-nearest intrinsics so that nvcc never contracts
// a multiply-add: the reference chain (numpy ufuncs, pykdtree's C loop) rounds after every operation.
template <typename T> struct Ar;
template <> struct Ar<float> {
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    static __device__ __forceinline__ float sqrt(float a) { return __fsqrt_rn(a); }
    static __device__ __forceinline__ float inf() { return __int_as_float(0x7f800000); }
    static __device__ __forceinline__ float nan() { return __int_as_float(0x7fc00000); }
    static __device__ __forceinline__ float exp(float a) { return expf(a); }
};
template <> struct Ar<double> {
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
    static __device__ __forceinline__ double sqrt(double a) { return __dsqrt_rn(a); }
    static __device__ __forceinline__ double inf() { return __longlong_as_double(0x7ff0000000000000LL); }
    static __device__ __forceinline__ double nan() { return __longlong_as_double(0x7ff8000000000000LL); }
    static __device__ __forceinline__ double exp(double a) { return ::exp(a); }
};

// numpy's float32 sin/cos (AVX2/AVX512F/FMA3 SIMD loop, numpy/_core/src/umath/loops_trigonometric.dispatch.*):
// Cody-Waite 3-term reduction by pi/2, degree-9 / degree-8 minimax polynomials, all in FMA float32.
// Reproduced operation for operation so the float32 ECEF chain of the reference (np.cos(np.deg2rad(f32)),
// SURVEY 9.3) is bit-identical on x86 hosts with FMA.  Valid for |x| < 71476 (lon/lat radians are < 4).
__device__ __forceinline__ void np_sincosf(float x, float &s_out, float &c_out)
{
    const float two_over_pi = 0x1.45f306p-1f;
    float q = rintf(__fmul_rn(x, two_over_pi));
    float r = __fmaf_rn(q, -0x1.921fb0p+00f, x);
    r = __fmaf_rn(q, -0x1.5110b4p-22f, r);
    r = __fmaf_rn(q, -0x1.846988p-48f, r);
    float r2 = __fmul_rn(r, r);
    float c = __fmaf_rn(0x1.98e616p-16f, r2, -0x1.6c06dcp-10f);
    c = __fmaf_rn(c, r2, 0x1.55553cp-05f);
    c = __fmaf_rn(c, r2, -0x1.000000p-01f);
    c = __fmaf_rn(c, r2, 1.0f);
    float s = __fmaf_rn(0x1.7d3bbcp-19f, r2, -0x1.a06bbap-13f);
    s = __fmaf_rn(s, r2, 0x1.11119ap-07f);
    s = __fmaf_rn(s, r2, -0x1.555556p-03f);
    s = __fmaf_rn(s, r2, 0.0f);
    s = __fmaf_rn(s, r, r);
    int iq = (int)q;
    float vs = (iq & 1) ? c : s;
    if (iq & 2) vs = __fsub_rn(0.0f, vs);
    int iqc = iq + 1;
    float vc = (iqc & 1) ? c : s;
    if (iqc & 2) vc = __fsub_rn(0.0f, vc);
    s_out = vs;
    c_out = vc;
}

// lon/lat (degrees) -> ECEF on the R = 6370997 m sphere, in the dtype of the inputs:
//   x = (R*cos(lat))*cos(lon), y = (R*cos(lat))*sin(lon), z = R*sin(lat)   (_spatial_mp.py:166-168)
// np.deg2rad(f32) multiplies by float32(pi/180); np.deg2rad(f64) by pi/180.
__device__ __forceinline__ void trig_f32(float deg, float &s, float &c)
{
    np_sincosf(__fmul_rn(deg, 0x1.1df46ap-6f), s, c);
}
__device__ __forceinline__ void trig_f64(double deg, double &s, double &c)
{
    sincos(__dmul_rn(deg, 0.017453292519943295), &s, &c);
}
__device__ __forceinline__ void lonlat_to_xyz(float lon, float lat, float &x, float &y, float &z)
{
    float sl, cl, sp, cp;
    trig_f32(lon, sl, cl);
    trig_f32(lat, sp, cp);
    float rc = __fmul_rn(6370997.0f, cp);
    x = __fmul_rn(rc, cl);
    y = __fmul_rn(rc, sl);
    z = __fmul_rn(6370997.0f, sp);
}
__device__ __forceinline__ void lonlat_to_xyz(double lon, double lat, double &x, double &y, double &z)
{
    double sl, cl, sp, cp;
    trig_f64(lon, sl, cl);
    trig_f64(lat, sp, cp);
    double rc = __dmul_rn(6370997.0, cp);
    x = __dmul_rn(rc, cl);
    y = __dmul_rn(rc, sl);
    z = __dmul_rn(6370997.0, sp);
}
template <typename T> __device__ __forceinline__ bool lonlat_in_range(T lon, T lat)
{
    // kd_tree.py:406,454 - closed intervals; NaN / inf compare false
    return (lon >= (T)-180) && (lon <= (T)180) && (lat <= (T)90) && (lat >= (T)-90);
}

// ------------------------------------------------------------------------------------------ point records
// One indexed source point: ECEF coordinates in the tree dtype + its raveled source index.
// float: 16 B (one LDG.128); double: 32 B (two LDG.128).
template <typename T> struct Rec;
template <> struct __align__(16) Rec<float> {
    float x, y, z;
    uint32_t idx;
};
template <> struct __align__(16) Rec<double> {
    double x, y, z;
    uint32_t idx;
    uint32_t pad;
};
__device__ __forceinline__ Rec<float> load_rec(const Rec<float> *p)
{
    float4 v = __ldg(reinterpret_cast<const float4 *>(p));
    Rec<float> r;
    r.x = v.x;
    r.y = v.y;
    r.z = v.z;
    r.idx = __float_as_uint(v.w);
    return r;
}
__device__ __forceinline__ Rec<double> load_rec(const Rec<double> *p)
{
    const double2 *q = reinterpret_cast<const double2 *>(p);
    double2 a = __ldg(q);
    double2 b = __ldg(q + 1);
    Rec<double> r;
    r.x = a.x;
    r.y = a.y;
    r.z = b.x;
    r.idx = (uint32_t)(__double_as_longlong(b.y) & 0xFFFFFFFFLL);
    r.pad = 0;
    return r;
}
__device__ __forceinline__ void store_rec(Rec<float> *p, float x, float y, float z, uint32_t idx)
{
    *reinterpret_cast<float4 *>(p) = make_float4(x, y, z, __uint_as_float(idx));
}
__device__ __forceinline__ void store_rec(Rec<double> *p, double x, double y, double z, uint32_t idx)
{
    double2 *q = reinterpret_cast<double2 *>(p);
    q[0] = make_double2(x, y);
    q[1] = make_double2(z, __longlong_as_double((long long)idx));
}

// pykdtree's distance: d2 = ((dx*dx) + dy*dy) + dz*dz in the tree dtype, no FMA (oracle/knn_ref.py).
template <typename T> __device__ __forceinline__ T dist2(T qx, T qy, T qz, T px, T py, T pz)
{
    T dx = Ar<T>::sub(qx, px), dy = Ar<T>::sub(qy, py), dz = Ar<T>::sub(qz, pz);
    return Ar<T>::add(Ar<T>::add(Ar<T>::mul(dx, dx), Ar<T>::mul(dy, dy)), Ar<T>::mul(dz, dz));
}

// ------------------------------------------------------------------------------------------ cube-face cells
// Source points live on a sphere, so the index is a 2-D grid on the six faces of the circumscribed cube
// (gnomonic chart + a monotone equal-angle-like warp), G x G cells per face, stored row-major per face:
//     cell = face*G*G + iv*G + iu
// A query cap maps to an axis-aligned (u, v) rectangle per face in closed form (bounding box of a conic),
// and one row of that rectangle is ONE contiguous run of point records.
__device__ __forceinline__ void face_uv(float x, float y, float z, int &face, float &u, float &v)
{
    float ax = fabsf(x), ay = fabsf(y), az = fabsf(z);
    if (ax >= ay && ax >= az) {
        float inv = 1.0f / ax;
        face = (x < 0.f) ? 1 : 0;
        u = y * inv;
        v = z * inv;
    } else if (ay >= az) {
        float inv = 1.0f / ay;
        face = (y < 0.f) ? 3 : 2;
        u = x * inv;
        v = z * inv;
    } else {
        float inv = 1.0f / az;
        face = (z < 0.f) ? 5 : 4;
        u = x * inv;
        v = y * inv;
    }
}
__device__ __forceinline__ float warp_uv(float u)
{
    u = fminf(fmaxf(u, -1.0f), 1.0f);
    return u * (1.27324f - 0.27324f * fabsf(u));  // ~ (4/pi)*atan(u): evens out the cell footprint
}
__device__ __forceinline__ int cell_coord(float w, int G)
{
    int i = (int)floorf((w + 1.0f) * 0.5f * (float)G);
    return min(max(i, 0), G - 1);
}

}  // namespace prs
